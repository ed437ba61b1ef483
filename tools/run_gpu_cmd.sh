#!/bin/bash
# gpurun helper: run the given commands, logging each to gpurun_out/<name>.log.  Usage: run_gpu_cmd.sh name 'cmd' [name 'cmd' ...]
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  name=$1; cmd=$2; shift 2
  echo "== $name: $cmd"
  timeout ${STEP_TIMEOUT:-900} bash -c "$cmd" > gpurun_out/$name.log 2>&1; echo "rc=$?"
  tail -${TAIL:-25} gpurun_out/$name.log
done
