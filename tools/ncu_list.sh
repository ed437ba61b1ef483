#!/bin/bash
# usage: tools/ncu_list.sh <tag> <python command...>  -> gpurun_out/launches_<tag>.csv + a compact per-kernel listing on stdout
tag=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv "$@" > gpurun_out/ncu_list_$tag.out 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$tag.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows:
    if "at::" not in r[4] and "synth" not in r[4]: print(r[4][:84], r[7], r[8], round(float(r[-1])/1e6,4), "ms")
PY
