#!/bin/bash
# round 2, call S (8 GPUs): bench.py at N=4 exactly as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L | head -8
S=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "rc=$? wall $(( $(date +%s) - S )) s"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n4.err | tail -5
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n4.json') if l.startswith('{')][-1])
print('N=4 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('columnar_host_table',{}).get('value'))
for arm,res in d['queries'].items():
    if not isinstance(res, dict): continue
    print('==', arm, res.get('exchange'), res.get('parity_ok'))
    for k,v in res.items():
        if isinstance(v, dict) and 'ms' in v: print('  ', k, round(v['ms'],2), 'ms', round(v['rows_per_s']/1e9,1), 'Grows/s', v['check_ok'], v.get('phases_ms'), (v.get('nvlink') or {}).get('gbs_per_gpu_per_direction'))
        elif isinstance(v, dict) and 'error' in v: print('  ', k, v['error'][:300])
PY
