#!/bin/bash
# round 2, call AR (2 GPUs): sharded GPU tests (spawned 2-rank run, both exchange paths) + bench.py at N=2 as the driver launches it
mkdir -p gpurun_out
echo "== sharded tests"; timeout 1200 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_sharded.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_sharded.log | cut -c1-250
S=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$? wall $(( $(date +%s) - S )) s"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n2.err | tail -5
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1])
print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('columnar_host_table',{}).get('value'))
for arm,res in d['queries'].items():
    if not isinstance(res, dict): continue
    print('==', arm, res.get('exchange'), res.get('parity_ok'))
    for k,v in res.items():
        if isinstance(v, dict) and 'ms' in v: print('  ', k, round(v['ms'],2), 'ms', v['ms_all'], round(v['rows_per_s']/1e9,1), 'Grows/s', v['check_ok'], v.get('phases_ms'))
        elif isinstance(v, dict) and 'error' in v: print('  ', k, v['error'][:300])
PY
