#!/bin/bash
# round 2, call H (2 GPUs): full bench at N=1 and N=2 + the 2-rank spawn tests
mkdir -p gpurun_out
echo "== bench N=1"; timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/bench_n1.err
echo "== bench N=2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n2.err | tail -5
echo "== ref arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json | cut -c1-300
python - <<'PY'
import json
def load(p):
    return json.loads([l for l in open(p) if l.startswith('{')][-1])
d=load('gpurun_out/bench_n1.json')
print('N=1 value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'columnar', d['e2e'].get('columnar_host_table'))
print('cpu', d.get('cpu_baseline'))
for k,v in d['queries'].items():
    if 'error' in v: print(k, v['error'][:200]); continue
    print(k, round(v['ms'],2), 'ms', round(v['rows_per_s']/1e9,1),'Grows/s frac', round(v['roofline']['frac'],3), v['check_ok'], {a:round(b['rows_per_s']/1e6) for a,b in (v.get('cpu_baseline') or {}).items() if isinstance(b,dict) and 'rows_per_s' in b})
d=load('gpurun_out/bench_n2.json')
print('N=2 value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('columnar_host_table',{}).get('value'))
for arm,res in d['queries'].items():
    if not isinstance(res, dict): continue
    print('==', arm, res.get('exchange'), res.get('parity_ok'))
    for k,v in res.items():
        if isinstance(v, dict) and 'ms' in v: print('  ', k, round(v['ms'],2), 'ms', round(v['rows_per_s']/1e9,1), 'Grows/s', v['check_ok'], v.get('phases_ms'))
        elif isinstance(v, dict) and 'error' in v: print('  ', k, v['error'][:300])
PY
echo "== spawn tests"; timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -p no:cacheprovider -k "two_ranks" > gpurun_out/pytest_spawn.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_spawn.log
