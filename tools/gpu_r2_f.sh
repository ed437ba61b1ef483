#!/bin/bash
# round 2, call F (2 GPUs): bench.py at N=2 through ShardedEnv (weak + strong, K8c + NCCL, parity_ok) and the sharded GPU tests
mkdir -p gpurun_out
echo "== bench N=2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print('value', d['value'], 'e2e', d['e2e']['value'])
for arm,res in d['queries'].items():
    if not isinstance(res, dict): continue
    print('==', arm, res.get('exchange'), res.get('parity_ok'))
    for k,v in res.items():
        if isinstance(v, dict) and 'ms' in v: print('  ', k, round(v['ms'],2), 'ms', round(v['rows_per_s']/1e9,1), 'Grows/s', v['check_ok'], v.get('phases_ms'))
        elif isinstance(v, dict) and 'error' in v: print('  ', k, v['error'][:300])
PY
echo "== sharded tests N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m pytest tests/test_gpu_sharded.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_sharded_n2.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_sharded_n2.log
