#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/sweep.txt
import torch
a=torch.empty(1<<30,dtype=torch.bfloat16,device='cuda'); b=torch.empty_like(a)
best=0
for i in range(12):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best=max(best, 2*a.numel()*2/(e0.elapsed_time(e1)*1e-3)/1e9)
print(f"box copy peak (torch copy_ 2 GiB): {best:.0f} GB/s")
PY
run() { echo -n "$1 impl=$2: "; HARK_LIB=$PWD/$1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --filter-impl $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(f\"{d['value']/1e9:.1f} Grows/s kernel {d['roofline']['kernel_ms']:.3f} ms {d['roofline']['achieved']:.0f} GB/s frac {d['roofline']['frac']:.3f} clocks {d['clocks']}\")"; }
(for lib in harkdb_b200/libhark.so harkdb_b200/libhark_*.so; do run $lib 0; done) 2>&1 | tee -a gpurun_out/sweep.txt
