#!/bin/bash
# round 2, call D: ncu --set full of the two K2 kernels (config 3); text summaries are written on the box as well
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hk_tpart_kernel|hk_dagg_tiles_kernel" -s 4 -c 2 -f -o gpurun_out/r02_k2_tiles python tools/ops_bench.py --ops groupby --reps 1 > gpurun_out/ncu_k2.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_k2_tiles.ncu-rep > gpurun_out/r02_k2_tiles_ncu.txt 2>&1
cat gpurun_out/r02_k2_tiles_ncu.txt
ls -la gpurun_out/
