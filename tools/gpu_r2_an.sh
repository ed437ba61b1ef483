#!/bin/bash
# round 2, call AN: ncu --set full of the two join expand kernels
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hk_hj_expand_kernel|hk_hj_count_kernel" -s 2 -c 2 -f -o gpurun_out/r02_jx python tools/ops_bench.py --ops join_hash --reps 1 > gpurun_out/ncu_jx.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_jx.ncu-rep > gpurun_out/r02_join_expand_ncu.txt 2>&1; cat gpurun_out/r02_join_expand_ncu.txt
ncu -i gpurun_out/r02_jx.ncu-rep --page source --csv > gpurun_out/jx_sass.csv 2>/dev/null; ls -la gpurun_out/jx_sass.csv
rm -f gpurun_out/r02_jx.ncu-rep
