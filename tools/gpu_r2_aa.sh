#!/bin/bash
# round 2, call AA: ncu --set full of the tie-repair scan (hk_sort_fix_find_kernel) at 0.5e9 rows
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hk_sort_fix" -s 2 -c 2 -f -o gpurun_out/r02_fix python tools/ops_bench.py --ops orderby --scale 0.25 --reps 1 > gpurun_out/ncu_fix.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_fix.ncu-rep > gpurun_out/r02_fix_find_ncu.txt 2>&1; cat gpurun_out/r02_fix_find_ncu.txt
ncu -i gpurun_out/r02_fix.ncu-rep --page source --csv > gpurun_out/fix_sass.csv 2>/dev/null; ls -la gpurun_out/fix_sass.csv
rm -f gpurun_out/r02_fix.ncu-rep
