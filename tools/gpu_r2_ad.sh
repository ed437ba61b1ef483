#!/bin/bash
# round 2, call AD: ncu --set full of K2 over K8t tiles on the Zipf GROUP BY (0.25e9 rows)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hk_dagg_tiles" -s 1 -c 1 -f -o gpurun_out/r02_zipf python tools/ops_bench.py --ops groupby_zipf --scale 0.25 --reps 1 > gpurun_out/ncu_zipf.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_zipf.ncu-rep > gpurun_out/r02_zipf_dagg_ncu.txt 2>&1; cat gpurun_out/r02_zipf_dagg_ncu.txt
ncu -i gpurun_out/r02_zipf.ncu-rep --page source --csv > gpurun_out/zipf_sass.csv 2>/dev/null; ls -la gpurun_out/zipf_sass.csv
rm -f gpurun_out/r02_zipf.ncu-rep
