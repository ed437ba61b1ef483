#!/bin/bash
# round 2, call X: ncu --set full of the chunked sweep16 pass and its histogram (0.5e9 rows), with the source page hot spots
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hk_sweep16" -s 10 -c 3 -f -o gpurun_out/r02_sweep16c python tools/ops_bench.py --ops orderby --scale 0.25 --reps 1 > gpurun_out/ncu_sweepc.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_sweep16c.ncu-rep > gpurun_out/r02_sweep16_chunked_ncu.txt 2>&1; cat gpurun_out/r02_sweep16_chunked_ncu.txt
ncu -i gpurun_out/r02_sweep16c.ncu-rep --page source --csv --print-source cuda > gpurun_out/sweep16c_source.csv 2>/dev/null; ls -la gpurun_out/sweep16c_source.csv
rm -f gpurun_out/r02_sweep16c.ncu-rep
