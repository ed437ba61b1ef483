#!/bin/bash
# round 2, call AV: ncu --set full of the FINAL K8t + K2 pair on config 3 (dealt units, first bin by bin share, 3 CTAs per SM)
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none -k regex:"hk_tpart_kernel|hk_dagg_tiles_kernel" -s 2 -c 2 -f -o gpurun_out/r02_k2f python tools/ops_bench.py --ops groupby --reps 1 > gpurun_out/ncu_k2f.log 2>&1; echo "rc=$?"
python tools/ncu_summary.py gpurun_out/r02_k2f.ncu-rep > gpurun_out/r02_k2_final2_ncu.txt 2>&1; grep -E "^==|gpu__time|dram__bytes|dram_throughput|issue_active|inst_executed.sum|bank_conflicts|wavefronts_mem_shared|registers|occupancy_limit|stalls" gpurun_out/r02_k2_final2_ncu.txt | cut -c1-230
rm -f gpurun_out/r02_k2f.ncu-rep
