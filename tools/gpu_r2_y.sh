#!/bin/bash
# round 2, call Y: sanitizer over the sort shapes (K3b included) + ncu of the chunked sweep pass (mid pass only)
mkdir -p gpurun_out
echo "== memcheck sort"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider -k orderby > gpurun_out/san_mem_sort.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/san_mem_sort.log | cut -c1-250
echo "== racecheck sort"; timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 3 python -m pytest tests/test_gpu_sanitize_shapes.py -m gpu -q -x -p no:cacheprovider -k orderby > gpurun_out/san_race_sort.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/san_race_sort.log | cut -c1-250
grep -c "ERROR SUMMARY: 0 errors\|RACECHECK SUMMARY: 0 hazards" gpurun_out/san_mem_sort.log gpurun_out/san_race_sort.log
for skip in 6 7; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hk_sweep16_kernel" -s $skip -c 1 -f -o gpurun_out/r02_sweep16c python tools/ops_bench.py --ops orderby --scale 0.25 --reps 1 > gpurun_out/ncu_sweepc_$skip.log 2>&1; echo "rc=$?"
grep -E "ERROR|Profiling" gpurun_out/ncu_sweepc_$skip.log | head -5
python tools/ncu_summary.py gpurun_out/r02_sweep16c.ncu-rep > gpurun_out/r02_sweep16_chunked_ncu_$skip.txt 2>&1; head -30 gpurun_out/r02_sweep16_chunked_ncu_$skip.txt
ncu -i gpurun_out/r02_sweep16c.ncu-rep --page source --csv --print-source cuda > gpurun_out/sweep16c_source_$skip.csv 2>/dev/null
rm -f gpurun_out/r02_sweep16c.ncu-rep
done
echo "== without source counters"
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --clock-control none -k regex:"hk_sweep16_kernel" -s 6 -c 2 python tools/ops_bench.py --ops orderby --scale 0.25 --reps 1 > gpurun_out/ncu_sweepc_sections.log 2>&1; echo "rc=$?"
grep -E "ERROR|hk_sweep16_kernel|Duration|DRAM Throughput|Issue Slots Busy|Stall|Theoretical Occ|Achieved Occ|Registers Per|Eligible|No Eligible" gpurun_out/ncu_sweepc_sections.log | head -60
