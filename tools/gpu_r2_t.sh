#!/bin/bash
# round 2, call T: ncu --set full summaries of the final kernels (text summaries are kept, reports deleted: 64 MB limit)
mkdir -p gpurun_out
cap() { # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 900 ncu --set full --clock-control none -k regex:"$rx" -s $skip -c $cnt -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/${name}_ncu.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep
}
cap r02_filter "hk_filter2_kernel" 3 1 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-queries
cap r02_k2_final "hk_tpart_kernel|hk_dagg_tiles_kernel" 4 2 python tools/ops_bench.py --ops groupby --reps 1
cap r02_k2_lut "hk_dagg_tiles_kernel" 0 1 python tools/ops_bench.py --ops join --reps 1
cap r02_k2_hash "hk_dagg_tiles_kernel|hk_hash_build" 0 2 python tools/ops_bench.py --ops join_sparse --reps 1
cap r02_sweep16_final "hk_sweep16_kernel|hk_sweep16_hist" 3 2 python tools/ops_bench.py --ops orderby --scale 0.25 --reps 1
cap r02_joins "hk_hj_expand_kernel|hk_hj_count_kernel|hk_mj_expand_kernel|hk_mj_bounds_kernel" 0 4 python tools/ops_bench.py --ops join_entry,join_hash --reps 1
grep -h -A4 "^==" gpurun_out/r02_*_ncu.txt | grep -E "^==|time_duration|dram__bytes|dram_throughput" | cut -c1-150
ls -la gpurun_out | head -30
