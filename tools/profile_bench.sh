#!/bin/bash
# Run on the GPU box (under gpurun): launch list and one full capture of the top kernel for the bench command.
# Only the library's own kernels are instrumented (-k regex); torch's synthetic-data kernels run natively.
set -u
OUT=${1:-gpurun_out}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e"
OURS='regex:^(depth_kernel|compact_kernel|selective_median_kernel|propagate_kernel|edge_confidence_kernel|downsample_kernel|nearest_valid_kernel|set_bounds_kernel|fuse_level_kernel|median3x3_kernel|valid_mask_kernel|fill_f32_kernel|normalise_f32_kernel|normalise_u8_kernel|stack_minmax_kernel|row_sum_kernel)'
$CMD > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
tail -c 600 $OUT/plain.log
# one step's worth of launches (the first warm-up step); cold-cache, serialised: compare SHARES
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 5400 --csv --log-file $OUT/launches_c3.csv $CMD > $OUT/ncu1.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $OUT/launches_c3.csv)"
# the dense first pass of level 0 of the depth kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^depth_kernel -s 0 -c 1 -o $OUT/prof_bench_c3_depth $CMD > $OUT/ncu2.log 2>&1
echo "full capture rc=$?"; tail -2 $OUT/ncu2.log
