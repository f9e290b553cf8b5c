#!/bin/bash
# GPU box (under gpurun, one GPU): (1) launch list of the bench command, (2) ncu --set full capture of the dominant
# kernel on the dense centre-line pass of a 270-row band of the C3 light field (the full C3 process is too large for
# ncu's kernel replay, see profiles/README.md).  Each ncu run only after the same command exited 0 without ncu.
set -u
OUT=${1:-gpurun_out}
TAG=${2:-r02}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-fast --no-stage-check"
OURS='regex:^(depth_kernel|compact_kernel|selective_median_kernel|propagate_kernel|edge_confidence_kernel|downsample_kernel|downsample_int_kernel|nearest_valid_kernel|set_bounds_kernel|fuse_level_kernel|median3x3_kernel|valid_mask_kernel|fill_f32_kernel|normalise_|stack_m|row_sum_kernel|p2p_push_halo_kernel|bal_apply_kernel|svu_push_halo_kernel|peer_wait_kernel|level_prepare_kernel|edge_)'
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
tail -c 400 $OUT/plain_$TAG.log; echo
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 4200 --csv --log-file $OUT/launches_c3_$TAG.csv $CMD > $OUT/ncu1_$TAG.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $OUT/launches_c3_$TAG.csv)"
python tools/quick_timing.py c3pile > $OUT/plain2_$TAG.log 2>&1 || { echo "plain c3pile failed"; tail -5 $OUT/plain2_$TAG.log; exit 1; }
tail -1 $OUT/plain2_$TAG.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:depth_kernel -s 1 -c 1 -o $OUT/prof_c3pile_$TAG python tools/quick_timing.py c3pile > $OUT/ncu2_$TAG.log 2>&1
echo "full capture rc=$?"; tail -2 $OUT/ncu2_$TAG.log
