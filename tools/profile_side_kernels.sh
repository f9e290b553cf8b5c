#!/bin/bash
# GPU box (under gpurun, one GPU): HBM traffic and duration of the stream kernels around the depth kernel (edge confidence,
# normalisation, compaction, selective median, propagation, pyramid, fusion) on a 256-row band of the C3 light field run
# through the fine-to-coarse pipeline.  One ncu pass with a handful of metrics, only after the same command exited 0 plain.
set -u
OUT=${1:-gpurun_out}
TAG=${2:-r02}
CMD="python tools/quick_timing.py c3band_ftc"
KER='regex:^(edge_confidence_kernel|normalise_f32_kernel|stack_minmax_kernel|compact_kernel|selective_median_kernel|propagate_kernel|downsample_kernel|nearest_valid_kernel|set_bounds_kernel|fuse_level_kernel|median3x3_kernel|valid_mask_kernel|row_sum_kernel)'
$CMD > $OUT/plain_side_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_side_$TAG.log; exit 1; }
tail -2 $OUT/plain_side_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct \
    --clock-control none -k "$KER" -c 2500 --csv --log-file $OUT/side_kernels_$TAG.csv $CMD > $OUT/ncu_side_$TAG.log 2>&1
echo "side kernels rc=$? lines=$(wc -l < $OUT/side_kernels_$TAG.csv)"
