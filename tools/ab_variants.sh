#!/bin/bash
# A/B timing of alternative builds of the library (remotesensingproject_b200/variants/*.so): parity tests, then C3 timings.
# usage (GPU box): tools/ab_variants.sh [workloads for quick_timing.py ...]
cd "$(dirname "$0")/.."
W="${@:-c3pile c3}"
for lib in base $(ls remotesensingproject_b200/variants/*.so 2>/dev/null); do
    if [ "$lib" = base ]; then unset RSLF_B200_LIB; else export RSLF_B200_LIB="$PWD/$lib"; fi
    echo "=== $lib"
    timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
    timeout 200 python tools/quick_timing.py $W 2>&1 | grep -E "^c[0-9]" | sed -E 's/"ms_(h2d|d2h)[^,]*, //g' | cut -c1-420
done
