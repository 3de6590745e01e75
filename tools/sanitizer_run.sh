#!/bin/bash
# compute-sanitizer over the kernel parity cases: memcheck on all of them, racecheck on the kernels that
# hand data between warps through shared memory / mbarriers.  Summaries go to gpurun_out/ (copied to profiles/ by hand).
set -u
cd "$(dirname "$0")/.."
SAN=/usr/local/cuda/bin/compute-sanitizer
OUT=gpurun_out
mkdir -p $OUT
run() {  # name, tool, pytest -k expression
    timeout 900 $SAN --tool "$2" --print-limit 20 --log-file $OUT/sanitizer_$1.log \
        python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$3" > $OUT/sanitizer_$1.pytest.log 2>&1
    echo "$1: rc=$? $(tail -1 $OUT/sanitizer_$1.pytest.log)"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $OUT/sanitizer_$1.log | tail -2
}
run memcheck_all memcheck "not es"
run racecheck_fused racecheck "layered_mlp or fused_train or scan or norm or adam or reduce_step"
run racecheck_tc racecheck "tc_modes"
