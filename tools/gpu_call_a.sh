#!/bin/bash
# first GPU call of round 2: the fast-kernel tests, then the bench with the new and the round-1 kernel
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast or heavy or kernel_variants or search_synthetic or wide_rows or randomised or glove_fixture") > gpurun_out/r2_pytest_a.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/r2_pytest_a.log
timeout 600 python bench.py --save-index /tmp/ix > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench_a.err
HNSWB200_NO_FAST=1 timeout 300 python bench.py --load-index /tmp/ix --no-cpu-baseline > gpurun_out/r2_bench_a_old.json 2> gpurun_out/r2_bench_a_old.err
echo "bench old exit $?"
