#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "fast or heavy or search_synthetic or wide_rows or randomised or kernel_variants or glove_fixture or c3 or c5 or save_load or insert or bruteforce") > gpurun_out/r2_pytest_k.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r2_pytest_k.log
timeout 400 python bench.py --save-index /tmp/ix > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err
echo "bench exit $?"; tail -2 gpurun_out/r2_bench_k.err
python tools/show_runs.py gpurun_out/r2_bench_k.json
