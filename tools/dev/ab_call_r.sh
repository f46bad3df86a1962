#!/bin/bash
# the in-tree kernel alone (compare with the previous call's numbers: box-to-box spread has been <= 0.1 %), plus the fast-path parity tests
mkdir -p gpurun_out
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_r_main.json 2> gpurun_out/r2_r_main.err
echo "bench main exit $?"
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast or heavy or search_synthetic or golden") > gpurun_out/r2_pytest_r.log 2>&1
echo "pytest exit $?"; tail -2 gpurun_out/r2_pytest_r.log
python tools/show_runs.py gpurun_out/r2_r_*.json | cut -c1-220
