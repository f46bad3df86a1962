#!/bin/bash
# A/B on top of call N: insertion path for up to 2 / 3 keys, no record prefetch before the visited test, 10 blocks per SM
mkdir -p gpurun_out
V=$PWD/hnsw_rs_b200/variants
(HNSWB200_LIB=$V/lib_b10.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast or heavy or search_synthetic or golden") > gpurun_out/r2_pytest_o_b10.log 2>&1
echo "pytest b10 exit $?"; tail -2 gpurun_out/r2_pytest_o_b10.log
(HNSWB200_LIB=$V/lib_ins3.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast or heavy or search_synthetic or golden") > gpurun_out/r2_pytest_o_ins3.log 2>&1
echo "pytest ins3 exit $?"; tail -2 gpurun_out/r2_pytest_o_ins3.log
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_o_main.json 2> gpurun_out/r2_o_main.err
echo "bench main exit $?"
for v in ins2 ins3 nopf b10; do
  HNSWB200_LIB=$V/lib_$v.so timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_o_$v.json 2> gpurun_out/r2_o_$v.err
  echo "bench $v exit $?"
done
python tools/show_runs.py gpurun_out/r2_o_*.json | cut -c1-220
