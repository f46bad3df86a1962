#!/bin/bash
# A/B of the two-lanes-per-record pre-filter and the single-key insertion (profiles/r02_ab_variants.txt, fourth group)
mkdir -p gpurun_out
V=$PWD/hnsw_rs_b200/variants
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "fast or heavy or search_synthetic or golden or c3") > gpurun_out/r2_pytest_n.log 2>&1
echo "pytest exit $?"; tail -2 gpurun_out/r2_pytest_n.log
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_n_main.json 2> gpurun_out/r2_n_main.err
echo "bench main exit $?"
for v in base f2only insonly; do
  HNSWB200_LIB=$V/lib_$v.so timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_n_$v.json 2> gpurun_out/r2_n_$v.err
  echo "bench $v exit $?"
done
timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_n_main2.json 2> gpurun_out/r2_n_main2.err
python tools/show_runs.py gpurun_out/r2_n_*.json | cut -c1-220
