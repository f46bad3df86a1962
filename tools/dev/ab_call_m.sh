#!/bin/bash
# A/B on top of the pre-filter kernel (profiles/r02_ab_variants.txt, third group)
mkdir -p gpurun_out
V=$PWD/hnsw_rs_b200/variants
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast or heavy") > gpurun_out/r2_pytest_m.log 2>&1
echo "pytest exit $?"; tail -2 gpurun_out/r2_pytest_m.log
(HNSWB200_LIB=$V/lib_b9.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast or heavy or search_synthetic") > gpurun_out/r2_pytest_m_b9.log 2>&1
echo "pytest b9 exit $?"; tail -2 gpurun_out/r2_pytest_m_b9.log
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_m_main.json 2> gpurun_out/r2_m_main.err
echo "bench main exit $?"
for v in nopf pipe b9 nofilter; do
  HNSWB200_LIB=$V/lib_$v.so timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_m_$v.json 2> gpurun_out/r2_m_$v.err
  echo "bench $v exit $?"
done
python tools/show_runs.py gpurun_out/r2_m_*.json | cut -c1-220
