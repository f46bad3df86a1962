#!/bin/bash
# A/B: insertion path for up to 3 (main) / 4 / 6 keys
mkdir -p gpurun_out
V=$PWD/hnsw_rs_b200/variants
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_p_main.json 2> gpurun_out/r2_p_main.err
echo "bench main exit $?"
for v in ins4 ins6; do
  HNSWB200_LIB=$V/lib_$v.so timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_p_$v.json 2> gpurun_out/r2_p_$v.err
  echo "bench $v exit $?"
done
python tools/show_runs.py gpurun_out/r2_p_*.json | cut -c1-220
