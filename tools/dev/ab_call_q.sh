#!/bin/bash
# A/B: groups without a candidate do not load (main) vs load record 0 (nopred)
mkdir -p gpurun_out
V=$PWD/hnsw_rs_b200/variants
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_q_main.json 2> gpurun_out/r2_q_main.err
echo "bench main exit $?"
for v in nopred; do
  HNSWB200_LIB=$V/lib_$v.so timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_q_$v.json 2> gpurun_out/r2_q_$v.err
  echo "bench $v exit $?"
done
timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_q_main2.json 2> gpurun_out/r2_q_main2.err
python tools/show_runs.py gpurun_out/r2_q_*.json | cut -c1-220
