#!/bin/bash
# A/B: row-only speculation (the adjacency row of the probable next expansion is loaded into a register one hop ahead)
mkdir -p gpurun_out
V=$PWD/hnsw_rs_b200/variants
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_t_main.json 2> gpurun_out/r2_t_main.err
echo "bench main exit $?"
for v in spec2 spec2b8; do
  HNSWB200_LIB=$V/lib_$v.so timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_t_$v.json 2> gpurun_out/r2_t_$v.err
  echo "bench $v exit $?"
done
python tools/show_runs.py gpurun_out/r2_t_*.json | cut -c1-220
