#!/bin/bash
# A/B: lane and the warp's shared-memory offset kept opaque (main) vs recomputed from %tid (noopaque)
mkdir -p gpurun_out
V=$PWD/hnsw_rs_b200/variants
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_s_main.json 2> gpurun_out/r2_s_main.err
echo "bench main exit $?"
HNSWB200_LIB=$V/lib_noopaque.so timeout 300 python bench.py --load-index /tmp/ix --ef 57 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_s_noopaque.json 2> gpurun_out/r2_s_noopaque.err
echo "bench noopaque exit $?"
python tools/show_runs.py gpurun_out/r2_s_*.json | cut -c1-220
