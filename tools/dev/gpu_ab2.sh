#!/bin/bash
# dev: occupancy experiment: visited table of 2048 slots (4 KB/warp) with builds compiled for 7 / 8 blocks per SM
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --save-index /tmp/ix > gpurun_out/b.log 2>&1
run() { for nq in 10000 100000; do echo "variant $1 slots=${HNSWB200_VIS_SLOTS:-auto} nq $nq"; timeout 300 python tools/dev/exp_search.py --load /tmp/ix --nq $nq --efs 64 --oracle-sample 0 2>&1 | grep "ef="; done; }
unset HNSWB200_LIB; unset HNSWB200_VIS_SLOTS; run default
export HNSWB200_VIS_SLOTS=2048; run default
export HNSWB200_LIB=$PWD/hnsw_rs_b200/variants/lib_b7.so; run b7
export HNSWB200_LIB=$PWD/hnsw_rs_b200/variants/lib_b8.so; run b8
