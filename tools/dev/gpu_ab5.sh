#!/bin/bash
# dev: occupancy at equal visited-table size: 2048-slot table with the 6-blocks/SM build and a 7-blocks/SM build (72 registers)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --save-index /tmp/ix > gpurun_out/b.log 2>&1
export EXP_NO_STATS=1
run() { echo "variant $1 slots=${HNSWB200_VIS_SLOTS:-auto}"; for nq in 10000 100000; do timeout 300 python tools/dev/exp_search.py --load /tmp/ix --nq $nq --efs 57,57 --oracle-sample 0 2>&1 | grep "ef=" | tail -1; done; }
for rep in 1 2; do
unset HNSWB200_LIB; unset HNSWB200_VIS_SLOTS; run default6
export HNSWB200_VIS_SLOTS=2048; run default6
export HNSWB200_LIB=$PWD/hnsw_rs_b200/variants/lib_B7.so; run b7
unset HNSWB200_VIS_SLOTS; run b7
done 2>&1 | tee gpurun_out/ab5.log
