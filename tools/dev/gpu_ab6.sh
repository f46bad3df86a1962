#!/bin/bash
# dev: 3584-entry visited table + 7 blocks/SM (default) against the 4096-entry table + 6 blocks/SM (HNSWB200_VIS_POW2=1)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --save-index /tmp/ix > gpurun_out/b.log 2>&1
export EXP_NO_STATS=1
export HNSWB200_DEBUG_LAUNCH=1
run() { echo "variant $1"; timeout 300 python tools/dev/exp_search.py --load /tmp/ix --nq 10000 --efs 57,57 --oracle-sample $2 2>&1 | grep "ef=\|parity\|hnswb200 search" | tail -4; timeout 300 python tools/dev/exp_search.py --load /tmp/ix --nq 100000 --efs 57 --oracle-sample 0 2>&1 | grep "ef=" | tail -1; }
for rep in 1 2; do
unset HNSWB200_VIS_POW2; run vis16n-7blocks $([ $rep = 1 ] && echo 300 || echo 0)
export HNSWB200_VIS_POW2=1; run vis16-6blocks 0
done 2>&1 | tee gpurun_out/ab6.log
