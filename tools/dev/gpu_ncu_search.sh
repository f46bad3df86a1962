#!/bin/bash
# dev: one ncu --set full capture of the search kernel on a saved/rebuilt C2 index.  $1 = tag, $2 = nq, $3 = ef
TAG=${1:-x}; NQ=${2:-100000}; EF=${3:-64}
mkdir -p gpurun_out
if [ ! -d /tmp/ix ]; then timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --save-index /tmp/ix > gpurun_out/b_$TAG.log 2>&1; fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 3 -c 1 -o gpurun_out/search_$TAG -f \
  python tools/dev/exp_search.py --load /tmp/ix --nq $NQ --efs $EF --oracle-sample 0 > gpurun_out/ncu_$TAG.log 2>&1
tail -5 gpurun_out/ncu_$TAG.log
