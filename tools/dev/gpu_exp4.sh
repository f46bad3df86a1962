#!/bin/bash
# dev: resident blocks per SM vs single-launch time of a 10,000-query batch (no overlap with a neighbour launch)
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --save-index /tmp/ix > gpurun_out/b.log 2>&1
export HNSWB200_NO_PDL=1
for b in 6 5 4 3; do
  export HNSWB200_SEARCH_BLOCKS_PER_SM=$b
  echo "blocks/SM $b"; timeout 300 python tools/dev/exp_search.py --load /tmp/ix --nq 10000 --efs 58 --oracle-sample 0 2>&1 | grep "ef="
done 2>&1 | tee gpurun_out/exp4.log
