#!/bin/bash
# dev experiment: build the C2 index once, check parity on a sample, time the search kernel
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --save-index /tmp/ix > gpurun_out/b.log 2>&1
tail -c 600 gpurun_out/b.log; echo
for nq in 10000 100000; do
  echo "nq $nq"; timeout 300 python tools/dev/exp_search.py --load /tmp/ix --nq $nq --efs ${EFS:-64,100} --oracle-sample ${OS:-300} 2>&1 | grep "ef=\|parity"
done 2>&1 | tee gpurun_out/exp2.log
