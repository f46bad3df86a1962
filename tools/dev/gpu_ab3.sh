#!/bin/bash
# dev: A/B of library builds with the same ABI on the bench workload (C2, ef=57), answers-only kernel variant,
# 10,000 and 100,000 queries per launch, oracle parity on 300 queries.  usage: gpu_ab3.sh <variant.so | default>...
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --save-index /tmp/ix > gpurun_out/b.log 2>&1
export EXP_NO_STATS=1
for rep in 1 2; do
for v in "$@"; do
  if [ $v = default ]; then unset HNSWB200_LIB; else export HNSWB200_LIB=$PWD/$v; fi
  echo "variant $v rep $rep"; timeout 300 python tools/dev/exp_search.py --load /tmp/ix --nq 10000 --efs 57,57 --oracle-sample $([ $rep = 1 ] && echo 300 || echo 0) 2>&1 | grep "ef=\|parity"
  timeout 300 python tools/dev/exp_search.py --load /tmp/ix --nq 100000 --efs 57 --oracle-sample 0 2>&1 | grep "ef="
done
done 2>&1 | tee gpurun_out/ab3.log
