#!/bin/bash
# dev: A/B of library builds with the same ABI.  usage: gpu_ab.sh <efs> <variant.so>...   ("default" = the in-tree library)
EFS=$1; shift
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --save-index /tmp/ix > gpurun_out/b.log 2>&1
for v in "$@"; do
  if [ $v = default ]; then unset HNSWB200_LIB; else export HNSWB200_LIB=$PWD/$v; fi
  for nq in 10000 100000; do
    echo "variant $v nq $nq"; timeout 300 python tools/dev/exp_search.py --load /tmp/ix --nq $nq --efs $EFS --oracle-sample 0 2>&1 | grep "ef="
  done
done 2>&1 | tee gpurun_out/ab.log
