#!/bin/bash
# dev experiment: parity tests, then kernel timings of the register-list path vs the older shared-memory-list path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --save-index /tmp/ix > gpurun_out/b.log 2>&1
tail -c 1500 gpurun_out/b.log
for mode in reg smem; do
  if [ $mode = smem ]; then export HNSWB200_SMEM_LIST=1; else unset HNSWB200_SMEM_LIST; fi
  for nq in 10000 100000; do
    echo "mode $mode nq $nq"; timeout 300 python tools/dev/exp_search.py --load /tmp/ix --nq $nq --efs 64,100 --oracle-sample 300 2>&1 | grep "ef=\|parity"
  done
done 2>&1 | tee gpurun_out/exp1.log
