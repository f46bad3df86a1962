#!/bin/bash
# dev: index quality vs insertion batch size (recall at small ef), C2 workload
for b in 4096 1024 256; do
  echo "== batch $b"; timeout 900 python tools/dev/exp_search.py --n 1183514 --nq 10000 --batch $b --efs 48,52,56,58,60,62,64 --oracle-sample 0 2>&1 | grep "build\|ef="
done 2>&1 | tee gpurun_out/exp3.log
