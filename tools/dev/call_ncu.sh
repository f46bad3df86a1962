#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_l_plain0.json 2> gpurun_out/r2_l_plain0.err
CMD="python bench.py --load-index /tmp/ix --ef 57 --steps 5 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r2_l_plain.json 2> gpurun_out/r2_l_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:search_kernel_fast -s 8 -c 1 -f -o gpurun_out/r2_search_fast_l $CMD > gpurun_out/r2_l_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r2_l_ncu.log
