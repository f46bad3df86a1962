#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --config c3 --save-index /tmp/ix3 --ef 100 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_i_plain0.json 2> gpurun_out/r2_i_plain0.err
echo "c3 save exit $?"
CMD="python bench.py --config c3 --load-index /tmp/ix3 --ef 100 --steps 5 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r2_i_plain.json 2> gpurun_out/r2_i_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:search_kernel_fast -s 6 -c 1 -o gpurun_out/r2_search_fast_c3 $CMD > gpurun_out/r2_i_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r2_i_ncu.log
