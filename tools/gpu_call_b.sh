#!/bin/bash
# full GPU suite, then launch list + one full ncu capture of the search kernel on the headline workload
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/r2_pytest_b.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/r2_pytest_b.log
timeout 300 python bench.py --save-index /tmp/ix --ef 57 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_plain0.json 2> gpurun_out/r2_b_plain0.err
echo "bench(save) exit $?"
CMD="python bench.py --load-index /tmp/ix --ef 57 --steps 5 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r2_b_plain.json 2> gpurun_out/r2_b_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_b.csv $CMD > gpurun_out/r2_b_ncu1.log 2>&1
echo "ncu launches exit $?"
$CMD > gpurun_out/r2_b_plain2.json 2> gpurun_out/r2_b_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:search_kernel_fast -s 8 -c 1 -o gpurun_out/r2_search_fast_b $CMD > gpurun_out/r2_b_ncu2.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/r2_b_ncu2.log
