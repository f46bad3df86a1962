#!/bin/bash
# the FullVec-mode measurement: bench line (same contract as the headline, --vec-type full) and one ncu --set full
# capture of its search kernel; then the other BASELINE configurations (tools/run_configs.py)
mkdir -p gpurun_out
timeout 600 python bench.py --vec-type full --save-index /tmp/ixf > gpurun_out/bench_fullvec.json 2> gpurun_out/bench_fullvec.err; echo "bench exit $?"
cut -c1-600 gpurun_out/bench_fullvec.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 8 -c 1 -f -o gpurun_out/search_fullvec \
  python bench.py --vec-type full --load-index /tmp/ixf --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fullvec.log 2>&1; echo "ncu exit $?"
timeout 900 python tools/run_configs.py > gpurun_out/other_configs.json 2> gpurun_out/other_configs.err; echo "configs exit $?"
cat gpurun_out/other_configs.json | cut -c1-700
