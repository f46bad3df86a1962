#!/bin/bash
# The measurement sequence of a round on the GPU box (B200_PROFILING.md): parity tests, the bench line,
# the ncu launch list of the same bench command, one ncu --set full capture of the dominant kernel.
# usage: bash tools/gpu_round.sh <tag>      outputs under gpurun_out/
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --save-index /tmp/ix > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
cat gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --load-index /tmp/ix --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"
cat gpurun_out/bench_ref_$TAG.json
# launch list (per-launch durations, cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --load-index /tmp/ix --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches exit $?"
# full capture of the search kernel (4th launch: after warm-up)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 8 -c 1 -f -o gpurun_out/search_$TAG \
  python bench.py --load-index /tmp/ix --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
