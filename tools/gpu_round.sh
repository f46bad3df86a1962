#!/bin/bash
# The measurement sequence of a round on one B200 (B200_PROFILING.md): parity tests, the bench line of both arms, the ncu launch
# list of the same bench command, one ncu --set full capture of the dominant kernel (each ncu run directly after a plain run of
# the same command that exited 0).   usage: bash tools/gpu_round.sh <tag>      outputs under gpurun_out/
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --save-index /tmp/ix > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --load-index /tmp/ix --steps 3 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"
EF=$(python -c "import json; print(json.load(open('gpurun_out/bench_$TAG.json'))['config']['ef'])")
CMD="python bench.py --load-index /tmp/ix --ef $EF --steps 5 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain1_$TAG.json 2> gpurun_out/plain1_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
$CMD > gpurun_out/plain2_$TAG.json 2> gpurun_out/plain2_$TAG.err && \
ncu --set full --clock-control none --import-source on -k regex:search_kernel_fast -s 8 -c 1 -f -o gpurun_out/search_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
python tools/show_runs.py gpurun_out/bench_$TAG.json
