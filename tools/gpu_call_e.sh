#!/bin/bash
# full GPU suite on the adopted defaults (+ new K5 filter, config tests), C4 timing with the per-chunk profile, launch list of C4
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x) > gpurun_out/r2_pytest_e.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/r2_pytest_e.log
HNSWB200_BF_PROFILE=1 C4_REPS=2 timeout 300 python tools/c4_profile.py > gpurun_out/r2_c4_profile.log 2>&1
echo "c4 exit $?"; grep -E "^rep|checksum" gpurun_out/r2_c4_profile.log
C4_REPS=2 timeout 300 python tools/c4_profile.py > gpurun_out/r2_c4_plain.log 2>&1 && \
C4_REPS=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_c4_launches.csv python tools/c4_profile.py > gpurun_out/r2_c4_ncu.log 2>&1
echo "ncu c4 exit $?"
timeout 400 python bench.py > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
echo "bench exit $?"; tail -2 gpurun_out/r2_bench_e.err
