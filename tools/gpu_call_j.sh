#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_fullvec.py tests/test_sharded_gpu.py -m gpu -q -x -k "bruteforce or c4 or c3 or recall or sharded or exchange or fullvec") > gpurun_out/r2_pytest_j.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/r2_pytest_j.log
HNSWB200_BF_PROFILE=1 C4_REPS=3 timeout 300 python tools/c4_profile.py > gpurun_out/r2_c4_profile_j.log 2>&1
echo "c4 exit $?"; grep -E "^rep|checksum" gpurun_out/r2_c4_profile_j.log; grep "bruteforce\]" gpurun_out/r2_c4_profile_j.log | tail -9
C4_K=10 C4_N=1183514 C4_REPS=3 timeout 300 python tools/c4_profile.py > gpurun_out/r2_gt_profile_j.log 2>&1
echo "gt(k=10, 1.18M) exit $?"; grep -E "^rep" gpurun_out/r2_gt_profile_j.log
C4_REPS=2 timeout 300 python tools/c4_profile.py > gpurun_out/r2_c4_plain_j.log 2>&1 && \
C4_REPS=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_c4_launches_j.csv python tools/c4_profile.py > gpurun_out/r2_c4_ncu_j.log 2>&1
echo "ncu c4 exit $?"
