#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -k "bruteforce or c4 or c3 or recall") > gpurun_out/r2_pytest_f.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/r2_pytest_f.log
HNSWB200_BF_PROFILE=1 C4_REPS=3 timeout 300 python tools/c4_profile.py > gpurun_out/r2_c4_profile_f.log 2>&1
echo "c4 exit $?"; grep -E "^rep|checksum" gpurun_out/r2_c4_profile_f.log; grep "bruteforce\]" gpurun_out/r2_c4_profile_f.log | tail -4
HNSWB200_LIB=$PWD/hnsw_rs_b200/variants/lib_tc_uncentred.so C4_REPS=3 timeout 300 python tools/c4_profile.py > gpurun_out/r2_c4_profile_f_unc.log 2>&1
echo "c4 uncentred exit $?"; grep -E "^rep|checksum" gpurun_out/r2_c4_profile_f_unc.log
C4_REPS=1 timeout 300 python tools/c4_profile.py > gpurun_out/r2_c4_plain_f.log 2>&1 && \
C4_REPS=1 ncu --set full --clock-control none --import-source on -k regex:bf_tc_filter -s 9 -c 1 -o gpurun_out/r2_bf_tc_filter python tools/c4_profile.py > gpurun_out/r2_c4_ncu_f.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r2_c4_ncu_f.log
