#!/bin/bash
# verification of the new default (SPILLNEST + SLOWINL + PRMTPOP): smoke() and every test that searches through search_kernel_fast
mkdir -p gpurun_out
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02b.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_r02b.log
timeout 95 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "search or fast_kernel or ann_by_vector or two_contexts or cosine" > gpurun_out/r2_pytest_x.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r2_pytest_x.log
