#!/bin/bash
# build-commit timing + build parity tests
mkdir -p gpurun_out
HNSWB200_BUILD_PROFILE=1 timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/buildprof.json 2> gpurun_out/buildprof.err
echo "bench rc=$?"; grep "hnswb200 build" gpurun_out/buildprof.err
timeout 600 python -m pytest tests -m gpu -x -q -k "build or insert or bulk or cpp or save or wide" > gpurun_out/pytest_build.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_build.log
