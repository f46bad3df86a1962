#!/bin/bash
mkdir -p gpurun_out
HNSWB200_BUILD_PROFILE=1 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/buildprof.json 2> gpurun_out/buildprof.err
echo "bench rc=$?"; grep "hnswb200 build" gpurun_out/buildprof.err; lscpu | grep -i "model name\|^CPU(s)\|L3\|L2\|NUMA node(s)"
