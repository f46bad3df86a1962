#!/bin/bash
# A/B (one process, one index): visited-set variants of search_kernel_fast -- HB_FAST_MATCH (free entries of a home bucket taken in
# lane order, one match.any, no claim read-back), HB_FAST_SPILLNEST (spill set consulted inside the slow-path branch), HB_FAST_SLOWINL
# (vis_slow inlined) and their combinations; every variant's ids / distance bits / counts / hops / evaluations are compared with the
# default build's on six batches (tools/dev/ab_multi.py)
mkdir -p gpurun_out
V=hnsw_rs_b200/variants
timeout 200 python tools/dev/ab_multi.py --out gpurun_out/r2_u.json main=hnsw_rs_b200/libhnsw_b200.so \
  nest=$V/lib_nest.so inl=$V/lib_inl.so match=$V/lib_match.so mn=$V/lib_mn.so mni=$V/lib_mni.so $EXTRA_VARIANTS 2>&1 | tee gpurun_out/r2_u.log
echo "ab_multi exit ${PIPESTATUS[0]}"
