#!/bin/bash
# A/B (one process, one index; tools/dev/ab_multi.py) of L1 / L2 eviction hints on the search's global loads, on top of nip
# (HB_FAST_SPILLNEST + HB_FAST_SLOWINL + HB_FAST_PRMTPOP): exact-round record loads no_allocate (x2) / evict_first (x3), filter
# record loads evict_last (f1) / evict_first (f3), row loads no_allocate (r2), combinations, row prefetch at admission L2::evict_last (el)
mkdir -p gpurun_out
V=hnsw_rs_b200/variants
timeout 200 python tools/dev/ab_multi.py --out gpurun_out/r2_w.json main=hnsw_rs_b200/libhnsw_b200.so nip=$V/lib_nip.so \
  x2=$V/lib_x2.so x3=$V/lib_x3.so f1=$V/lib_f1.so r2=$V/lib_r2.so f1x3=$V/lib_f1x3.so f1x3r2=$V/lib_f1x3r2.so f3=$V/lib_f3.so \
  el=$V/lib_el.so 2>&1 | tee gpurun_out/r2_w.log
echo "ab_multi exit ${PIPESTATUS[0]}"
