#!/bin/bash
# A/B (one process, one index; tools/dev/ab_multi.py): HB_FAST_SPILLNEST + HB_FAST_SLOWINL together (ni), plus the free marks of a
# bucket counted through one byte permute (HB_FAST_PRMTPOP: nip) and record loads that do not allocate in L1 (HB_FAST_LDNA: nil)
mkdir -p gpurun_out
V=hnsw_rs_b200/variants
timeout 200 python tools/dev/ab_multi.py --out gpurun_out/r2_v.json main=hnsw_rs_b200/libhnsw_b200.so \
  inl=$V/lib_inl.so ni=$V/lib_ni.so nip=$V/lib_nip.so nil=$V/lib_nil.so nipl=$V/lib_nipl.so 2>&1 | tee gpurun_out/r2_v.log
echo "ab_multi exit ${PIPESTATUS[0]}"
