#!/bin/bash
# dev: brute-force parity tests + ground-truth timing (tensor-core filter vs CUDA-core path)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "bruteforce or recall" 2>&1 | tail -5
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/bf.log
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import hnsw_rs_b200 as H
from bench import synth
base = synth(1183514, 100, 2048, 1)
q = synth(10000, 100, 2048, 2)
pts = H.SimplePoints.new(base)
for k in (10, 100):
    for mode in ("tc", "cuda"):
        if mode == "cuda": os.environ["HNSWB200_BF_NO_TC"] = "1"
        else: os.environ.pop("HNSWB200_BF_NO_TC", None)
        H.bruteforce_topk(pts, q[:256], k)
        t = time.time(); ids, d = H.bruteforce_topk(pts, q, k); dt = time.time() - t
        print(f"k={k} {mode}: {dt*1e3:.1f} ms  checksum {int(ids.astype(np.uint64).sum())} {float(d.astype(np.float64).sum()):.6f}", flush=True)
        if mode == "tc": ref = (ids.copy(), d.copy())
        else: print("  identical to tc:", np.array_equal(ids, ref[0]) and np.array_equal(d.view(np.uint32), ref[1].view(np.uint32)))
PY
