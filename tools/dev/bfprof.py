"""dev: ground-truth timing on the C2 workload (10,000 queries x 1,183,514 base rows)."""
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
        ts = []
        for rep in range(3):
            t = time.time(); ids, d = H.bruteforce_topk(pts, q, k); ts.append(time.time() - t)
        print(f"k={k} {mode}: " + " ".join(f"{x*1e3:.1f}" for x in ts) + " ms", flush=True)
