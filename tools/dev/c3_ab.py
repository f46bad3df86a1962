"""dev: C3 (1M x 128, ef = 100, 4 keys per lane) with the 3584-entry visited table (6 blocks/SM) against the 4096-entry one (5 blocks)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import hnsw_rs_b200 as H
from run_configs import sift_like, time_search
from bench import oracle_from_index, recall_at_k
ctx = H.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
base = sift_like(1000000, 128, 1024, 3)
q = sift_like(10000, 128, 1024, 4)
ix = H.HNSW.new(16, 200, 128, ctx=ctx).insert_bulk(base)
gt, _ = H.bruteforce_topk(ix._points(), q, 10, ctx=ctx)
os.environ["HNSWB200_DEBUG_LAUNCH"] = "1"
for rep in range(2):
    for mode in ("n3584", "pow2"):
        if mode == "pow2": os.environ["HNSWB200_VIS_POW2"] = "1"
        else: os.environ.pop("HNSWB200_VIS_POW2", None)
        for ef in (100, 128):
            ms, ids = time_search(ix, ctx, q, 10, ef)
            print(f"{mode} ef={ef} ms={ms:.3f} qps={10000/ms*1e3:.0f} recall={recall_at_k(ids, gt):.4f}", flush=True)
os.environ.pop("HNSWB200_VIS_POW2", None)
orc = oracle_from_index(ix)
for ef in (100, 128):
    ids, d, c, st = ix.ann_batch(q[:300], 10, ef, with_stats=True)
    oi, od, oc, oh, oe = orc.search_batch(q[:300], 10, ef, threads=os.cpu_count())
    ok = st["flags"] == 0
    print("parity ef", ef, bool(np.array_equal(ids, oi) and np.array_equal(d.view(np.uint32), od.view(np.uint32)) and np.array_equal(st["hops"], oh)
                          and np.array_equal(st["evals"][ok], oe[ok])), "overflow-flagged", int((~ok).sum()))
