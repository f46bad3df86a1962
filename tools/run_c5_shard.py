#!/usr/bin/env python3
"""One shard of BASELINE config 5 at its true size: 12,500,000 x 96 (unit-norm, 65,536-centre mixture) on one GPU:
device build, exact ground truth on the tensor cores, ef sweep, oracle parity on a sample of the queries."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hnsw_rs_b200 as H  # noqa: E402
from bench import oracle_from_index, recall_at_k  # noqa: E402
from tools.run_configs import time_search  # noqa: E402


def synth_big(n, dim, ncent, seed, sigma=0.35, chunk=2000000):
    rc = np.random.default_rng(1234)
    cent = rc.standard_normal((ncent, dim), dtype=np.float32)
    r = np.random.default_rng(seed)
    out = np.empty((n, dim), np.float32)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        x = cent[r.integers(0, ncent, b - a)] + np.float32(sigma) * r.standard_normal((b - a, dim), dtype=np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        out[a:b] = x
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12500000
    ctx = H.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    base = synth_big(n, 96, 65536, 5)
    q = synth_big(10000, 96, 65536, 6)
    t = time.time()
    ix = H.HNSW.new(16, 200, 96, ctx=ctx).insert_bulk(base)
    build_s = time.time() - t
    del base
    t = time.time()
    gt, _ = H.bruteforce_topk(ix._points(), q, 10, ctx=ctx)
    gt_s = time.time() - t
    res = {"config": f"C5 one shard at true size: {n} x 96, 10,000 queries, M=16 ef_cons=200, one GPU", "build_seconds": round(build_s, 1),
           "inserts_per_second": round(n / build_s), "ground_truth_seconds_tensor_core": round(gt_s, 3), "layers": ix.nb_layers(), "sweep": []}
    for ef in (20, 40, 60, 80, 120, 160):
        ms, ids = time_search(ix, ctx, q, 10, ef)
        r = recall_at_k(ids, gt)
        res["sweep"].append({"ef": ef, "recall_at_10": round(r, 5), "ms_per_10k": round(ms, 3), "qps": round(10000 / ms * 1e3)})
        if r >= 0.99:
            break
    ef = res["sweep"][-1]["ef"]
    t = time.time()
    orc = oracle_from_index(ix)
    res["oracle_import_seconds"] = round(time.time() - t, 1)
    ids, d, c, st = ix.ann_batch(q[:200], 10, ef, with_stats=True)
    oi, od, oc, oh, oe = orc.search_batch(q[:200], 10, ef, threads=os.cpu_count())
    ok = st["flags"] == 0
    res["oracle_parity_200_queries"] = bool(np.array_equal(ids, oi) and np.array_equal(d.view(np.uint32), od.view(np.uint32))
                                            and np.array_equal(st["hops"], oh) and np.array_equal(st["evals"][ok], oe[ok]))
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
