#!/usr/bin/env python3
"""The other BASELINE.json configurations at full size on one GPU (bench.py measures configs[1]):
  C3  synthetic SIFT shape, 1M x 128, L2 (no normalisation), 10k queries: build, ef for recall@10 >= 0.99, QPS, oracle parity on a sample
  C4  brute-force exact top-100, 1M x 100 base x 10k queries (tensor-core filter + exact re-rank), checked against the CUDA-core path
  C5' base-sharded HNSW, down-scaled: 8 shards x 125k x 96 on one GPU, K6 merge == brute force over the union (recall) and == oracle merge on a sample
Writes one JSON object per config to stdout."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hnsw_rs_b200 as H  # noqa: E402
from hnsw_rs_b200 import _ffi  # noqa: E402
from bench import oracle_from_index, recall_at_k, synth  # noqa: E402


def sift_like(n, dim, ncent, seed):
    rc = np.random.default_rng(4321)
    cent = np.abs(rc.standard_normal((ncent, dim), dtype=np.float32)) * 40
    r = np.random.default_rng(seed)
    x = cent[r.integers(0, ncent, n)] + np.abs(r.standard_normal((n, dim), dtype=np.float32)) * 20
    return np.minimum(np.floor(x), 218).astype(np.float32)


def time_search(ix, ctx, q, k, ef, reps=10):
    lib = _ffi.lib()
    dq = torch.from_numpy(q).cuda()
    nq = q.shape[0]
    ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    def run():
        _ffi.check(lib.hnswb200_search_dev(ctx.h, ix.h, dq.data_ptr(), nq, k, ef, ids.data_ptr(), None, None, None, None, None, None))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, ids.cpu().numpy().astype(np.uint32)


def main():
    ctx = H.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    out = []
    # ---- C3 ----
    base = sift_like(1000000, 128, 1024, 3)
    q = sift_like(10000, 128, 1024, 4)
    t = time.time()
    ix = H.HNSW.new(16, 200, 128, ctx=ctx).insert_bulk(base)
    build_s = time.time() - t
    t = time.time()
    gt, _ = H.bruteforce_topk(ix._points(), q, 10, ctx=ctx)
    gt_s = time.time() - t
    res = {"config": "C3 synthetic SIFT shape 1,000,000 x 128 (L2), 10,000 queries, M=16 ef_cons=200", "build_seconds": round(build_s, 1),
           "ground_truth_seconds": round(gt_s, 2), "sweep": []}
    for ef in (20, 40, 60, 80, 100, 140, 200):
        ms, ids = time_search(ix, ctx, q, 10, ef)
        r = recall_at_k(ids, gt)
        res["sweep"].append({"ef": ef, "recall_at_10": round(r, 5), "ms_per_10k": round(ms, 3), "qps": round(10000 / ms * 1e3)})
        if r >= 0.99:
            break
    orc = oracle_from_index(ix)
    ef = res["sweep"][-1]["ef"]
    ids, d, c, st = ix.ann_batch(q[:500], 10, ef, with_stats=True)
    oi, od, oc, oh, oe = orc.search_batch(q[:500], 10, ef, threads=os.cpu_count())
    res["oracle_parity_500_queries"] = bool(np.array_equal(ids, oi) and np.array_equal(d.view(np.uint32), od.view(np.uint32))
                                            and np.array_equal(st["hops"], oh) and np.array_equal(st["evals"][st["flags"] == 0], oe[st["flags"] == 0]))
    out.append(res)
    del ix, orc, base
    # ---- C4 ----
    base = synth(1000000, 100, 2048, 1)
    q = synth(10000, 100, 2048, 2)
    pts = H.SimplePoints.new(base, ctx=ctx)
    H.bruteforce_topk(pts, q[:128], 100, ctx=ctx)
    t = time.time()
    ids, d = H.bruteforce_topk(pts, q, 100, ctx=ctx)
    tc_s = time.time() - t
    os.environ["HNSWB200_BF_NO_TC"] = "1"
    t = time.time()
    ids2, d2 = H.bruteforce_topk(pts, q, 100, ctx=ctx)
    cc_s = time.time() - t
    os.environ.pop("HNSWB200_BF_NO_TC")
    out.append({"config": "C4 brute-force exact top-100, 1,000,000 x 100 base x 10,000 queries, one GPU",
                "tensor_core_path_ms": round(tc_s * 1e3, 1), "cuda_core_path_ms": round(cc_s * 1e3, 1),
                "paths_bit_identical": bool(np.array_equal(ids, ids2) and np.array_equal(d.view(np.uint32), d2.view(np.uint32))),
                "useful_int8_tops": round(2 * 1e6 * 1e4 * 100 / tc_s / 1e12, 1)})
    del pts, base
    # ---- C5 down-scaled ----
    from hnsw_rs_b200 import sharded
    G, per, dim = 8, 125000, 96
    base = synth(G * per, dim, 4096, 5)
    q = synth(2000, dim, 4096, 6)
    t = time.time()
    shards = [H.HNSW.new(16, 200, dim, ctx=ctx).insert_bulk(base[g * per:(g + 1) * per]) for g in range(G)]
    build_s = time.time() - t
    allp = H.SimplePoints.new(base, ctx=ctx)
    gt, _ = H.bruteforce_topk(allp, q, 10, ctx=ctx)
    res = {"config": f"C5 down-scaled: {G} shards x {per} x {dim} on one GPU, all queries to all shards, K6 merge", "build_seconds": round(build_s, 1)}
    for ef in (40, 80):
        t = time.time()
        parts = []
        for g, ix in enumerate(shards):
            i, d, _ = ix.ann_batch(q, 10, ef)
            parts.append((sharded.BaseShardedSearch(ix, g * per)._globalise(i), d))
        mi, md = H.topk_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), ctx=ctx)
        res[f"ef{ef}"] = {"recall_at_10_vs_bruteforce_over_union": round(recall_at_k(mi, gt), 5), "seconds_host_path": round(time.time() - t, 3)}
    # merged result == oracle merge of per-shard oracle searches (sample)
    qs = q[:100]
    exp = []
    for g, ix in enumerate(shards):
        o = oracle_from_index(ix)
        i, d = o.search_batch(qs, 10, 80, threads=os.cpu_count())[:2]
        exp.append((np.where(i != 0xFFFFFFFF, i + np.uint32(g * per), i), d))
    key = lambda i, d: (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | i.astype(np.uint64)
    allk = np.concatenate([key(i, d) for i, d in exp], axis=1)
    allk.sort(axis=1)
    got = []
    for g, ix in enumerate(shards):
        i, d, _ = ix.ann_batch(qs, 10, 80)
        got.append((sharded.BaseShardedSearch(ix, g * per)._globalise(i), d))
    mi, md = H.topk_merge(np.stack([p[0] for p in got]), np.stack([p[1] for p in got]), ctx=ctx)
    res["merge_equals_oracle_merge_100_queries"] = bool(np.array_equal(key(mi, md), allk[:, :10]))
    out.append(res)
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
