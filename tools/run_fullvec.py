#!/usr/bin/env python3
"""The f32 (FullVec) index mode at BASELINE configs[1] size on one GPU: the C2 workload (1,183,514 x 100 unit-norm clustered
mixture, 10,000 queries, M=16, ef_cons=200) with `type VecType = FullVec` (points/src/point.rs:4): device build, exact ground
truth, smallest ef with recall@10 >= 0.99, queries/s, counted algorithmic bytes (4*dim per evaluation in f32 mode)
against the measured HBM peak, and parity with the oracle in the same mode on a sample.  One JSON object on stdout."""
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


def main():
    n = int(os.environ.get("FULLVEC_N", 1183514))
    ctx = H.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    lib = _ffi.lib()
    base = synth(n, 100, 2048, 1)
    q = synth(10000, 100, 2048, 2)
    nq, k = q.shape[0], 10
    t = time.time()
    ix = H.HNSW.new(16, 200, 100, ctx=ctx, vec_type="full").insert_bulk(base)
    build_s = time.time() - t
    t = time.time()
    gt, _ = H.bruteforce_topk(ix._points(), q, k, ctx=ctx)
    gt_s = time.time() - t
    dq = torch.from_numpy(q).cuda()
    ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")

    def run(ef, stats=None):
        s = stats or [None] * 4
        _ffi.check(lib.hnswb200_search_dev(ctx.h, ix.h, dq.data_ptr(), nq, k, ef, ids.data_ptr(), None, None, s[0], s[1], s[2], s[3]))

    res = {"config": f"C2 workload in FullVec mode: {n} x 100 f32 records (448 B), 10,000 queries, M=16 ef_cons=200",
           "build_seconds": round(build_s, 1), "ground_truth_seconds_cuda_core_exact": round(gt_s, 2), "sweep": []}
    for ef in (20, 40, 48, 56, 64, 80, 100):
        for _ in range(3):
            run(ef)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run(ef)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        r = recall_at_k(ids.cpu().numpy().astype(np.uint32), gt)
        res["sweep"].append({"ef": ef, "recall_at_10": round(r, 5), "ms_per_10k": round(ms, 3), "qps": round(nq / ms * 1e3)})
        if r >= 0.99:
            break
    ef = res["sweep"][-1]["ef"]
    ms = res["sweep"][-1]["ms_per_10k"]
    cnt = [torch.zeros(nq, dtype=torch.int32, device="cuda") for _ in range(4)]
    run(ef, [c.data_ptr() for c in cnt])
    torch.cuda.synchronize()
    hops, evals, flags, nbrs = [c.cpu().numpy().astype(np.float64) for c in cnt]
    alg = hops.sum() * 8 + nbrs.sum() * 4 + evals.sum() * 4 * 100 + nq * (4 * 100 + 8 * k)
    peak = 6550.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    res["roofline"] = {"bound": "hbm", "algorithmic_bytes_per_query": round(alg / nq), "achieved_gbs": round(alg / (ms * 1e-3) / 1e9, 1),
                       "peak_gbs": peak, "frac": round(alg / (ms * 1e-3) / 1e9 / peak, 4),
                       "per_query": {"hops": round(hops.mean(), 2), "evals": round(evals.mean(), 2), "nbr_ids": round(nbrs.mean(), 2)}}
    orc = oracle_from_index(ix)
    a, d, c, st = ix.ann_batch(q[:300], k, ef, with_stats=True)
    oi, od, oc, oh, oe = orc.search_batch(q[:300], k, ef, threads=os.cpu_count())
    ok = st["flags"] == 0
    res["oracle_parity_300_queries"] = bool(np.array_equal(a, oi) and np.array_equal(d.view(np.uint32), od.view(np.uint32))
                                            and np.array_equal(st["hops"], oh) and np.array_equal(st["evals"][ok], oe[ok]))
    t = time.time()
    orc.search_batch(q[:2000], k, ef, threads=os.cpu_count())
    res["cpu_oracle_qps"] = {"value": round(2000 / (time.time() - t)), "cores": os.cpu_count(), "sample": "2000 queries, all host threads"}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
