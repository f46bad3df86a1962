#!/usr/bin/env python3
"""Development experiment (not the bench): build an index on the GPU, sweep ef, report recall,
kernel time (CUDA events) and end-to-end time, and check parity with the oracle on a sample."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hnsw_rs_b200 as H  # noqa: E402
from hnsw_rs_b200 import _ffi  # noqa: E402


def synth(n, dim, ncent, seed, sigma=0.35, normalise=True):
    rc = np.random.default_rng(1234)
    cent = rc.standard_normal((ncent, dim), dtype=np.float32)
    r = np.random.default_rng(seed)
    x = cent[r.integers(0, ncent, n)]
    x += np.float32(sigma) * r.standard_normal((n, dim), dtype=np.float32)
    if normalise:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x, np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200000)
    ap.add_argument("--dim", type=int, default=100)
    ap.add_argument("--m", type=int, default=16)
    ap.add_argument("--efc", type=int, default=200)
    ap.add_argument("--nq", type=int, default=10000)
    ap.add_argument("--ncent", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--oracle-sample", type=int, default=500)
    ap.add_argument("--efs", type=str, default="10,20,40,80,100,160,320")
    ap.add_argument("--load", default="")
    ap.add_argument("--save", default="")
    a = ap.parse_args()
    import torch
    queries = synth(a.nq, a.dim, a.ncent, 2)
    ctx = H.Context.default()
    t = time.time()
    if a.load:
        ix = H.HNSW.load(a.load)
    else:
        base = synth(a.n, a.dim, a.ncent, 1)
        t = time.time()
        ix = H.HNSW.new(a.m, a.efc, a.dim).insert_bulk(base, batch=a.batch or None)
    print(f"build/load: {time.time() - t:.2f}s  n={ix.len()} layers={ix.nb_layers()}", flush=True)
    if a.save:
        ix.save(a.save)
    for l in range(ix.nb_layers()):
        ids, off, _ = ix.export_layer(l)
        deg = np.diff(off.astype(np.int64))
        print(f"  layer {l}: nodes={len(ids)} deg min/mean/max={deg.min()}/{deg.mean():.2f}/{deg.max()} over_cap={(deg > ix.layer_cap(l)).sum()}")
    t = time.time()
    gt, _ = H.bruteforce_topk(ix._points(), queries, 10)
    print(f"bruteforce top-10 of {a.nq} x {a.n}: {time.time() - t:.2f}s", flush=True)
    # device-resident timing through the _dev entry point on torch's stream
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    dq = torch.from_numpy(queries).cuda()
    nq = a.nq
    d_ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda")
    d_d = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
    d_cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_h = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_e = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_f = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_nb = torch.empty(nq, dtype=torch.int32, device="cuda")
    rec = 8 + a.dim
    for ef in [int(x) for x in a.efs.split(",")]:
        def run(stats=not os.environ.get("EXP_NO_STATS")):
            _ffi.check(_ffi.lib().hnswb200_search_dev(ctx.h, ix.h, dq.data_ptr(), nq, 10, ef, d_ids.data_ptr(),
                                                      d_d.data_ptr(), d_cnt.data_ptr(), d_h.data_ptr() if stats else None,
                                                      d_e.data_ptr() if stats else None, d_f.data_ptr() if stats else None,
                                                      d_nb.data_ptr() if stats else None))
        run(True)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        ids = d_ids.cpu().numpy().astype(np.uint32)
        hits = sum(len(set(gt[i].tolist()) & set(ids[i].tolist())) for i in range(nq))
        hops, evals, nbrs = d_h.cpu().numpy(), d_e.cpu().numpy(), d_nb.cpu().numpy()
        ovf = int((d_f.cpu().numpy() & 2).sum())
        bytes_alg = float((hops.astype(np.float64) * 8 + nbrs.astype(np.float64) * 4 + evals.astype(np.float64) * rec).sum()
                          + nq * (4 * a.dim + 8 * 10))
        t0 = time.time()
        for _ in range(3):
            ix.ann_batch(queries, 10, ef)
        e2e = (time.time() - t0) / 3
        print(f"    evals max={evals.max()} p99={np.percentile(evals, 99):.0f} hops max={hops.max()} p99={np.percentile(hops, 99):.0f}")
        print(f"ef={ef:4d} recall={hits / (10 * nq):.4f} kernel={ms:8.3f} ms  qps={nq / ms * 1e3:12.0f}  e2e_qps={nq / e2e:12.0f} "
              f"hops={hops.mean():7.1f} evals={evals.mean():8.1f} nbrs={nbrs.mean():8.1f} ovf={ovf} "
              f"alg_GBps={bytes_alg / ms / 1e6:8.1f}", flush=True)
    if a.oracle_sample:
        from oracle import pyoracle as O
        codes, mins, deltas, levels = ix._points().download()
        p = ix.params
        layers = [ix.export_layer(l) for l in range(ix.nb_layers())]
        orc = O.Index.from_parts(p.m, p.ef_cons, p.dim, p.ep, codes, mins, deltas, levels, layers)
        qs = queries[:a.oracle_sample]
        for ef in (10, 100):
            ids, dists, counts, st = ix.ann_batch(qs, 10, ef, with_stats=True)
            t0 = time.time()
            oids, od, oc, oh, oe = orc.search_batch(qs, 10, ef, threads=1)
            t1 = time.time() - t0
            t0 = time.time()
            orc.search_batch(qs, 10, ef, threads=os.cpu_count())
            tn = time.time() - t0
            ok = (np.array_equal(ids, oids) and np.array_equal(dists.view(np.uint32), od.view(np.uint32))
                  and np.array_equal(st["hops"], oh) and np.array_equal(st["evals"], oe))
            if not ok:
                fl = st["flags"] if "flags" in st else None
                bad_ids = np.flatnonzero((ids != oids).any(axis=1))
                bad_h = np.flatnonzero(st["hops"] != oh)
                bad_e = np.flatnonzero(st["evals"] != oe)
                print(f"   mismatch detail: ids rows={len(bad_ids)} dists={int((dists.view(np.uint32) != od.view(np.uint32)).sum())} "
                      f"hops rows={len(bad_h)} evals rows={len(bad_e)} "
                      f"overflow-flagged among evals-mismatch={int((fl[bad_e] & 2 != 0).sum()) if fl is not None else 'n/a'}")
            print(f"oracle parity ef={ef}: {'OK' if ok else 'MISMATCH'}  cpu qps 1thr={len(qs) / t1:.0f} {os.cpu_count()}thr={len(qs) / tn:.0f}",
                  flush=True)


if __name__ == "__main__":
    main()
