"""bench.py --config c3 | c4 | c5: the multi-GPU configurations of BASELINE.json (configs[2..4]) under the same contract
as the headline (one JSON line on rank 0, max-over-ranks device timing, clocks, e2e, roofline, cpu_baseline).

  c3  synthetic SIFT shape 1,000,000 x 128 (L2), 10,000 queries, index replicated, the 10,000 queries SPLIT over the N
      ranks (strong scaling).  The all-gather of the id rows is fused into the search kernel (peer stores over NVLink),
      followed by flag words in peer memory: every rank holds all 10,000 result rows at the end of every step.
  c4  exact top-100 ground truth, 1,000,000 x 100 base x 10,000 queries on the tensor cores, base rows split over the N
      ranks; per-rank exact top-100 with global ids, peer stores, flags, K6 merge (strong scaling).
  c5  synthetic Deep shape, N shards of --shard-size x 96 (default 12,500,000: 100M over 8 GPUs), one HNSW per GPU built
      on the device, all 10,000 queries to all shards, rows stored to all peers by the search kernel, flags, K6 merge
      (weak scaling in the base).

No NCCL call sits in any timed loop; torch.distributed only bootstraps (IPC handles) and reduces the timings.
"""
import ctypes as C
import json
import os
import time

import numpy as np

K10 = 10


def sift_like(n, dim, ncent, seed):
    """non-negative, heavy-tailed integer-valued coordinates around cluster centres, capped at 218."""
    rc = np.random.default_rng(4321)
    cent = np.abs(rc.standard_normal((ncent, dim), dtype=np.float32)) * 40
    r = np.random.default_rng(seed)
    x = cent[r.integers(0, ncent, n)] + np.abs(r.standard_normal((n, dim), dtype=np.float32)) * 20
    return np.minimum(np.floor(x), 218).astype(np.float32)


def deep_like(n, dim, ncent, seed, row0=0, chunk=1 << 20):
    """unit-norm rows of a 65,536-centre mixture; rows [row0, row0+n) of the stream with this seed,
    generated chunk by chunk (one generator per chunk index, so any shard can be produced on its own)."""
    rc = np.random.default_rng(9876)
    cent = rc.standard_normal((ncent, dim), dtype=np.float32)
    out = np.empty((n, dim), np.float32)
    done = 0
    while done < n:
        g = row0 + done
        ci, off = divmod(g, chunk)
        r = np.random.default_rng([seed, ci])
        take = min(chunk - off, n - done)
        idx = r.integers(0, ncent, chunk)
        noise = r.standard_normal((chunk, dim), dtype=np.float32)
        x = cent[idx[off:off + take]] + np.float32(0.35) * noise[off:off + take]
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        out[done:done + take] = x
        done += take
    return out


def _max_over_ranks(v, world, dist, torch):
    if world == 1:
        return float(v)
    t = torch.tensor([v], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _peaks(root):
    try:
        return json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def run(a, env):
    """env: dict with torch, dist (or None), H, _ffi, ctx, rank, local_rank, world, ROOT, emit, ClockSampler, recall_at_k,
    alg_bytes, oracle_from_index, synth"""
    return {"c3": run_c3, "c4": run_c4, "c5": run_c5}[a.config](a, env)


# ------------------------------------------------------------------------------------------------------------------
def run_c3(a, E):
    torch, dist, H, _ffi, ctx = E["torch"], E["dist"], E["H"], E["_ffi"], E["ctx"]
    rank, world, lib = E["rank"], E["world"], E["_ffi"].lib()
    from hnsw_rs_b200 import sharded
    n_base, dim, nq_all, K = a.n_base or 1000000, 128, a.n_queries, K10
    workload = (f"C3 synthetic SIFT shape: {n_base}x{dim} L2 (non-negative heavy-tailed, 1024 centres, seed 3), {nq_all} queries "
                f"(seed 4) SPLIT over {world} GPU(s), replicated index, k={K}, quantised-L2")
    queries = sift_like(nq_all, dim, 1024, 4)
    t0 = time.time()
    if a.load_index:
        ix = H.HNSW.load(a.load_index, ctx=ctx)
    else:
        ix = H.HNSW.new(a.m, a.ef_cons, dim, ctx=ctx).insert_bulk(sift_like(n_base, dim, 1024, 3))
    build_s = time.time() - t0
    if a.save_index and rank == 0:
        ix.save(a.save_index)
    t0 = time.time()
    gt, _ = H.bruteforce_topk(ix._points(), queries, K, ctx=ctx)
    gt_s = time.time() - t0
    dq_all = torch.from_numpy(queries).cuda()
    d_ids_all = torch.empty((nq_all, K), dtype=torch.int32, device="cuda")

    def search_all(ef):
        _ffi.check(lib.hnswb200_search_dev(ctx.h, ix.h, dq_all.data_ptr(), nq_all, K, ef, d_ids_all.data_ptr(), None, None,
                                           None, None, None, None))
        torch.cuda.synchronize()
        return d_ids_all.cpu().numpy().astype(np.uint32)

    sweep, ef, rec = [], None, 0.0
    for e in ([a.ef] if a.ef else [40, 60, 80, 90, 100, 110, 120, 128, 160, 200, 256]):
        ids_full = search_all(e)
        rec = E["recall_at_k"](ids_full, gt)
        sweep.append((e, round(rec, 5)))
        ef = e
        if rec >= 0.99:
            break
    ids_full = search_all(ef)  # the single-GPU answer every rank checks the gathered rows against
    lo, hi = sharded.split_range(nq_all, rank, world)
    nq = hi - lo
    per = -(-nq_all // world)
    dq = dq_all[lo:hi].contiguous()
    d_ids = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    d_d = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    d_c = torch.empty(nq, dtype=torch.int32, device="cuda")
    pg = sharded.PeerGather(ctx, per, K)

    def step():
        pg.search(ix, dq.data_ptr(), nq, ef, d_ids.data_ptr(), d_d.data_ptr(), d_c.data_ptr())
        pg.signal_wait()

    sampler = E["ClockSampler"](E["local_rank"])
    if rank == 0:
        sampler.start()
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    total_ms = e0.elapsed_time(e1)
    kern_variant = _ffi.last_search_variant()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.barrier()
    got = pg.download()
    for r in range(world):
        x, y = sharded.split_range(nq_all, r, world)
        assert np.array_equal(got[r * per:r * per + (y - x)], ids_full[x:y]), "gathered rows differ from the single-GPU search"
    my_ms = total_ms
    total_ms = _max_over_ranks(total_ms, world, dist, torch)
    value = nq_all * a.steps / (total_ms / 1e3)
    # counters of this rank's slice -> algorithmic bytes of one launch
    d_h = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_e = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_f = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_nb = torch.empty(nq, dtype=torch.int32, device="cuda")
    _ffi.check(lib.hnswb200_search_dev(ctx.h, ix.h, dq.data_ptr(), nq, K, ef, d_ids.data_ptr(), d_d.data_ptr(), d_c.data_ptr(),
                                       d_h.data_ptr(), d_e.data_ptr(), d_f.data_ptr(), d_nb.data_ptr()))
    torch.cuda.synchronize()
    hops, evals, nbrs, flags = (x.cpu().numpy().astype(np.uint32) for x in (d_h, d_e, d_nb, d_f))
    ab = E["alg_bytes"](hops, nbrs, evals, nq, dim, K)
    # e2e: hnswb200_search on this rank's slice, host buffers
    hq = torch.from_numpy(queries[lo:hi].copy()).pin_memory()
    h_ids = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    h_d = torch.empty((nq, K), dtype=torch.float32).pin_memory()
    h_c = torch.empty(nq, dtype=torch.int32).pin_memory()

    def search_host():
        _ffi.check(lib.hnswb200_search(ctx.h, ix.h, C.cast(hq.data_ptr(), _ffi.f32p), nq, dim, K, ef,
                                       C.cast(h_ids.data_ptr(), _ffi.u32p), C.cast(h_d.data_ptr(), _ffi.f32p),
                                       C.cast(h_c.data_ptr(), _ffi.u32p), None))
    for _ in range(a.warmup):
        search_host()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        search_host()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, world, dist, torch)
    assert np.array_equal(h_ids.numpy().view(np.uint32), ids_full[lo:hi])
    if world > 1:
        dist.barrier()
    pg.close()
    if rank != 0:
        return 0
    peaks = _peaks(E["ROOT"])
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kern_ms = my_ms / a.steps
    cfg = {"workload": workload, "index": "HNSW M=%d ef_cons=%d, built on the device, replicated per GPU" % (a.m, a.ef_cons),
           "ef": ef, "recall_at_10": round(rec, 5), "ef_sweep": sweep, "queries_per_gpu": nq,
           "l2": "index (records + adjacency) is %.0f MB > 126 MB L2; no flush between steps" % (ix.len() * (144 + 128) / 1e6)}
    line = {"metric": "queries/sec at recall@10>=0.99", "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "e2e": {"value": nq_all * a.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(nq * dim * 4),
                    "d2h_bytes_per_step": int(nq * K * 8 + nq * 4),
                    "transfer": "hnswb200_search on each rank's slice, page-locked host buffers, synchronous steps"},
            "gpu_launches": 3 * a.steps, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kern_variant, "achieved": round(ab / (kern_ms / 1e3) / 1e9, 1), "peak": peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s", "unit": "GB/s",
                         "frac": round(ab / (kern_ms / 1e3) / 1e9 / peak, 4), "traffic": None,
                         "algorithmic_bytes_per_launch": ab, "kernel_ms": round(kern_ms, 4),
                         "note": "rank 0's slice; a step = search kernel (fused peer stores) + signal + wait kernels",
                         "per_query": {"hops": float(hops.mean()), "evals": float(evals.mean()), "nbr_ids": float(nbrs.mean()),
                                       "evals_p99": float(np.percentile(evals, 99)), "evals_max": int(evals.max())},
                         "visited_spill_queries": int(((flags & 4) != 0).sum()),
                         "visited_overflow_queries": int(((flags & 2) != 0).sum())},
            "setup": {"build_seconds": round(build_s, 2), "ground_truth_seconds": round(gt_s, 2)},
            "checks": {"gathered_rows_equal_single_gpu_search": True}}
    if not a.no_cpu_baseline:
        orc = E["oracle_from_index"](ix)
        cores = os.cpu_count() or 1
        ns = min(2000, nq)
        orc.search_batch(queries[lo:lo + 64], K, ef, threads=cores)
        t0 = time.perf_counter()
        oi, od, oc, oh, oe = orc.search_batch(queries[lo:lo + ns], K, ef, threads=cores)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                                "sample": f"first {ns} queries of rank 0's slice at ef={ef}, {cores} threads",
                                "parity_vs_gpu": {"queries_compared": ns,
                                                  "ids_identical": bool(np.array_equal(oi, ids_full[lo:lo + ns])),
                                                  "hops_identical": bool(np.array_equal(oh, hops[:ns])),
                                                  "evals_identical": bool(np.array_equal(oe, evals[:ns]))}}
        assert line["cpu_baseline"]["parity_vs_gpu"]["ids_identical"]
    E["emit"](line)
    return 0


# ------------------------------------------------------------------------------------------------------------------
def run_c4(a, E):
    torch, dist, H, _ffi, ctx = E["torch"], E["dist"], E["H"], E["_ffi"], E["ctx"]
    rank, world = E["rank"], E["world"]
    from hnsw_rs_b200 import sharded
    n_base, dim, nq, K = a.n_base or 1000000, 100, a.n_queries, 100
    workload = (f"C4 exact top-{K} ground truth (brute_force_nns under the quantised metric): {n_base}x{dim} base (C2 law, seed 1) "
                f"x {nq} queries (seed 2), base rows split over {world} GPU(s)")
    base = E["synth"](n_base, dim, a.ncent, 1)
    queries = E["synth"](nq, dim, a.ncent, 2)
    lo, hi = sharded.split_range(n_base, rank, world)
    full_pts = H.SimplePoints.new(base, ctx=ctx)                       # the unsharded checker (every rank holds 128 MB)
    pts = full_pts if world == 1 else H.SimplePoints.new(base[lo:hi], ctx=ctx)
    dq = torch.from_numpy(queries).cuda()
    ref_i = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    ref_d = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    lib = _ffi.lib()
    _ffi.check(lib.hnswb200_bruteforce_topk_dev(ctx.h, full_pts.h, dq.data_ptr(), nq, K, 0, ref_i.data_ptr(), ref_d.data_ptr()))
    px = sharded.PeerExchange(ctx, nq, K)
    out_i = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    out_d = torch.empty((nq, K), dtype=torch.float32, device="cuda")

    def step():
        px.shard_bruteforce(pts, dq.data_ptr(), nq, lo)
        px.signal_wait()
        px.merge(out_i.data_ptr(), out_d.data_ptr())

    sampler = E["ClockSampler"](E["local_rank"])
    if rank == 0:
        sampler.start()
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    total_ms = _max_over_ranks(e0.elapsed_time(e1), world, dist, torch)
    clocks = sampler.stop() if rank == 0 else None
    assert torch.equal(out_i, ref_i) and torch.equal(out_d.view(torch.int32), ref_d.view(torch.int32)), \
        "merged base-sharded top-k differs from the unsharded brute force"
    # e2e: host queries in, merged ids / distances out, every step
    hq = torch.from_numpy(queries).pin_memory()
    h_i = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    h_d = torch.empty((nq, K), dtype=torch.float32).pin_memory()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        dq.copy_(hq, non_blocking=True)
        step()
        h_i.copy_(out_i, non_blocking=True)
        h_d.copy_(out_d, non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, world, dist, torch)
    px.close()
    if rank != 0:
        return 0
    peaks = _peaks(E["ROOT"])
    peak = float(peaks.get("bf16_tflops_sustained", 1409.6))
    ms = total_ms / a.steps
    ops = 2.0 * nq * n_base * dim
    line = {"metric": "queries/sec of exact top-100 ground truth (recall 1.0 by construction)", "value": nq * a.steps / (total_ms / 1e3),
            "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8 codes, s32 accumulate (tcgen05 kind::i8) + exact f32 re-rank",
            "data": "synthetic", "config": {"workload": workload, "k": K, "rows_per_gpu": hi - lo},
            "e2e": {"value": nq * a.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(nq * dim * 4),
                    "d2h_bytes_per_step": int(nq * K * 8)},
            "gpu_launches": "per step: quantise + prepare + per chunk (filter, re-rank/merge) + 2 peer puts + signal + wait + K6",
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "hb::bf_tc_filter_kernel (+ exact re-rank, merge, exchange: the whole step is timed)",
                         "achieved": round(ops / world / (ms / 1e3) / 1e12, 2), "peak": peak,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (dense bf16; the int8 pipe peaks at 2x)" if peaks else "fallback",
                         "unit": "TFLOP/s", "frac": round(ops / world / (ms / 1e3) / 1e12 / peak, 4), "traffic": None,
                         "note": "useful ops per GPU = 2*Q*(N/G)*dim over the WHOLE step (first exact chunk, filter chunks, re-rank, merges, exchange)"},
            "checks": {"merged_equals_unsharded_bruteforce": True}}
    if not a.no_cpu_baseline:
        from oracle import pyoracle as O
        codes, mins, deltas, levels = full_pts.download()
        flat = [(np.arange(n_base, dtype=np.uint32), np.zeros(n_base + 1, np.uint64), np.zeros(0, np.uint32))]
        orc = O.Index.from_parts(4, 8, dim, 0, codes, mins, deltas, np.zeros(n_base, np.uint8), flat)
        ns = 16
        t0 = time.perf_counter()
        oi, od = orc.bruteforce(queries[:ns], K)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns / dt, "unit": "queries/s", "cores": 1, "kind": "port",
                                "sample": f"first {ns} queries against all {n_base} rows, the oracle's brute force (one thread)",
                                "parity_vs_gpu": {"ids_identical": bool(np.array_equal(oi, out_i[:ns].cpu().numpy().view(np.uint32))),
                                                  "dist_bits_identical": bool(np.array_equal(od.view(np.uint32), out_d[:ns].cpu().numpy().view(np.uint32)))}}
        assert line["cpu_baseline"]["parity_vs_gpu"]["ids_identical"]
    E["emit"](line)
    return 0


# ------------------------------------------------------------------------------------------------------------------
def run_c5(a, E):
    torch, dist, H, _ffi, ctx = E["torch"], E["dist"], E["H"], E["_ffi"], E["ctx"]
    rank, world = E["rank"], E["world"]
    from hnsw_rs_b200 import sharded
    per, dim, nq, K = a.shard_size, 96, a.n_queries, K10
    workload = (f"C5 synthetic Deep shape: {world} shard(s) x {per} x {dim} unit-norm rows (65,536-centre mixture, seed 5), one HNSW "
                f"per GPU, all {nq} queries (seed 6) to all shards, k={K}, quantised-L2")
    queries = deep_like(nq, dim, 65536, 6)
    lo = rank * per
    t0 = time.time()
    ix = H.HNSW.new(a.m, a.ef_cons, dim, ctx=ctx)
    chunk = 1 << 21
    for s in range(0, per, chunk):
        ix = ix.insert_bulk(deep_like(min(chunk, per - s), dim, 65536, 5, row0=lo + s))
    build_s = time.time() - t0
    dq = torch.from_numpy(queries).cuda()
    px = sharded.PeerExchange(ctx, nq, K)
    out_i = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    out_d = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    # exact ground truth over ALL shards: per-shard brute force + the same exchange + merge
    t0 = time.time()
    px.shard_bruteforce(ix._points(), dq.data_ptr(), nq, lo)
    px.signal_wait()
    px.merge(out_i.data_ptr(), out_d.data_ptr())
    ctx.sync()
    gt = out_i.cpu().numpy().view(np.uint32).copy()
    gt_s = time.time() - t0

    def step(ef):
        px.shard_search(ix, dq.data_ptr(), nq, ef, lo)
        px.signal_wait()
        px.merge(out_i.data_ptr(), out_d.data_ptr())

    sweep, ef, rec = [], None, 0.0
    for e in ([a.ef] if a.ef else [40, 64, 100, 128, 160, 200, 256, 320, 400, 512, 640, 800, 1024]):
        step(e)
        ctx.sync()
        rec = E["recall_at_k"](out_i.cpu().numpy().view(np.uint32), gt)
        sweep.append((e, round(rec, 5)))
        ef = e
        if rec >= 0.99:
            break
    sampler = E["ClockSampler"](E["local_rank"])
    if rank == 0:
        sampler.start()
    for _ in range(a.warmup):
        step(ef)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    for _ in range(a.steps):
        step(ef)
    e1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    my_ms = e0.elapsed_time(e1)
    total_ms = _max_over_ranks(my_ms, world, dist, torch)
    kern_variant = _ffi.last_search_variant()
    clocks = sampler.stop() if rank == 0 else None
    merged = out_i.cpu().numpy().view(np.uint32).copy()
    # counters of this rank's shard search
    lib = _ffi.lib()
    d_ids = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    d_h = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_e = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_f = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_nb = torch.empty(nq, dtype=torch.int32, device="cuda")
    _ffi.check(lib.hnswb200_search_dev(ctx.h, ix.h, dq.data_ptr(), nq, K, ef, d_ids.data_ptr(), None, None, d_h.data_ptr(),
                                       d_e.data_ptr(), d_f.data_ptr(), d_nb.data_ptr()))
    torch.cuda.synchronize()
    hops, evals, nbrs, sflags = (x.cpu().numpy().astype(np.uint32) for x in (d_h, d_e, d_nb, d_f))
    local_ids = d_ids.cpu().numpy().view(np.uint32)
    ab = E["alg_bytes"](hops, nbrs, evals, nq, dim, K)
    # e2e: host queries in, merged ids out, every step
    hq = torch.from_numpy(queries).pin_memory()
    h_i = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        dq.copy_(hq, non_blocking=True)
        step(ef)
        h_i.copy_(out_i, non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, world, dist, torch)
    assert np.array_equal(h_i.numpy().view(np.uint32), merged)
    px.close()
    if rank != 0:
        return 0
    peaks = _peaks(E["ROOT"])
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kern_ms = my_ms / a.steps
    line = {"metric": "queries/sec at recall@10>=0.99", "value": nq * a.steps / (total_ms / 1e3), "unit": "queries/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "index": "one HNSW per GPU, M=%d ef_cons=%d, built on the device" % (a.m, a.ef_cons),
                       "ef": ef, "recall_at_10": round(rec, 5), "ef_sweep": sweep, "total_base_rows": per * world},
            "e2e": {"value": nq * a.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(nq * dim * 4),
                    "d2h_bytes_per_step": int(nq * K * 4)},
            "gpu_launches": 4 * a.steps, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": kern_variant, "achieved": round(ab / (kern_ms / 1e3) / 1e9, 1), "peak": peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s", "unit": "GB/s",
                         "frac": round(ab / (kern_ms / 1e3) / 1e9 / peak, 4), "traffic": None, "algorithmic_bytes_per_launch": ab,
                         "kernel_ms": round(kern_ms, 4),
                         "note": "rank 0's shard; a step = search kernel (fused peer stores of id + distance rows) + signal + wait + K6 merge",
                         "per_query": {"hops": float(hops.mean()), "evals": float(evals.mean()), "nbr_ids": float(nbrs.mean())}},
            "setup": {"build_seconds": round(build_s, 1), "ground_truth_seconds": round(gt_s, 2),
                      "inserts_per_second": round(per / build_s)}}
    if not a.no_cpu_baseline:
        orc = E["oracle_from_index"](ix)
        cores = os.cpu_count() or 1
        ns = min(500, nq)
        t0 = time.perf_counter()
        oi, od, oc, oh, oe = orc.search_batch(queries[:ns], K, ef, threads=cores)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ns / dt, "unit": "queries/s", "cores": cores, "kind": "port",
                                "sample": f"first {ns} queries on rank 0's shard at ef={ef}, {cores} threads (one shard of {world})",
                                "parity_vs_gpu": {"queries_compared": ns, "ids_identical": bool(np.array_equal(oi, local_ids[:ns])),
                                                  "hops_identical": bool(np.array_equal(oh, hops[:ns])),
                                                  "evals_identical": bool(np.array_equal(oe, evals[:ns])),
                                                  # shards of more than 2^21 rows run the round-1 kernel, whose evaluation counter
                                                  # may over-count for queries that raise flag bit 1 (visited-table overflow)
                                                  "evals_identical_where_not_flagged": bool(np.array_equal(
                                                      oe[(sflags[:ns] & 2) == 0], evals[:ns][(sflags[:ns] & 2) == 0])),
                                                  "flagged_queries": int(((sflags[:ns] & 2) != 0).sum())}}
        assert line["cpu_baseline"]["parity_vs_gpu"]["ids_identical"]
    E["emit"](line)
    return 0
