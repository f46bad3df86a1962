#!/usr/bin/env python3
"""bench.py -- queries/sec at recall@10 >= 0.99 on the synthetic GloVe-100-shaped workload
(BASELINE.json configs[1]: 1,183,514 x 100 unit-norm clustered mixture, 10,000 queries, ef sweep),
1 B200 per process, one JSON line on rank 0.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU restatement of the reference, same config

A step = one pass of HNSW::ann_by_vector over the whole query batch (hnswb200_search_dev).
value  = queries/s with the queries already resident in HBM (CUDA events, max over ranks).
e2e    = the same through the host-buffer entry point hnswb200_search (pinned host queries in,
         ids / distances / counts out), host<->device copies inside the timed region.
N > 1: the index is replicated, every rank searches its own 10,000 queries (weak scaling), and
the ids are all-gathered over NCCL inside the timed region so that every rank holds all results
(fused into the search kernel: peer stores over NVLink into every rank's buffer; NCCL only checks it).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EF_SWEEP = [10, 20, 40, 48, 52, 56, 57, 58, 59, 60, 62, 64, 72, 80, 100, 160, 320, 640]  # BASELINE's sweep, refined between 40 and 80
K = 10


def synth(n, dim, ncent, seed, sigma=0.35):
    """Clustered Gaussian mixture, rows L2-normalised (cosine == L2 on unit vectors)."""
    rc = np.random.default_rng(1234)
    cent = rc.standard_normal((ncent, dim), dtype=np.float32)
    r = np.random.default_rng(seed)
    x = cent[r.integers(0, ncent, n)]
    x += np.float32(sigma) * r.standard_normal((n, dim), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x, np.float32)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  The timed region of a
    default run is only some tens of milliseconds, far shorter than nvidia-smi's start-up, so the samples come from an
    NVML polling thread (pynvml: the same counters nvidia-smi prints) that is already running when the region starts;
    only the samples whose time stamps fall inside [mark_begin, mark_end] are reported.  nvidia-smi -lms is the fallback."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []      # (t, sm_mhz, reasons bit mask)
        self.max_mhz = None
        self.stop_flag = False
        self.t0 = self.t1 = None
        self.kind = None
        self.t = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.gpu).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].strip().isdigit() else self.gpu
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]

            def poll():
                while not self.stop_flag:
                    try:
                        mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                        r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.rows.append((time.perf_counter(), mhz, sum(1 << i for i, b in enumerate(bits) if r & b)))
                    except Exception:
                        pass
                    time.sleep(0.0005)
            # the first queries after nvmlInit are slow (lazy initialisation, tens of milliseconds): pay for them here,
            # synchronously, so that the polling thread is at its steady rate (~1 ms per sample) when the timed region starts
            for _ in range(3):
                nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.kind = "nvml"
            self.t = threading.Thread(target=poll, daemon=True)
            self.t.start()
            return
        except Exception:
            pass
        try:
            q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def read():
                for line in self.p.stdout:
                    f = [x.strip() for x in line.split(",")]
                    try:
                        self.max_mhz = float(f[2])
                        self.rows.append((time.perf_counter(), float(f[1]),
                                          sum(1 << i for i, v in enumerate(f[3:7]) if v.lower().startswith("active"))))
                    except (ValueError, IndexError):
                        continue
            self.kind = "nvidia-smi"
            self.t = threading.Thread(target=read, daemon=True)
            self.t.start()
            time.sleep(0.5)  # nvidia-smi needs a moment before its first line
        except Exception:
            self.kind = None

    def samples_inside(self):
        return sum(1 for r in list(self.rows) if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or r[0]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.kind:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no NVML and no nvidia-smi"]}
        self.stop_flag = True
        if self.kind == "nvidia-smi":
            self.p.terminate()
        self.t.join(timeout=2)
        inside = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or r[0])]
        near = inside
        if not near and self.rows and self.t0 is not None:
            # a region shorter than one polling period: the samples bracketing it
            mid = 0.5 * (self.t0 + (self.t1 or self.t0))
            near = sorted(self.rows, key=lambda r: abs(r[0] - mid))[:2]
        mask = 0
        for r in near:
            mask |= r[2]
        return {"sm_mhz": float(np.median([r[1] for r in near])) if near else None, "sm_max_mhz": self.max_mhz,
                "samples": len(inside), "samples_total": len(self.rows), "source": self.kind,
                "window": "samples taken between the first launch of the timed region and its final synchronize",
                "reasons": [n for i, n in enumerate(self.NAMES) if mask >> i & 1]}


def recall_at_k(ids, gt):
    hits = 0
    for i in range(gt.shape[0]):
        hits += len(set(gt[i].tolist()) & set(ids[i].tolist()))
    return hits / gt.size


def alg_bytes(hops, nbrs, evals, nq, dim, k, rec=None):
    """B_q = sum_layers[hops*8 + 4*deg(expanded) + evals*rec] + 4*dim + 8*k (counted, not modelled);
    rec = 8 + dim bytes per evaluated QuantVec record, 4*dim per FullVec record."""
    rec = (8 + dim) if rec is None else rec
    return float(hops.astype(np.float64).sum() * 8 + nbrs.astype(np.float64).sum() * 4 +
                 evals.astype(np.float64).sum() * rec + nq * (4 * dim + 8 * k))


def oracle_from_index(ix):
    from oracle import pyoracle as O
    p = ix.params
    layers = [ix.export_layer(l) for l in range(ix.nb_layers())]
    if ix.vec_type == "full":
        vals, levels = ix._points().values()
        return O.Index.from_parts(p.m, p.ef_cons, p.dim, p.ep, vals, None, None, levels, layers)
    codes, mins, deltas, levels = ix._points().download()
    return O.Index.from_parts(p.m, p.ef_cons, p.dim, p.ep, codes, mins, deltas, levels, layers)


def pick_ef(search_fn, gt, efs):
    table = []
    chosen = None
    for ef in efs:
        ids = search_fn(ef)
        r = recall_at_k(ids, gt)
        table.append((ef, round(r, 5)))
        if r >= 0.99 and chosen is None:
            chosen = (ef, r)
            break
    if chosen is None:
        chosen = (efs[-1], table[-1][1])
    return chosen, table


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line, on the real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout carries exactly one JSON line: everything libraries print while we run (NCCL's version banner, ...)
    # is sent to stderr by pointing fd 1 at fd 2 for the duration of the run
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)  # 100 steps = 64 ms of kernels: long enough for the clock sampler
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE.json configuration: c2 = the headline (configs[1]); c3 / c4 / c5 = configs[2..4], see "
                         "bench_configs.py")
    ap.add_argument("--shard-size", type=int, default=12500000, help="c5: rows per GPU shard")
    ap.add_argument("--n-base", type=int, default=0, help="base rows (default: what the configuration names)")
    ap.add_argument("--n-queries", type=int, default=10000)
    ap.add_argument("--dim", type=int, default=100)
    ap.add_argument("--m", type=int, default=16)
    ap.add_argument("--ef-cons", type=int, default=200)
    ap.add_argument("--ncent", type=int, default=2048)
    ap.add_argument("--ef", type=int, default=0, help="skip the sweep and use this ef")
    ap.add_argument("--vec-type", default="quant", choices=["quant", "full"],
                    help="the reference's `type VecType` (points/src/point.rs:4): quant = QuantVec (the reference as committed, the "
                         "headline); full = FullVec (f32 vectors, strictly sequential distance)")
    ap.add_argument("--save-index", default="", help="HNSW::save the built index here (profiling helper)")
    ap.add_argument("--load-index", default="", help="HNSW::load instead of building (profiling helper)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle timing (profiling helper)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.config == "c2" and not a.n_base:
        a.n_base = 1183514

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if a.impl == "reference" and rank != 0:
        return 0  # the CPU arm runs on rank 0 alone

    import torch
    import hnsw_rs_b200 as H
    from hnsw_rs_b200 import _ffi

    have_gpu = torch.cuda.is_available()
    if not have_gpu:
        raise SystemExit("bench.py needs a CUDA device: hnsw_rs_b200 has no CPU fallback "
                         "(the reference arm also builds its index with the device builder)")
    torch.cuda.set_device(local_rank)

    dist = None
    if world > 1 and a.impl == "b200":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = H.Context(local_rank)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    # adopted stream: consecutive searches may only overlap with our promise that nothing else is enqueued between them
    # (true for every loop below: the query buffers are written once, before the first search)
    ctx.set_overlap(True)

    if a.config != "c2":
        import bench_configs
        if a.impl == "reference":
            emit({"impl": "reference", "unavailable": "--impl reference is defined for the headline configuration (c2); "
                  "c3/c4/c5 report the oracle as cpu_baseline inside their own line"})
            return 0
        env = {"torch": torch, "dist": dist if world > 1 else None, "H": H, "_ffi": _ffi, "ctx": ctx, "rank": rank,
               "local_rank": local_rank, "world": world, "ROOT": ROOT, "emit": emit, "ClockSampler": ClockSampler,
               "recall_at_k": recall_at_k, "alg_bytes": alg_bytes, "oracle_from_index": oracle_from_index, "synth": synth}
        rc = bench_configs.run(a, env)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return rc

    workload = (f"C2 synthetic GloVe-100 shape: {a.n_base}x{a.dim} unit-norm clustered mixture ({a.ncent} centres, "
                f"sigma 0.35, seed 1), {a.n_queries} queries/GPU (seed 2+rank), k={K}, " +
                ("quantised-L2 (== cosine rank)" if a.vec_type == "quant" else "f32 L2, VecType = FullVec (== cosine rank)"))
    queries = synth(a.n_queries, a.dim, a.ncent, 2 + rank)
    nq, dim = queries.shape

    t0 = time.time()
    if a.load_index:
        ix = H.HNSW.load(a.load_index, ctx=ctx)
    else:
        base = synth(a.n_base, a.dim, a.ncent, 1)
        t0 = time.time()
        ix = H.HNSW.new(a.m, a.ef_cons, a.dim, ctx=ctx, vec_type=a.vec_type).insert_bulk(base)
        del base
    build_s = time.time() - t0
    if a.save_index and rank == 0:
        ix.save(a.save_index)

    # exact ground truth under the quantised metric on the device (K5)
    t0 = time.time()
    gt, _ = H.bruteforce_topk(ix._points(), queries, K, ctx=ctx)
    gt_s = time.time() - t0

    dq = torch.from_numpy(queries).cuda()
    d_ids = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    d_d = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    d_cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_h = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_e = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_f = torch.empty(nq, dtype=torch.int32, device="cuda")
    d_nb = torch.empty(nq, dtype=torch.int32, device="cuda")
    lib = _ffi.lib()

    def search_dev(ef, counters=False):
        """ids, distances and counts always; the diagnostic counters (hops, evaluations, neighbour ids read, flags) only
        on request -- they are optional outputs of the entry point and the kernel variant without them is what a caller
        who wants answers runs, so that is what is timed; the roofline's byte count comes from one extra launch with
        counters on the same batch."""
        _ffi.check(lib.hnswb200_search_dev(ctx.h, ix.h, dq.data_ptr(), nq, K, ef, d_ids.data_ptr(), d_d.data_ptr(),
                                           d_cnt.data_ptr(), d_h.data_ptr() if counters else None,
                                           d_e.data_ptr() if counters else None, d_f.data_ptr() if counters else None,
                                           d_nb.data_ptr() if counters else None))

    def search_ids(ef):
        search_dev(ef)
        torch.cuda.synchronize()
        return d_ids.cpu().numpy().astype(np.uint32)

    if a.ef:
        ef, rec = a.ef, recall_at_k(search_ids(a.ef), gt)
        sweep = [(ef, round(rec, 5))]
    else:
        (ef, rec), sweep = pick_ef(search_ids, gt, EF_SWEEP)
    if world > 1 and a.impl == "b200":
        # every rank must run the same ef: take the largest any rank needs
        t = torch.tensor([ef], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ef = int(t.item())
        rec = recall_at_k(search_ids(ef), gt)

    cfg = {"workload": workload, "index": "HNSW M=%d ef_cons=%d, built on the device (batched inserts), replicated per GPU"
           % (a.m, a.ef_cons), "ef": ef, "recall_at_10": round(rec, 5), "ef_sweep": sweep,
           "l2": "index (records + adjacency) is %.0f MB > 126 MB L2; no flush between steps"
                 % ((ix.len() * (128 + 4 * 2 * a.m)) / 1e6)}
    setup = {"build_seconds": round(build_s, 2), "ground_truth_seconds": round(gt_s, 2),
             "index_built_by": "device builder (hnswb200_build), untimed setup"}

    if a.impl == "reference":
        return run_reference(a, ix, queries, gt, ef, rec, cfg, setup)

    # ---------------- device-resident throughput (value) ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # already polling during the warm-up, so the timed region is covered from its first launch
    for _ in range(a.warmup):
        search_dev(ef)
    torch.cuda.synchronize()
    if world > 1:
        # the all-gather is fused into the search kernel: every rank stores its id rows straight into the result
        # buffer of every other rank (CUDA IPC mapping, NVLink / NVSwitch peer stores) while its other queries compute
        from hnsw_rs_b200 import sharded
        pg = sharded.PeerGather(ctx, nq, K)
        reference_gather = torch.empty((world * nq, K), dtype=torch.int32, device="cuda")
        dist.all_gather_into_tensor(reference_gather, d_ids)  # NCCL, outside the timed region: the checker of the fused path
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    for s in range(a.steps):
        # consecutive searches are programmatic dependent launches: the blocks of step s+1 take over the SMs
        # that the last long queries of step s leave idle
        if world > 1:
            pg.search(ix, dq.data_ptr(), nq, ef, d_ids.data_ptr(), d_d.data_ptr(), d_cnt.data_ptr())
        else:
            search_dev(ef)
    e1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    if world > 1:
        dist.barrier()
    total_ms = e0.elapsed_time(e1)
    replayed = False
    if rank == 0 and sampler.kind and not sampler.samples_inside():
        # a timed region shorter than the sampler's period (few steps, slow NVML): replay the very same launches, untimed,
        # for half a second under the sampler, so that the clocks line still describes this loop under load
        replayed = True
        sampler.mark_begin()
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            for _ in range(20):
                search_dev(ef)
            torch.cuda.synchronize()
        sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None and replayed:
        clocks["window"] = "no sample fell inside the timed region: untimed replay of the same launch loop (0.5 s) right after it"
    if world > 1:
        got = torch.from_numpy(pg.download().view(np.int32)).cuda()
        assert torch.equal(got, reference_gather), "fused all-gather differs from the NCCL all-gather"
        dist.barrier()
        pg.close()
    # the search kernel's launch duration, on its own (not overlapped with a neighbour): mean over the same
    # number of launches, each bracketed by events (an event between two launches serialises them)
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for s in range(a.steps):
        k_ev[s][0].record()
        search_dev(ef)
        k_ev[s][1].record()
    torch.cuda.synchronize()
    kern_alone_ms = float(np.mean([x.elapsed_time(y) for x, y in k_ev]))
    # average launch duration inside the timed region (launches overlap at their boundaries): the same definition at
    # every N (this rank's own timed region; `value` uses the max over ranks)
    kern_ms = total_ms / a.steps
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * nq * a.steps / (total_ms / 1e3)

    kernel_variant = _ffi.last_search_variant()  # of the timed launches (asked of the library, not composed here)
    search_dev(ef, counters=True)
    torch.cuda.synchronize()
    stats_ids = d_ids.cpu().numpy().astype(np.uint32)
    stats_dist_bits = d_d.cpu().numpy().view(np.uint32).copy()
    stats_counts = d_cnt.cpu().numpy().astype(np.uint32)
    hops = d_h.cpu().numpy().astype(np.uint32)
    evals = d_e.cpu().numpy().astype(np.uint32)
    nbrs = d_nb.cpu().numpy().astype(np.uint32)
    flags = d_f.cpu().numpy()
    full = ix.vec_type == "full"
    ab = alg_bytes(hops, nbrs, evals, nq, dim, K, 4 * dim if full else None)

    # ---------------- end to end through the host-buffer entry point ----------------
    hq = torch.from_numpy(queries).pin_memory()
    h_ids = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    h_d = torch.empty((nq, K), dtype=torch.float32).pin_memory()
    h_c = torch.empty(nq, dtype=torch.int32).pin_memory()

    # the ctypes pointer objects are built once: the timed loop is the call itself
    p_hq, p_ids, p_d, p_c = (C.cast(hq.data_ptr(), _ffi.f32p), C.cast(h_ids.data_ptr(), _ffi.u32p),
                             C.cast(h_d.data_ptr(), _ffi.f32p), C.cast(h_c.data_ptr(), _ffi.u32p))

    def search_host():
        _ffi.check(lib.hnswb200_search(ctx.h, ix.h, p_hq, nq, dim, K, ef, p_ids, p_d, p_c, None))

    for _ in range(a.warmup):
        search_host()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        search_host()  # synchronous: returns when the results are in host memory
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    search_dev(ef)
    torch.cuda.synchronize()
    assert np.array_equal(h_ids.numpy(), d_ids.cpu().numpy()), "host and device entry points disagree"
    e2e = {"value": world * nq * a.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(nq * dim * 4),
           "d2h_bytes_per_step": int(nq * K * 8 + nq * 4),
           "transfer": "hnswb200_search with page-locked host buffers: first wave of queries by DMA, the others read in "
                       "place by the kernel; ids / distances / counts written in place over PCIe; every step is synchronous"}

    # the same from host memory with several batches in flight (hnswb200_search_async + one hnswb200_ctx_sync): what a
    # server that pipelines its batches sees; reported next to the synchronous number, not instead of it
    nbuf = 4
    h_ids_p = [torch.empty((nq, K), dtype=torch.int32).pin_memory() for _ in range(nbuf)]

    p_ids_p = [C.cast(t.data_ptr(), _ffi.u32p) for t in h_ids_p]

    def search_async(i):
        _ffi.check(lib.hnswb200_search_async(ctx.h, ix.h, p_hq, nq, dim, K, ef, p_ids_p[i % nbuf], None, None))
    for i in range(a.warmup):
        search_async(i)
    ctx.sync()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        search_async(i)
    ctx.sync()
    pipe_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([pipe_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pipe_s = float(t.item())
    assert np.array_equal(h_ids_p[(a.steps - 1) % nbuf].numpy(), h_ids.numpy()), "asynchronous and synchronous host paths disagree"
    e2e["pipelined"] = {"value": world * nq * a.steps / pipe_s, "unit": "queries/s", "batches_in_flight": "all steps, one sync",
                        "note": "ids only (ann_by_vector returns ids); every step reads its queries from and writes its ids "
                                "to page-locked host memory"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = ab / (kern_ms / 1e3) / 1e9
    # dram bytes per launch of THIS kernel variant from its ncu --set full capture, if one is committed for it
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "search_kernel_traffic.json")))
        if t.get("kernel_variant") == kernel_variant and t.get("ef") == ef:
            traffic = t.get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel_variant, "achieved": round(achieved, 1),
                "peak": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "frac_launched_alone": round(ab / (kern_alone_ms / 1e3) / 1e9 / peak, 4),
                "algorithmic_bytes_per_launch": ab, "kernel_ms": round(kern_ms, 4),
                "kernel_ms_launched_alone": round(kern_alone_ms, 4),
                "timing": "achieved = algorithmic bytes per launch / (timed region / launches); consecutive launches "
                          "overlap at their boundaries (programmatic dependent launch); kernel_ms_launched_alone brackets "
                          "every launch with events, which serialises them",
                "counters": "hops / evaluations / neighbour ids are counted by one extra launch of the same batch with the "
                            "optional counter outputs; the timed launches return ids, distances and counts only",
                "per_query": {"hops": float(hops.mean()), "evals": float(evals.mean()), "nbr_ids": float(nbrs.mean()),
                              "bytes": ab / nq},
                "visited_spill_queries": int(((flags & 4) != 0).sum()),      # exact spill list used: everything exact
                "visited_overflow_queries": int(((flags & 2) != 0).sum())}   # evaluation counter may over-count

    # ---------------- CPU baseline: the oracle (port of the reference) on this box's cores ----------------
    if a.no_cpu_baseline:
        line = {"metric": "queries/sec at recall@10>=0.99", "value": value, "unit": "queries/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "config": cfg, "e2e": e2e,
                "roofline": roofline, "clocks": clocks, "setup": setup, "note": "profiling helper run, no cpu_baseline"}
        emit(line)
        if world > 1:
            dist.destroy_process_group()
        return 0
    orc = oracle_from_index(ix)
    cores = os.cpu_count() or 1
    sample = queries
    orc.search_batch(sample[:256], K, ef, threads=cores)  # warm-up
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        oids, odists, ocounts, oh, oe = orc.search_batch(sample, K, ef, threads=cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    # parity of the headline batch against the oracle on the same graph, over ALL queries at the bench's ef:
    # the timed variant's ids (d_ids holds the last timed-variant launch) and the counter variant's everything
    timed_ids = d_ids.cpu().numpy().astype(np.uint32)
    parity = {
        "queries_compared": int(nq), "ef": int(ef),
        "ids_identical": bool(np.array_equal(oids, timed_ids) and np.array_equal(oids, stats_ids)),
        "dist_bits_identical": bool(np.array_equal(odists.view(np.uint32), stats_dist_bits)),
        "counts_identical": bool(np.array_equal(ocounts.astype(np.uint32), stats_counts)),
        "hops_identical": bool(np.array_equal(oh, hops)),
        "evals_identical": bool(np.array_equal(oe, evals)),
        "evals_mismatch_queries": int((oe != evals).sum()),
        "ids_mismatch_queries": int((oids != timed_ids).any(axis=1).sum()),
    }
    assert parity["ids_identical"] and parity["dist_bits_identical"] and parity["counts_identical"] and \
        parity["hops_identical"], f"GPU search differs from the oracle on the headline batch: {parity}"
    n1 = min(2000, nq)
    t0 = time.perf_counter()
    orc.search_batch(sample[:n1], K, ef, threads=1)
    dt1 = time.perf_counter() - t0
    cpu = {"value": len(sample) / best, "unit": "queries/s", "cores": cores, "kind": "port",
           "sample": f"all {len(sample)} queries of the step at ef={ef}, best of 3 passes, {cores} threads over a shared "
                     f"read-only index (the graph the device built)",
           "single_thread_qps": n1 / dt1, "parity_vs_gpu": parity}

    line = {"metric": "queries/sec at recall@10>=0.99", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "e2e": e2e, "gpu_launches": a.steps, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "setup": setup}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_reference(a, ix, queries, gt, ef, rec, cfg, setup):
    """The reference's own CPU implementation of the path (the oracle port: no Rust toolchain here),
    all host threads, same index / queries / ef.  One step = all queries of the batch."""
    orc = oracle_from_index(ix)
    cores = os.cpu_count() or 1
    nq = queries.shape[0]
    for _ in range(a.warmup):
        orc.search_batch(queries, K, ef, threads=cores)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ids, _, _, _, _ = orc.search_batch(queries, K, ef, threads=cores)
    dt = time.perf_counter() - t0
    value = nq * a.steps / dt
    setup = dict(setup)
    setup["recall_at_10_cpu"] = round(recall_at_k(ids, gt), 5)
    setup["index_built_by"] = "device builder (setup, untimed); the oracle searches that graph"
    line = {"impl": "reference", "metric": "queries/sec at recall@10>=0.99", "value": value, "unit": "queries/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"all {nq} queries per step at ef={ef}, {cores} threads"},
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "setup": setup}
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
