#!/usr/bin/env python3
"""A/B of several builds of libhnsw_b200.so in ONE process (one box, one index, one set of queries): every build loads the
same saved index, answers the same batches with the diagnostic counters on (ids, distance bits, counts, hops, evaluations,
flags compared with the first build's, which is the tree's verified default), and is then timed like bench.py times
`value` (device-resident, launches overlapping).  Not a bench value: it ranks variants, profiles/r02_ab_variants.txt
records the outcome.

usage: python tools/dev/ab_multi.py --out gpurun_out/r2_u.json name=path/to/lib.so [name=path ...]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def load(path):
    from hnsw_rs_b200 import _ffi
    L = C.CDLL(os.path.abspath(path))
    for name, (res, args) in _ffi.SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    return L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--out", default="gpurun_out/ab_multi.json")
    ap.add_argument("--index", default="/tmp/ix")
    ap.add_argument("--ef", type=int, default=57)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--n-base", type=int, default=1183514)
    a = ap.parse_args()

    import torch
    import bench
    from hnsw_rs_b200 import _ffi
    vp = _ffi.vp
    K = 10
    nq, dim = 10000, 100
    queries = bench.synth(nq, dim, 2048, 2)
    names = [s.split("=", 1)[0] for s in a.libs]
    paths = [s.split("=", 1)[1] for s in a.libs]

    torch.cuda.set_device(0)
    if not os.path.isdir(a.index):
        import hnsw_rs_b200 as H  # the tree's default build makes the index
        t0 = time.time()
        base = bench.synth(a.n_base, dim, 2048, 1)
        ix = H.HNSW.new(16, 200, dim).insert_bulk(base)
        ix.save(a.index)
        del ix, base
        print(f"index built and saved in {time.time() - t0:.1f} s", flush=True)

    dq = torch.from_numpy(queries).cuda()
    bufs = {k: torch.empty((nq, K) if k in ("ids", "d") else (nq,), dtype=torch.int32 if k != "d" else torch.float32,
                           device="cuda") for k in ("ids", "d", "cnt", "h", "e", "f", "nb")}
    stream = torch.cuda.current_stream().cuda_stream
    # (label, ef, HNSWB200_FAST_NB): the headline, a short and the longest two-keys-per-lane list, the four-keys-per-lane
    # variant, and a table so small that the slow path and the exact spill set carry the query
    cases = [("ef%d" % a.ef, a.ef, None), ("ef10", 10, None), ("ef64", 64, None), ("ef100", 100, None),
             ("ef%d/nb258" % a.ef, a.ef, "258"), ("ef100/nb300", 100, "300")]
    ref = {}
    out = {"cases": [c[0] for c in cases], "runs": []}

    def dump():
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)

    def one(name, path):
        L = load(path)

        def chk(rc):
            if rc != 0:
                raise RuntimeError(f"{name}: [{rc}] {L.hnswb200_last_error().decode(errors='replace')}")
        ctx, ix = vp(), vp()
        chk(L.hnswb200_ctx_create(0, C.byref(ctx)))
        chk(L.hnswb200_ctx_set_stream(ctx, vp(stream)))
        chk(L.hnswb200_ctx_set_overlap(ctx, 1))
        chk(L.hnswb200_index_load_dir(ctx, a.index.encode(), C.byref(ix)))

        def search(ef, counters):
            p = lambda k: bufs[k].data_ptr() if counters else None
            chk(L.hnswb200_search_dev(ctx, ix, dq.data_ptr(), nq, K, ef, bufs["ids"].data_ptr(), bufs["d"].data_ptr(),
                                      bufs["cnt"].data_ptr(), p("h"), p("e"), p("f"), p("nb")))
        row = {"name": name, "path": path, "parity": {}}
        for label, ef, nb in cases:
            if nb is None:
                os.environ.pop("HNSWB200_FAST_NB", None)
            else:
                os.environ["HNSWB200_FAST_NB"] = nb
            for k in bufs:
                bufs[k].fill_(-7)
            search(ef, True)
            torch.cuda.synchronize()
            got = {k: bufs[k].cpu().numpy().copy() for k in bufs}
            got["d"] = got["d"].view(np.uint32)
            spill = int(((got["f"] & 4) != 0).sum())
            ovf = int(((got["f"] & 2) != 0).sum())
            if label not in ref:
                ref[label] = got
                row["parity"][label] = {"reference": True, "spill_queries": spill, "overflow_queries": ovf,
                                        "mean_evals": float(got["e"].mean())}
            else:
                r = ref[label]
                row["parity"][label] = {k: bool(np.array_equal(got[k], r[k])) for k in ("ids", "d", "cnt", "h", "e", "nb")}
                row["parity"][label]["flag_overflow_equal"] = bool(np.array_equal(got["f"] & 2, r["f"] & 2))
                row["parity"][label]["spill_queries"] = spill
                row["parity"][label]["overflow_queries"] = ovf
        os.environ.pop("HNSWB200_FAST_NB", None)
        row["variant"] = (L.hnswb200_last_search_variant() or b"").decode()
        ms = []
        for _ in range(a.repeats):
            for _ in range(a.warmup):
                search(a.ef, False)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(a.steps):
                search(a.ef, False)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / a.steps)
        row["ms_per_step"] = [round(m, 5) for m in ms]
        row["mqps_best"] = round(nq / min(ms) / 1e3, 3)
        row["mqps_median"] = round(nq / sorted(ms)[len(ms) // 2] / 1e3, 3)
        L.hnswb200_index_destroy(ix)
        L.hnswb200_ctx_destroy(ctx)
        ok = all(all(v for k, v in p.items() if isinstance(v, bool)) for p in row["parity"].values())
        row["parity_all_equal"] = ok
        print(f"{name:14s} {row['mqps_median']:8.3f} M q/s (best {row['mqps_best']:.3f})  ms {row['ms_per_step']}  parity "
              f"{'OK' if ok else 'DIFFERS'}  spill/ovf {[(p['spill_queries'], p['overflow_queries']) for p in row['parity'].values()]}",
              flush=True)
        if not ok:
            print("   ", json.dumps(row["parity"]), flush=True)
        out["runs"].append(row)
        dump()

    for n, p in zip(names, paths):
        try:
            one(n, p)
        except Exception as ex:  # a variant that fails must not take the others with it
            print(f"{n}: FAILED {ex}", flush=True)
            out["runs"].append({"name": n, "path": p, "error": str(ex)})
            dump()
            torch.cuda.synchronize()
    one(names[0] + "(again)", paths[0])  # drift of the box over the call
    return 0


if __name__ == "__main__":
    sys.exit(main())
