"""BASELINE.json configurations 3, 4 and 5 at sizes that run in well under a minute each, through the C ABI, against the
oracle.  The full-size runs are `bench.py --config c3|c4|c5` (profiles/r02_*.json)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def to_oracle(oracle, ix):
    codes, mins, deltas, levels = ix._points().download()
    p = ix.params
    layers = [ix.export_layer(l) for l in range(ix.nb_layers())]
    return oracle.Index.from_parts(p.m, p.ef_cons, p.dim, p.ep, codes, mins, deltas, levels, layers)


def recall(ids, gt):
    return sum(len(set(gt[i].tolist()) & set(ids[i].tolist())) for i in range(len(gt))) / gt.size


def test_c3_sift_shape_ids_equal_oracle_at_the_recall_ef(oracle):
    """C3 at 200,000 x 128 (L2, SIFT-shaped values, 144-byte records), 2,000 queries: the ef the sweep stops at (recall@10 >=
    0.99 against the exact ground truth) is the ef the comparison runs at -- ids, distance bits, counts, hops and evaluations
    of ALL queries equal the oracle's on the same (device-built) graph."""
    import hnsw_rs_b200 as H
    from bench_configs import sift_like
    base = sift_like(200000, 128, 1024, 3)
    queries = sift_like(2000, 128, 1024, 4)
    ix = H.HNSW.new(16, 200, 128).insert_bulk(base)
    gt, _ = H.bruteforce_topk(ix._points(), queries, 10)
    ef = None
    for e in (40, 60, 80, 100, 128, 160, 200):
        ids, dists, counts, st = ix.ann_batch(queries, 10, e, with_stats=True)
        if recall(ids, gt) >= 0.99:
            ef = e
            break
    assert ef is not None, "no ef of the sweep reaches recall 0.99"
    orc = to_oracle(oracle, ix)
    oi, od, oc, oh, oe = orc.search_batch(queries, 10, ef, threads=os.cpu_count())
    assert np.array_equal(ids, oi) and np.array_equal(bits(dists), bits(od)) and np.array_equal(counts, oc)
    assert np.array_equal(st["hops"], oh)
    assert not (st["flags"] & 2).any() and np.array_equal(st["evals"], oe)
    ogt, _ = orc.bruteforce(queries[:50], 10)
    assert np.array_equal(gt[:50], ogt)


def test_c4_bruteforce_top100_base_sharded_equals_oracle(oracle):
    """C4 at 100,000 x 100 x 400 queries, k = 100: eight base shards with global ids (id_offset), K6 merge of the [8][nq][k]
    rows == the unsharded device result == the oracle's brute force (ids and distance bits)."""
    import torch
    import hnsw_rs_b200 as H
    from hnsw_rs_b200 import _ffi, sharded
    from bench import synth
    n, nq, k, G = 100000, 400, 100, 8
    base = synth(n, 100, 512, 1)
    queries = synth(nq, 100, 512, 2)
    ctx = H.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)  # one stream for torch's fills and the library's kernels
    lib = _ffi.lib()
    full = H.SimplePoints.new(base, ctx=ctx)
    fi, fd = H.bruteforce_topk(full, queries, k, ctx=ctx)
    dq = torch.from_numpy(queries).cuda()
    gi = torch.empty((G, nq, k), dtype=torch.int32, device="cuda")
    gd = torch.empty((G, nq, k), dtype=torch.float32, device="cuda")
    shards = []
    for g in range(G):
        lo, hi = sharded.split_range(n, g, G)
        pts = H.SimplePoints.new(base[lo:hi], ctx=ctx)
        shards.append(pts)
        _ffi.check(lib.hnswb200_bruteforce_topk_dev(ctx.h, pts.h, dq.data_ptr(), nq, k, lo, gi[g].data_ptr(), gd[g].data_ptr()))
    oi = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    _ffi.check(lib.hnswb200_topk_merge_dev(ctx.h, gi.data_ptr(), gd.data_ptr(), G, nq, k, oi.data_ptr(), od.data_ptr()))
    ctx.sync()
    mi, md = oi.cpu().numpy().view(np.uint32), od.cpu().numpy()
    assert np.array_equal(mi, fi) and np.array_equal(bits(md), bits(fd))
    codes, mins, deltas = oracle.quantise_rows(base)
    flat = [(np.arange(n, dtype=np.uint32), np.zeros(n + 1, np.uint64), np.zeros(0, np.uint32))]
    orc = oracle.Index.from_parts(4, 8, 100, 0, codes, mins, deltas, np.zeros(n, np.uint8), flat)
    ei, ed = orc.bruteforce(queries[:40], k)
    assert np.array_equal(mi[:40], ei) and np.array_equal(bits(md[:40]), bits(ed))


def test_c5_deep_shape_sharded_hnsw_merge_equals_oracle_merge(oracle):
    """C5 down-scaled to 8 shards x 50,000 x 96: one HNSW per shard built on the device; hnswb200_search_dev_shard writes the
    (global id, distance) rows of every shard into slot g of one [8][nq][k] gather buffer (the peer-store path, here with
    the local buffer as the only 'peer'), K6 merges them == numpy merge of the oracle's per-shard searches on the same graphs."""
    import torch
    import hnsw_rs_b200 as H
    from hnsw_rs_b200 import _ffi
    from bench_configs import deep_like
    G, per, dim, nq, k, ef = 8, 50000, 96, 500, 10, 80
    queries = deep_like(nq, dim, 65536, 6)
    ctx = H.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)  # one stream for torch's fills and the library's kernels
    lib = _ffi.lib()
    dq = torch.from_numpy(queries).cuda()
    gi = torch.full((G, nq, k), -1, dtype=torch.int32, device="cuda")
    gd = torch.full((G, nq, k), float("inf"), dtype=torch.float32, device="cuda")
    li = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    ld = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    vp = _ffi.vp
    exp = []
    for g in range(G):
        ix = H.HNSW.new(16, 100, dim, ctx=ctx).insert_bulk(deep_like(per, dim, 65536, 5, row0=g * per))
        pid, pdd = (vp * 1)(vp(gi.data_ptr())), (vp * 1)(vp(gd.data_ptr()))
        _ffi.check(lib.hnswb200_search_dev_shard(ctx.h, ix.h, dq.data_ptr(), nq, k, ef, g * per, li.data_ptr(), ld.data_ptr(), None,
                                                 1, pid, pdd, g * nq))
        ctx.sync()
        assert torch.equal(gi[g], li) and torch.equal(gd[g].view(torch.int32), ld.view(torch.int32))  # peer rows == local rows
        orc = to_oracle(oracle, ix)
        oi, od = orc.search_batch(queries, k, ef, threads=os.cpu_count())[:2]
        exp.append((np.where(oi != 0xFFFFFFFF, oi + np.uint32(g * per), oi), od))
    oi = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    _ffi.check(lib.hnswb200_topk_merge_dev(ctx.h, gi.data_ptr(), gd.data_ptr(), G, nq, k, oi.data_ptr(), od.data_ptr()))
    ctx.sync()
    key = lambda i, d: (d.view(np.uint32).astype(np.uint64) << np.uint64(32)) | i.astype(np.uint64)
    allk = np.concatenate([key(i, d) for i, d in exp], axis=1)
    allk.sort(axis=1)
    got = key(oi.cpu().numpy().view(np.uint32), od.cpu().numpy())
    assert np.array_equal(got, allk[:, :k])
