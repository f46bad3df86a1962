"""Multi-GPU paths on real devices: the CUDA engine as the local search, NCCL for the one all_gather,
the K6 kernel for the merge.  The 2-rank tests need two GPUs and skip otherwise; the world-1 tests
exercise the same code on one device by searching the shards one after the other."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def to_gpu(H, orc, ctx=None):
    codes, mins, deltas, levels = orc.export_points()
    p = orc.params()
    prm = H.Params(p["ep"], p["m"], p["mmax"], p["mmax0"], p["ml"], p["ef_cons"], p["dim"])
    caps = [orc.layer_cap(l) for l in range(orc.nb_layers)]
    return H.HNSW.from_parts(prm, codes, mins, deltas, levels, orc.export_layers(), caps, ctx=ctx)


def np_merge(ids, dists):
    G, nq, k = ids.shape
    oi = np.full((nq, k), 0xFFFFFFFF, np.uint32)
    od = np.full((nq, k), np.inf, np.float32)
    for q in range(nq):
        i, d = ids[:, q, :].reshape(-1), dists[:, q, :].reshape(-1)
        keep = i != 0xFFFFFFFF
        i, d = i[keep], d[keep]
        order = np.lexsort((i, d.view(np.uint32)))[:k]
        oi[q, :len(order)], od[q, :len(order)] = i[order], d[order]
    return oi, od


def test_base_sharded_search_one_device(oracle, glove):
    """Two shard indexes on one GPU, searched in turn, merged by the K6 kernel == numpy merge of the oracle's
    per-shard searches == (for brute force) the unsharded oracle."""
    import hnsw_rs_b200 as H
    from hnsw_rs_b200 import sharded
    store, queries = glove
    ids_g, d_g, oi_g, od_g = [], [], [], []
    for r in range(2):
        lo, hi = sharded.split_range(len(store), r, 2)
        orc = oracle.Index(12, None, store.shape[1]).insert_bulk(store[lo:hi])
        ix = to_gpu(H, orc)
        s = sharded.BaseShardedSearch(ix, lo)
        i, d = ix.ann_batch(queries, 10, 50)[:2]
        ids_g.append(s._globalise(i)); d_g.append(d)
        oi, od = orc.search_batch(queries, 10, 50)[:2]
        oi_g.append(np.where(oi != 0xFFFFFFFF, oi + np.uint32(lo), oi)); od_g.append(od)
    mi, md = H.topk_merge(np.stack(ids_g), np.stack(d_g))
    ei, ed = np_merge(np.stack(oi_g), np.stack(od_g))
    assert np.array_equal(mi, ei) and np.array_equal(bits(md), bits(ed))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import hnsw_rs_b200 as H
    from hnsw_rs_b200 import sharded
    from oracle import pyoracle as O

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    ctx = H.Context(rank)
    store = O.load_glove(os.path.join(GOLDEN, "store.txt"))
    queries = O.load_glove(os.path.join(GOLDEN, "queries.txt"))[:37]
    # query-sharded, replicated index
    orc = O.Index(12, None, store.shape[1]).insert_bulk(store)
    ix = to_gpu(H, orc, ctx)
    ids, dists, counts = sharded.QueryShardedSearch(ix, device=dev).search(queries, 10, 40)
    ref = orc.search_batch(queries, 10, 40)
    ok = np.array_equal(ids, ref[0]) and np.array_equal(bits(dists), bits(ref[1])) and np.array_equal(counts, ref[2])
    # base-sharded HNSW and brute force
    lo, hi = sharded.split_range(len(store), rank, world)
    so = O.Index(12, None, store.shape[1]).insert_bulk(store[lo:hi])
    six = to_gpu(H, so, ctx)
    s = sharded.BaseShardedSearch(six, lo, device=dev)
    mi, md = s.search(queries, 10, 60)
    parts = []
    for r in range(world):
        a, b = sharded.split_range(len(store), r, world)
        sh = O.Index(12, None, store.shape[1]).insert_bulk(store[a:b])
        i, d = sh.search_batch(queries, 10, 60)[:2]
        parts.append((np.where(i != 0xFFFFFFFF, i + np.uint32(a), i), d))
    ei, ed = np_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    ok = ok and np.array_equal(mi, ei) and np.array_equal(bits(md), bits(ed))
    bi, bd = s.bruteforce(queries, 10)
    fi, fd = orc.bruteforce(queries, 10)
    ok = ok and np.array_equal(bi, fi) and np.array_equal(bits(bd), bits(fd))
    # fused all-gather: every rank stores its id rows straight into every peer's buffer (CUDA IPC over NVLink)
    qlo, qhi = sharded.split_range(len(queries), rank, world)
    per = -(-len(queries) // world)
    pg = sharded.PeerGather(ctx, per, 10)
    dq = torch.from_numpy(queries[qlo:qhi].copy()).to(dev)
    d_ids = torch.empty((qhi - qlo, 10), dtype=torch.int32, device=dev)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):  # back-to-back launches overlap; the buffers end up the same
        pg.search(ix, dq.data_ptr(), qhi - qlo, 40, d_ids.data_ptr())
    torch.cuda.synchronize()
    dist.barrier()
    got = pg.download()
    for r in range(world):
        a, b = sharded.split_range(len(queries), r, world)
        ok = ok and np.array_equal(got[r * per:r * per + (b - a)], ref[0][a:b])
    ok = ok and np.array_equal(d_ids.cpu().numpy().view(np.uint32), ref[0][qlo:qhi])
    dist.barrier()
    pg.close()
    open(os.path.join(out_dir, f"ok_{rank}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_two_gpus_nccl(oracle, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"ok_{r}").read() == "1"


# ---- device-resident exchange: peer stores + flag words + K6, no collective library in the data path -------------
def _px_worker(rank, world, port, out_dir, ndev):
    """Two ranks on two GPUs, gloo for the bootstrap only (no NCCL anywhere).  Never two ranks on ONE GPU: a kernel that
    spins on a flag another process must raise is not guaranteed to run at the same time as that process's kernels
    (B200_PROFILING.md: Xid 109)."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import hnsw_rs_b200 as H
    from hnsw_rs_b200 import sharded
    from oracle import pyoracle as O

    dev_i = rank % ndev
    torch.cuda.set_device(dev_i)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dev = torch.device("cuda", dev_i)
    ctx = H.Context(dev_i)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    store = O.load_glove(os.path.join(GOLDEN, "store.txt"))
    queries = O.load_glove(os.path.join(GOLDEN, "queries.txt"))[:64]
    nq, k = len(queries), 10
    dq = torch.from_numpy(queries.copy()).to(dev)
    # expected: numpy merge of the oracle's per-shard searches, and the unsharded oracle for the brute force
    parts, shards = [], []
    for r in range(world):
        a, b = sharded.split_range(len(store), r, world)
        sh = O.Index(12, None, store.shape[1]).insert_bulk(store[a:b])
        shards.append((a, sh))
        i, d = sh.search_batch(queries, k, 60)[:2]
        parts.append((np.where(i != 0xFFFFFFFF, i + np.uint32(a), i), d))
    ei, ed = np_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    full = O.Index(12, None, store.shape[1]).insert_bulk(store)
    fi, fd = full.bruteforce(queries, k)
    lo, so = shards[rank]
    six = to_gpu(H, so, ctx)
    px = sharded.PeerExchange(ctx, nq, k)
    out_i = torch.empty((nq, k), dtype=torch.int32, device=dev)
    out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    ok = True
    for step in range(5):  # several steps back to back: the double buffering and the step flags hold
        px.shard_search(six, dq.data_ptr(), nq, 60, lo)
        px.signal_wait()
        px.merge(out_i.data_ptr(), out_d.data_ptr())
        if step % 2 == 0:
            ctx.sync()
            ok = ok and np.array_equal(out_i.cpu().numpy().view(np.uint32), ei) and np.array_equal(bits(out_d.cpu().numpy()), bits(ed))
    ctx.sync()
    ok = ok and np.array_equal(out_i.cpu().numpy().view(np.uint32), ei) and np.array_equal(bits(out_d.cpu().numpy()), bits(ed))
    px.shard_bruteforce(six._points(), dq.data_ptr(), nq, lo)
    px.signal_wait()
    px.merge(out_i.data_ptr(), out_d.data_ptr())
    ctx.sync()
    ok = ok and np.array_equal(out_i.cpu().numpy().view(np.uint32), fi) and np.array_equal(bits(out_d.cpu().numpy()), bits(fd))
    px.close()
    # query-sharded: fused id gather + flags (PeerGather.signal_wait), no barrier before the read
    orc = O.Index(12, None, store.shape[1]).insert_bulk(store)
    ix = to_gpu(H, orc, ctx)
    ref = orc.search_batch(queries, k, 40)
    qlo, qhi = sharded.split_range(nq, rank, world)
    per = -(-nq // world)
    pg = sharded.PeerGather(ctx, per, k)
    d_ids = torch.empty((qhi - qlo, k), dtype=torch.int32, device=dev)
    for _ in range(3):
        pg.search(ix, dq[qlo:qhi].contiguous().data_ptr(), qhi - qlo, 40, d_ids.data_ptr())
        pg.signal_wait()
    ctx.sync()
    got = pg.download()
    for r in range(world):
        a, b = sharded.split_range(nq, r, world)
        ok = ok and np.array_equal(got[r * per:r * per + (b - a)], ref[0][a:b])
    dist.barrier()
    pg.close()
    open(os.path.join(out_dir, f"px_ok_{rank}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_peer_exchange_two_ranks(oracle, tmp_path):
    """PeerExchange / PeerGather.signal_wait across two processes: base-sharded HNSW (C5 shape) == numpy merge of the
    oracle's per-shard searches, base-sharded brute force (C4 shape) == the unsharded oracle, query-sharded gather (C3
    shape) == the oracle; everything device-resident, synchronised by flag words in peer memory."""
    import torch
    import torch.multiprocessing as mp
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs (ranks that wait for each other's flags must not share one GPU)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_px_worker, args=(2, port, str(tmp_path), ndev), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"px_ok_{r}").read() == "1"


def test_peer_exchange_single_rank(oracle, glove):
    """world = 1: the same four stages on one rank (no peers): shard search with an id offset + merge == the oracle's
    search with the offset added."""
    import torch
    import hnsw_rs_b200 as H
    from hnsw_rs_b200 import sharded
    store, queries = glove
    orc = oracle.Index(12, None, store.shape[1]).insert_bulk(store)
    ctx = H.Context(0)
    ix = to_gpu(H, orc, ctx)
    nq, k = len(queries), 10
    dq = torch.from_numpy(queries.copy()).cuda()
    out_i = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    out_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    px = sharded.PeerExchange(ctx, nq, k)
    oi, od = orc.search_batch(queries, k, 50)[:2]
    for _ in range(3):
        px.shard_search(ix, dq.data_ptr(), nq, 50, 1000)
        px.signal_wait()
        px.merge(out_i.data_ptr(), out_d.data_ptr())
    ctx.sync()
    assert np.array_equal(out_i.cpu().numpy().view(np.uint32), oi + np.uint32(1000))
    assert np.array_equal(bits(out_d.cpu().numpy()), bits(od))
    px.close()
