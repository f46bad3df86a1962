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
