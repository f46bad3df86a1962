"""Host-side logic of the multi-GPU paths (hnsw_rs_b200/sharded.py) with world_size = 2 over gloo on CPU.
The local engine is the oracle (test infrastructure) and the merge is a numpy restatement of the
(dist, id) order; what is under test is the partitioning, padding, gather layout and id offsets."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def np_merge(ids, dists):
    G, nq, k = ids.shape
    oi = np.full((nq, k), 0xFFFFFFFF, np.uint32)
    od = np.full((nq, k), np.inf, np.float32)
    for q in range(nq):
        i = ids[:, q, :].reshape(-1)
        d = dists[:, q, :].reshape(-1)
        keep = i != 0xFFFFFFFF
        i, d = i[keep], d[keep]
        order = np.lexsort((i, d.view(np.uint32)))[:k]
        oi[q, :len(order)], od[q, :len(order)] = i[order], d[order]
    return oi, od


def _worker(rank, world, port, mode, out_dir, full=False):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import pyoracle as O
    from hnsw_rs_b200 import sharded

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    store = O.load_glove(os.path.join(GOLDEN, "store.txt"))
    queries = O.load_glove(os.path.join(GOLDEN, "queries.txt"))[:37]  # 37: not divisible by 2 -> padding path
    if mode == "query":
        ix = O.Index(12, None, store.shape[1], full=full).insert_bulk(store)
        s = sharded.QueryShardedSearch(None, local_search=lambda q, n, ef: ix.search_batch(q, n, ef)[:3])
        ids, dists, counts = s.search(queries, 10, 40)
        ref = ix.search_batch(queries, 10, 40)
        ok = np.array_equal(ids, ref[0]) and np.array_equal(dists.view(np.uint32), ref[1].view(np.uint32)) \
            and np.array_equal(counts, ref[2])
        # ef < n: short lists keep their padding through the gather
        ids, dists, counts = s.search(queries, 10, 4)
        ref = ix.search_batch(queries, 10, 4)
        ok = ok and np.array_equal(ids, ref[0]) and np.array_equal(counts, ref[2]) and (counts == 4).all()
    else:
        lo, hi = sharded.split_range(len(store), rank, world)
        shard = O.Index(12, None, store.shape[1], full=full).insert_bulk(store[lo:hi])
        s = sharded.BaseShardedSearch(None, lo, local_search=lambda q, n, ef: shard.search_batch(q, n, ef)[:2],
                                      merge=np_merge)
        ids, dists = s.search(queries, 10, 60)
        # expected: the merge of the per-shard oracle searches (the reference has no sharded mode)
        parts = []
        for r in range(world):
            a, b = sharded.split_range(len(store), r, world)
            sh = O.Index(12, None, store.shape[1], full=full).insert_bulk(store[a:b])
            i, d = sh.search_batch(queries, 10, 60)[:2]
            parts.append((np.where(i != 0xFFFFFFFF, i + np.uint32(a), i), d))
        ei, ed = np_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
        ok = np.array_equal(ids, ei) and np.array_equal(dists.view(np.uint32), ed.view(np.uint32))
        # base-sharded exact brute force == unsharded brute force
        whole = O.Index(12, None, store.shape[1], full=full).insert_bulk(store)
        bi, bd = s.bruteforce(queries, 10, local_bruteforce=lambda q, k: shard.bruteforce(q, k))
        fi, fd = whole.bruteforce(queries, 10)
        ok = ok and np.array_equal(bi, fi) and np.array_equal(bd.view(np.uint32), fd.view(np.uint32))
    open(os.path.join(out_dir, f"ok_{mode}_{rank}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("full", [False, True])  # QuantVec (the reference as committed) and FullVec shards
@pytest.mark.parametrize("mode", ["query", "base"])
def test_sharded_world2_gloo(oracle, tmp_path, mode, full):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, mode, str(tmp_path), full), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"ok_{mode}_{r}").read() == "1"


def test_split_range_partitions():
    from hnsw_rs_b200 import sharded
    for n in (0, 1, 7, 10000, 1183514):
        for w in (1, 2, 3, 8):
            cuts = [sharded.split_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
