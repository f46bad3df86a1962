"""Multi-GPU search: one process per GPU, torch.distributed for the plumbing.

The path shards without a data-path collective -- queries are independent and top-k lists are
mergeable -- so the only exchange is ONE all_gather of fixed-size per-rank result buffers:

  query-sharded, replicated index (BASELINE config 3): rank r searches queries [lo_r, hi_r);
      the gathered blocks are the answer (no merge).
  base-sharded (configs 4, 5): every rank searches ALL queries on its own shard, adds its id
      offset (global ids), and the G gathered lists of n per query are merged under the
      reference's (dist, id) order (graph/src/dist.rs:30-37) by the K6 kernel
      (hnswb200_topk_merge).  Each global top-n member is in its shard's local top-n, so the
      merged result equals the unsharded one.

The reference has no sharded mode; its per-shard semantics are unchanged (HNSW::ann_by_vector,
hnsw/src/template.rs:306-335; brute_force_nns, hnsw/src/helpers/glove.rs:73-109).

`local_search` / `merge` are injectable so that the host-side logic (partitioning, padding,
gather layout, id offsets) can be exercised by world_size-2 gloo tests on CPU with the oracle as the
local engine; the defaults are the CUDA engine and there is no CPU fallback.
"""
import numpy as np

from . import _ffi
from ._ffi import NO_ID


def split_range(n, rank, world):
    """Contiguous partition of range(n): the first n % world ranks get one extra unit."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist


def _world(group):
    dist = _dist()
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def all_gather_rows(arr, group=None, device=None):
    """all_gather of equally shaped numpy blocks -> array [world, *arr.shape] on every rank.
    With NCCL the blocks travel as CUDA tensors (device = this rank's GPU); with gloo as CPU tensors."""
    import torch
    dist = _dist()
    rank, world = _world(group)
    if world == 1:
        return arr[None].copy()
    view = arr.view(np.int32) if arr.dtype == np.uint32 else arr
    t = torch.from_numpy(np.ascontiguousarray(view))
    if device is not None:
        t = t.to(device)
    if t.dim() == 0:
        t = t.reshape(1)
    # concatenation along dim 0 is the form both NCCL and gloo accept
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    res = out.cpu().numpy().reshape((world,) + tuple(t.shape))
    return res.view(np.uint32) if arr.dtype == np.uint32 else res


class QueryShardedSearch:
    """Replicated index, queries split across the ranks; every rank ends up with all results."""

    def __init__(self, index, group=None, device=None, local_search=None):
        self.index, self.group, self.device = index, group, device
        self.local_search = local_search or (lambda q, n, ef: index.ann_batch(q, n, ef))

    def search(self, queries, n, ef):
        queries = np.ascontiguousarray(queries, np.float32)
        nq = queries.shape[0]
        rank, world = _world(self.group)
        lo, hi = split_range(nq, rank, world)
        per = -(-nq // world)  # every rank sends the same block size; the tail is padding
        ids = np.full((per, n), NO_ID, np.uint32)
        dists = np.full((per, n), np.inf, np.float32)
        counts = np.zeros(per, np.uint32)
        if hi > lo:
            i, d, c = self.local_search(queries[lo:hi], n, ef)
            ids[:hi - lo], dists[:hi - lo], counts[:hi - lo] = i, d, c
        gi = all_gather_rows(ids, self.group, self.device)
        gd = all_gather_rows(dists, self.group, self.device)
        gc = all_gather_rows(counts, self.group, self.device)
        oi = np.empty((nq, n), np.uint32)
        od = np.empty((nq, n), np.float32)
        oc = np.empty(nq, np.uint32)
        for r in range(world):
            a, b = split_range(nq, r, world)
            oi[a:b], od[a:b], oc[a:b] = gi[r, :b - a], gd[r, :b - a], gc[r, :b - a]
        return oi, od, oc


def _merge_cuda(ids, dists, ctx=None):
    from .helpers import topk_merge
    return topk_merge(ids, dists, ctx)


class BaseShardedSearch:
    """Every rank holds the index (or the points) of one base shard whose local id i is the global
    id i + id_offset; all ranks see all queries."""

    def __init__(self, shard, id_offset, group=None, device=None, local_search=None, merge=None):
        self.shard, self.id_offset, self.group, self.device = shard, int(id_offset), group, device
        self.local_search = local_search or (lambda q, n, ef: shard.ann_batch(q, n, ef)[:2])
        self.merge = merge or (lambda i, d: _merge_cuda(i, d, getattr(shard, "ctx", None)))

    def _globalise(self, ids):
        out = ids.astype(np.uint32, copy=True)
        real = out != NO_ID
        out[real] += np.uint32(self.id_offset)
        return out

    def search(self, queries, n, ef):
        """HNSW search of every shard + merge: ids[q, n], dists[q, n] with global ids."""
        queries = np.ascontiguousarray(queries, np.float32)
        ids, dists = self.local_search(queries, n, ef)
        gi = all_gather_rows(self._globalise(ids), self.group, self.device)
        gd = all_gather_rows(np.ascontiguousarray(dists, np.float32), self.group, self.device)
        return self.merge(gi, gd)

    def bruteforce(self, queries, k, local_bruteforce=None):
        """Exact top-k over all shards (base-sharded brute_force_nns)."""
        from .helpers import bruteforce_topk
        queries = np.ascontiguousarray(queries, np.float32)
        if local_bruteforce is None:
            pts = self.shard._points() if hasattr(self.shard, "_points") else self.shard
            ids, dists = bruteforce_topk(pts, queries, k, self.id_offset)
        else:
            ids, dists = local_bruteforce(queries, k)
            ids = self._globalise(ids)
        gi = all_gather_rows(np.ascontiguousarray(ids, np.uint32), self.group, self.device)
        gd = all_gather_rows(np.ascontiguousarray(dists, np.float32), self.group, self.device)
        return self.merge(gi, gd)


class PeerGather:
    """Fused all-gather for the query-sharded path: every rank owns one device buffer of world * rows_per_rank result
    rows and maps the buffers of all the other ranks of the box (CUDA IPC, peer access over NVLink / NVSwitch).
    hnswb200_search_dev_gather then stores the id row of local query q to row rank * rows_per_rank + q of EVERY buffer
    while the other queries keep computing: no collective kernel, no extra launch.  After the ranks have synchronised
    (stream sync + barrier) every buffer holds all results."""

    def __init__(self, ctx, rows_per_rank, n, group=None):
        import ctypes as C
        from ._ffi import check, lib, vp
        dist = _dist()
        self.ctx, self.group, self.rows, self.n = ctx, group, int(rows_per_rank), int(n)
        self.rank, self.world = _world(group)
        if self.world > 8:
            raise ValueError("PeerGather: at most 8 ranks (one box)")
        self.bytes = self.world * self.rows * self.n * 4
        self.local = vp()
        check(lib().hnswb200_dev_alloc(ctx.h, self.bytes, C.byref(self.local)))
        handle = (C.c_uint8 * 64)()
        check(lib().hnswb200_ipc_export(ctx.h, self.local, handle))
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(handle), group=group)
        self._opened = []
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(self.local.value)
                continue
            p = vp()
            check(lib().hnswb200_ipc_open(ctx.h, (C.c_uint8 * 64).from_buffer_copy(handles[r]), C.byref(p)))
            self._opened.append(p)
            ptrs.append(p.value)
        self.ptrs = (vp * self.world)(*ptrs)

    def search(self, index, d_queries_ptr, nq, ef, d_ids_ptr, d_dists_ptr=None, d_counts_ptr=None):
        """Asynchronous on the context's stream: search nq (<= rows_per_rank) device-resident queries, results to the
        local buffers given AND to this rank's rows of every peer buffer."""
        from ._ffi import check, lib
        if nq > self.rows:
            raise ValueError("PeerGather.search: more queries than rows per rank")
        check(lib().hnswb200_search_dev_gather(self.ctx.h, index.h, d_queries_ptr, nq, self.n, ef, d_ids_ptr, d_dists_ptr,
                                               d_counts_ptr, self.world, self.ptrs, self.rank * self.rows))

    def download(self):
        """ids[world * rows_per_rank, n] as this rank sees them (call after stream sync + barrier)."""
        import ctypes as C
        from ._ffi import check, lib, u32p
        out = np.empty((self.world * self.rows, self.n), np.uint32)
        check(lib().hnswb200_dev_download(self.ctx.h, self.local, out.ctypes.data_as(C.c_void_p), self.bytes))
        return out

    def close(self):
        """Unmap the peers and free the local buffer; every rank must be past its last use (barrier first)."""
        from ._ffi import check, lib
        for p in self._opened:
            check(lib().hnswb200_ipc_close(self.ctx.h, p))
        self._opened = []
        if self.local:
            check(lib().hnswb200_dev_free(self.ctx.h, self.local))
            self.local = None
