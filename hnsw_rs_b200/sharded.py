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
        # one flag word per rank (0xFF-filled = step -1), for signal_wait()
        self.flag_local = vp()
        check(lib().hnswb200_dev_alloc(ctx.h, 256, C.byref(self.flag_local)))
        fh = (C.c_uint8 * 64)()
        check(lib().hnswb200_ipc_export(ctx.h, self.flag_local, fh))
        fhs = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(fhs, bytes(fh), group=group)
        fptrs = []
        for r in range(self.world):
            if r == self.rank:
                fptrs.append(self.flag_local.value)
                continue
            p = vp()
            check(lib().hnswb200_ipc_open(ctx.h, (C.c_uint8 * 64).from_buffer_copy(fhs[r]), C.byref(p)))
            self._opened.append(p)
            fptrs.append(p.value)
        self.fptrs = (vp * self.world)(*fptrs)
        self.step = 0
        if self.world > 1:
            dist.barrier(group=group)

    def signal_wait(self):
        """After search(): tell every rank this rank's rows of the step are stored, and hold the stream until the rows of
        all ranks have arrived here (hnswb200_peer_signal_dev / _peer_wait_dev) -- no host synchronisation, no NCCL."""
        from ._ffi import check, lib
        self.step += 1
        check(lib().hnswb200_peer_signal_dev(self.ctx.h, self.world, self.fptrs, self.rank, self.step))
        check(lib().hnswb200_peer_wait_dev(self.ctx.h, self.flag_local, self.world, self.step))

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
        if getattr(self, "flag_local", None):
            check(lib().hnswb200_dev_free(self.ctx.h, self.flag_local))
            self.flag_local = None


class PeerExchange:
    """Device-resident all-gather + merge of per-rank top-k rows over peer memory -- the multi-GPU data path without a
    collective library and without a host round trip (the numpy path above stays for the gloo tests).

    Every rank owns two gather buffers [G][rows][k] (ids u32, distances f32; double-buffered by step parity) and one
    flag word per rank, all plain device memory that the other ranks of the box map with CUDA IPC.  One step on rank r:
      1. produce the local rows straight into slot r of the local buffer, and into slot r of every peer's buffer --
         fused into the search kernel (hnswb200_search_dev_shard / _gather: peer stores over NVLink / NVSwitch while the
         other queries compute), or with hnswb200_peer_put_dev after a brute force;
      2. hnswb200_peer_signal_dev: store the step number to flag r of every rank (stream order: after the rows);
      3. hnswb200_peer_wait_dev: hold the stream until all G flags of this rank have reached the step number;
      4. hnswb200_topk_merge_dev (K6) over the G slots.
    Everything is asynchronous on the context's stream.  A rank can run at most one step ahead of the slowest peer (its
    step s+1 merge waits for every peer's step s+1 rows, which a peer produces only after its own step s merge), so two
    buffers are enough."""

    def __init__(self, ctx, rows, k, group=None):
        import ctypes as C
        from ._ffi import check, lib, vp
        self.ctx, self.group, self.rows, self.k = ctx, group, int(rows), int(k)
        self.rank, self.world = _world(group)
        if self.world > 8:
            raise ValueError("PeerExchange: at most 8 ranks (one box)")
        self.slot_bytes = self.rows * self.k * 4
        self.buf_bytes = self.world * self.slot_bytes            # one [G][rows][k] buffer
        self.step = 0
        self._local, self._opened = [], []

        def shared(nbytes):
            """allocate locally, exchange IPC handles, map the peers; returns the G base addresses (own one included)"""
            p = vp()
            check(lib().hnswb200_dev_alloc(ctx.h, nbytes, C.byref(p)))
            self._local.append(p)
            handle = (C.c_uint8 * 64)()
            check(lib().hnswb200_ipc_export(ctx.h, p, handle))
            handles = [bytes(handle)] * self.world
            if self.world > 1:
                _dist().all_gather_object(handles, bytes(handle), group=group)
            out = []
            for r in range(self.world):
                if r == self.rank:
                    out.append(p.value)
                    continue
                q = vp()
                check(lib().hnswb200_ipc_open(ctx.h, (C.c_uint8 * 64).from_buffer_copy(handles[r]), C.byref(q)))
                self._opened.append(q)
                out.append(q.value)
            return out

        self.ids = shared(2 * self.buf_bytes)      # [2][G][rows][k] u32 on every rank
        self.dists = shared(2 * self.buf_bytes)    # [2][G][rows][k] f32
        # word r: last step rank r has completed.  dev_alloc fills with 0xFF = step -1 under the wrap-safe compare of
        # hnswb200_peer_wait_dev, so no reset is needed
        self.flags = shared(256)
        if self.world > 1:
            _dist().barrier(group=group)
        self.others = [r for r in range(self.world) if r != self.rank]

    # ---- addresses ----
    def _slot(self, bases, rank_of_buffer, parity, slot):
        return bases[rank_of_buffer] + parity * self.buf_bytes + slot * self.slot_bytes

    def local_ids(self, parity=None):
        """device address of this step's local [G][rows][k] id buffer"""
        parity = self.step & 1 if parity is None else parity
        return self.ids[self.rank] + parity * self.buf_bytes

    def local_dists(self, parity=None):
        parity = self.step & 1 if parity is None else parity
        return self.dists[self.rank] + parity * self.buf_bytes

    def _arr(self, values):
        from ._ffi import vp
        return (vp * max(1, len(values)))(*[vp(v) for v in values]) if values else (vp * 1)()

    # ---- the four stages ----
    def begin(self):
        self.step += 1
        return self.step & 1

    def shard_search(self, index, d_queries_ptr, nq, ef, id_offset):
        """stage 1 for a base-sharded HNSW (BASELINE config 5): rows of all nq queries, global ids"""
        from ._ffi import check, lib
        if nq != self.rows:
            raise ValueError("PeerExchange.shard_search: nq must equal rows")
        par = self.begin()
        peer_ids = self._arr([self.ids[r] + par * self.buf_bytes for r in self.others])
        peer_d = self._arr([self.dists[r] + par * self.buf_bytes for r in self.others])
        check(lib().hnswb200_search_dev_shard(self.ctx.h, index.h, d_queries_ptr, nq, self.k, ef, id_offset,
                                              self._slot(self.ids, self.rank, par, self.rank),
                                              self._slot(self.dists, self.rank, par, self.rank), None,
                                              len(self.others), peer_ids, peer_d, self.rank * self.rows))

    def shard_bruteforce(self, points, d_queries_ptr, nq, id_offset):
        """stage 1 for a base-sharded brute force (config 4): exact local top-k, then peer stores of the two row blocks"""
        from ._ffi import check, lib
        if nq != self.rows:
            raise ValueError("PeerExchange.shard_bruteforce: nq must equal rows")
        par = self.begin()
        li, ld = self._slot(self.ids, self.rank, par, self.rank), self._slot(self.dists, self.rank, par, self.rank)
        check(lib().hnswb200_bruteforce_topk_dev(self.ctx.h, points.h, d_queries_ptr, nq, self.k, id_offset, li, ld))
        if self.others:
            check(lib().hnswb200_peer_put_dev(self.ctx.h, li, self.slot_bytes, len(self.others),
                                              self._arr([self._slot(self.ids, r, par, self.rank) for r in self.others])))
            check(lib().hnswb200_peer_put_dev(self.ctx.h, ld, self.slot_bytes, len(self.others),
                                              self._arr([self._slot(self.dists, r, par, self.rank) for r in self.others])))

    def signal_wait(self):
        """stages 2 + 3"""
        from ._ffi import check, lib
        flags = self._arr(self.flags)
        check(lib().hnswb200_peer_signal_dev(self.ctx.h, self.world, flags, self.rank, self.step))
        check(lib().hnswb200_peer_wait_dev(self.ctx.h, self.flags[self.rank], self.world, self.step))

    def merge(self, d_out_ids_ptr, d_out_dists_ptr):
        """stage 4: top-k of the G slots under (dist, id) (graph/src/dist.rs:30-37) for every row"""
        from ._ffi import check, lib
        check(lib().hnswb200_topk_merge_dev(self.ctx.h, self.local_ids(), self.local_dists(), self.world, self.rows, self.k,
                                            d_out_ids_ptr, d_out_dists_ptr))

    def close(self):
        from ._ffi import check, lib
        self.ctx.sync()
        if self.world > 1:
            _dist().barrier(group=self.group)
        for p in self._opened:
            check(lib().hnswb200_ipc_close(self.ctx.h, p))
        for p in self._local:
            check(lib().hnswb200_dev_free(self.ctx.h, p))
        self._opened, self._local = [], []
