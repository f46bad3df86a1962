"""hnsw/src/helpers/glove.rs: GloVe text loader and brute-force ground truth."""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import Context, check, f32, lib, ptr
from .points import SimplePoints


def load_glove_array(lim, path, verbose=False):
    """helpers/glove.rs:14-71.  Returns (words, embeddings[n, dim] f32); values are parsed
    straight to f32 (correctly rounded) by the library, like Rust's parse::<f32>()."""
    p = str(path).encode()
    dim = C.c_uint64()
    rows = lib().hnswb200_load_glove(p, lim, None, 0, C.byref(dim))
    if rows < 0:
        check(int(rows))
    out = np.zeros((rows, dim.value), np.float32)
    lib().hnswb200_load_glove(p, lim, ptr(out, _ffi.f32p), out.size, C.byref(dim))
    words = []
    with open(path, "r", encoding="utf-8", errors="replace") as f:
        for i, line in enumerate(f):
            if lim and i >= lim:
                break
            words.append(line.split(" ", 1)[0])
    return words, out


def bruteforce_topk(base, queries, k, id_offset=0, ctx=None):
    """Exact top-k of f32 queries against device-resident SimplePoints under the quantised
    metric with (dist, id) ties: ids[q,k], dists[q,k]."""
    ctx = ctx or base.ctx
    q = f32(queries)
    nq = q.shape[0]
    ids = np.zeros((nq, k), np.uint32)
    dists = np.zeros((nq, k), np.float32)
    check(lib().hnswb200_bruteforce_topk(ctx.h, base.h, ptr(q, _ffi.f32p), nq, k, id_offset,
                                         ptr(ids, _ffi.u32p), ptr(dists, _ffi.f32p)))
    return ids, dists


def brute_force_nns(nb_nns, train_set, test_vectors, ids, bar=None):
    """helpers/glove.rs:73-92: {query id -> ids of its nb_nns nearest train points}.
    train_set: SimplePoints; test_vectors: f32 rows indexed by `ids`."""
    tv = f32(test_vectors)
    sel = np.asarray(list(ids), dtype=np.int64)
    nn, _ = bruteforce_topk(train_set, tv[sel], nb_nns)
    return {int(i): [int(x) for x in nn[j]] for j, i in enumerate(sel)}


def topk_merge(ids, dists, ctx=None):
    """Merge per-shard sorted top-k lists ids/dists[G, q, k] into [q, k] under (dist, id)."""
    ctx = ctx or Context.default()
    ids = np.ascontiguousarray(ids, np.uint32)
    dists = f32(dists)
    G, nq, k = ids.shape
    oi = np.zeros((nq, k), np.uint32)
    od = np.zeros((nq, k), np.float32)
    check(lib().hnswb200_topk_merge(ctx.h, ptr(ids, _ffi.u32p), ptr(dists, _ffi.f32p), G, nq, k,
                                    ptr(oi, _ffi.u32p), ptr(od, _ffi.f32p)))
    return oi, od
