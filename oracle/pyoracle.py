"""ctypes binding of the CPU oracle (oracle/libhnsw_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the
product package hnsw_rs_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhnsw_oracle.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)


def build(force=False):
    src = os.path.join(_HERE, "hnsw_oracle.cpp")
    if force or not os.path.exists(_SO) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_dist_quant.restype = C.c_float
        L.oracle_dist_quant.argtypes = [u8p, C.c_float, C.c_float, u8p, C.c_float, C.c_float, C.c_uint64]
        L.oracle_dist_quant_generic.restype = C.c_float
        L.oracle_dist_quant_generic.argtypes = L.oracle_dist_quant.argtypes
        L.oracle_dist_full.restype = C.c_float
        L.oracle_dist_full.argtypes = [f32p, f32p, C.c_uint64]
        L.oracle_quantise.argtypes = [f32p, C.c_uint64, u8p, f32p, f32p]
        L.oracle_dequantise.argtypes = [u8p, C.c_float, C.c_float, C.c_uint64, f32p]
        L.oracle_dist_cmp.argtypes = [C.c_uint32, C.c_float, C.c_uint32, C.c_float]
        L.oracle_levels.argtypes = [C.c_uint64, C.c_uint64, u8p]
        L.oracle_chacha_block.argtypes = [u32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, u32p]
        L.oracle_index_new.restype = C.c_void_p
        L.oracle_index_new.argtypes = [C.c_uint64, C.c_int64, C.c_uint64]
        L.oracle_index_free.argtypes = [C.c_void_p]
        L.oracle_set_vec_type.argtypes = [C.c_void_p, C.c_int]
        L.oracle_vec_type.argtypes = [C.c_void_p]
        L.oracle_export_values.argtypes = [C.c_void_p, f32p]
        L.oracle_insert_bulk.argtypes = [C.c_void_p, f32p, C.c_uint64, C.c_uint64, u8p, u64p]
        L.oracle_insert_vec.restype = C.c_int64
        L.oracle_insert_vec.argtypes = [C.c_void_p, f32p, C.c_uint64]
        for name in ("oracle_len", "oracle_nb_layers", "oracle_dim"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p]
        L.oracle_ep.restype = C.c_uint32
        L.oracle_ep.argtypes = [C.c_void_p]
        L.oracle_set_ep.argtypes = [C.c_void_p, C.c_uint32]
        L.oracle_params.argtypes = [C.c_void_p, u64p, f32p]
        L.oracle_distance.restype = C.c_float
        L.oracle_distance.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.oracle_export_points.argtypes = [C.c_void_p, u8p, f32p, f32p, u8p]
        for name in ("oracle_layer_nb_nodes", "oracle_layer_nb_edges", "oracle_layer_cap"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_export_layer.argtypes = [C.c_void_p, C.c_uint64, u32p, u64p, u32p]
        L.oracle_index_from_parts.restype = C.c_void_p
        L.oracle_index_from_parts.argtypes = [
            C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, u8p, f32p, f32p, u8p,
            C.c_uint64, u64p, C.POINTER(u32p), C.POINTER(u64p), C.POINTER(u32p)]
        L.oracle_save.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_load.restype = C.c_void_p
        L.oracle_load.argtypes = [C.c_char_p]
        L.oracle_search_batch.argtypes = [C.c_void_p, f32p, C.c_uint64, C.c_uint64, C.c_uint64,
                                          C.c_uint32, u32p, f32p, u32p, u32p, u32p]
        L.oracle_bruteforce.argtypes = [C.c_void_p, f32p, C.c_uint64, C.c_uint64, C.c_uint32, u32p, f32p]
        L.oracle_dist_query_many.argtypes = [C.c_void_p, f32p, u32p, C.c_uint64, f32p]
        L.oracle_layer_degree_range.argtypes = [C.c_void_p, C.c_uint64, u64p, u64p]
        L.oracle_graph_new.restype = C.c_void_p
        L.oracle_graph_new.argtypes = [C.c_uint64, C.c_uint64]
        L.oracle_graph_free.argtypes = [C.c_void_p]
        L.oracle_graph_add_node.argtypes = [C.c_void_p, C.c_uint32]
        L.oracle_graph_add_edge.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.oracle_graph_remove_edge.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.oracle_graph_contains.argtypes = [C.c_void_p, C.c_uint32]
        L.oracle_graph_degree.restype = C.c_int64
        L.oracle_graph_degree.argtypes = [C.c_void_p, C.c_uint32]
        L.oracle_graph_neighbors.restype = C.c_int64
        L.oracle_graph_neighbors.argtypes = [C.c_void_p, C.c_uint32, u32p, C.c_uint64]
        L.oracle_graph_replace_neighbors.argtypes = [C.c_void_p, C.c_uint32, u32p, C.c_uint64]
        L.oracle_graph_nb_nodes.restype = C.c_uint64
        L.oracle_graph_nb_nodes.argtypes = [C.c_void_p]
        L.oracle_load_glove.restype = C.c_int64
        L.oracle_load_glove.argtypes = [C.c_char_p, C.c_uint64, f32p, C.c_uint64, u64p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def last_error():
    return lib().oracle_last_error().decode()


class OracleError(RuntimeError):
    pass


# ---- vectors ---------------------------------------------------------------
def quantise(v):
    v = np.ascontiguousarray(v, dtype=np.float32)
    codes = np.zeros(v.shape[0], dtype=np.uint8)
    mn, dl = C.c_float(), C.c_float()
    if lib().oracle_quantise(_p(v, f32p), v.shape[0], _p(codes, u8p), C.byref(mn), C.byref(dl)):
        raise OracleError(last_error())
    return codes, np.float32(mn.value), np.float32(dl.value)


def quantise_rows(rows):
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n, d = rows.shape
    codes = np.zeros((n, d), np.uint8)
    mins = np.zeros(n, np.float32)
    deltas = np.zeros(n, np.float32)
    for i in range(n):
        codes[i], mins[i], deltas[i] = quantise(rows[i])
    return codes, mins, deltas


def dequantise(codes, mn, delta):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    out = np.zeros(codes.shape[0], np.float32)
    lib().oracle_dequantise(_p(codes, u8p), C.c_float(delta), C.c_float(mn), codes.shape[0], _p(out, f32p))
    return out


def dist_quant(a, b, generic=False):
    """a, b = (codes, min, delta) triples; distance_unrolled (or the generic zip sum)."""
    ac = np.ascontiguousarray(a[0], np.uint8)
    bc = np.ascontiguousarray(b[0], np.uint8)
    fn = lib().oracle_dist_quant_generic if generic else lib().oracle_dist_quant
    d = min(ac.shape[0], bc.shape[0])
    return np.float32(fn(_p(ac, u8p), C.c_float(a[2]), C.c_float(a[1]), _p(bc, u8p), C.c_float(b[2]),
                         C.c_float(b[1]), d))


def dist_full(x, y):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    return np.float32(lib().oracle_dist_full(_p(x, f32p), _p(y, f32p), min(x.shape[0], y.shape[0])))


def dist_cmp(ida, da, idb, db):
    return lib().oracle_dist_cmp(ida, C.c_float(da), idb, C.c_float(db))


def levels(m, n):
    out = np.zeros(n, np.uint8)
    lib().oracle_levels(m, n, _p(out, u8p))
    return out


def chacha_block(key_words, counter, rounds):
    key = np.ascontiguousarray(key_words, np.uint32)
    out = np.zeros(16, np.uint32)
    lib().oracle_chacha_block(_p(key, u32p), counter, 0, 0, rounds, _p(out, u32p))
    return out


def load_glove(path, lim=0):
    dim = C.c_uint64()
    rows = lib().oracle_load_glove(path.encode(), lim, None, 0, C.byref(dim))
    if rows < 0:
        raise OracleError(last_error())
    out = np.zeros((rows, dim.value), np.float32)
    lib().oracle_load_glove(path.encode(), lim, _p(out, f32p), out.size, C.byref(dim))
    return out


# ---- graph (bare) ----------------------------------------------------------
class Graph:
    def __init__(self, level, m):
        self.h = C.c_void_p(lib().oracle_graph_new(level, m))

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_graph_free(self.h)
            self.h = None

    def add_node(self, i): lib().oracle_graph_add_node(self.h, i)
    def add_edge(self, a, b): return lib().oracle_graph_add_edge(self.h, a, b)
    def remove_edge(self, a, b): return lib().oracle_graph_remove_edge(self.h, a, b)
    def contains(self, i): return bool(lib().oracle_graph_contains(self.h, i))
    def degree(self, i): return lib().oracle_graph_degree(self.h, i)
    def nb_nodes(self): return lib().oracle_graph_nb_nodes(self.h)

    def neighbors(self, i):
        buf = np.zeros(4096, np.uint32)
        n = lib().oracle_graph_neighbors(self.h, i, _p(buf, u32p), buf.size)
        if n < 0:
            return None
        return set(int(x) for x in buf[:n])

    def replace_neighbors(self, i, nn):
        a = np.ascontiguousarray(list(nn), np.uint32)
        return lib().oracle_graph_replace_neighbors(self.h, i, _p(a, u32p), a.size)


# ---- index -----------------------------------------------------------------
class Index:
    """Mirror of hnsw::template::HNSW (hnsw/src/template.rs) on the CPU oracle."""

    def __init__(self, m=12, ef_cons=None, dim=0, _handle=None, full=False):
        """full=True: the index the reference builds with `type VecType = FullVec;` (points/src/point.rs:4)."""
        if _handle is not None:
            self.h = C.c_void_p(_handle)
        else:
            self.h = C.c_void_p(lib().oracle_index_new(m, -1 if ef_cons is None else ef_cons, dim))
            if full and lib().oracle_set_vec_type(self.h, 1):
                raise OracleError(last_error())

    @property
    def full(self): return bool(lib().oracle_vec_type(self.h))

    def export_values(self):
        """FullVec store: (values[n, dim], levels[n])"""
        n, d = len(self), self.dim
        vals = np.zeros((n, d), np.float32)
        lv = np.zeros(n, np.uint8)
        lib().oracle_export_values(self.h, _p(vals, f32p))
        lib().oracle_export_points(self.h, None, None, None, _p(lv, u8p))
        return vals, lv

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_index_free(self.h)
            self.h = None

    def __len__(self): return lib().oracle_len(self.h)
    @property
    def nb_layers(self): return lib().oracle_nb_layers(self.h)
    @property
    def ep(self): return lib().oracle_ep(self.h)
    @property
    def dim(self): return lib().oracle_dim(self.h)

    def params(self):
        a = np.zeros(6, np.uint64)
        ml = C.c_float()
        lib().oracle_params(self.h, _p(a, u64p), C.byref(ml))
        return dict(m=int(a[0]), mmax=int(a[1]), mmax0=int(a[2]), ef_cons=int(a[3]), dim=int(a[4]),
                    ep=int(a[5]), ml=np.float32(ml.value))

    def insert_bulk(self, rows, levels=None):
        rows = np.ascontiguousarray(rows, np.float32)
        n, d = rows.shape
        lv = None if levels is None else np.ascontiguousarray(levels, np.uint8)
        ev = C.c_uint64()
        if lib().oracle_insert_bulk(self.h, _p(rows, f32p), n, d, _p(lv, u8p), C.byref(ev)):
            raise OracleError(last_error())
        self.build_evals = ev.value
        return self

    def insert_vec(self, row):
        row = np.ascontiguousarray(row, np.float32)
        r = lib().oracle_insert_vec(self.h, _p(row, f32p), row.shape[0])
        if r < 0:
            raise OracleError(last_error())
        return int(r)

    def distance(self, a, b):
        return np.float32(lib().oracle_distance(self.h, a, b))

    def export_points(self):
        n, d = len(self), self.dim
        codes = np.zeros((n, d), np.uint8)
        mins = np.zeros(n, np.float32)
        deltas = np.zeros(n, np.float32)
        lv = np.zeros(n, np.uint8)
        lib().oracle_export_points(self.h, _p(codes, u8p), _p(mins, f32p), _p(deltas, f32p), _p(lv, u8p))
        return codes, mins, deltas, lv

    def export_layer(self, l):
        nn = lib().oracle_layer_nb_nodes(self.h, l)
        ne = lib().oracle_layer_nb_edges(self.h, l)
        ids = np.zeros(nn, np.uint32)
        off = np.zeros(nn + 1, np.uint64)
        nb = np.zeros(max(ne, 1), np.uint32)
        lib().oracle_export_layer(self.h, l, _p(ids, u32p), _p(off, u64p), _p(nb, u32p))
        return ids, off, nb[:ne]

    def export_layers(self):
        return [self.export_layer(l) for l in range(self.nb_layers)]

    def layer_cap(self, l): return lib().oracle_layer_cap(self.h, l)

    def layer_degree_range(self, l):
        a, b = C.c_uint64(), C.c_uint64()
        lib().oracle_layer_degree_range(self.h, l, C.byref(a), C.byref(b))
        return a.value, b.value

    @staticmethod
    def from_parts(m, ef_cons, dim, ep, codes, mins, deltas, levels, layers):
        """mins is None and deltas is None: `codes` holds the f32 values of a FullVec index."""
        full = mins is None and deltas is None
        codes = np.ascontiguousarray(codes, np.float32 if full else np.uint8)
        if not full:
            mins = np.ascontiguousarray(mins, np.float32)
            deltas = np.ascontiguousarray(deltas, np.float32)
        levels = np.ascontiguousarray(levels, np.uint8)
        L = len(layers)
        keep = []
        nn = np.zeros(L, np.uint64)
        ids_arr = (u32p * L)()
        off_arr = (u64p * L)()
        nb_arr = (u32p * L)()
        for l, (ids, off, nb) in enumerate(layers):
            ids = np.ascontiguousarray(ids, np.uint32)
            off = np.ascontiguousarray(off, np.uint64)
            nb = np.ascontiguousarray(nb if len(nb) else np.zeros(1, np.uint32), np.uint32)
            keep += [ids, off, nb]
            nn[l] = ids.shape[0]
            ids_arr[l] = _p(ids, u32p)
            off_arr[l] = _p(off, u64p)
            nb_arr[l] = _p(nb, u32p)
        h = lib().oracle_index_from_parts(m, ef_cons, dim, ep, levels.shape[0], _p(codes, u8p),
                                          None if full else _p(mins, f32p),
                                          None if full else _p(deltas, f32p), _p(levels, u8p), L, _p(nn, u64p), ids_arr,
                                          off_arr, nb_arr)
        return Index(_handle=h)

    def save(self, path):
        if lib().oracle_save(self.h, str(path).encode()):
            raise OracleError(last_error())

    @staticmethod
    def load(path):
        h = lib().oracle_load(str(path).encode())
        if not h:
            raise OracleError(last_error())
        return Index(_handle=h)

    def search_batch(self, queries, n, ef, threads=1):
        """Returns ids[q,n] (0xFFFFFFFF pad), dists[q,n], counts[q], hops[q], evals[q]."""
        queries = np.ascontiguousarray(queries, np.float32)
        q = queries.shape[0]
        ids = np.zeros((q, n), np.uint32)
        dists = np.zeros((q, n), np.float32)
        counts = np.zeros(q, np.uint32)
        hops = np.zeros(q, np.uint32)
        evals = np.zeros(q, np.uint32)
        if lib().oracle_search_batch(self.h, _p(queries, f32p), q, n, ef, threads, _p(ids, u32p),
                                     _p(dists, f32p), _p(counts, u32p), _p(hops, u32p), _p(evals, u32p)):
            raise OracleError(last_error())
        return ids, dists, counts, hops, evals

    def ann_by_vector(self, vector, n, ef):
        ids, _, counts, _, _ = self.search_batch(np.asarray(vector, np.float32)[None, :], n, ef)
        return [int(x) for x in ids[0, :counts[0]]]

    def bruteforce(self, queries, k, threads=1):
        queries = np.ascontiguousarray(queries, np.float32)
        q = queries.shape[0]
        ids = np.zeros((q, k), np.uint32)
        dists = np.zeros((q, k), np.float32)
        if lib().oracle_bruteforce(self.h, _p(queries, f32p), q, k, threads, _p(ids, u32p), _p(dists, f32p)):
            raise OracleError(last_error())
        return ids, dists

    def dist_query_many(self, query, ids):
        query = np.ascontiguousarray(query, np.float32)
        ids = np.ascontiguousarray(ids, np.uint32)
        out = np.zeros(ids.shape[0], np.float32)
        if lib().oracle_dist_query_many(self.h, _p(query, f32p), _p(ids, u32p), ids.shape[0], _p(out, f32p)):
            raise OracleError(last_error())
        return out
