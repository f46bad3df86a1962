"""hnsw crate: HNSW (hnsw/src/template.rs) on the B200 engine.

Same public surface as the reference type -- new, insert_bulk, insert_vec, ann_by_vector,
save, load, len, distance, get_point, get_layer, layer_degrees, assert_param_compliance,
print_index, `params` -- plus ann_batch, the batched form the device is built for (one
C-ABI call = many queries).  Errors the reference reports as Err(String) or panics
surface as HnswB200Error carrying the same text.
"""
import ctypes as C
import math
import os

import numpy as np

from . import _ffi
from ._ffi import Context, HnswB200Error, check, f32, lib, ptr
from .graph import Graph
from .params import Params
from .points import SimplePoints


class HNSW:
    METRICS = {"l2": 0, "cosine": 1}  # HNSWB200_METRIC_*
    VEC_TYPES = {"quant": 0, "full": 1}  # HNSWB200_VEC_*: the reference's `type VecType` (points/src/point.rs:4)

    def __init__(self, m=12, ef_cons=None, dim=0, ctx=None, _handle=None, metric="l2", vec_type="quant"):
        self.ctx = ctx or Context.default()
        self._params0 = Params.from_m(m, dim) if ef_cons is None else Params.from_m_efcons(m, ef_cons, dim)
        self.h = _handle  # hnswb200_index*, created by the first insert_bulk / load
        self._metric = self.METRICS[metric]
        self._vec_type = self.VEC_TYPES[vec_type]

    @staticmethod
    def new(m, ef_cons, dim, ctx=None, metric="l2", vec_type="quant"):
        """template.rs:133-144.  metric="cosine" is an addition (unit-norm rows and queries); vec_type="full" is the index
        the reference builds when its VecType alias is flipped to FullVec (f32 vectors, strictly sequential distance)."""
        return HNSW(m, ef_cons, dim, ctx, metric=metric, vec_type=vec_type)

    @property
    def vec_type(self):
        code = int(lib().hnswb200_points_vec_type(lib().hnswb200_index_points(self.h))) if self.h else self._vec_type
        return "full" if code == 1 else "quant"

    @property
    def metric(self):
        code = int(lib().hnswb200_index_metric(self.h)) if self.h else self._metric
        return "cosine" if code == 1 else "l2"

    def set_metric(self, metric):
        """Only while the index is empty; needed again after load() (the reference's save format has no metric field)."""
        self._metric = self.METRICS[metric]
        if self.h:
            check(lib().hnswb200_index_set_metric(self.h, self._metric))
        return self

    def __del__(self):
        if getattr(self, "h", None):
            try:
                lib().hnswb200_index_destroy(self.h)
            except Exception:  # interpreter shutdown: the loader's modules may be gone already
                pass
            self.h = None

    # ---- params -----------------------------------------------------------
    @property
    def params(self):
        if self.h is None:
            return self._params0
        c = _ffi.Params()
        check(lib().hnswb200_index_params(self.h, C.byref(c)))
        return Params.from_c(c)

    def len(self):  # template.rs:146-148
        return int(lib().hnswb200_index_len(self.h)) if self.h else 0

    __len__ = len

    # ---- build ------------------------------------------------------------
    def insert_bulk(self, vectors, nb_threads=1, verbose=False, levels=None, batch=None):
        """template.rs:388-444.  nb_threads == 1 (the reference's deterministic mode) maps to
        batch = 1 only if `batch` says so explicitly; by default the device inserts points in
        batches against a frozen snapshot, the analogue of nb_threads > 1."""
        rows = f32(vectors)
        if rows.ndim != 2:
            raise ValueError("vectors must be n x dim")
        n, d = rows.shape
        lv = None if levels is None else np.ascontiguousarray(levels, np.uint8)
        b = 0 if batch is None else int(batch)
        if self.h is None:
            prm = self._params0.to_c()
            if d != prm.dim:
                raise HnswB200Error(_ffi.C.c_int(-1).value,
                                    f"The current index dimension is {prm.dim}, but tried inserting points of dimension {d}")
            h = _ffi.vp()
            # HNSW::new (empty index) with this index's VecType, choose the metric, then insert_bulk
            prev = int(lib().hnswb200_ctx_vec_type(self.ctx.h))
            check(lib().hnswb200_ctx_set_vec_type(self.ctx.h, self._vec_type))
            try:
                check(lib().hnswb200_build(self.ctx.h, None, 0, d, C.byref(prm), None, b, C.byref(h)))
            finally:
                lib().hnswb200_ctx_set_vec_type(self.ctx.h, prev)
            self.h = h
            if self._metric:
                check(lib().hnswb200_index_set_metric(self.h, self._metric))
            check(lib().hnswb200_index_insert_bulk(self.ctx.h, self.h, ptr(rows, _ffi.f32p), n, d, ptr(lv, _ffi.u8p), b))
        else:
            check(lib().hnswb200_index_insert_bulk(self.ctx.h, self.h, ptr(rows, _ffi.f32p), n, d,
                                                   ptr(lv, _ffi.u8p), b))
        return self

    def insert_vec(self, vector):  # template.rs:165-173
        row = f32(vector)
        if self.h is None:
            self.insert_bulk(row[None, :], batch=1)
            return 0
        out = C.c_uint32()
        check(lib().hnswb200_index_insert_vec(self.ctx.h, self.h, ptr(row, _ffi.f32p), row.shape[0], C.byref(out)))
        return int(out.value)

    # ---- query ------------------------------------------------------------
    def ann_batch(self, queries, n, ef, with_stats=False):
        """ann_by_vector for many queries in one call.  Returns (ids[q,n] padded with NO_ID,
        dists[q,n], counts[q]) and, with_stats, a dict of hops / evals / flags per query."""
        if self.h is None:
            raise HnswB200Error(-5, "search: the index holds no points")
        q = f32(queries)
        if q.ndim != 2:
            raise ValueError("queries must be nq x dim")
        nq, d = q.shape
        ids = np.full((nq, n), _ffi.NO_ID, np.uint32)
        dists = np.full((nq, n), np.inf, np.float32)
        counts = np.zeros(nq, np.uint32)
        st = None
        stats = None
        if with_stats:
            stats = dict(hops=np.zeros(nq, np.uint32), evals=np.zeros(nq, np.uint32), flags=np.zeros(nq, np.uint32),
                         nbrs=np.zeros(nq, np.uint32))
            st = _ffi.SearchStats(ptr(stats["hops"], _ffi.u32p), ptr(stats["evals"], _ffi.u32p),
                                  ptr(stats["flags"], _ffi.u32p), ptr(stats["nbrs"], _ffi.u32p))
        check(lib().hnswb200_search(self.ctx.h, self.h, ptr(q, _ffi.f32p), nq, d, n, ef, ptr(ids, _ffi.u32p),
                                    ptr(dists, _ffi.f32p), ptr(counts, _ffi.u32p),
                                    C.byref(st) if st is not None else None))
        if with_stats:
            return ids, dists, counts, stats
        return ids, dists, counts

    def ann_by_vector(self, vector, n, ef):  # template.rs:306-335: ids only, <= n of them
        ids, _, counts = self.ann_batch(f32(vector)[None, :], n, ef)
        return [int(x) for x in ids[0, :counts[0]]]

    # ---- accessors ----------------------------------------------------------
    def _points(self):
        return SimplePoints(self.ctx, lib().hnswb200_index_points(self.h), _owned=False)

    def distance(self, a, b):  # template.rs:150-152
        return self._points().distance(a, b) if self.h else None

    def get_point(self, point_id):  # template.rs:154-156
        return self._points().get_point(point_id) if self.h else None

    def nb_layers(self):
        return int(lib().hnswb200_graph_nb_layers(lib().hnswb200_index_graph(self.h))) if self.h else 0

    def export_layer(self, layer_nb):
        g = lib().hnswb200_index_graph(self.h)
        if layer_nb >= self.nb_layers():
            raise IndexError(f"Layer {layer_nb} not found in the structure.")  # layers.rs:28 panics
        nn = int(lib().hnswb200_graph_layer_nb_nodes(g, layer_nb))
        ne = int(lib().hnswb200_graph_layer_nb_edges(g, layer_nb))
        ids = np.zeros(nn, np.uint32)
        off = np.zeros(nn + 1, np.uint64)
        nb = np.zeros(max(ne, 1), np.uint32)
        check(lib().hnswb200_graph_export_layer(g, layer_nb, ptr(ids, _ffi.u32p), ptr(off, _ffi.u64p),
                                                ptr(nb, _ffi.u32p)))
        return ids, off, nb[:ne]

    def layer_cap(self, layer_nb):
        return int(lib().hnswb200_graph_layer_cap(lib().hnswb200_index_graph(self.h), layer_nb))

    def get_layer(self, layer_nb):  # template.rs:192-194 -> &Graph
        ids, off, nb = self.export_layer(layer_nb)
        return Graph.from_csr(layer_nb, self.layer_cap(layer_nb), ids, off, nb)

    def layer_degrees(self, layer_nb):  # template.rs:158-163 (prints; we also return them)
        ids, off, _ = self.export_layer(layer_nb)
        deg = np.diff(off.astype(np.int64))
        for dgr in deg:
            print(int(dgr))
        return deg

    def assert_param_compliance(self):  # template.rs:341-370
        p = self.params
        is_ok = True
        for l in range(self.nb_layers()):
            max_degree = p.mmax if l > 0 else p.mmax0
            ids, off, _ = self.export_layer(l)
            deg = np.diff(off.astype(np.int64))
            lim = math.ceil(float(np.float32(max_degree) * np.float32(1.1)))
            for node, dgr in zip(ids, deg):
                if dgr > lim:
                    is_ok = False
                    print(f"layer {l}, {node} degree = {dgr}, limit = {max_degree}")
                if dgr == 0 and len(ids) > 1:
                    is_ok = False
                    print(f"layer {l}, {node} degree = 0")
        if is_ok:
            print("Index complies with params.")
        return is_ok

    def print_index(self):  # template.rs:372-384
        p = self.params
        print(f"m = {p.m}\nmmax = {p.mmax}\nmmax0 = {p.mmax0}\nml = {p.ml}\nef_cons = {p.ef_cons}")
        print(f"Nb. layers = {self.nb_layers()}\nNb. of points = {self.len()}")
        g = lib().hnswb200_index_graph(self.h) if self.h else None
        for l in range(self.nb_layers()):
            print(f"NB. nodes in layer {l}: {int(lib().hnswb200_graph_layer_nb_nodes(g, l))}")
        print(f"ep: {p.ep}")

    # ---- persistence --------------------------------------------------------
    def save(self, path):  # template.rs:43-73
        if self.h is None:
            raise HnswB200Error(-5, "save: the index holds no points")
        check(lib().hnswb200_index_save_dir(self.ctx.h, self.h, os.fspath(path).encode()))

    @staticmethod
    def load(path, ctx=None):  # template.rs:75-131
        ctx = ctx or Context.default()
        h = _ffi.vp()
        check(lib().hnswb200_index_load_dir(ctx.h, os.fspath(path).encode(), C.byref(h)))
        c = _ffi.Params()
        check(lib().hnswb200_index_params(h, C.byref(c)))
        ix = HNSW(int(c.m), int(c.ef_cons), int(c.dim), ctx, _handle=h)
        return ix

    # ---- flat-array import (an index built elsewhere, e.g. by the reference) ----
    @staticmethod
    def from_parts(params, codes, mins, deltas, levels, layers, caps=None, ctx=None):
        """layers: list of (node_ids, offsets, nbrs) CSR triples, layer 0 first.  mins is None and deltas is None:
        `codes` holds the f32 values of a FullVec index."""
        ctx = ctx or Context.default()
        pts = SimplePoints.from_parts(codes, mins, deltas, levels, ctx)
        L = len(layers)
        if caps is None:
            caps = [2 * params.m if l == 0 else params.m for l in range(L)]
        caps = np.ascontiguousarray(caps, np.uint32)
        nn = np.zeros(L, np.uint64)
        keep = []
        ids_arr = (_ffi.u32p * L)()
        off_arr = (_ffi.u64p * L)()
        nb_arr = (_ffi.u32p * L)()
        for l, (ids, off, nb) in enumerate(layers):
            ids = np.ascontiguousarray(ids, np.uint32)
            off = np.ascontiguousarray(off, np.uint64)
            nb = np.ascontiguousarray(nb if len(nb) else np.zeros(1, np.uint32), np.uint32)
            keep += [ids, off, nb]
            nn[l] = ids.shape[0]
            ids_arr[l], off_arr[l], nb_arr[l] = ptr(ids, _ffi.u32p), ptr(off, _ffi.u64p), ptr(nb, _ffi.u32p)
        g = _ffi.vp()
        check(lib().hnswb200_graph_upload(ctx.h, len(pts), L, ptr(caps, _ffi.u32p), ptr(nn, _ffi.u64p), ids_arr,
                                          off_arr, nb_arr, C.byref(g)))
        h = _ffi.vp()
        prm = params.to_c()
        rc = lib().hnswb200_index_from_parts(ctx.h, pts.h, g, C.byref(prm), C.byref(h))
        if rc != 0:
            lib().hnswb200_graph_destroy(g)
            check(rc)
        pts._owned = False  # ownership moved into the index
        return HNSW(params.m, params.ef_cons, params.dim, ctx, _handle=h)
