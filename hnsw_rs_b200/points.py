"""points crate: Point and SimplePoints (points/src/point.rs, points/src/points.rs).

SimplePoints keeps its vectors on the device (hnswb200_points, lane-sliced records,
csrc/layout.h); Point is the host-side value type.
"""
import ctypes as C
import struct

import numpy as np

from . import _ffi
from ._ffi import Context, check, f32, lib, ptr, u32
from .vectors import FullVec, QuantVec

VecType = QuantVec  # points/src/point.rs:4 (the reference's default; SimplePoints(vec_type="full") stores FullVec)
VEC_TYPES = {"quant": 0, "full": 1}  # HNSWB200_VEC_*


class Point:
    """points/src/point.rs:5-10"""

    def __init__(self, id, level, vector):
        self.id = int(id)
        self.level = int(level) & 0xFF
        self.vector = vector

    @staticmethod
    def new(vector, ctx=None):  # point.rs:24-30: id 0, level 0
        return Point(0, 0, VecType.new(vector, ctx))

    @staticmethod
    def with_level_and_id(vector, level, id, ctx=None):  # point.rs:13-18
        p = Point.new(vector, ctx)
        p.id, p.level = int(id), int(level) & 0xFF
        return p

    def distance(self, other, ctx=None):
        return self.vector.distance(other.vector if isinstance(other, Point) else other, ctx)

    def dist2other(self, other, ctx=None):  # point.rs:35-37
        return self.vector.dist2other(other.vector, ctx)

    def dist2many(self, others, ctx=None):
        return self.vector.dist2many([o.vector for o in others], ctx)

    def iter_vals(self):
        return self.vector.iter_vals()

    def get_vals(self):
        return self.vector.get_vals()

    def dim(self):
        return self.vector.dim()

    # Serializer, point.rs:46-76: level byte + vector; the id is positional
    def size(self):
        return 1 + self.vector.size()

    def serialize(self):
        return struct.pack(">B", self.level) + self.vector.serialize()

    @staticmethod
    def deserialize(data):
        return Point(0, data[0], VecType.deserialize(data[1:]))


class SimplePoints:
    """points/src/points.rs:33-116, device resident."""

    def __init__(self, ctx=None, _handle=None, _owned=True):
        self.ctx = ctx or Context.default()
        self.h = _handle
        self._owned = _owned

    def __del__(self):
        if getattr(self, "h", None) and self._owned:
            try:
                lib().hnswb200_points_destroy(self.h)
            except Exception:  # interpreter shutdown
                pass
            self.h = None

    @staticmethod
    def new(vecs, ml=None, levels=None, ctx=None, vec_type="quant"):
        """points.rs:39-48.  Levels: `levels` if given, else all 0 (the level draw belongs to
        HNSW::store_points in this design; see hnswb200_build).  vec_type="full": `type VecType = FullVec`."""
        ctx = ctx or Context.default()
        rows = f32(vecs)
        if rows.ndim != 2:
            raise ValueError("vecs must be n x dim")
        n, d = rows.shape
        lv = None if levels is None else np.ascontiguousarray(levels, np.uint8)
        h = _ffi.vp()
        if VEC_TYPES[vec_type]:
            check(lib().hnswb200_points_upload_f32(ctx.h, ptr(rows, _ffi.f32p), ptr(lv, _ffi.u8p), n, d, C.byref(h)))
        else:
            check(lib().hnswb200_points_from_f32(ctx.h, ptr(rows, _ffi.f32p), n, d, ptr(lv, _ffi.u8p), C.byref(h)))
        return SimplePoints(ctx, h)

    @property
    def vec_type(self):
        return "full" if int(lib().hnswb200_points_vec_type(self.h)) == 1 else "quant"

    def values(self):
        """get_vals of every point (vectors/src/lib.rs:24-26): (rows[n, dim], levels[n])"""
        n, d = self.len(), int(lib().hnswb200_points_dim(self.h))
        rows = np.zeros((n, d), np.float32)
        levels = np.zeros(n, np.uint8)
        check(lib().hnswb200_points_values(self.ctx.h, self.h, ptr(rows, _ffi.f32p), ptr(levels, _ffi.u8p)))
        return rows, levels

    @staticmethod
    def from_parts(codes, mins, deltas, levels=None, ctx=None):
        """mins is None and deltas is None: `codes` holds the f32 values of FullVec points."""
        ctx = ctx or Context.default()
        if mins is None and deltas is None:
            return SimplePoints.new(codes, levels=levels, ctx=ctx, vec_type="full")
        codes = np.ascontiguousarray(codes, np.uint8)
        n, d = codes.shape
        mins, deltas = f32(mins), f32(deltas)
        lv = None if levels is None else np.ascontiguousarray(levels, np.uint8)
        h = _ffi.vp()
        check(lib().hnswb200_points_upload(ctx.h, ptr(codes, _ffi.u8p), ptr(mins, _ffi.f32p),
                                           ptr(deltas, _ffi.f32p), ptr(lv, _ffi.u8p), n, d, C.byref(h)))
        return SimplePoints(ctx, h)

    def len(self):
        return int(lib().hnswb200_points_len(self.h))

    __len__ = len

    def ids(self):
        return iter(range(self.len()))

    def dim(self):
        return int(lib().hnswb200_points_dim(self.h)) if self.len() else None

    def download(self):
        n, d = self.len(), int(lib().hnswb200_points_dim(self.h))
        codes = np.zeros((n, d), np.uint8)
        mins = np.zeros(n, np.float32)
        deltas = np.zeros(n, np.float32)
        levels = np.zeros(n, np.uint8)
        check(lib().hnswb200_points_download(self.ctx.h, self.h, ptr(codes, _ffi.u8p), ptr(mins, _ffi.f32p),
                                             ptr(deltas, _ffi.f32p), ptr(levels, _ffi.u8p)))
        return codes, mins, deltas, levels

    def get_point(self, idx):  # points.rs:75-77 (None when out of range)
        if idx < 0 or idx >= self.len():
            return None
        if self.vec_type == "full":
            rows, levels = self.values()
            return Point(idx, levels[idx], FullVec(rows[idx]))
        codes, mins, deltas, levels = self.download()
        return Point(idx, levels[idx], QuantVec(deltas[idx], mins[idx], codes[idx]))

    def get_points_iter(self, indices):
        if self.vec_type == "full":
            rows, levels = self.values()
            for i in indices:
                yield Point(i, levels[i], FullVec(rows[i]))
            return
        codes, mins, deltas, levels = self.download()
        for i in indices:
            yield Point(i, levels[i], QuantVec(deltas[i], mins[i], codes[i]))

    def distance(self, a_idx, b_idx):  # points.rs:86-93
        n = self.len()
        if not (0 <= a_idx < n and 0 <= b_idx < n):
            return None
        return self.distances(np.array([a_idx]), np.array([b_idx]))[0]

    def distances(self, a, b):
        a, b = u32(a), u32(b)
        out = np.zeros(a.shape[0], np.float32)
        check(lib().hnswb200_dist_pairs(self.ctx.h, self.h, ptr(a, _ffi.u32p), ptr(b, _ffi.u32p), a.shape[0],
                                        ptr(out, _ffi.f32p)))
        return out

    def distance2point(self, point, idx):  # points.rs:95-101; `point` is an f32 vector or a Point
        if idx < 0 or idx >= self.len():
            return None
        return self.dist_query_many(point, np.array([idx]))[0]

    def dist_query_many(self, query, ids):
        """dist2many of the (re-quantised) f32 query against stored ids."""
        q = f32(query.get_vals() if isinstance(query, Point) else query)
        ids = u32(ids)
        out = np.zeros(ids.shape[0], np.float32)
        check(lib().hnswb200_dist_query_many(self.ctx.h, self.h, ptr(q, _ffi.f32p), ptr(ids, _ffi.u32p),
                                             ids.shape[0], ptr(out, _ffi.f32p)))
        return out

    # Serializer, points.rs:119-146
    def size(self):
        d = self.dim() or 0
        return 16 + self.len() * ((1 + 4 * d) if self.vec_type == "full" else (9 + d))

    def serialize(self):
        if self.vec_type == "full":  # point.rs:55-61 + full.rs:55-61
            rows, levels = self.values()
            n, d = rows.shape
            rec = np.zeros((n, 1 + 4 * d), np.uint8)
            rec[:, 0] = levels
            rec[:, 1:] = rows.astype(">f4").view(np.uint8).reshape(n, 4 * d)
            return struct.pack(">QQ", n, 1 + 4 * d) + rec.tobytes()
        codes, mins, deltas, levels = self.download()
        n, d = codes.shape
        rec = np.zeros((n, 9 + d), np.uint8)
        rec[:, 0] = levels
        rec[:, 1:5] = mins.astype(">f4").view(np.uint8).reshape(n, 4)
        rec[:, 5:9] = deltas.astype(">f4").view(np.uint8).reshape(n, 4)
        rec[:, 9:] = codes
        return struct.pack(">QQ", n, 9 + d) + rec.tobytes()

    @staticmethod
    def deserialize(data, ctx=None, vec_type="quant"):
        """The byte string does not name its VecType (the reference fixes it at compile time): say which."""
        n, psz = struct.unpack(">QQ", bytes(data[:16]))
        rec = np.frombuffer(bytes(data[16:16 + n * psz]), np.uint8).reshape(n, psz)
        levels = rec[:, 0].copy()
        if VEC_TYPES[vec_type]:
            rows = rec[:, 1:].copy().view(">f4").astype(np.float32).reshape(n, (psz - 1) // 4)
            return SimplePoints.new(rows, levels=levels, ctx=ctx, vec_type="full")
        mins = rec[:, 1:5].copy().view(">f4").astype(np.float32).reshape(n)
        deltas = rec[:, 5:9].copy().view(">f4").astype(np.float32).reshape(n)
        return SimplePoints.from_parts(rec[:, 9:].copy(), mins, deltas, levels, ctx)


def new_layer(ml, rng):  # points.rs:148-160 (rng: any object with .random() in [0,1))
    r = np.float32(0.0)
    while r == 0.0 or r == 1.0:
        r = np.float32(rng.random())
    return int(np.floor(-np.log(r) * np.float32(ml)))
