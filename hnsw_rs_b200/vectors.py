"""vectors crate: VecBase over FullVec / QuantVec, Serializer, gen_rand_vecs.

Distances and the quantiser run on the GPU through the C ABI (vectors/src/quant.rs:14-66,
vectors/src/full.rs:23-29); this module only holds the values and the byte formats.
"""
import struct

import numpy as np

from . import _ffi
from ._ffi import Context, check, f32, lib, ptr


class FullVec:
    """vectors/src/full.rs:3-6"""

    def __init__(self, vector):
        self.vector = f32(vector).copy()

    @staticmethod
    def new(vector):
        return FullVec(vector)

    def dim(self):
        return int(self.vector.shape[0])

    def iter_vals(self):
        return iter(self.vector)

    def get_vals(self):
        return self.vector.copy()

    def distance(self, other, ctx=None):  # full.rs:23-29: zip truncates to the shorter
        a, b = self.get_vals(), f32(other.get_vals())
        d = min(a.shape[0], b.shape[0])
        return _dist_full(a[:d], b[:d], ctx)

    def dist2other(self, other, ctx=None):
        return self.distance(other, ctx)

    def dist2many(self, others, ctx=None):
        others = list(others)
        if not others:
            return np.zeros(0, np.float32)
        d = min([self.dim()] + [o.dim() for o in others])
        x = np.repeat(self.vector[None, :d], len(others), axis=0)
        y = np.stack([o.get_vals()[:d] for o in others])
        return _dist_full_rows(x, y, ctx)

    # Serializer, full.rs:44-70
    def size(self):
        return self.dim() * 4

    def serialize(self):
        return self.vector.astype(">f4").tobytes()

    @staticmethod
    def deserialize(data):
        n = len(data) // 4
        return FullVec(np.frombuffer(bytes(data[:4 * n]), dtype=">f4").astype(np.float32))


class QuantVec:
    """vectors/src/quant.rs:6-11 (delta, min, codes); quantised on the device."""

    def __init__(self, delta, mn, codes):
        self.delta = np.float32(delta)
        self.min = np.float32(mn)
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)

    @staticmethod
    def new(vector, ctx=None):  # quant.rs:41-66
        v = f32(vector)
        codes, mins, deltas = quantise_rows(v[None, :], ctx)
        return QuantVec(deltas[0], mins[0], codes[0])

    def dim(self):
        return int(self.codes.shape[0])

    def get_vals(self):  # quant.rs:79-83: (c as f32) * delta + min, two roundings
        return (self.codes.astype(np.float32) * self.delta + self.min).astype(np.float32)

    def iter_vals(self):
        return iter(self.get_vals())

    def distance(self, other, ctx=None):  # quant.rs:67-73 generic zip distance
        a, b = self.get_vals(), f32(other.get_vals())
        d = min(a.shape[0], b.shape[0])
        return _dist_full(a[:d], b[:d], ctx)

    def dist2other(self, other, ctx=None):  # quant.rs:75-77 -> distance_unrolled
        return self.dist2many([other], ctx)[0]

    def dist2many(self, others, ctx=None):  # vectors/src/lib.rs:17-22
        others = list(others)
        if not others:
            return np.zeros(0, np.float32)
        ctx = ctx or Context.default()
        allv = [self] + others
        n, d = len(allv), self.dim()
        codes = np.stack([o.codes for o in allv])
        mins = np.array([o.min for o in allv], np.float32)
        deltas = np.array([o.delta for o in allv], np.float32)
        h = _ffi.vp()
        check(lib().hnswb200_points_upload(ctx.h, ptr(codes, _ffi.u8p), ptr(mins, _ffi.f32p),
                                           ptr(deltas, _ffi.f32p), None, n, d, _ffi.C.byref(h)))
        try:
            a = np.zeros(n - 1, np.uint32)
            b = np.arange(1, n, dtype=np.uint32)
            out = np.zeros(n - 1, np.float32)
            check(lib().hnswb200_dist_pairs(ctx.h, h, ptr(a, _ffi.u32p), ptr(b, _ffi.u32p), n - 1,
                                            ptr(out, _ffi.f32p)))
        finally:
            lib().hnswb200_points_destroy(h)
        return out

    # Serializer, quant.rs:90-125: min, delta (BE f32), then the codes
    def size(self):
        return 8 + self.dim()

    def serialize(self):
        return struct.pack(">ff", float(self.min), float(self.delta)) + self.codes.tobytes()

    @staticmethod
    def deserialize(data):
        mn, delta = struct.unpack(">ff", bytes(data[:8]))
        return QuantVec(delta, mn, np.frombuffer(bytes(data[8:]), dtype=np.uint8).copy())


def normalise_rows(rows, ctx=None):
    """rows / |row| on the device (what a cosine index applies to rows and queries); a zero row stays zero."""
    ctx = ctx or Context.default()
    r = f32(rows)
    if r.ndim != 2:
        raise ValueError("rows must be n x dim")
    out = np.empty_like(r)
    check(lib().hnswb200_normalise(ctx.h, ptr(r, _ffi.f32p), r.shape[0], r.shape[1], ptr(out, _ffi.f32p)))
    return out


def quantise_rows(rows, ctx=None):
    """QuantVec::new for every row -> (codes[n,dim] u8, mins[n], deltas[n])."""
    ctx = ctx or Context.default()
    rows = f32(rows)
    n, d = rows.shape
    codes = np.zeros((n, d), np.uint8)
    mins = np.zeros(n, np.float32)
    deltas = np.zeros(n, np.float32)
    check(lib().hnswb200_quantise(ctx.h, ptr(rows, _ffi.f32p), n, d, ptr(codes, _ffi.u8p),
                                  ptr(mins, _ffi.f32p), ptr(deltas, _ffi.f32p)))
    return codes, mins, deltas


def _dist_full_rows(x, y, ctx=None):
    ctx = ctx or Context.default()
    x, y = f32(x), f32(y)
    n, d = x.shape
    out = np.zeros(n, np.float32)
    check(lib().hnswb200_dist_full_pairs(ctx.h, ptr(x, _ffi.f32p), ptr(y, _ffi.f32p), n, d, ptr(out, _ffi.f32p)))
    return out


def _dist_full(a, b, ctx=None):
    return _dist_full_rows(a[None, :], b[None, :], ctx)[0]


def gen_rand_vecs(dim, n, rng=None):  # vectors/src/lib.rs:29-37
    assert n > 0
    rng = rng or np.random.default_rng()
    return [rng.random(dim, dtype=np.float32) for _ in range(n)]
