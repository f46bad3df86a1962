"""hnsw/src/params.rs: Params + its byte format (52 bytes, big-endian)."""
import struct
from dataclasses import dataclass

import numpy as np

from . import _ffi


def get_default_ml(m):  # params.rs:15-17
    return np.float32(1.0) / np.log(np.float32(m))


@dataclass
class Params:
    ep: int
    m: int
    mmax: int
    mmax0: int
    ml: np.float32
    ef_cons: int
    dim: int

    @staticmethod
    def from_m(m, dim):  # params.rs:20-30
        return Params(0, m, m, m * 2, get_default_ml(m), m * 2, dim)

    @staticmethod
    def from_m_efcons(m, ef_cons, dim):  # params.rs:32-42
        return Params(0, m, m, m * 2, get_default_ml(m), ef_cons, dim)

    @staticmethod
    def from_(m, ef_cons=None, mmax=None, mmax0=None, ml=None, dim=0):  # params.rs:44-61
        return Params(0, m, m if mmax is None else mmax, m * 2 if mmax0 is None else mmax0,
                      get_default_ml(m) if ml is None else np.float32(ml),
                      m * 2 if ef_cons is None else ef_cons, dim)

    def size(self):  # params.rs:74-76 (claims 58; 52 bytes are written)
        return 58

    def serialize(self):  # params.rs:78-91
        return (struct.pack(">QQQ", self.m, self.mmax, self.mmax0) + struct.pack(">f", float(self.ml)) +
                struct.pack(">QQQ", self.ef_cons, self.dim, self.ep))

    @staticmethod
    def deserialize(data):  # params.rs:93-114
        m, mmax, mmax0 = struct.unpack(">QQQ", data[0:24])
        (ml,) = struct.unpack(">f", data[24:28])
        ef_cons, dim, ep = struct.unpack(">QQQ", data[28:52])
        return Params(ep & 0xFFFFFFFF, m, mmax, mmax0, np.float32(ml), ef_cons, dim)

    def to_c(self):
        return _ffi.Params(self.ep, self.m, self.mmax, self.mmax0, float(self.ml), self.ef_cons, self.dim)

    @staticmethod
    def from_c(c):
        return Params(c.ep, c.m, c.mmax, c.mmax0, np.float32(c.ml), c.ef_cons, c.dim)
