"""dev: where the end-to-end step time goes (host call overhead vs kernel)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import hnsw_rs_b200 as H
from hnsw_rs_b200 import _ffi
from bench import synth
ix = H.HNSW.load("/tmp/ix")
ctx = ix.ctx
lib = _ffi.lib()
q = synth(10000, 100, 2048, 2)
for nq in (1, 100, 1000, 10000):
    hq = torch.from_numpy(q[:nq].copy()).pin_memory()
    hid = torch.zeros((nq, 10), dtype=torch.int32).pin_memory()
    hd = torch.zeros((nq, 10), dtype=torch.float32).pin_memory()
    hc = torch.zeros(nq, dtype=torch.int32).pin_memory()
    def call():
        _ffi.check(lib.hnswb200_search(ctx.h, ix.h, C.cast(hq.data_ptr(), _ffi.f32p), nq, 100, 10, 57,
                                       C.cast(hid.data_ptr(), _ffi.u32p), C.cast(hd.data_ptr(), _ffi.f32p),
                                       C.cast(hc.data_ptr(), _ffi.u32p), None))
    for mode in ("zero-copy", "staged"):
        if mode == "staged": os.environ["HNSWB200_NO_ZERO_COPY"] = "1"
        else: os.environ.pop("HNSWB200_NO_ZERO_COPY", None)
        for _ in range(10): call()
        t = time.perf_counter()
        for _ in range(50): call()
        dt = (time.perf_counter() - t) / 50
        print(f"nq={nq:6d} {mode:10s}: {dt*1e6:8.1f} us per call", flush=True)
