"""dev: does an event record (what torch's async collectives enqueue) between two searches inhibit their overlap?"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import hnsw_rs_b200 as H
from hnsw_rs_b200 import _ffi
from bench import synth
ix = H.HNSW.load("/tmp/ix")
ctx = ix.ctx
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
q = synth(10000, 100, 2048, 2)
dq = torch.from_numpy(q).cuda()
nq = 10000
ids = [torch.empty((nq, 10), dtype=torch.int32, device="cuda") for _ in range(20)]
lib = _ffi.lib()
def run(i):
    _ffi.check(lib.hnswb200_search_dev(ctx.h, ix.h, dq.data_ptr(), nq, 10, 57, ids[i].data_ptr(), None, None, None, None, None, None))
side = torch.cuda.Stream()
for mode in ("plain", "event_record", "event_record+side_stream_wait"):
    for _ in range(3): run(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        run(i)
        if mode != "plain":
            ev = torch.cuda.Event()
            ev.record()
            if mode.endswith("wait"):
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    ids[i].add_(0)  # stand-in for the collective reading the step's result
    e1.record()
    torch.cuda.synchronize()
    print(f"{mode}: {e0.elapsed_time(e1) / 20:.3f} ms per step", flush=True)
