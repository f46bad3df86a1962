#!/usr/bin/env python3
"""C4 on one GPU (exact top-100, 1,000,000 x 100 base x 10,000 queries): wall time of hnswb200_bruteforce_topk_dev and, with
HNSWB200_BF_PROFILE=1, the per-chunk host timings on stderr.  Run under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hnsw_rs_b200 as H  # noqa: E402
from hnsw_rs_b200 import _ffi  # noqa: E402
from bench import synth  # noqa: E402

n, nq, k = int(os.environ.get("C4_N", 1000000)), 10000, int(os.environ.get("C4_K", 100))
ctx = H.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
base = synth(n, 100, 2048, 1)
q = synth(nq, 100, 2048, 2)
pts = H.SimplePoints.new(base, ctx=ctx)
dq = torch.from_numpy(q).cuda()
oi = torch.empty((nq, k), dtype=torch.int32, device="cuda")
od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
lib = _ffi.lib()
for rep in range(int(os.environ.get("C4_REPS", 3))):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _ffi.check(lib.hnswb200_bruteforce_topk_dev(ctx.h, pts.h, dq.data_ptr(), nq, k, 0, oi.data_ptr(), od.data_ptr()))
    torch.cuda.synchronize()
    print(f"rep {rep}: {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
print("checksum", int(oi.to(torch.int64).sum().item()))
