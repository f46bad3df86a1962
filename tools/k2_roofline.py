#!/usr/bin/env python3
"""K2 (batched distances) against the HBM roofline: one f32 query vs all 1,183,514 stored points of the C2 base,
ids in storage order (streaming) and in random order (gather).  Run under
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:dist_query_many --csv
to get the kernel's own duration; algorithmic bytes per pair = (8 + dim) record + 4 id + 4 distance."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hnsw_rs_b200 as H  # noqa: E402
from bench import synth  # noqa: E402

base = synth(1183514, 100, 2048, 1)
pts = H.SimplePoints.new(base)
q = synth(1, 100, 2048, 2)[0]
ids = np.arange(len(base), dtype=np.uint32)
for order in ("storage", "random"):
    if order == "random":
        np.random.default_rng(0).shuffle(ids)
    for _ in range(3):
        d = pts.dist_query_many(q, ids)
    print(order, float(d.astype(np.float64).sum()))
