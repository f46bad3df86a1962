#!/usr/bin/env python3
"""Writes tests/golden/expected_{quant,full}.npz: what the CPU oracle (oracle/hnsw_oracle.cpp, pinned to the reference's own
tests) answers on the reference's fixture (store.txt / queries.txt, copied from test-data/): the index
HNSW::new(12, None, 50).insert_bulk(store, 1, ..) in the single-thread order, ann_by_vector(q, 10, ef) for ef in (10, 100) with
distances as bit patterns and the hop / evaluation counters, and the exact top-10 (brute_force_nns).  quant = the reference as
committed (VecType = QuantVec), full = the alias flipped to FullVec.  Run from the repository root:
    python tests/golden/make_expected.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as O  # noqa: E402


def expected(full):
    store = O.load_glove(os.path.join(HERE, "store.txt"))
    queries = O.load_glove(os.path.join(HERE, "queries.txt"))
    ix = O.Index(12, None, store.shape[1], full=full).insert_bulk(store)
    out = {"ep": np.uint32(ix.ep), "nb_layers": np.uint32(ix.nb_layers)}
    for l, (ids, off, nb) in enumerate(ix.export_layers()):
        out[f"layer{l}_ids"], out[f"layer{l}_off"], out[f"layer{l}_nbrs"] = ids, off, nb
    for ef in (10, 100):
        ids, dists, counts, hops, evals = ix.search_batch(queries, 10, ef)
        out[f"ef{ef}_ids"], out[f"ef{ef}_dist_bits"], out[f"ef{ef}_counts"] = ids, dists.view(np.uint32), counts
        out[f"ef{ef}_hops"], out[f"ef{ef}_evals"] = hops, evals
    gt, gd = ix.bruteforce(queries, 10)
    out["bf_ids"], out["bf_dist_bits"] = gt, gd.view(np.uint32)
    return out


if __name__ == "__main__":
    for name, full in (("quant", False), ("full", True)):
        np.savez_compressed(os.path.join(HERE, f"expected_{name}.npz"), **expected(full))
        print("wrote", f"expected_{name}.npz")
