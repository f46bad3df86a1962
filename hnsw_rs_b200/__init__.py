"""hnsw_rs_b200: B200-native (sm_100a) HNSW query / distance / build engine behind the API
surface of the Rust workspace Gumo-A/hnsw_rs.  Everything that computes runs in
libhnsw_b200.so (hand-written CUDA, C ABI in include/hnsw_b200.h); this package is the
host-side mirror of the reference's crates: vectors, points, graph, hnsw.
"""
from ._ffi import Context, HnswB200Error, NO_ID, LIB_PATH, lib
from .graph import Dist, Graph, GraphError, Layers
from .helpers import brute_force_nns, bruteforce_topk, load_glove_array, topk_merge
from .hnsw import HNSW
from .params import Params, get_default_ml
from .points import Point, SimplePoints, new_layer
from .vectors import FullVec, QuantVec, gen_rand_vecs, normalise_rows, quantise_rows

__all__ = ["Context", "HnswB200Error", "NO_ID", "LIB_PATH", "lib", "Dist", "Graph", "GraphError", "Layers",
           "brute_force_nns", "bruteforce_topk", "load_glove_array", "topk_merge", "HNSW", "Params",
           "get_default_ml", "Point", "SimplePoints", "new_layer", "FullVec", "QuantVec", "gen_rand_vecs",
           "quantise_rows", "normalise_rows"]
