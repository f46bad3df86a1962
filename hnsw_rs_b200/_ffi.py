"""ctypes binding of libhnsw_b200.so (include/hnsw_b200.h).

This is the same C ABI a Rust `-sys` crate would bind.  There is no
fallback of any kind: if the shared library is missing the import fails, and if no
CUDA device is present every compute call raises HnswB200Error.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HNSWB200_LIB", os.path.join(_HERE, "libhnsw_b200.so"))  # env: A/B builds of the same ABI

NO_ID = 0xFFFFFFFF

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
vp = C.c_void_p


class Params(C.Structure):
    """hnswb200_params == hnsw/src/params.rs:4-12"""
    _fields_ = [("ep", C.c_uint32), ("m", C.c_uint64), ("mmax", C.c_uint64), ("mmax0", C.c_uint64),
                ("ml", C.c_float), ("ef_cons", C.c_uint64), ("dim", C.c_uint64)]


class SearchStats(C.Structure):
    _fields_ = [("hops", u32p), ("evals", u32p), ("flags", u32p), ("nbrs", u32p)]


class HnswB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


# every symbol include/hnsw_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hnswb200_last_error": (C.c_char_p, []),
    "hnswb200_version": (C.c_int, []),
    "hnswb200_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "hnswb200_ctx_destroy": (None, [vp]),
    "hnswb200_ctx_set_stream": (C.c_int, [vp, vp]),
    "hnswb200_ctx_set_overlap": (C.c_int, [vp, C.c_int]),
    "hnswb200_ctx_sync": (C.c_int, [vp]),
    "hnswb200_ctx_device": (C.c_int, [vp]),
    "hnswb200_ctx_set_vec_type": (C.c_int, [vp, C.c_int]),
    "hnswb200_ctx_vec_type": (C.c_int, [vp]),
    "hnswb200_params_default": (None, [C.c_uint64, C.c_int64, C.c_uint64, C.POINTER(Params)]),
    "hnswb200_quantise": (C.c_int, [vp, f32p, C.c_uint64, C.c_uint32, u8p, f32p, f32p]),
    "hnswb200_normalise": (C.c_int, [vp, f32p, C.c_uint64, C.c_uint32, f32p]),
    "hnswb200_points_set_metric": (C.c_int, [vp, C.c_int]),
    "hnswb200_points_metric": (C.c_int, [vp]),
    "hnswb200_index_set_metric": (C.c_int, [vp, C.c_int]),
    "hnswb200_index_metric": (C.c_int, [vp]),
    "hnswb200_dist_full_pairs": (C.c_int, [vp, f32p, f32p, C.c_uint64, C.c_uint32, f32p]),
    "hnswb200_points_upload": (C.c_int, [vp, u8p, f32p, f32p, u8p, C.c_uint64, C.c_uint32, C.POINTER(vp)]),
    "hnswb200_points_from_f32": (C.c_int, [vp, f32p, C.c_uint64, C.c_uint32, u8p, C.POINTER(vp)]),
    "hnswb200_points_download": (C.c_int, [vp, vp, u8p, f32p, f32p, u8p]),
    "hnswb200_points_upload_f32": (C.c_int, [vp, f32p, u8p, C.c_uint64, C.c_uint32, C.POINTER(vp)]),
    "hnswb200_points_values": (C.c_int, [vp, vp, f32p, u8p]),
    "hnswb200_points_vec_type": (C.c_int, [vp]),
    "hnswb200_points_len": (C.c_uint64, [vp]),
    "hnswb200_points_dim": (C.c_uint32, [vp]),
    "hnswb200_points_destroy": (None, [vp]),
    "hnswb200_dist_pairs": (C.c_int, [vp, vp, u32p, u32p, C.c_uint64, f32p]),
    "hnswb200_dist_query_many": (C.c_int, [vp, vp, f32p, u32p, C.c_uint64, f32p]),
    "hnswb200_graph_upload": (C.c_int, [vp, C.c_uint64, C.c_uint32, u32p, u64p, C.POINTER(u32p),
                                        C.POINTER(u64p), C.POINTER(u32p), C.POINTER(vp)]),
    "hnswb200_graph_nb_layers": (C.c_uint32, [vp]),
    "hnswb200_graph_layer_nb_nodes": (C.c_uint64, [vp, C.c_uint32]),
    "hnswb200_graph_layer_nb_edges": (C.c_uint64, [vp, C.c_uint32]),
    "hnswb200_graph_layer_cap": (C.c_uint32, [vp, C.c_uint32]),
    "hnswb200_graph_export_layer": (C.c_int, [vp, C.c_uint32, u32p, u64p, u32p]),
    "hnswb200_graph_destroy": (None, [vp]),
    "hnswb200_index_from_parts": (C.c_int, [vp, vp, vp, C.POINTER(Params), C.POINTER(vp)]),
    "hnswb200_build": (C.c_int, [vp, f32p, C.c_uint64, C.c_uint32, C.POINTER(Params), u8p, C.c_uint32,
                                 C.POINTER(vp)]),
    "hnswb200_index_insert_bulk": (C.c_int, [vp, vp, f32p, C.c_uint64, C.c_uint32, u8p, C.c_uint32]),
    "hnswb200_index_insert_vec": (C.c_int, [vp, vp, f32p, C.c_uint32, u32p]),
    "hnswb200_index_save_dir": (C.c_int, [vp, vp, C.c_char_p]),
    "hnswb200_index_load_dir": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
    "hnswb200_index_destroy": (None, [vp]),
    "hnswb200_index_params": (C.c_int, [vp, C.POINTER(Params)]),
    "hnswb200_index_len": (C.c_uint64, [vp]),
    "hnswb200_index_points": (vp, [vp]),
    "hnswb200_index_graph": (vp, [vp]),
    "hnswb200_search": (C.c_int, [vp, vp, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u32p, f32p,
                                  u32p, C.POINTER(SearchStats)]),
    "hnswb200_search_async": (C.c_int, [vp, vp, f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u32p, f32p, u32p]),
    "hnswb200_search_dev": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, vp, vp]),
    "hnswb200_search_dev_gather": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_uint32,
                                             C.POINTER(vp), C.c_uint64]),
    "hnswb200_last_search_variant": (C.c_char_p, []),
    "hnswb200_search_dev_shard": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, vp,
                                            C.c_uint32, C.POINTER(vp), C.POINTER(vp), C.c_uint64]),
    "hnswb200_peer_put_dev": (C.c_int, [vp, vp, C.c_uint64, C.c_uint32, C.POINTER(vp)]),
    "hnswb200_peer_signal_dev": (C.c_int, [vp, C.c_uint32, C.POINTER(vp), C.c_uint32, C.c_uint32]),
    "hnswb200_peer_wait_dev": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32]),
    "hnswb200_dev_alloc": (C.c_int, [vp, C.c_uint64, C.POINTER(vp)]),
    "hnswb200_dev_free": (C.c_int, [vp, vp]),
    "hnswb200_dev_download": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "hnswb200_ipc_export": (C.c_int, [vp, vp, u8p]),
    "hnswb200_ipc_open": (C.c_int, [vp, u8p, C.POINTER(vp)]),
    "hnswb200_ipc_close": (C.c_int, [vp, vp]),
    "hnswb200_bruteforce_topk": (C.c_int, [vp, vp, f32p, C.c_uint64, C.c_uint32, C.c_uint32, u32p, f32p]),
    "hnswb200_bruteforce_topk_dev": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp]),
    "hnswb200_topk_merge": (C.c_int, [vp, u32p, f32p, C.c_uint32, C.c_uint64, C.c_uint32, u32p, f32p]),
    "hnswb200_topk_merge_dev": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint64, C.c_uint32, vp, vp]),
    "hnswb200_load_glove": (C.c_int64, [C.c_char_p, C.c_uint64, f32p, C.c_uint64, u64p]),
}

_lib = None


def last_search_variant():
    """Name of the kernel variant the calling thread's last search ran (hnswb200_last_search_variant)."""
    s = lib().hnswb200_last_search_variant()
    return s.decode() if s else ""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  hnsw_rs_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise HnswB200Error(rc, lib().hnswb200_last_error().decode(errors="replace"))


def ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


class Context:
    """One per GPU (hnswb200_ctx): owns the stream all calls run on."""
    _default = {}

    def __init__(self, device=0):
        h = vp()
        check(lib().hnswb200_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def set_stream(self, cuda_stream_ptr):
        check(lib().hnswb200_ctx_set_stream(self.h, vp(cuda_stream_ptr)))

    def set_overlap(self, allow=True):
        """Opt in to overlapping consecutive searches on an adopted stream (hnswb200_ctx_set_overlap): the caller
        promises that no kernel producing a search's query buffer is enqueued between two searches."""
        check(lib().hnswb200_ctx_set_overlap(self.h, 1 if allow else 0))

    def sync(self):
        check(lib().hnswb200_ctx_sync(self.h))

    def close(self):
        if getattr(self, "h", None):
            lib().hnswb200_ctx_destroy(self.h)
            self.h = None

    @classmethod
    def default(cls, device=0):
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]
