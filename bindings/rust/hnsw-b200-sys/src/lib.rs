//! Raw declarations of the C ABI in `include/hnsw_b200.h`, one for one.
//! Written for the Rust workspace Gumo-A/hnsw_rs; NOT compiled in the engine's own CI
//! (there is no Rust toolchain in that image) -- the same entry points are exercised
//! through ctypes by the engine's parity tests.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const HNSWB200_OK: c_int = 0;
pub const HNSWB200_EINVAL: c_int = -1;
pub const HNSWB200_ECUDA: c_int = -2;
pub const HNSWB200_EIO: c_int = -3;
pub const HNSWB200_ENOMEM: c_int = -4;
pub const HNSWB200_ESTATE: c_int = -5;
pub const HNSWB200_NO_ID: u32 = 0xFFFF_FFFF;
pub const HNSWB200_METRIC_L2: c_int = 0;
pub const HNSWB200_METRIC_COSINE: c_int = 1;
pub const HNSWB200_VEC_QUANT: c_int = 0;
pub const HNSWB200_VEC_FULL: c_int = 1;

#[repr(C)] pub struct hnswb200_ctx { _p: [u8; 0] }
#[repr(C)] pub struct hnswb200_points { _p: [u8; 0] }
#[repr(C)] pub struct hnswb200_graph { _p: [u8; 0] }
#[repr(C)] pub struct hnswb200_index { _p: [u8; 0] }

/// hnsw/src/params.rs:4-12
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct hnswb200_params {
    pub ep: u32,
    pub m: u64,
    pub mmax: u64,
    pub mmax0: u64,
    pub ml: f32,
    pub ef_cons: u64,
    pub dim: u64,
}

#[repr(C)]
pub struct hnswb200_search_stats {
    pub hops: *mut u32,
    pub evals: *mut u32,
    pub flags: *mut u32,
    pub nbrs: *mut u32,
}

extern "C" {
    pub fn hnswb200_last_error() -> *const c_char;
    pub fn hnswb200_version() -> c_int;

    pub fn hnswb200_ctx_create(device: c_int, out: *mut *mut hnswb200_ctx) -> c_int;
    pub fn hnswb200_ctx_destroy(ctx: *mut hnswb200_ctx);
    pub fn hnswb200_ctx_set_stream(ctx: *mut hnswb200_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn hnswb200_ctx_sync(ctx: *mut hnswb200_ctx) -> c_int;
    pub fn hnswb200_ctx_device(ctx: *const hnswb200_ctx) -> c_int;
    // `type VecType` (points/src/point.rs:4) as a property of the context: HNSWB200_VEC_QUANT (0) / HNSWB200_VEC_FULL (1)
    pub fn hnswb200_ctx_set_vec_type(ctx: *mut hnswb200_ctx, vec_type: c_int) -> c_int;
    pub fn hnswb200_ctx_vec_type(ctx: *const hnswb200_ctx) -> c_int;

    pub fn hnswb200_params_default(m: u64, ef_cons: i64, dim: u64, out: *mut hnswb200_params);

    // vectors crate
    pub fn hnswb200_quantise(ctx: *mut hnswb200_ctx, rows: *const f32, n: u64, dim: u32, codes: *mut u8,
                             mins: *mut f32, deltas: *mut f32) -> c_int;
    pub fn hnswb200_normalise(ctx: *mut hnswb200_ctx, rows: *const f32, n: u64, dim: u32, out: *mut f32) -> c_int;
    pub fn hnswb200_points_set_metric(p: *mut hnswb200_points, metric: c_int) -> c_int;
    pub fn hnswb200_points_metric(p: *const hnswb200_points) -> c_int;
    pub fn hnswb200_index_set_metric(ix: *mut hnswb200_index, metric: c_int) -> c_int;
    pub fn hnswb200_index_metric(ix: *const hnswb200_index) -> c_int;
    pub fn hnswb200_dist_full_pairs(ctx: *mut hnswb200_ctx, x: *const f32, y: *const f32, n: u64, dim: u32,
                                    out: *mut f32) -> c_int;

    // points crate
    pub fn hnswb200_points_upload(ctx: *mut hnswb200_ctx, codes: *const u8, mins: *const f32, deltas: *const f32,
                                  levels: *const u8, n: u64, dim: u32, out: *mut *mut hnswb200_points) -> c_int;
    pub fn hnswb200_points_from_f32(ctx: *mut hnswb200_ctx, rows: *const f32, n: u64, dim: u32, levels: *const u8,
                                    out: *mut *mut hnswb200_points) -> c_int;
    pub fn hnswb200_points_download(ctx: *mut hnswb200_ctx, p: *const hnswb200_points, codes: *mut u8, mins: *mut f32,
                                    deltas: *mut f32, levels: *mut u8) -> c_int;
    pub fn hnswb200_points_upload_f32(ctx: *mut hnswb200_ctx, rows: *const f32, levels: *const u8, n: u64, dim: u32,
                                      out: *mut *mut hnswb200_points) -> c_int;
    pub fn hnswb200_points_values(ctx: *mut hnswb200_ctx, p: *const hnswb200_points, rows: *mut f32, levels: *mut u8) -> c_int;
    pub fn hnswb200_points_vec_type(p: *const hnswb200_points) -> c_int;
    pub fn hnswb200_points_len(p: *const hnswb200_points) -> u64;
    pub fn hnswb200_points_dim(p: *const hnswb200_points) -> u32;
    pub fn hnswb200_points_destroy(p: *mut hnswb200_points);
    pub fn hnswb200_dist_pairs(ctx: *mut hnswb200_ctx, p: *const hnswb200_points, a: *const u32, b: *const u32, n: u64,
                               out: *mut f32) -> c_int;
    pub fn hnswb200_dist_query_many(ctx: *mut hnswb200_ctx, p: *const hnswb200_points, query: *const f32,
                                    ids: *const u32, n: u64, out: *mut f32) -> c_int;

    // graph crate
    pub fn hnswb200_graph_upload(ctx: *mut hnswb200_ctx, n_points: u64, n_layers: u32, caps: *const u32,
                                 n_nodes: *const u64, node_ids: *const *const u32, offsets: *const *const u64,
                                 nbrs: *const *const u32, out: *mut *mut hnswb200_graph) -> c_int;
    pub fn hnswb200_graph_nb_layers(g: *const hnswb200_graph) -> u32;
    pub fn hnswb200_graph_layer_nb_nodes(g: *const hnswb200_graph, layer: u32) -> u64;
    pub fn hnswb200_graph_layer_nb_edges(g: *const hnswb200_graph, layer: u32) -> u64;
    pub fn hnswb200_graph_layer_cap(g: *const hnswb200_graph, layer: u32) -> u32;
    pub fn hnswb200_graph_export_layer(g: *const hnswb200_graph, layer: u32, node_ids: *mut u32, offsets: *mut u64,
                                       nbrs: *mut u32) -> c_int;
    pub fn hnswb200_graph_destroy(g: *mut hnswb200_graph);

    // hnsw crate
    pub fn hnswb200_index_from_parts(ctx: *mut hnswb200_ctx, points: *mut hnswb200_points, graph: *mut hnswb200_graph,
                                     params: *const hnswb200_params, out: *mut *mut hnswb200_index) -> c_int;
    pub fn hnswb200_build(ctx: *mut hnswb200_ctx, rows: *const f32, n: u64, dim: u32, params: *const hnswb200_params,
                          levels: *const u8, batch: u32, out: *mut *mut hnswb200_index) -> c_int;
    pub fn hnswb200_index_insert_bulk(ctx: *mut hnswb200_ctx, ix: *mut hnswb200_index, rows: *const f32, n: u64,
                                      dim: u32, levels: *const u8, batch: u32) -> c_int;
    pub fn hnswb200_index_insert_vec(ctx: *mut hnswb200_ctx, ix: *mut hnswb200_index, row: *const f32, dim: u32,
                                     id_out: *mut u32) -> c_int;
    pub fn hnswb200_index_save_dir(ctx: *mut hnswb200_ctx, ix: *const hnswb200_index, dir: *const c_char) -> c_int;
    pub fn hnswb200_index_load_dir(ctx: *mut hnswb200_ctx, dir: *const c_char, out: *mut *mut hnswb200_index) -> c_int;
    pub fn hnswb200_index_destroy(ix: *mut hnswb200_index);
    pub fn hnswb200_index_params(ix: *const hnswb200_index, out: *mut hnswb200_params) -> c_int;
    pub fn hnswb200_index_len(ix: *const hnswb200_index) -> u64;
    pub fn hnswb200_index_points(ix: *const hnswb200_index) -> *const hnswb200_points;
    pub fn hnswb200_index_graph(ix: *const hnswb200_index) -> *const hnswb200_graph;

    pub fn hnswb200_search(ctx: *mut hnswb200_ctx, ix: *const hnswb200_index, queries: *const f32, nq: u64, dim: u32,
                           n: u32, ef: u32, out_ids: *mut u32, out_dists: *mut f32, out_counts: *mut u32,
                           stats: *const hnswb200_search_stats) -> c_int;
    pub fn hnswb200_search_async(ctx: *mut hnswb200_ctx, ix: *const hnswb200_index, queries: *const f32, nq: u64, dim: u32,
                                 n: u32, ef: u32, out_ids: *mut u32, out_dists: *mut f32, out_counts: *mut u32) -> c_int;
    pub fn hnswb200_search_dev(ctx: *mut hnswb200_ctx, ix: *const hnswb200_index, d_queries: *const f32, nq: u64, n: u32,
                               ef: u32, d_out_ids: *mut u32, d_out_dists: *mut f32, d_out_counts: *mut u32,
                               d_hops: *mut u32, d_evals: *mut u32, d_flags: *mut u32, d_nbrs: *mut u32) -> c_int;

    pub fn hnswb200_ctx_set_overlap(ctx: *mut hnswb200_ctx, allow: c_int) -> c_int;
    pub fn hnswb200_last_search_variant() -> *const c_char;
    pub fn hnswb200_search_dev_shard(ctx: *mut hnswb200_ctx, ix: *const hnswb200_index, d_queries: *const f32, nq: u64,
                                     n: u32, ef: u32, id_offset: u32, d_out_ids: *mut u32, d_out_dists: *mut f32,
                                     d_out_counts: *mut u32, n_peers: u32, peer_ids: *const *mut u32,
                                     peer_dists: *const *mut f32, row_offset: u64) -> c_int;
    pub fn hnswb200_peer_put_dev(ctx: *mut hnswb200_ctx, d_src: *const c_void, bytes: u64, n_peers: u32,
                                 peer_dst: *const *mut c_void) -> c_int;
    pub fn hnswb200_peer_signal_dev(ctx: *mut hnswb200_ctx, n_peers: u32, peer_flags: *const *mut u32, slot: u32,
                                    epoch: u32) -> c_int;
    pub fn hnswb200_peer_wait_dev(ctx: *mut hnswb200_ctx, d_flags: *const u32, n_slots: u32, epoch: u32) -> c_int;
    pub fn hnswb200_search_dev_gather(ctx: *mut hnswb200_ctx, ix: *const hnswb200_index, d_queries: *const f32, nq: u64,
                                      n: u32, ef: u32, d_out_ids: *mut u32, d_out_dists: *mut f32, d_out_counts: *mut u32,
                                      n_peers: u32, peer_ids: *const *mut u32, row_offset: u64) -> c_int;
    pub fn hnswb200_dev_alloc(ctx: *mut hnswb200_ctx, bytes: u64, out: *mut *mut c_void) -> c_int;
    pub fn hnswb200_dev_free(ctx: *mut hnswb200_ctx, ptr: *mut c_void) -> c_int;
    pub fn hnswb200_dev_download(ctx: *mut hnswb200_ctx, d_src: *const c_void, host_dst: *mut c_void, bytes: u64) -> c_int;
    pub fn hnswb200_ipc_export(ctx: *mut hnswb200_ctx, d_ptr: *mut c_void, handle: *mut u8) -> c_int;
    pub fn hnswb200_ipc_open(ctx: *mut hnswb200_ctx, handle: *const u8, out: *mut *mut c_void) -> c_int;
    pub fn hnswb200_ipc_close(ctx: *mut hnswb200_ctx, ptr: *mut c_void) -> c_int;

    pub fn hnswb200_bruteforce_topk(ctx: *mut hnswb200_ctx, base: *const hnswb200_points, queries: *const f32, nq: u64,
                                    k: u32, id_offset: u32, out_ids: *mut u32, out_dists: *mut f32) -> c_int;
    pub fn hnswb200_bruteforce_topk_dev(ctx: *mut hnswb200_ctx, base: *const hnswb200_points, d_queries: *const f32,
                                        nq: u64, k: u32, id_offset: u32, d_out_ids: *mut u32,
                                        d_out_dists: *mut f32) -> c_int;
    pub fn hnswb200_topk_merge(ctx: *mut hnswb200_ctx, ids: *const u32, dists: *const f32, g: u32, nq: u64, k: u32,
                               out_ids: *mut u32, out_dists: *mut f32) -> c_int;
    pub fn hnswb200_topk_merge_dev(ctx: *mut hnswb200_ctx, d_ids: *const u32, d_dists: *const f32, g: u32, nq: u64,
                                   k: u32, d_out_ids: *mut u32, d_out_dists: *mut f32) -> c_int;
    pub fn hnswb200_load_glove(path: *const c_char, lim: u64, out: *mut f32, cap: u64, dim_out: *mut u64) -> i64;
}


/// The process-wide default context (device 0) behind a mutex: a context serves one host thread at a time
/// (INTEGRATION.md), so every safe wrapper goes through `engine::with_ctx`.
pub mod engine {
    use super::*;
    use std::ffi::CStr;
    use std::sync::{Mutex, OnceLock};

    struct Ctx(*mut hnswb200_ctx);
    unsafe impl Send for Ctx {}

    static DEFAULT: OnceLock<Mutex<Ctx>> = OnceLock::new();

    pub fn last_error() -> String {
        unsafe { CStr::from_ptr(hnswb200_last_error()).to_string_lossy().into_owned() }
    }
    pub fn check(rc: c_int) -> Result<(), String> {
        if rc == HNSWB200_OK { Ok(()) } else { Err(last_error()) }
    }
    /// Runs `f` with the default context locked.  Panics when there is no CUDA device: the engine has no CPU fallback.
    pub fn with_ctx<R>(f: impl FnOnce(*mut hnswb200_ctx) -> R) -> R {
        let m = DEFAULT.get_or_init(|| {
            let mut c = std::ptr::null_mut();
            check(unsafe { hnswb200_ctx_create(0, &mut c) }).expect("no CUDA device: this engine has no CPU fallback");
            Mutex::new(Ctx(c))
        });
        let g = m.lock().unwrap_or_else(|e| e.into_inner());
        f(g.0)
    }
}
