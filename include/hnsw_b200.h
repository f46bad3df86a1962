/* hnsw_b200.h -- C ABI of libhnsw_b200.so, the B200 (sm_100a) engine for the search /
 * distance / build hot path of the Rust workspace Gumo-A/hnsw_rs.
 *
 * The reference has no FFI boundary of its own; its seams are Rust
 * traits and HNSW's public methods.  Each entry point below names the reference
 * item (file:line, relative to the reference repository) whose work it replaces;
 * INTEGRATION.md shows the Rust `extern "C"` block that binds them.
 *
 * Conventions
 *  - every function returns 0 on success or a negative HNSWB200_E* code;
 *    hnswb200_last_error() returns a thread-local message mirroring the reference's
 *    Err(String) / panic text.  Nothing throws across the boundary.
 *  - pointers are caller-owned HOST buffers unless the function name ends in _dev
 *    (then they are device pointers on the context's device and the call is
 *    asynchronous on the context's stream).
 *  - ids are u32 (graph/src/lib.rs:1 `type NodeID = u32`) and must be < 2^31.
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef HNSW_B200_H
#define HNSW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HNSWB200_OK 0
#define HNSWB200_EINVAL (-1)   /* bad argument (dimension mismatch, NaN in a vector, ...) */
#define HNSWB200_ECUDA (-2)    /* CUDA runtime error, text in last_error */
#define HNSWB200_EIO (-3)      /* file problem in save/load */
#define HNSWB200_ENOMEM (-4)
#define HNSWB200_ESTATE (-5)   /* object not in a state that allows the call */

#define HNSWB200_NO_ID 0xFFFFFFFFu /* padding id in result arrays */

/* Metric of a points object / index.  The reference only has Euclidean L2 (vectors/src/quant.rs:14-37); COSINE is an
 * addition: every stored row and every query is L2-normalised on the device before it is quantised, and the reference's
 * L2 path runs unchanged on the unit vectors (|a-b|^2 = 2 - 2 cos(a,b): the same ranking).  Reported distances are those
 * L2 distances.  The metric is not part of the reference's save format: set it again after hnswb200_index_load_dir. */
#define HNSWB200_METRIC_L2 0
#define HNSWB200_METRIC_COSINE 1
/* the reference's `type VecType` (points/src/point.rs:4): what a Point stores and which distance it evaluates */
#define HNSWB200_VEC_QUANT 0 /* QuantVec: u8 codes + min + delta, distance_unrolled (vectors/src/quant.rs:14-37); the default */
#define HNSWB200_VEC_FULL 1  /* FullVec: the f32 values, strictly sequential sum (vectors/src/full.rs:23-29) */

typedef struct hnswb200_ctx hnswb200_ctx;
typedef struct hnswb200_points hnswb200_points;
typedef struct hnswb200_graph hnswb200_graph;
typedef struct hnswb200_index hnswb200_index;

/* hnsw/src/params.rs:4-12 (Params) */
typedef struct hnswb200_params {
    uint32_t ep;
    uint64_t m, mmax, mmax0;
    float ml;
    uint64_t ef_cons, dim;
} hnswb200_params;

/* per-batch search statistics (all optional, length nq) */
typedef struct hnswb200_search_stats {
    uint32_t* hops;   /* expansions = iterations of the loop at searcher.rs:36-95 that did not break */
    uint32_t* evals;  /* dist2other calls: 1 (entry point) + every unvisited neighbour */
    uint32_t* flags;  /* bit0: NaN in query (reference panics); bit1: the visited set could not record an id, answers stay
                         exact through list membership but `evals` may over-count for this query; bit2: the exact spill
                         list of the visited set was used (everything exact) */
    uint32_t* nbrs;   /* neighbour ids read = sum of degrees of the expanded nodes (roofline accounting) */
} hnswb200_search_stats;

const char* hnswb200_last_error(void);
int hnswb200_version(void);

/* ---- context: one per GPU; owns a stream and scratch ---- */
int hnswb200_ctx_create(int device, hnswb200_ctx** out);
void hnswb200_ctx_destroy(hnswb200_ctx* ctx);
/* run all later work of this context on an existing cudaStream_t (e.g. torch's current stream) */
int hnswb200_ctx_set_stream(hnswb200_ctx* ctx, void* cuda_stream);
/* Consecutive searches of one context may overlap on the device: a search is launched as the programmatic dependent
 * of the search enqueued directly before it, so its blocks take over the SMs the previous batch's last queries leave
 * idle (a search reads nothing a previous search writes).  On the context's own stream the library knows what precedes a
 * search and does this by itself.  On an ADOPTED stream (hnswb200_ctx_set_stream) the caller may enqueue work the
 * library cannot see, so overlap is off until the caller opts in with allow = 1 and thereby promises that the query
 * buffer of a search is complete before the PREVIOUS search of this context was enqueued (no producer kernel between two
 * searches).  No reference analogue (ann_by_vector is one synchronous call, hnsw/src/template.rs:306-335). */
int hnswb200_ctx_set_overlap(hnswb200_ctx* ctx, int allow);
int hnswb200_ctx_sync(hnswb200_ctx* ctx);
int hnswb200_ctx_device(const hnswb200_ctx* ctx);
/* HNSWB200_VEC_*: the reference chooses its vector type at compile time (`type VecType = QuantVec;`,
 * points/src/point.rs:4); here it is a property of the context.  Point sets created through the context afterwards
 * (points_from_f32, build, an index loaded from a directory keeps the type its file holds) store that type;
 * existing point sets keep theirs, and every distance, search, build and brute-force entry point follows the type of
 * the points it is given.  FullVec points must be finite (a NaN distance makes the reference panic). */
int hnswb200_ctx_set_vec_type(hnswb200_ctx* ctx, int vec_type);
int hnswb200_ctx_vec_type(const hnswb200_ctx* ctx);

/* Params::from_m / from_m_efcons (params.rs:19-44); ef_cons < 0 -> 2*m */
void hnswb200_params_default(uint64_t m, int64_t ef_cons, uint64_t dim, hnswb200_params* out);

/* ---- vectors crate ---- */
/* QuantVec::new for n rows (vectors/src/quant.rs:41-66). rows[n*dim] -> codes[n*dim], mins[n], deltas[n] */
int hnswb200_quantise(hnswb200_ctx* ctx, const float* rows, uint64_t n, uint32_t dim, uint8_t* codes,
                      float* mins, float* deltas);
/* rows[i] / |rows[i]| with a strictly sequential f32 sum of squares (the arithmetic style of FullVec::distance,
 * vectors/src/full.rs:23-29); a zero row stays zero.  What a COSINE index applies to rows and queries. */
int hnswb200_normalise(hnswb200_ctx* ctx, const float* rows, uint64_t n, uint32_t dim, float* out);
/* FullVec::distance for n pairs of f32 vectors (vectors/src/full.rs:23-29): out[i] = d(x[i], y[i]) */
int hnswb200_dist_full_pairs(hnswb200_ctx* ctx, const float* x, const float* y, uint64_t n, uint32_t dim,
                             float* out);

/* ---- points crate: SimplePoints (points/src/points.rs:33-116), device resident ---- */
/* from already quantised parts (levels may be NULL = all 0) */
int hnswb200_points_upload(hnswb200_ctx* ctx, const uint8_t* codes, const float* mins, const float* deltas,
                           const uint8_t* levels, uint64_t n, uint32_t dim, hnswb200_points** out);
/* SimplePoints::new without the level draw: VecType::new of every row on the device (quantised, or kept as f32 when the
 * context's vector type is HNSWB200_VEC_FULL), device resident */
int hnswb200_points_from_f32(hnswb200_ctx* ctx, const float* rows, uint64_t n, uint32_t dim,
                             const uint8_t* levels, hnswb200_points** out);
int hnswb200_points_download(hnswb200_ctx* ctx, const hnswb200_points* p, uint8_t* codes, float* mins,
                             float* deltas, uint8_t* levels);
/* FullVec points (VecType = FullVec, vectors/src/full.rs:18-22) whatever the context's type (levels may be NULL) */
int hnswb200_points_upload_f32(hnswb200_ctx* ctx, const float* rows, const uint8_t* levels, uint64_t n, uint32_t dim,
                               hnswb200_points** out);
/* VecBase::get_vals of every point (vectors/src/lib.rs:24-26): rows[n*dim] = the stored f32 values of FullVec
 * points, the dequantised values of QuantVec points; rows and levels may be NULL */
int hnswb200_points_values(hnswb200_ctx* ctx, const hnswb200_points* p, float* rows, uint8_t* levels);
int hnswb200_points_vec_type(const hnswb200_points* p);
/* HNSWB200_METRIC_*; may only change while the object holds no points */
int hnswb200_points_set_metric(hnswb200_points* p, int metric);
int hnswb200_points_metric(const hnswb200_points* p);
uint64_t hnswb200_points_len(const hnswb200_points* p);
uint32_t hnswb200_points_dim(const hnswb200_points* p);
void hnswb200_points_destroy(hnswb200_points* p);
/* Points::distance(a[i], b[i]) (points.rs:86-93) */
int hnswb200_dist_pairs(hnswb200_ctx* ctx, const hnswb200_points* p, const uint32_t* a, const uint32_t* b,
                        uint64_t n, float* out);
/* Points::distance2point / VecBase::dist2many (points.rs:95-101, vectors/src/lib.rs:17-22):
 * the f32 query becomes a Point like a stored one (Point::new: quantised for QuantVec points, as it is for FullVec
 * points) and is compared with ids[n] */
int hnswb200_dist_query_many(hnswb200_ctx* ctx, const hnswb200_points* p, const float* query,
                             const uint32_t* ids, uint64_t n, float* out);

/* ---- graph crate: Layers / Graph (graph/src/layers.rs, graph.rs) as per-layer CSR ---- */
/* layer l: node_ids[l][n_nodes[l]], offsets[l][n_nodes[l]+1], nbrs[l][...]; caps[l] = Graph.m */
int hnswb200_graph_upload(hnswb200_ctx* ctx, uint64_t n_points, uint32_t n_layers, const uint32_t* caps,
                          const uint64_t* n_nodes, const uint32_t* const* node_ids,
                          const uint64_t* const* offsets, const uint32_t* const* nbrs,
                          hnswb200_graph** out);
uint32_t hnswb200_graph_nb_layers(const hnswb200_graph* g);
uint64_t hnswb200_graph_layer_nb_nodes(const hnswb200_graph* g, uint32_t layer);
uint64_t hnswb200_graph_layer_nb_edges(const hnswb200_graph* g, uint32_t layer); /* sum of degrees */
uint32_t hnswb200_graph_layer_cap(const hnswb200_graph* g, uint32_t layer);
/* CSR export, rows in ascending node id, neighbours ascending */
int hnswb200_graph_export_layer(const hnswb200_graph* g, uint32_t layer, uint32_t* node_ids,
                                uint64_t* offsets, uint32_t* nbrs);
void hnswb200_graph_destroy(hnswb200_graph* g);

/* ---- hnsw crate ---- */
/* takes ownership of points and graph */
int hnswb200_index_from_parts(hnswb200_ctx* ctx, hnswb200_points* points, hnswb200_graph* graph,
                              const hnswb200_params* params, hnswb200_index** out);
/* HNSW::new + insert_bulk (hnsw/src/template.rs:133-144, 388-444): VecType::new (the context's vector type), draw levels, build.
 * levels may be NULL (drawn like points.rs:148-160 from the library's own seeded generator).
 * batch = max points inserted concurrently against one frozen graph snapshot (1 = the
 * reference's single-thread order; 0 = library default). */
int hnswb200_build(hnswb200_ctx* ctx, const float* rows, uint64_t n, uint32_t dim,
                   const hnswb200_params* params, const uint8_t* levels, uint32_t batch,
                   hnswb200_index** out);
/* HNSW::insert_bulk on an existing index (template.rs:493-504) / HNSW::insert_vec (template.rs:165-173) */
int hnswb200_index_insert_bulk(hnswb200_ctx* ctx, hnswb200_index* ix, const float* rows, uint64_t n,
                               uint32_t dim, const uint8_t* levels, uint32_t batch);
int hnswb200_index_insert_vec(hnswb200_ctx* ctx, hnswb200_index* ix, const float* row, uint32_t dim,
                              uint32_t* id_out);
/* HNSW::save / HNSW::load in the reference's byte formats (template.rs:43-131) */
int hnswb200_index_save_dir(hnswb200_ctx* ctx, const hnswb200_index* ix, const char* dir);
int hnswb200_index_load_dir(hnswb200_ctx* ctx, const char* dir, hnswb200_index** out);
void hnswb200_index_destroy(hnswb200_index* ix);
int hnswb200_index_params(const hnswb200_index* ix, hnswb200_params* out);
int hnswb200_index_set_metric(hnswb200_index* ix, int metric); /* empty index only (after hnswb200_build with n = 0) */
int hnswb200_index_metric(const hnswb200_index* ix);
uint64_t hnswb200_index_len(const hnswb200_index* ix);                 /* HNSW::len */
const hnswb200_points* hnswb200_index_points(const hnswb200_index* ix);
const hnswb200_graph* hnswb200_index_graph(const hnswb200_index* ix);

/* HNSW::ann_by_vector for a batch (template.rs:306-335 + searcher.rs:23-103).
 * queries[nq*dim] f32; out_ids[nq*n] (HNSWB200_NO_ID padded), out_dists[nq*n] (+inf padded, may be NULL),
 * out_counts[nq] = number of valid ids (< n when ef < n, results.rs:59-61; may be NULL). */
int hnswb200_search(hnswb200_ctx* ctx, const hnswb200_index* ix, const float* queries, uint64_t nq,
                    uint32_t dim, uint32_t n, uint32_t ef, uint32_t* out_ids, float* out_dists,
                    uint32_t* out_counts, const hnswb200_search_stats* stats);
/* The host-buffer search without waiting: every buffer must be page-locked; the kernel reads the queries and writes
 * the results in place over PCIe, the call only launches.  Consecutive calls overlap on the device.  Results (and a
 * NaN-query error) are delivered by hnswb200_ctx_sync. */
int hnswb200_search_async(hnswb200_ctx* ctx, const hnswb200_index* ix, const float* queries, uint64_t nq,
                          uint32_t dim, uint32_t n, uint32_t ef, uint32_t* out_ids, float* out_dists,
                          uint32_t* out_counts);
/* same with device buffers, asynchronous on the context stream */
int hnswb200_search_dev(hnswb200_ctx* ctx, const hnswb200_index* ix, const float* d_queries, uint64_t nq,
                        uint32_t n, uint32_t ef, uint32_t* d_out_ids, float* d_out_dists,
                        uint32_t* d_out_counts, uint32_t* d_hops, uint32_t* d_evals, uint32_t* d_flags,
                        uint32_t* d_nbrs);

/* name of the kernel variant the last search launched from the calling thread ran (diagnostics / bench labels);
 * valid until the thread's next search */
const char* hnswb200_last_search_variant(void);

/* The same search with the all-gather of the results fused into it (query-sharded, replicated index): the id row of
 * query q is also stored to row row_offset + q of each of the n_peers (<= 8) buffers in peer_ids -- device pointers
 * valid on this context's device, typically the other ranks' result buffers opened with hnswb200_ipc_open, so the
 * stores travel over NVLink / NVSwitch while the other queries keep computing.  No reference analogue. */
int hnswb200_search_dev_gather(hnswb200_ctx* ctx, const hnswb200_index* ix, const float* d_queries, uint64_t nq,
                               uint32_t n, uint32_t ef, uint32_t* d_out_ids, float* d_out_dists,
                               uint32_t* d_out_counts, uint32_t n_peers, uint32_t* const* peer_ids,
                               uint64_t row_offset);
/* Base-sharded index (one HNSW per GPU over its slice of the base; the reference's per-shard semantics are those of
 * ann_by_vector, template.rs:306-335): the same search, with id_offset added to every returned id (global ids) and the
 * (id, distance) rows of query q also stored to row row_offset + q of the peers' gather buffers [G][nq][n] -- typically
 * row_offset = rank * nq -- so that hnswb200_topk_merge_dev can merge them once every rank has signalled. */
int hnswb200_search_dev_shard(hnswb200_ctx* ctx, const hnswb200_index* ix, const float* d_queries, uint64_t nq,
                              uint32_t n, uint32_t ef, uint32_t id_offset, uint32_t* d_out_ids, float* d_out_dists,
                              uint32_t* d_out_counts, uint32_t n_peers, uint32_t* const* peer_ids,
                              float* const* peer_dists, uint64_t row_offset);
/* Peer exchange over NVLink / NVSwitch without a collective library, asynchronous on the context stream:
 *  peer_put    copy `bytes` (multiple of 16) from d_src to each of the n_peers (<= 8) device-visible destinations;
 *  peer_signal after everything enqueued before it on this stream has completed, store `epoch` to word `slot` of each
 *              peer's flag array (release, system scope);
 *  peer_wait   hold the stream until the n_slots (<= 32) words of d_flags are all >= epoch (acquire, system scope;
 *              wrap-safe compare).  A wait of more than ~10 s raises an error that hnswb200_ctx_sync reports.
 * Flag arrays and gather buffers are hnswb200_dev_alloc memory opened on the peers with hnswb200_ipc_open. */
int hnswb200_peer_put_dev(hnswb200_ctx* ctx, const void* d_src, uint64_t bytes, uint32_t n_peers, void* const* peer_dst);
int hnswb200_peer_signal_dev(hnswb200_ctx* ctx, uint32_t n_peers, uint32_t* const* peer_flags, uint32_t slot,
                             uint32_t epoch);
int hnswb200_peer_wait_dev(hnswb200_ctx* ctx, const uint32_t* d_flags, uint32_t n_slots, uint32_t epoch);
/* device buffers that the other processes of the box can write: plain cudaMalloc memory (filled with 0xFF) and its
 * CUDA IPC handle (64 bytes); ipc_open maps a peer's buffer into this context's device with peer access enabled */
int hnswb200_dev_alloc(hnswb200_ctx* ctx, uint64_t bytes, void** out);
int hnswb200_dev_free(hnswb200_ctx* ctx, void* ptr);
int hnswb200_dev_download(hnswb200_ctx* ctx, const void* d_src, void* host_dst, uint64_t bytes);
int hnswb200_ipc_export(hnswb200_ctx* ctx, void* d_ptr, uint8_t handle[64]);
int hnswb200_ipc_open(hnswb200_ctx* ctx, const uint8_t handle[64], void** out);
int hnswb200_ipc_close(hnswb200_ctx* ctx, void* ptr);

/* brute_force_nns (hnsw/src/helpers/glove.rs:73-109): exact top-k under the metric of the points (quantised, or f32) with
 * (dist, id) order.  id_offset is added to every returned id (global ids of a base shard). */
int hnswb200_bruteforce_topk(hnswb200_ctx* ctx, const hnswb200_points* base, const float* queries,
                             uint64_t nq, uint32_t k, uint32_t id_offset, uint32_t* out_ids,
                             float* out_dists);
int hnswb200_bruteforce_topk_dev(hnswb200_ctx* ctx, const hnswb200_points* base, const float* d_queries,
                                 uint64_t nq, uint32_t k, uint32_t id_offset, uint32_t* d_out_ids,
                                 float* d_out_dists);

/* merge G sorted top-k lists per query ([G][nq][k], e.g. the result of an NCCL all-gather of
 * per-shard results) into one top-k under (dist, id) order.  No reference analogue. */
int hnswb200_topk_merge(hnswb200_ctx* ctx, const uint32_t* ids, const float* dists, uint32_t G,
                        uint64_t nq, uint32_t k, uint32_t* out_ids, float* out_dists);
int hnswb200_topk_merge_dev(hnswb200_ctx* ctx, const uint32_t* d_ids, const float* d_dists, uint32_t G,
                            uint64_t nq, uint32_t k, uint32_t* d_out_ids, float* d_out_dists);

/* helpers/glove.rs:14-71 load_glove_array: `word v1 .. vd` text rows straight to f32.
 * Call with out == NULL to size: returns rows, *dim_out = values per row. */
int64_t hnswb200_load_glove(const char* path, uint64_t lim, float* out, uint64_t cap, uint64_t* dim_out);

#ifdef __cplusplus
}
#endif
#endif /* HNSW_B200_H */
