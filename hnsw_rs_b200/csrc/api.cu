// C ABI of libhnsw_b200.so (see include/hnsw_b200.h).  Host-side glue only: argument
// checks, device memory, format I/O.  All arithmetic of the hot path runs in the
// kernels of kernels.cu / builder.cu; there is deliberately no CPU fallback.
#include <dirent.h>
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <time.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "engine.h"

namespace hb {
static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return HNSWB200_ECUDA;
}
static int fail(int code, const std::string& s) {
    g_err = s;
    return code;
}
}  // namespace hb
using namespace hb;

int hnswb200_ctx::ws_reserve(size_t bytes) {
    if (bytes <= ws_bytes) return 0;
    if (d_ws) { cudaStreamSynchronize(stream); cudaFree(d_ws); d_ws = nullptr; ws_bytes = 0; }
    size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&d_ws, want);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(workspace)");
    ws_bytes = want;
    return 0;
}

int hnswb200_ctx::norm_ws_reserve(size_t bytes) {
    if (bytes <= norm_ws_bytes) return 0;
    if (d_norm_ws) { cudaStreamSynchronize(stream); cudaFree(d_norm_ws); d_norm_ws = nullptr; norm_ws_bytes = 0; }
    cudaError_t e = cudaMalloc(&d_norm_ws, bytes + bytes / 4);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(normalised-query workspace)");
    norm_ws_bytes = bytes + bytes / 4;
    return 0;
}

int hnswb200_ctx::bf_ws_reserve(size_t bytes) {
    if (bytes <= bf_ws_bytes) return 0;
    if (d_bf_ws) { cudaStreamSynchronize(stream); cudaFree(d_bf_ws); d_bf_ws = nullptr; bf_ws_bytes = 0; }
    cudaError_t e = cudaMalloc(&d_bf_ws, bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(brute-force workspace)");
    bf_ws_bytes = bytes;
    return 0;
}

int hnswb200_ctx::use() const {
    last_was_search = false;  // whatever the caller enqueues next is not known to be a search (the search entry points re-arm it)
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return 0;
}

extern "C" {

const char* hnswb200_last_error(void) { return g_err.c_str(); }
int hnswb200_version(void) { return 100; }

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
int hnswb200_ctx_create(int device, hnswb200_ctx** out) {
    if (!out) return fail(HNSWB200_EINVAL, "ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(HNSWB200_ECUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                        cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(HNSWB200_EINVAL, "ctx_create: bad device index");
    hnswb200_ctx* c = new hnswb200_ctx();
    c->device = device;
    // any failure below releases what was allocated so far (hnswb200_ctx_destroy copes with a half-built context)
    auto init = [&]() -> int {
        HB_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        HB_CUDA(cudaGetDeviceProperties(&prop, device));
        c->num_sms = prop.multiProcessorCount;
        HB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
        HB_CUDA(cudaMalloc((void**)&c->d_scratch, 64 * sizeof(uint32_t)));
        HB_CUDA(cudaMemset(c->d_scratch, 0, 64 * sizeof(uint32_t)));
        HB_CUDA(cudaHostAlloc((void**)&c->h_status, 64, cudaHostAllocMapped));
        HB_CUDA(cudaHostGetDevicePointer((void**)&c->d_status, c->h_status, 0));
        c->h_status[0] = 0;
        HB_CUDA(cudaMalloc((void**)&c->d_counters, hnswb200_ctx::COUNTER_RING * sizeof(uint32_t)));
        HB_CUDA(cudaMemset(c->d_counters, 0, hnswb200_ctx::COUNTER_RING * sizeof(uint32_t)));
        return 0;
    };
    const int rc = init();
    if (rc) {
        hnswb200_ctx_destroy(c);
        return rc;
    }
    *out = c;
    return 0;
}

void hnswb200_ctx_destroy(hnswb200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    if (c->d_scratch) cudaFree(c->d_scratch);
    if (c->d_counters) cudaFree(c->d_counters);
    if (c->d_spill_ws) cudaFree(c->d_spill_ws);
    if (c->h_status) cudaFreeHost(c->h_status);
    if (c->d_ws) cudaFree(c->d_ws);
    if (c->d_bf_ws) cudaFree(c->d_bf_ws);
    if (c->d_norm_ws) cudaFree(c->d_norm_ws);
    delete c;
}

int hnswb200_ctx_set_stream(hnswb200_ctx* c, void* s) {
    if (!c) return fail(HNSWB200_EINVAL, "ctx is NULL");
    if (c->use()) return HNSWB200_ECUDA;
    // work on the old stream may still use the counter ring and the workspaces: drain it, and make the next search
    // start a fresh ring (its memset is ordered on the new stream)
    cudaStreamSynchronize(c->stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->search_seq = 0;
    c->stream = (cudaStream_t)s;
    c->own_stream = false;
    c->stream_adopted = true;
    c->last_was_search = false;
    return 0;
}

int hnswb200_ctx_set_overlap(hnswb200_ctx* c, int allow) {
    if (!c) return fail(HNSWB200_EINVAL, "ctx is NULL");
    c->overlap_opt_in = allow != 0;
    c->last_was_search = false;
    return 0;
}

const char* hnswb200_last_search_variant(void) { return hb::last_search_variant(); }

int hnswb200_ctx_sync(hnswb200_ctx* c) {
    if (!c) return fail(HNSWB200_EINVAL, "ctx is NULL");
    if (c->use()) return HNSWB200_ECUDA;
    HB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->h_status && c->h_status[0]) {  // raised by an asynchronous search (hnswb200_search_async)
        c->h_status[0] = 0;
        return fail(HNSWB200_EINVAL, "search: NaN in a query (the reference panics in partial_cmp().unwrap())");
    }
    return 0;
}
int hnswb200_ctx_device(const hnswb200_ctx* c) { return c ? c->device : -1; }

// `type VecType = QuantVec;` (points/src/point.rs:4) as a property of the context: every point set created through it
// afterwards (points_from_f32, build, insert) stores that vector type.  Existing point sets keep theirs.
int hnswb200_ctx_set_vec_type(hnswb200_ctx* c, int vec_type) {
    if (!c) return fail(HNSWB200_EINVAL, "ctx is NULL");
    if (vec_type != HNSWB200_VEC_QUANT && vec_type != HNSWB200_VEC_FULL) return fail(HNSWB200_EINVAL, "unknown vector type");
    c->vec_type = vec_type;
    return 0;
}
int hnswb200_ctx_vec_type(const hnswb200_ctx* c) { return c ? c->vec_type : -1; }

void hnswb200_params_default(uint64_t m, int64_t ef_cons, uint64_t dim, hnswb200_params* p) {
    p->ep = 0;
    p->m = m;
    p->mmax = m;
    p->mmax0 = 2 * m;
    p->ml = 1.0f / logf((float)m);  // params.rs:15-17
    p->ef_cons = ef_cons >= 0 ? (uint64_t)ef_cons : 2 * m;
    p->dim = dim;
}

// ---------------------------------------------------------------------------
// vectors
// ---------------------------------------------------------------------------
int hnswb200_quantise(hnswb200_ctx* c, const float* rows, uint64_t n, uint32_t dim, uint8_t* codes,
                      float* mins, float* deltas) {
    if (!c || !rows || !codes || !mins || !deltas) return fail(HNSWB200_EINVAL, "quantise: NULL argument");
    if (dim == 0) return fail(HNSWB200_EINVAL, "quantise: cannot quantise an empty vector");
    if (n == 0) return 0;
    if (c->use()) return HNSWB200_ECUDA;
    RecLayout L = hb_make_layout(dim);
    DevBuf<float> d_rows, d_mins, d_deltas;
    DevBuf<uint8_t> d_codes;
    HB_CUDA(d_rows.alloc(n * dim));
    HB_CUDA(d_codes.alloc(n * dim));
    HB_CUDA(d_mins.alloc(n));
    HB_CUDA(d_deltas.alloc(n));
    uint32_t* nan_flag = c->d_scratch + 1;
    HB_CUDA(cudaMemsetAsync(nan_flag, 0, 4, c->stream));
    HB_CUDA(cudaMemcpyAsync(d_rows.p, rows, n * dim * 4, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(launch_quantise(d_rows.p, n, L, nullptr, d_codes.p, d_mins.p, d_deltas.p, nan_flag, c->stream));
    HB_CUDA(cudaMemcpyAsync(codes, d_codes.p, n * dim, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaMemcpyAsync(mins, d_mins.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaMemcpyAsync(deltas, d_deltas.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    uint32_t flag = 0;
    HB_CUDA(cudaMemcpyAsync(&flag, nan_flag, 4, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    if (flag) return fail(HNSWB200_EINVAL, "quantise: NaN in vector (the reference panics in partial_cmp().unwrap())");
    return 0;
}

int hnswb200_normalise(hnswb200_ctx* c, const float* rows, uint64_t n, uint32_t dim, float* out) {
    if (!c || (n && (!rows || !out)) || dim == 0) return fail(HNSWB200_EINVAL, "normalise: bad argument");
    if (n == 0) return 0;
    if (c->use()) return HNSWB200_ECUDA;
    DevBuf<float> d;
    HB_CUDA(d.alloc(n * dim));
    HB_CUDA(cudaMemcpyAsync(d.p, rows, n * dim * 4, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(launch_normalise(d.p, n, dim, d.p, c->stream));
    HB_CUDA(cudaMemcpyAsync(out, d.p, n * dim * 4, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int hnswb200_points_set_metric(hnswb200_points* p, int metric) {
    if (!p) return fail(HNSWB200_EINVAL, "points_set_metric: NULL argument");
    if (metric != HNSWB200_METRIC_L2 && metric != HNSWB200_METRIC_COSINE) return fail(HNSWB200_EINVAL, "unknown metric");
    if (p->n != 0 && metric != p->metric)
        return fail(HNSWB200_ESTATE, "the metric can only be chosen while the points object is empty (stored rows are already quantised)");
    p->metric = metric;
    return 0;
}
int hnswb200_points_metric(const hnswb200_points* p) { return p ? p->metric : -1; }

int hnswb200_dist_full_pairs(hnswb200_ctx* c, const float* x, const float* y, uint64_t n, uint32_t dim,
                             float* out) {
    if (!c || !x || !y || !out) return fail(HNSWB200_EINVAL, "dist_full_pairs: NULL argument");
    if (n == 0) return 0;
    if (c->use()) return HNSWB200_ECUDA;
    DevBuf<float> dx, dy, dout;
    HB_CUDA(dx.alloc(n * dim));
    HB_CUDA(dy.alloc(n * dim));
    HB_CUDA(dout.alloc(n));
    HB_CUDA(cudaMemcpyAsync(dx.p, x, n * dim * 4, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(cudaMemcpyAsync(dy.p, y, n * dim * 4, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(launch_dist_full_pairs(dx.p, dy.p, n, dim, dout.p, c->stream));
    HB_CUDA(cudaMemcpyAsync(out, dout.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------------------
// points
// ---------------------------------------------------------------------------
}  // extern "C"

int hnswb200_points::reserve(uint64_t want) {
    if (want <= cap) return 0;
    uint64_t ncap = std::max<uint64_t>(want, cap + cap / 2);
    uint8_t* nrec = nullptr;
    HB_CUDA(cudaMalloc((void**)&nrec, std::max<uint64_t>(ncap, 1) * L.stride));
    HB_CUDA(cudaMemsetAsync(nrec, 0, ncap * L.stride, ctx->stream));
    if (d_rec && n) HB_CUDA(cudaMemcpyAsync(nrec, d_rec, n * L.stride, cudaMemcpyDeviceToDevice, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (d_rec) cudaFree(d_rec);
    d_rec = nrec;
    cap = ncap;
    return 0;
}

static int points_new(hnswb200_ctx* c, uint32_t dim, uint64_t n, hnswb200_points** out, int vec_type = -1) {
    if (dim == 0) return fail(HNSWB200_EINVAL, "points: dimension 0");
    if (n >= (1ull << 31)) return fail(HNSWB200_EINVAL, "points: more than 2^31-1 points");
    hnswb200_points* p = new hnswb200_points();
    p->ctx = c;
    if (vec_type < 0) vec_type = c->vec_type;
    p->L = hb_make_layout_kind(dim, vec_type == HNSWB200_VEC_FULL ? HB_REC_F32 : HB_REC_QUANT);
    int rc = p->reserve(n);
    if (rc) { delete p; return rc; }
    *out = p;
    return 0;
}

// append n rows, quantised on the device
int hb::points_append_f32(hnswb200_ctx* c, hnswb200_points* p, const float* rows, uint64_t n,
                          const uint8_t* levels) {
    if (n == 0) return 0;
    if (p->n + n >= (1ull << 31)) return fail(HNSWB200_EINVAL, "points: more than 2^31-1 points");
    int rc = p->reserve(p->n + n);
    if (rc) return rc;
    uint32_t* nan_flag = c->d_scratch + 1;
    HB_CUDA(cudaMemsetAsync(nan_flag, 0, 4, c->stream));
    const uint64_t CH = 1u << 20;  // stage through a bounded device buffer
    DevBuf<float> d_rows;
    HB_CUDA(d_rows.alloc(std::min(n, CH) * p->L.dim));
    for (uint64_t s = 0; s < n; s += CH) {
        uint64_t cnt = std::min(CH, n - s);
        HB_CUDA(cudaMemcpyAsync(d_rows.p, rows + s * p->L.dim, cnt * p->L.dim * 4, cudaMemcpyHostToDevice, c->stream));
        if (p->metric == HNSWB200_METRIC_COSINE) HB_CUDA(launch_normalise(d_rows.p, cnt, p->L.dim, d_rows.p, c->stream));
        if (p->L.kind == HB_REC_F32)
            HB_CUDA(launch_pack_f32(d_rows.p, cnt, p->L, p->d_rec + (p->n + s) * p->L.stride, nan_flag, c->stream));
        else
            HB_CUDA(launch_quantise(d_rows.p, cnt, p->L, p->d_rec + (p->n + s) * p->L.stride, nullptr, nullptr,
                                    nullptr, nan_flag, c->stream));
    }
    uint32_t flag = 0;
    HB_CUDA(cudaMemcpyAsync(&flag, nan_flag, 4, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    if (flag)
        return fail(HNSWB200_EINVAL, p->L.kind == HB_REC_F32
                                         ? "non-finite value in vector (a NaN distance makes the reference panic in partial_cmp().unwrap())"
                                         : "NaN in vector (the reference panics in partial_cmp().unwrap())");
    for (uint64_t i = 0; i < n; ++i) p->levels.push_back(levels ? levels[i] : 0);
    p->n += n;
    return 0;
}

extern "C" {

int hnswb200_points_upload(hnswb200_ctx* c, const uint8_t* codes, const float* mins, const float* deltas,
                           const uint8_t* levels, uint64_t n, uint32_t dim, hnswb200_points** out) {
    if (!c || !out || (n && (!codes || !mins || !deltas))) return fail(HNSWB200_EINVAL, "points_upload: NULL argument");
    if (c->use()) return HNSWB200_ECUDA;
    hnswb200_points* p = nullptr;
    int rc = points_new(c, dim, n, &p, HNSWB200_VEC_QUANT);  // codes, min, delta: a QuantVec set whatever the context's type
    if (rc) return rc;
    if (n) {
        DevBuf<uint8_t> d_codes;
        DevBuf<float> d_mins, d_deltas;
        cudaError_t e;
        if ((e = d_codes.alloc(n * dim)) != cudaSuccess || (e = d_mins.alloc(n)) != cudaSuccess ||
            (e = d_deltas.alloc(n)) != cudaSuccess) { hnswb200_points_destroy(p); return cuda_fail(e, "cudaMalloc"); }
        cudaMemcpyAsync(d_codes.p, codes, n * dim, cudaMemcpyHostToDevice, c->stream);
        cudaMemcpyAsync(d_mins.p, mins, n * 4, cudaMemcpyHostToDevice, c->stream);
        cudaMemcpyAsync(d_deltas.p, deltas, n * 4, cudaMemcpyHostToDevice, c->stream);
        launch_pack(d_codes.p, d_mins.p, d_deltas.p, n, p->L, p->d_rec, c->stream);
        e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { hnswb200_points_destroy(p); return cuda_fail(e, "points_upload"); }
    }
    p->levels.assign(n, 0);
    if (levels) memcpy(p->levels.data(), levels, n);
    p->n = n;
    *out = p;
    return 0;
}

int hnswb200_points_from_f32(hnswb200_ctx* c, const float* rows, uint64_t n, uint32_t dim,
                             const uint8_t* levels, hnswb200_points** out) {
    if (!c || !out || (n && !rows)) return fail(HNSWB200_EINVAL, "points_from_f32: NULL argument");
    if (c->use()) return HNSWB200_ECUDA;
    hnswb200_points* p = nullptr;
    int rc = points_new(c, dim, n, &p);
    if (rc) return rc;
    rc = points_append_f32(c, p, rows, n, levels);
    if (rc) { hnswb200_points_destroy(p); return rc; }
    *out = p;
    return 0;
}

int hnswb200_points_download(hnswb200_ctx* c, const hnswb200_points* p, uint8_t* codes, float* mins,
                             float* deltas, uint8_t* levels) {
    if (!c || !p) return fail(HNSWB200_EINVAL, "points_download: NULL argument");
    if (c->use()) return HNSWB200_ECUDA;
    uint64_t n = p->n;
    if (levels && n) memcpy(levels, p->levels.data(), n);
    if (n == 0 || (!codes && !mins && !deltas)) return 0;
    if (p->L.kind != HB_REC_QUANT)
        return fail(HNSWB200_ESTATE, "points_download: these points store f32 vectors (FullVec); use hnswb200_points_values");
    DevBuf<uint8_t> d_codes;
    DevBuf<float> d_mins, d_deltas;
    HB_CUDA(d_codes.alloc(n * p->L.dim));
    HB_CUDA(d_mins.alloc(n));
    HB_CUDA(d_deltas.alloc(n));
    HB_CUDA(launch_unpack(p->d_rec, n, p->L, d_codes.p, d_mins.p, d_deltas.p, c->stream));
    if (codes) HB_CUDA(cudaMemcpyAsync(codes, d_codes.p, n * p->L.dim, cudaMemcpyDeviceToHost, c->stream));
    if (mins) HB_CUDA(cudaMemcpyAsync(mins, d_mins.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (deltas) HB_CUDA(cudaMemcpyAsync(deltas, d_deltas.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int hnswb200_points_upload_f32(hnswb200_ctx* c, const float* rows, const uint8_t* levels, uint64_t n, uint32_t dim,
                               hnswb200_points** out) {
    if (!c || !out || (n && !rows)) return fail(HNSWB200_EINVAL, "points_upload_f32: NULL argument");
    if (c->use()) return HNSWB200_ECUDA;
    hnswb200_points* p = nullptr;
    int rc = points_new(c, dim, n, &p, HNSWB200_VEC_FULL);
    if (rc) return rc;
    rc = points_append_f32(c, p, rows, n, levels);
    if (rc) { hnswb200_points_destroy(p); return rc; }
    *out = p;
    return 0;
}

int hnswb200_points_values(hnswb200_ctx* c, const hnswb200_points* p, float* rows, uint8_t* levels) {
    if (!c || !p) return fail(HNSWB200_EINVAL, "points_values: NULL argument");
    if (c->use()) return HNSWB200_ECUDA;
    const uint64_t n = p->n;
    if (levels && n) memcpy(levels, p->levels.data(), n);
    if (n == 0 || !rows) return 0;
    DevBuf<float> d_rows;
    HB_CUDA(d_rows.alloc(n * p->L.dim));
    HB_CUDA(launch_record_values(p->d_rec, n, p->L, d_rows.p, c->stream));
    HB_CUDA(cudaMemcpyAsync(rows, d_rows.p, n * p->L.dim * 4, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int hnswb200_points_vec_type(const hnswb200_points* p) {
    return p ? (p->L.kind == HB_REC_F32 ? HNSWB200_VEC_FULL : HNSWB200_VEC_QUANT) : -1;
}
uint64_t hnswb200_points_len(const hnswb200_points* p) { return p ? p->n : 0; }
uint32_t hnswb200_points_dim(const hnswb200_points* p) { return p ? p->L.dim : 0; }
void hnswb200_points_destroy(hnswb200_points* p) {
    if (!p) return;
    if (p->ctx) cudaSetDevice(p->ctx->device);
    if (p->d_rec) cudaFree(p->d_rec);
    delete p;
}

int hnswb200_dist_pairs(hnswb200_ctx* c, const hnswb200_points* p, const uint32_t* a, const uint32_t* b,
                        uint64_t n, float* out) {
    if (!c || !p || (n && (!a || !b || !out))) return fail(HNSWB200_EINVAL, "dist_pairs: NULL argument");
    if (n == 0) return 0;
    for (uint64_t i = 0; i < n; ++i)
        if (a[i] >= p->n || b[i] >= p->n) return fail(HNSWB200_EINVAL, "dist_pairs: point id out of range (Points::distance returns None)");
    if (c->use()) return HNSWB200_ECUDA;
    DevBuf<uint32_t> da, db;
    DevBuf<float> dout;
    HB_CUDA(da.alloc(n));
    HB_CUDA(db.alloc(n));
    HB_CUDA(dout.alloc(n));
    HB_CUDA(cudaMemcpyAsync(da.p, a, n * 4, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(cudaMemcpyAsync(db.p, b, n * 4, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(launch_dist_pairs(p->d_rec, p->L, da.p, db.p, n, dout.p, c->stream));
    HB_CUDA(cudaMemcpyAsync(out, dout.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int hnswb200_dist_query_many(hnswb200_ctx* c, const hnswb200_points* p, const float* query,
                             const uint32_t* ids, uint64_t n, float* out) {
    if (!c || !p || !query || (n && (!ids || !out))) return fail(HNSWB200_EINVAL, "dist_query_many: NULL argument");
    if (n == 0) return 0;
    for (uint64_t i = 0; i < n; ++i)
        if (ids[i] >= p->n) return fail(HNSWB200_EINVAL, "dist_query_many: point id out of range (distance2point returns None)");
    if (c->use()) return HNSWB200_ECUDA;
    DevBuf<uint32_t> dids;
    DevBuf<float> dq, dout;
    HB_CUDA(dids.alloc(n));
    HB_CUDA(dq.alloc(p->L.dim));
    HB_CUDA(dout.alloc(n));
    uint32_t* nan_flag = c->d_scratch + 1;
    HB_CUDA(cudaMemsetAsync(nan_flag, 0, 4, c->stream));
    HB_CUDA(cudaMemcpyAsync(dids.p, ids, n * 4, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(cudaMemcpyAsync(dq.p, query, p->L.dim * 4, cudaMemcpyHostToDevice, c->stream));
    if (p->metric == HNSWB200_METRIC_COSINE) HB_CUDA(launch_normalise(dq.p, 1, p->L.dim, dq.p, c->stream));
    HB_CUDA(launch_dist_query_many(p->d_rec, p->L, dq.p, dids.p, n, dout.p, nan_flag, c->stream));
    HB_CUDA(cudaMemcpyAsync(out, dout.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    uint32_t flag = 0;
    HB_CUDA(cudaMemcpyAsync(&flag, nan_flag, 4, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    if (flag) return fail(HNSWB200_EINVAL, "dist_query_many: NaN in query");
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// graph: device mirror of hb::HostGraph
// ---------------------------------------------------------------------------
hb::DevGraph hnswb200_graph::view() const {
    hb::DevGraph g;
    g.adj0 = d_adj0;
    g.S0 = h.a0.S;
    g.upper_off = d_upper_off;
    g.upper_adj = d_adju;
    g.SU = h.au.S;
    g.n_layers = h.n_layers();
    return g;
}

void hnswb200_graph::free_device() {
    if (ctx) cudaSetDevice(ctx->device);
    if (d_adj0) cudaFree(d_adj0);
    if (d_upper_off) cudaFree(d_upper_off);
    if (d_adju) cudaFree(d_adju);
    d_adj0 = d_upper_off = d_adju = nullptr;
    device_valid = false;
    if (h_stage_rows) cudaFreeHost(h_stage_rows);
    if (h_stage_data) cudaFreeHost(h_stage_data);
    if (d_stage_rows) cudaFree(d_stage_rows);
    if (d_stage_data) cudaFree(d_stage_data);
    h_stage_rows = h_stage_data = d_stage_rows = d_stage_data = nullptr;
    stage_cap = 0;
    stage_S = 0;
}

namespace {
// Write row `row` of store `s` in device form into dst (S slots) and, when the degree
// exceeds S, into continuation rows (the last slot of a full row is CHAIN | next_row).
// `alloc_chain` hands out continuation row indices.  Returns false if none are left.
struct RowSink {
    virtual void put(uint64_t dev_row, const uint32_t* slots) = 0;
};
bool materialise_row(const hb::AdjStore& s, uint32_t row, uint64_t chain_base, uint64_t chain_cap,
                     uint64_t& chain_used, std::unordered_map<uint32_t, std::vector<uint32_t>>& chains,
                     RowSink& sink) {
    const uint32_t S = s.S, d = s.deg[row];
    if (d <= S) {  // the host row is already in device form (unused slots are EMPTY)
        sink.put(row, &s.data[(size_t)row * S]);
        return true;
    }
    std::vector<uint32_t> slots(S, hb::H_EMPTY);
    std::vector<uint32_t>& ch = chains[row];
    uint32_t i = 0;
    uint64_t cur = row;
    size_t ci = 0;
    while (true) {
        uint32_t left = d - i;
        std::fill(slots.begin(), slots.end(), hb::H_EMPTY);
        if (left <= S) {
            for (uint32_t j = 0; j < left; ++j) slots[j] = s.get(row, i + j);
            sink.put(cur, slots.data());
            return true;
        }
        for (uint32_t j = 0; j < S - 1; ++j) slots[j] = s.get(row, i + j);
        i += S - 1;
        if (ci >= ch.size()) {
            if (chain_used >= chain_cap) return false;
            ch.push_back((uint32_t)(chain_base + chain_used++));
        }
        uint32_t nxt = ch[ci++];
        slots[S - 1] = 0x80000000u | nxt;
        sink.put(cur, slots.data());
        cur = nxt;
    }
}
struct VecSink : RowSink {
    std::vector<uint32_t>& v;
    uint32_t S;
    VecSink(std::vector<uint32_t>& v_, uint32_t S_) : v(v_), S(S_) {}
    void put(uint64_t r, const uint32_t* slots) override { memcpy(&v[r * S], slots, S * 4); }
};
struct ListSink : RowSink {
    std::vector<uint32_t>& rows;
    std::vector<uint32_t>& data;
    uint32_t S;
    ListSink(std::vector<uint32_t>& r, std::vector<uint32_t>& d, uint32_t S_) : rows(r), data(d), S(S_) {}
    void put(uint64_t r, const uint32_t* slots) override {
        rows.push_back((uint32_t)r);
        data.insert(data.end(), slots, slots + S);
    }
};
}  // namespace

int hnswb200_graph::upload_full() {
    if (ctx->use()) return HNSWB200_ECUDA;
    free_device();
    chains0.clear();
    chainsu.clear();
    chain0_used = chainu_used = 0;
    const uint64_t n = h.n_points();
    // leave head-room so appended points / new continuation rows do not force a rebuild
    // continuation rows: twice what the rows wider than S need now (a row keeps the ones it was given)
    auto chain_rows = [](const hb::AdjStore& s) {
        uint64_t c = 0;
        for (uint32_t d : s.deg)
            if (d > s.S) c += (d - 2) / (s.S - 1);
        return c;
    };
    rows0_cap = n + n / 8 + 1024;
    chain0_cap = std::max<uint64_t>(1024, chain_rows(h.a0) * 2 + n / 64);
    rowsu_cap = h.au.rows() + h.au.rows() / 8 + 1024;
    chainu_cap = std::max<uint64_t>(1024, chain_rows(h.au) * 2 + h.au.rows() / 64);
    ++n_full_uploads;
    upper_off_cap = rows0_cap;
    const uint32_t S0 = h.a0.S, SU = h.au.S;
    std::vector<uint32_t> st0((rows0_cap + chain0_cap) * S0, hb::H_EMPTY);
    {
        VecSink sink(st0, S0);
        for (uint64_t r = 0; r < n; ++r)
            if (!materialise_row(h.a0, (uint32_t)r, rows0_cap, chain0_cap, chain0_used, chains0, sink))
                return fail(HNSWB200_ENOMEM, "graph: continuation rows exhausted");
    }
    std::vector<uint32_t> stu((rowsu_cap + chainu_cap) * SU, hb::H_EMPTY);
    {
        VecSink sink(stu, SU);
        for (uint64_t r = 0; r < h.au.rows(); ++r)
            if (!materialise_row(h.au, (uint32_t)r, rowsu_cap, chainu_cap, chainu_used, chainsu, sink))
                return fail(HNSWB200_ENOMEM, "graph: continuation rows exhausted");
    }
    std::vector<uint32_t> uo(upper_off_cap, hb::H_EMPTY);
    std::copy(h.upper_off.begin(), h.upper_off.end(), uo.begin());
    HB_CUDA(cudaMalloc((void**)&d_adj0, st0.size() * 4));
    HB_CUDA(cudaMalloc((void**)&d_adju, stu.size() * 4));
    HB_CUDA(cudaMalloc((void**)&d_upper_off, uo.size() * 4));
    HB_CUDA(cudaMemcpyAsync(d_adj0, st0.data(), st0.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    HB_CUDA(cudaMemcpyAsync(d_adju, stu.data(), stu.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    HB_CUDA(cudaMemcpyAsync(d_upper_off, uo.data(), uo.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    device_valid = true;
    return 0;
}

int hnswb200_graph::sync_new_nodes(uint64_t first_new) {
    if (!device_valid || h.n_points() > rows0_cap || h.au.rows() > rowsu_cap || h.n_points() > upper_off_cap)
        return upload_full();
    if (ctx->use()) return HNSWB200_ECUDA;
    // rows past the old end are already EMPTY on the device; only upper_off is new
    uint64_t cnt = h.n_points() - first_new;
    if (cnt) {
        HB_CUDA(cudaMemcpyAsync(d_upper_off + first_new, h.upper_off.data() + first_new, cnt * 4,
                                cudaMemcpyHostToDevice, ctx->stream));
        HB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

int hnswb200_graph::upload_rows(std::vector<uint32_t>& dirty0, std::vector<uint32_t>& dirtyu) {
    if (!device_valid || h.n_points() > rows0_cap || h.au.rows() > rowsu_cap) {
        dirty0.clear();
        dirtyu.clear();
        return upload_full();
    }
    if (ctx->use()) return HNSWB200_ECUDA;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (int which = 0; which < 2; ++which) {
        std::vector<uint32_t>& dirty = which ? dirtyu : dirty0;
        if (dirty.empty()) continue;
        const double ts0 = now();
        const hb::AdjStore& s = which ? h.au : h.a0;
        const uint32_t S = s.S;
        // de-duplicate with a mark array (the list holds every touched row, often many times)
        std::vector<uint8_t>& mark = which ? marku : mark0;
        if (mark.size() < s.rows()) mark.resize(s.rows(), 0);
        size_t uniq = 0;
        for (uint32_t r : dirty)
            if (!mark[r]) { mark[r] = 1; dirty[uniq++] = r; }
        dirty.resize(uniq);
        for (uint32_t r : dirty) mark[r] = 0;
        const double ta = now();
        t_up[0] += ta - ts0;
        // stage: rows whose degree fits are already in device form on the host
        size_t need = 2 * uniq + 64;  // a row wider than S is staged as head + continuation row(s)
        const uint32_t Smax = std::max(h.a0.S, h.au.S);
        if (stage_cap < need || stage_S < Smax) {
            if (h_stage_rows) cudaFreeHost(h_stage_rows);
            if (h_stage_data) cudaFreeHost(h_stage_data);
            if (d_stage_rows) cudaFree(d_stage_rows);
            if (d_stage_data) cudaFree(d_stage_data);
            h_stage_rows = h_stage_data = d_stage_rows = d_stage_data = nullptr;
            stage_cap = std::max(stage_cap, need + need / 2);
            stage_S = Smax;
            HB_CUDA(cudaHostAlloc((void**)&h_stage_rows, stage_cap * 4, cudaHostAllocDefault));
            HB_CUDA(cudaHostAlloc((void**)&h_stage_data, stage_cap * (size_t)stage_S * 4, cudaHostAllocDefault));
            HB_CUDA(cudaMalloc((void**)&d_stage_rows, stage_cap * 4));
            HB_CUDA(cudaMalloc((void**)&d_stage_data, stage_cap * (size_t)stage_S * 4));
        }
        const double tb = now();
        t_up[1] += tb - ta;
        size_t cnt = 0;
        std::vector<uint32_t> big;  // rows that need continuation rows: rare
        for (size_t di = 0; di < dirty.size(); ++di) {
            const uint32_t r = dirty[di];
            if (di + 8 < dirty.size()) {  // the rows are scattered over the whole store: fetch ahead
                const char* nx = (const char*)&s.data[(size_t)dirty[di + 8] * S];
                for (uint32_t b = 0; b < S * 4; b += 64) __builtin_prefetch(nx + b);
                __builtin_prefetch(&s.deg[dirty[di + 8]]);
            }
            if (s.deg[r] <= S) {
                h_stage_rows[cnt] = r;
                memcpy(h_stage_data + cnt * S, &s.data[(size_t)r * S], (size_t)S * 4);
                ++cnt;
            } else {
                big.push_back(r);
            }
        }
        const double tc = now();
        t_up[2] += tc - tb;
        n_up_rows += cnt;
        n_up_big += big.size();
        if (!big.empty()) {
            std::vector<uint32_t> rows, data;
            ListSink sink(rows, data, S);
            for (uint32_t r : big) {
                bool ok = which ? materialise_row(s, r, rowsu_cap, chainu_cap, chainu_used, chainsu, sink)
                                : materialise_row(s, r, rows0_cap, chain0_cap, chain0_used, chains0, sink);
                if (!ok) {  // out of continuation rows: rebuild with more head-room
                    dirty0.clear();
                    dirtyu.clear();
                    return upload_full();
                }
            }
            if (cnt + rows.size() > stage_cap) {  // only with rows wider than 2S-1
                dirty0.clear();
                dirtyu.clear();
                return upload_full();
            }
            for (size_t i = 0; i < rows.size(); ++i) {
                h_stage_rows[cnt] = rows[i];
                memcpy(h_stage_data + cnt * S, &data[i * S], (size_t)S * 4);
                ++cnt;
            }
        }
        const double ts1 = now();
        HB_CUDA(cudaMemcpyAsync(d_stage_rows, h_stage_rows, cnt * 4, cudaMemcpyHostToDevice, ctx->stream));
        HB_CUDA(cudaMemcpyAsync(d_stage_data, h_stage_data, cnt * (size_t)S * 4, cudaMemcpyHostToDevice, ctx->stream));
        HB_CUDA(hb::launch_scatter_rows(which ? d_adju : d_adj0, S, d_stage_rows, d_stage_data, (uint32_t)cnt, ctx->stream));
        HB_CUDA(cudaStreamSynchronize(ctx->stream));
        t_stage += ts1 - ts0;
        t_xfer += now() - ts1;
        dirty.clear();
    }
    return 0;
}

extern "C" {

int hnswb200_graph_upload(hnswb200_ctx* c, uint64_t n_points, uint32_t n_layers, const uint32_t* caps,
                          const uint64_t* n_nodes, const uint32_t* const* node_ids,
                          const uint64_t* const* offsets, const uint32_t* const* nbrs,
                          hnswb200_graph** out) {
    if (!c || !out || !caps || !n_nodes || !node_ids || !offsets || !nbrs)
        return fail(HNSWB200_EINVAL, "graph_upload: NULL argument");
    if (n_layers == 0) return fail(HNSWB200_EINVAL, "graph_upload: an index has at least one layer");
    if (n_points >= (1ull << 31)) return fail(HNSWB200_EINVAL, "graph_upload: more than 2^31-1 points");
    for (uint32_t l = 2; l < n_layers; ++l)
        if (caps[l] != caps[1]) return fail(HNSWB200_EINVAL, "graph_upload: upper layers must share one cap (Layers::add_level)");
    if (n_nodes[0] != n_points) return fail(HNSWB200_EINVAL, "graph_upload: layer 0 must hold every point");
    // level of a node = highest layer that lists it; membership must be nested
    std::vector<int> lvl(n_points, -1);
    for (uint32_t l = 0; l < n_layers; ++l) {
        for (uint64_t r = 0; r < n_nodes[l]; ++r) {
            uint32_t id = node_ids[l][r];
            if (id >= n_points) return fail(HNSWB200_EINVAL, "graph_upload: node id out of range");
            if (lvl[id] != (int)l - 1) return fail(HNSWB200_EINVAL, "graph_upload: layer membership is not nested / node listed twice");
            lvl[id] = (int)l;
        }
    }
    hnswb200_graph* g = new hnswb200_graph();
    g->ctx = c;
    uint32_t capu = n_layers > 1 ? caps[1] : (caps[0] + 1) / 2;
    g->h.init(capu, caps[0], capu);
    for (uint64_t i = 0; i < n_points; ++i) g->h.add_node((uint32_t)lvl[i]);
    while (g->h.layer_nodes.size() < n_layers) g->h.layer_nodes.push_back(0);
    for (uint32_t l = 0; l < n_layers; ++l) {
        hb::AdjStore& s = g->h.store(l);
        for (uint64_t r = 0; r < n_nodes[l]; ++r) {
            uint32_t id = node_ids[l][r];
            uint32_t row = g->h.row(id, l);
            for (uint64_t e = offsets[l][r]; e < offsets[l][r + 1]; ++e) {
                uint32_t nb = nbrs[l][e];
                if (nb >= n_points || !g->h.in_layer(nb, l)) {
                    delete g;
                    return fail(HNSWB200_EINVAL, "graph_upload: neighbour is not a node of the layer (NodeNotInGraph)");
                }
                if (nb == id) { delete g; return fail(HNSWB200_EINVAL, "graph_upload: self connection"); }
                s.insert(row, nb);
            }
        }
    }
    g->h.weights_valid = false;  // edge lengths are recomputed on the device if the index is extended
    int rc = g->upload_full();
    if (rc) { hnswb200_graph_destroy(g); return rc; }
    *out = g;
    return 0;
}

uint32_t hnswb200_graph_nb_layers(const hnswb200_graph* g) { return g ? g->h.n_layers() : 0; }
uint64_t hnswb200_graph_layer_nb_nodes(const hnswb200_graph* g, uint32_t l) {
    return (g && l < g->h.n_layers()) ? g->h.layer_nodes[l] : 0;
}
uint64_t hnswb200_graph_layer_nb_edges(const hnswb200_graph* g, uint32_t l) {
    if (!g || l >= g->h.n_layers()) return 0;
    uint64_t e = 0;
    for (uint64_t i = 0; i < g->h.n_points(); ++i)
        if (g->h.level[i] >= l) e += g->h.degree((uint32_t)i, l);
    return e;
}
uint32_t hnswb200_graph_layer_cap(const hnswb200_graph* g, uint32_t l) { return g ? g->h.cap(l) : 0; }

int hnswb200_graph_export_layer(const hnswb200_graph* g, uint32_t l, uint32_t* node_ids, uint64_t* offsets,
                                uint32_t* nbrs) {
    if (!g || l >= g->h.n_layers()) return fail(HNSWB200_EINVAL, "graph_export_layer: Layer not found in the structure.");
    uint64_t r = 0, off = 0;
    std::vector<uint32_t> tmp;
    for (uint64_t i = 0; i < g->h.n_points(); ++i) {
        if (g->h.level[i] < l) continue;
        if (node_ids) node_ids[r] = (uint32_t)i;
        if (offsets) offsets[r] = off;
        g->h.store(l).list(g->h.row((uint32_t)i, l), tmp);
        std::sort(tmp.begin(), tmp.end());
        if (nbrs) for (uint32_t v : tmp) nbrs[off++] = v;
        else off += tmp.size();
        ++r;
    }
    if (offsets) offsets[r] = off;
    return 0;
}

void hnswb200_graph_destroy(hnswb200_graph* g) {
    if (!g) return;
    g->free_device();
    delete g;
}

// ---------------------------------------------------------------------------
// index
// ---------------------------------------------------------------------------
int hnswb200_index_from_parts(hnswb200_ctx* c, hnswb200_points* points, hnswb200_graph* graph,
                              const hnswb200_params* params, hnswb200_index** out) {
    if (!c || !points || !graph || !params || !out) return fail(HNSWB200_EINVAL, "index_from_parts: NULL argument");
    if (points->n != graph->h.n_points()) return fail(HNSWB200_EINVAL, "index_from_parts: points and graph disagree on the number of points");
    if (params->dim != points->L.dim) return fail(HNSWB200_EINVAL, "index_from_parts: params.dim != points dimension");
    if (points->n && params->ep >= points->n) return fail(HNSWB200_EINVAL, "index_from_parts: entry point out of range");
    if (points->n && graph->h.level[params->ep] + 1u != graph->h.n_layers())
        return fail(HNSWB200_EINVAL, "index_from_parts: entry point is not a node of the top layer");
    hnswb200_index* ix = new hnswb200_index();
    ix->ctx = c;
    ix->points = points;
    ix->graph = graph;
    ix->params = *params;
    *out = ix;
    return 0;
}

void hnswb200_index_destroy(hnswb200_index* ix) {
    if (!ix) return;
    hnswb200_points_destroy(ix->points);
    hnswb200_graph_destroy(ix->graph);
    delete ix;
}
int hnswb200_index_set_metric(hnswb200_index* ix, int metric) {
    if (!ix) return fail(HNSWB200_EINVAL, "index_set_metric: NULL argument");
    return hnswb200_points_set_metric(ix->points, metric);
}
int hnswb200_index_metric(const hnswb200_index* ix) { return ix ? ix->points->metric : -1; }

int hnswb200_index_params(const hnswb200_index* ix, hnswb200_params* out) {
    if (!ix || !out) return fail(HNSWB200_EINVAL, "index_params: NULL argument");
    *out = ix->params;
    return 0;
}
uint64_t hnswb200_index_len(const hnswb200_index* ix) { return ix ? ix->points->n : 0; }
const hnswb200_points* hnswb200_index_points(const hnswb200_index* ix) { return ix ? ix->points : nullptr; }
const hnswb200_graph* hnswb200_index_graph(const hnswb200_index* ix) { return ix ? ix->graph : nullptr; }

// ---------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------
static int search_check(const hnswb200_index* ix, uint64_t nq, uint32_t n, uint32_t ef) {
    if (ix->points->n == 0) return fail(HNSWB200_ESTATE, "search: the index holds no points");
    if (n == 0 || ef == 0) return fail(HNSWB200_EINVAL, "search: n and ef must be >= 1");
    if (nq >= (1ull << 32)) return fail(HNSWB200_EINVAL, "search: too many queries in one call");
    if (ef > 16384) return fail(HNSWB200_EINVAL, "search: ef above 16384 is not supported by the shared-memory result list");
    if (!ix->graph->device_valid) return fail(HNSWB200_ESTATE, "search: graph is not resident on the device");
    return 0;
}

struct SearchExtras {
    bool prev_was_search = false;  // captured by the entry point BEFORE ctx->use() (which clears the flag)
    const float* queries_tail = nullptr;
    uint32_t split = 0;
    uint32_t n_peers = 0;
    uint32_t* const* peer_ids = nullptr;
    float* const* peer_dists = nullptr;
    uint64_t peer_row0 = 0;
    uint32_t id_offset = 0;
};
static int search_dev_impl(hnswb200_ctx* c, const hnswb200_index* ix, const float* d_queries, uint64_t nq,
                           uint32_t n, uint32_t ef, uint32_t* d_out_ids, float* d_out_dists,
                           uint32_t* d_out_counts, uint32_t* d_hops, uint32_t* d_evals, uint32_t* d_flags,
                           uint32_t* d_nbrs, uint32_t* nan_any, const SearchExtras& x);

int hnswb200_search_dev(hnswb200_ctx* c, const hnswb200_index* ix, const float* d_queries, uint64_t nq,
                        uint32_t n, uint32_t ef, uint32_t* d_out_ids, float* d_out_dists,
                        uint32_t* d_out_counts, uint32_t* d_hops, uint32_t* d_evals, uint32_t* d_flags,
                        uint32_t* d_nbrs) {
    SearchExtras x;
    x.prev_was_search = c && c->last_was_search;
    return search_dev_impl(c, ix, d_queries, nq, n, ef, d_out_ids, d_out_dists, d_out_counts, d_hops, d_evals, d_flags,
                           d_nbrs, nullptr, x);
}

static int search_dev_impl(hnswb200_ctx* c, const hnswb200_index* ix, const float* d_queries, uint64_t nq,
                           uint32_t n, uint32_t ef, uint32_t* d_out_ids, float* d_out_dists,
                           uint32_t* d_out_counts, uint32_t* d_hops, uint32_t* d_evals, uint32_t* d_flags,
                           uint32_t* d_nbrs, uint32_t* nan_any, const SearchExtras& x) {
    const float* queries_tail = x.queries_tail;
    uint32_t split = x.split;
    bool prev_search = x.prev_was_search;
    if (!c || !ix || (nq && (!d_queries || !d_out_ids))) return fail(HNSWB200_EINVAL, "search_dev: NULL argument");
    if (nq == 0) return 0;
    int rc = search_check(ix, nq, n, ef);
    if (rc) return rc;
    if (c->use()) return HNSWB200_ECUDA;
    SearchLaunch a;
    if (ix->points->metric == HNSWB200_METRIC_COSINE) {
        // cosine: search the unit-norm copy of the queries (both sources of a split batch are merged into it)
        const uint32_t dim = ix->points->L.dim;
        if (c->norm_ws_reserve(nq * dim * 4)) return HNSWB200_ECUDA;
        float* nqz = (float*)c->d_norm_ws;
        const uint64_t head = queries_tail ? std::min<uint64_t>(split, nq) : nq;
        HB_CUDA(launch_normalise(d_queries, head, dim, nqz, c->stream));
        if (head < nq) HB_CUDA(launch_normalise(queries_tail + head * dim, nq - head, dim, nqz + head * dim, c->stream));
        d_queries = nqz;
        queries_tail = nullptr;
        split = 0;
        prev_search = false;  // the search reads what launch_normalise writes: an ordinary, fully ordered launch
    }
    a.rec = ix->points->d_rec;
    a.L = ix->points->L;
    a.g = ix->graph->view();
    a.ep = ix->params.ep;
    a.n_points = ix->points->n;
    a.queries = d_queries;
    a.nq = (uint32_t)nq;
    a.topn = n;
    a.ef = ef;
    a.out_ids = d_out_ids;
    a.out_dists = d_out_dists;
    a.out_counts = d_out_counts;
    a.out_hops = d_hops;
    a.out_evals = d_evals;
    a.out_flags = d_flags;
    a.out_nbrs = d_nbrs;
    a.nan_any = nan_any;
    a.queries_tail = queries_tail;
    a.split = split;
    a.n_peers = x.n_peers;
    for (uint32_t g = 0; g < x.n_peers && g < hb::HB_MAX_PEERS; ++g) {
        a.peer_ids[g] = x.peer_ids[g];
        a.peer_dists[g] = x.peer_dists ? x.peer_dists[g] : nullptr;
    }
    a.peer_row0 = x.peer_row0;
    a.id_offset = x.id_offset;
    // counter ring (engine.h): slot 0 follows a memset of the whole ring and is an ordinary launch; the other
    // slots are launched as programmatic dependents of whatever kernel precedes them in the stream
    const uint32_t slot = (uint32_t)(c->search_seq++ % hnswb200_ctx::COUNTER_RING);
    if (slot == 0) HB_CUDA(cudaMemsetAsync(c->d_counters, 0, hnswb200_ctx::COUNTER_RING * sizeof(uint32_t), c->stream));
    a.work_counter = c->d_counters + slot;
    a.counter_is_fresh = true;
    if (!c->d_spill_ws) {
        const size_t words = (size_t)hnswb200_ctx::SPILL_SLICES * (1 + hnswb200_ctx::SPILL_CAP);
        HB_CUDA(cudaMalloc((void**)&c->d_spill_ws, words * 4));
        HB_CUDA(cudaMemsetAsync(c->d_spill_ws, 0, (size_t)hnswb200_ctx::SPILL_SLICES * 4, c->stream));  // owner words: free
        HB_CUDA(cudaMemsetAsync(c->d_spill_ws + hnswb200_ctx::SPILL_SLICES, 0xFF, (size_t)hnswb200_ctx::SPILL_SLICES * hnswb200_ctx::SPILL_CAP * 4,
                                c->stream));
        c->spill_warps = hnswb200_ctx::SPILL_SLICES;
    }
    a.spill_ws = c->d_spill_ws;
    a.spill_cap = hnswb200_ctx::SPILL_CAP;
    a.spill_warps = c->spill_warps;
    // programmatic dependent launch only directly behind another search of this context (any other
    // predecessor may write what this search reads), and on an adopted stream only with the caller's opt-in
    a.overlap_previous = slot != 0 && prev_search && (!c->stream_adopted || c->overlap_opt_in) && !getenv("HNSWB200_NO_PDL");
    HB_CUDA(launch_search(a, c->num_sms, c->stream));
    c->last_was_search = true;
    return 0;
}

int hnswb200_search_dev_gather(hnswb200_ctx* c, const hnswb200_index* ix, const float* d_queries, uint64_t nq,
                               uint32_t n, uint32_t ef, uint32_t* d_out_ids, float* d_out_dists, uint32_t* d_out_counts,
                               uint32_t n_peers, uint32_t* const* peer_ids, uint64_t row_offset) {
    if (n_peers > hb::HB_MAX_PEERS) return fail(HNSWB200_EINVAL, "search_dev_gather: at most 8 peer buffers");
    if (n_peers && !peer_ids) return fail(HNSWB200_EINVAL, "search_dev_gather: NULL peer list");
    for (uint32_t g = 0; g < n_peers; ++g)
        if (!peer_ids[g]) return fail(HNSWB200_EINVAL, "search_dev_gather: NULL peer buffer");
    SearchExtras x;
    x.prev_was_search = c && c->last_was_search;
    x.n_peers = n_peers;
    x.peer_ids = peer_ids;
    x.peer_row0 = row_offset;
    return search_dev_impl(c, ix, d_queries, nq, n, ef, d_out_ids, d_out_dists, d_out_counts, nullptr, nullptr, nullptr,
                           nullptr, nullptr, x);
}

int hnswb200_search_dev_shard(hnswb200_ctx* c, const hnswb200_index* ix, const float* d_queries, uint64_t nq, uint32_t n,
                              uint32_t ef, uint32_t id_offset, uint32_t* d_out_ids, float* d_out_dists,
                              uint32_t* d_out_counts, uint32_t n_peers, uint32_t* const* peer_ids,
                              float* const* peer_dists, uint64_t row_offset) {
    if (n_peers > hb::HB_MAX_PEERS) return fail(HNSWB200_EINVAL, "search_dev_shard: at most 8 peer buffers");
    if (n_peers && (!peer_ids || !peer_dists)) return fail(HNSWB200_EINVAL, "search_dev_shard: NULL peer list");
    for (uint32_t g = 0; g < n_peers; ++g)
        if (!peer_ids[g] || !peer_dists[g]) return fail(HNSWB200_EINVAL, "search_dev_shard: NULL peer buffer");
    if (ix && (uint64_t)id_offset + ix->points->n > (1ull << 31))
        return fail(HNSWB200_EINVAL, "search_dev_shard: global ids must stay below 2^31");
    SearchExtras x;
    x.prev_was_search = c && c->last_was_search;
    x.n_peers = n_peers;
    x.peer_ids = peer_ids;
    x.peer_dists = peer_dists;
    x.peer_row0 = row_offset;
    x.id_offset = id_offset;
    return search_dev_impl(c, ix, d_queries, nq, n, ef, d_out_ids, d_out_dists, d_out_counts, nullptr, nullptr, nullptr,
                           nullptr, nullptr, x);
}

// ---- peer exchange without a collective library (kernels.cu) ----
int hnswb200_peer_put_dev(hnswb200_ctx* c, const void* d_src, uint64_t bytes, uint32_t n_peers, void* const* peer_dst) {
    if (!c || (bytes && !d_src) || (n_peers && !peer_dst)) return fail(HNSWB200_EINVAL, "peer_put: NULL argument");
    if (n_peers > hb::HB_MAX_PEERS) return fail(HNSWB200_EINVAL, "peer_put: at most 8 peer buffers");
    if (bytes % 16 || ((uintptr_t)d_src & 15)) return fail(HNSWB200_EINVAL, "peer_put: size and pointers must be multiples of 16 bytes");
    for (uint32_t g = 0; g < n_peers; ++g)
        if (!peer_dst[g] || ((uintptr_t)peer_dst[g] & 15)) return fail(HNSWB200_EINVAL, "peer_put: bad peer pointer");
    if (c->use()) return HNSWB200_ECUDA;
    HB_CUDA(hb::launch_peer_put(d_src, bytes, n_peers, peer_dst, c->num_sms, c->stream));
    return 0;
}
int hnswb200_peer_signal_dev(hnswb200_ctx* c, uint32_t n_peers, uint32_t* const* peer_flags, uint32_t slot, uint32_t epoch) {
    if (!c || (n_peers && !peer_flags)) return fail(HNSWB200_EINVAL, "peer_signal: NULL argument");
    if (n_peers > hb::HB_MAX_PEERS) return fail(HNSWB200_EINVAL, "peer_signal: at most 8 peers");
    for (uint32_t g = 0; g < n_peers; ++g)
        if (!peer_flags[g]) return fail(HNSWB200_EINVAL, "peer_signal: NULL flag array");
    if (c->use()) return HNSWB200_ECUDA;
    HB_CUDA(hb::launch_peer_signal(n_peers, peer_flags, slot, epoch, c->stream));
    return 0;
}
int hnswb200_peer_wait_dev(hnswb200_ctx* c, const uint32_t* d_flags, uint32_t n_slots, uint32_t epoch) {
    if (!c || (n_slots && !d_flags)) return fail(HNSWB200_EINVAL, "peer_wait: NULL argument");
    if (n_slots > 32) return fail(HNSWB200_EINVAL, "peer_wait: at most 32 flag words");
    if (c->use()) return HNSWB200_ECUDA;
    HB_CUDA(hb::launch_peer_wait(d_flags, n_slots, epoch, c->d_status, c->stream));
    return 0;
}

// ---- device buffers that other processes of the box can write (CUDA IPC over NVLink / NVSwitch) ----
int hnswb200_dev_alloc(hnswb200_ctx* c, uint64_t bytes, void** out) {
    if (!c || !out) return fail(HNSWB200_EINVAL, "dev_alloc: NULL argument");
    if (c->use()) return HNSWB200_ECUDA;
    HB_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    HB_CUDA(cudaMemsetAsync(*out, 0xFF, bytes, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int hnswb200_dev_free(hnswb200_ctx* c, void* p) {
    if (!c) return fail(HNSWB200_EINVAL, "dev_free: NULL context");
    if (c->use()) return HNSWB200_ECUDA;
    if (p) HB_CUDA(cudaFree(p));
    return 0;
}
int hnswb200_dev_download(hnswb200_ctx* c, const void* d_src, void* host_dst, uint64_t bytes) {
    if (!c || !d_src || !host_dst) return fail(HNSWB200_EINVAL, "dev_download: NULL argument");
    if (c->use()) return HNSWB200_ECUDA;
    HB_CUDA(cudaMemcpyAsync(host_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int hnswb200_ipc_export(hnswb200_ctx* c, void* d_ptr, uint8_t handle[64]) {
    if (!c || !d_ptr || !handle) return fail(HNSWB200_EINVAL, "ipc_export: NULL argument");
    if (c->use()) return HNSWB200_ECUDA;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    HB_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, 64);
    return 0;
}
int hnswb200_ipc_open(hnswb200_ctx* c, const uint8_t handle[64], void** out) {
    if (!c || !handle || !out) return fail(HNSWB200_EINVAL, "ipc_open: NULL argument");
    if (c->use()) return HNSWB200_ECUDA;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    HB_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int hnswb200_ipc_close(hnswb200_ctx* c, void* p) {
    if (!c) return fail(HNSWB200_EINVAL, "ipc_close: NULL context");
    if (c->use()) return HNSWB200_ECUDA;
    if (p) HB_CUDA(cudaIpcCloseMemHandle(p));
    return 0;
}

// device alias of a page-locked (mapped) host buffer, or NULL for pageable memory
static void* mapped_alias(const void* host) {
    if (!host || getenv("HNSWB200_NO_ZERO_COPY")) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (at.type == cudaMemoryTypeHost) ? at.devicePointer : nullptr;
}

int hnswb200_search_async(hnswb200_ctx* c, const hnswb200_index* ix, const float* queries, uint64_t nq, uint32_t dim,
                          uint32_t n, uint32_t ef, uint32_t* out_ids, float* out_dists, uint32_t* out_counts) {
    if (!c || !ix || (nq && (!queries || !out_ids))) return fail(HNSWB200_EINVAL, "search_async: NULL argument");
    const bool prev_search = c->last_was_search;
    if (dim != ix->points->L.dim)
        return fail(HNSWB200_EINVAL, "search: query dimension " + std::to_string(dim) + " != index dimension " +
                                         std::to_string(ix->points->L.dim));
    if (nq == 0) return 0;
    int rc = search_check(ix, nq, n, ef);
    if (rc) return rc;
    if (c->use()) return HNSWB200_ECUDA;
    void* dq = mapped_alias(queries);
    void* di = mapped_alias(out_ids);
    void* dd = out_dists ? mapped_alias(out_dists) : nullptr;
    void* dc = out_counts ? mapped_alias(out_counts) : nullptr;
    if (!dq || !di || (out_dists && !dd) || (out_counts && !dc))
        return fail(HNSWB200_EINVAL, "search_async: every buffer must be page-locked (cudaHostAlloc / cudaHostRegister): "
                                     "the kernel reads and writes them in place");
    // no staging and no copies: the call only launches; consecutive calls overlap on the device
    SearchExtras x;
    x.prev_was_search = prev_search;
    return search_dev_impl(c, ix, (const float*)dq, nq, n, ef, (uint32_t*)di, (float*)dd, (uint32_t*)dc, nullptr, nullptr,
                           nullptr, nullptr, c->d_status, x);
}

int hnswb200_search(hnswb200_ctx* c, const hnswb200_index* ix, const float* queries, uint64_t nq,
                    uint32_t dim, uint32_t n, uint32_t ef, uint32_t* out_ids, float* out_dists,
                    uint32_t* out_counts, const hnswb200_search_stats* stats) {
    if (!c || !ix || (nq && (!queries || !out_ids))) return fail(HNSWB200_EINVAL, "search: NULL argument");
    if (dim != ix->points->L.dim)
        return fail(HNSWB200_EINVAL, "search: query dimension " + std::to_string(dim) + " != index dimension " +
                                         std::to_string(ix->points->L.dim));
    if (nq == 0) return 0;
    int rc = search_check(ix, nq, n, ef);
    if (rc) return rc;
    if (c->use()) return HNSWB200_ECUDA;
    const bool prof_call = getenv("HNSWB200_PROFILE_CALL") != nullptr;
    auto now_us = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
    const double t_enter = prof_call ? now_us() : 0;
    // Page-locked caller buffers are used in place: the kernel reads the queries and writes the results over
    // PCIe itself (88 + 4*dim bytes per query, spread over the whole launch), so there is no staging copy before
    // or after it.  Pageable buffers are staged through the grow-only per-context workspace (no cudaMalloc on
    // the steady-state path either way).
    const size_t b_q = nq * dim * 4, b_ids = nq * (size_t)n * 4, b_u = nq * 4;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    auto mapped = [](const void* host) -> void* { return mapped_alias(host); };
    struct Buf { void* host; void* dev; size_t bytes; bool staged; };
    uint32_t* st_hops = stats ? stats->hops : nullptr;
    uint32_t* st_evals = stats ? stats->evals : nullptr;
    uint32_t* st_flags = stats ? stats->flags : nullptr;
    uint32_t* st_nbrs = stats ? stats->nbrs : nullptr;
    Buf q{(void*)queries, mapped(queries), b_q, false};
    Buf outs[7] = {{out_ids, mapped(out_ids), b_ids, false},   {out_dists, mapped(out_dists), b_ids, false},
                   {out_counts, mapped(out_counts), b_u, false}, {st_hops, mapped(st_hops), b_u, false},
                   {st_evals, mapped(st_evals), b_u, false},     {st_flags, mapped(st_flags), b_u, false},
                   {st_nbrs, mapped(st_nbrs), b_u, false}};
    // The first wave of queries (one per resident warp) is wanted all at once when the kernel starts: those are
    // copied by DMA; the rest of a page-locked query buffer is read in place, one query at a time, while others compute.
    const uint64_t first_wave = std::min<uint64_t>(nq, (uint64_t)c->num_sms * 32);
    const float* q_tail = nullptr;
    if (q.dev && first_wave < nq) q_tail = (const float*)q.dev;
    const size_t b_head = q.dev ? (q_tail ? first_wave * dim * 4 : 0) : b_q;
    size_t total = al(b_head);
    for (Buf& b : outs)
        if (b.host && !b.dev) total += al(b.bytes);
    if (total && c->ws_reserve(total)) return HNSWB200_ECUDA;
    unsigned char* w = (unsigned char*)c->d_ws;
    if (b_head) { q.dev = w; w += al(b_head); q.staged = true; }
    for (Buf& b : outs)
        if (b.host && !b.dev) { b.dev = w; w += al(b.bytes); b.staged = true; }
    c->h_status[0] = 0;
    const double t_setup = prof_call ? now_us() : 0;
    if (q.staged) HB_CUDA(cudaMemcpyAsync(q.dev, queries, b_head, cudaMemcpyHostToDevice, c->stream));
    SearchExtras sx;
    sx.queries_tail = q_tail;
    sx.split = (uint32_t)first_wave;
    rc = search_dev_impl(c, ix, (const float*)q.dev, nq, n, ef, (uint32_t*)outs[0].dev, (float*)outs[1].dev,
                         (uint32_t*)outs[2].dev, (uint32_t*)outs[3].dev, (uint32_t*)outs[4].dev, (uint32_t*)outs[5].dev,
                         (uint32_t*)outs[6].dev, c->d_status, sx);
    if (rc) return rc;
    for (Buf& b : outs)
        if (b.staged) HB_CUDA(cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, c->stream));
    const double t_enq = prof_call ? now_us() : 0;
    HB_CUDA(cudaStreamSynchronize(c->stream));
    if (prof_call) {
        const double t_done = now_us();
        fprintf(stderr, "[hnswb200 search] nq=%llu setup %.1f us, enqueue %.1f us, wait %.1f us\n", (unsigned long long)nq,
                t_setup - t_enter, t_enq - t_setup, t_done - t_enq);
    }
    if (c->h_status[0]) {
        c->h_status[0] = 0;
        return fail(HNSWB200_EINVAL, "search: NaN in a query (the reference panics in partial_cmp().unwrap())");
    }
    return 0;
}

// ---------------------------------------------------------------------------
// brute force
// ---------------------------------------------------------------------------
int hnswb200_bruteforce_topk_dev(hnswb200_ctx* c, const hnswb200_points* base, const float* d_queries,
                                 uint64_t nq, uint32_t k, uint32_t id_offset, uint32_t* d_out_ids,
                                 float* d_out_dists) {
    if (!c || !base || (nq && (!d_queries || !d_out_ids))) return fail(HNSWB200_EINVAL, "bruteforce: NULL argument");
    if (k == 0 || k > 2048) return fail(HNSWB200_EINVAL, "bruteforce: k must be in 1..2048");
    if (nq == 0) return 0;
    if (nq >= (1ull << 31)) return fail(HNSWB200_EINVAL, "bruteforce: too many queries");
    if (c->use()) return HNSWB200_ECUDA;
    const RecLayout& L = base->L;
    const uint32_t cap = 2048;
    const uint64_t N = base->n;
    // Tensor-core filter (bf_tc.cu) for every chunk after the first: the first (exact) chunk establishes
    // the thresholds.  HNSWB200_BF_NO_TC forces the CUDA-core path (test knob).
    const bool use_tc = bf_tc_supported(L) && N > 2 * cap && !getenv("HNSWB200_BF_NO_TC");
    const uint64_t nq_pad = (nq + 127) / 128 * 128;
    // all scratch comes from one grow-only workspace
    struct Carve {
        unsigned char* base = nullptr;
        size_t off = 0;
        size_t take(size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; }
    } cv;
    const size_t o_qrec = cv.take(nq * L.stride), o_topk = cv.take(nq * k * 8), o_tau = cv.take(nq * 8),
                 o_buf = cv.take(nq * (size_t)cap * 8), o_cnt = cv.take(nq * 4),
                 o_bconst = cv.take(use_tc ? N * 16 : 0), o_qstat = cv.take(use_tc ? nq * 16 : 0),
                 o_qconst = cv.take(use_tc ? nq_pad * 16 : 0), o_amask = cv.take(use_tc ? nq_pad * 128 : 0),
                 o_qshift = cv.take(use_tc ? nq_pad * 4 : 0),
                 o_qnorm = cv.take(base->metric == HNSWB200_METRIC_COSINE ? nq * L.dim * 4 : 0);
    if (c->bf_ws_reserve(cv.off)) return HNSWB200_ECUDA;
    unsigned char* W = (unsigned char*)c->d_bf_ws;
    struct { uint8_t* p; } qrec{W + o_qrec}, amask{W + o_amask};
    struct { uint64_t* p; } topk{(uint64_t*)(W + o_topk)}, tau{(uint64_t*)(W + o_tau)}, buf{(uint64_t*)(W + o_buf)};
    struct { uint32_t* p; } cnt{(uint32_t*)(W + o_cnt)};
    struct { float4* p; } bconst{(float4*)(W + o_bconst)}, qstat{(float4*)(W + o_qstat)}, qconst{(float4*)(W + o_qconst)};
    uint32_t* nan_flag = c->d_scratch + 1;
    uint32_t* ovf_flag = c->d_scratch + 2;
    HB_CUDA(cudaMemsetAsync(nan_flag, 0, 8, c->stream));
    HB_CUDA(cudaMemsetAsync(qrec.p, 0, nq * L.stride, c->stream));
    HB_CUDA(cudaMemsetAsync(topk.p, 0xFF, nq * k * 8, c->stream));
    HB_CUDA(cudaMemsetAsync(tau.p, 0xFF, nq * 8, c->stream));
    HB_CUDA(cudaMemsetAsync(cnt.p, 0, nq * 4, c->stream));
    if (base->metric == HNSWB200_METRIC_COSINE) {  // cosine: unit-norm copy of the queries first
        float* qn = (float*)(W + o_qnorm);
        HB_CUDA(launch_normalise(d_queries, nq, L.dim, qn, c->stream));
        d_queries = qn;
    }
    // queries become points like the stored ones (Point::new) and are laid out as records
    if (L.kind == HB_REC_F32) HB_CUDA(launch_pack_f32(d_queries, nq, L, qrec.p, nan_flag, c->stream));
    else HB_CUDA(launch_quantise(d_queries, nq, L, qrec.p, nullptr, nullptr, nullptr, nan_flag, c->stream));
    uint32_t flags[2] = {0, 0};
    HB_CUDA(cudaMemcpyAsync(flags, nan_flag, 8, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    if (flags[0]) return fail(HNSWB200_EINVAL, L.kind == HB_REC_F32 ? "bruteforce: non-finite value in a query" : "bruteforce: NaN in a query");
    // Chunks double in size: with tau = current k-th best, the expected number of survivors of a
    // chunk as large as everything seen before is <= k per query.  A chunk that overflows the
    // per-query buffer (adversarial order) is redone in pieces of `cap` rows, which cannot overflow.
    if (use_tc) {
        HB_CUDA(cudaMemsetAsync(amask.p, 0, nq_pad * 128, c->stream));
        HB_CUDA(cudaMemsetAsync(W + o_qshift, 0, nq_pad * 4, c->stream));
        HB_CUDA(bf_tc_prepare(base->d_rec, N, L, qrec.p, (uint32_t)nq, bconst.p, amask.p, qstat.p, (int*)(W + o_qshift), c->stream));
    }
    const bool prof = getenv("HNSWB200_BF_PROFILE") != nullptr;
    bool tc_done = false;
    if (use_tc) {
        // Optimistic pass, no host synchronisation between the chunks: the first rows are ranked exactly (every pair of
        // them is a survivor), then chunks that multiply the rows seen by `grow`.  With tau = the k-th best of n rows in
        // random order, a chunk of (grow-1)*n rows leaves about k*ln(grow) survivors per query, far below `cap`.  The
        // overflow flag is read once at the end; an order that overflows a list (sorted by decreasing distance, say)
        // sends the whole call through the exact path below.
        uint64_t first = 256;
        while (first < 4ull * k && first < cap) first <<= 1;
        first = std::min<uint64_t>(first, N);
        const uint64_t grow = (4ull * k <= cap) ? 4 : 2, TC_MAXCH = 1ull << 20;
        HB_CUDA(bf_tc_first(base->d_rec, L, (uint32_t)first, id_offset, qrec.p, (uint32_t)nq,
                            reinterpret_cast<unsigned long long*>(buf.p), cap, cnt.p, c->stream));
        HB_CUDA(launch_bf_merge(topk.p, k, tau.p, buf.p, cap, cnt.p, (uint32_t)nq, c->stream));
        uint64_t done = first;
        while (done < N) {
            const uint64_t ch = std::min<uint64_t>(std::min<uint64_t>(done * (grow - 1), TC_MAXCH), N - done);
            HB_CUDA(bf_tc_chunk(base->d_rec, N, L, done, done + ch, id_offset, qrec.p, amask.p, qstat.p, bconst.p, qconst.p,
                                (const int*)(W + o_qshift), (uint32_t)nq, reinterpret_cast<const unsigned long long*>(tau.p),
                                reinterpret_cast<unsigned long long*>(buf.p), cap, cnt.p, ovf_flag, c->num_sms, c->stream));
            HB_CUDA(launch_bf_merge(topk.p, k, tau.p, buf.p, cap, cnt.p, (uint32_t)nq, c->stream));
            if (prof) fprintf(stderr, "[hnswb200 bruteforce] rows [%llu, %llu) tensor-core filter (enqueued)\n",
                              (unsigned long long)done, (unsigned long long)(done + ch));
            done += ch;
        }
        uint32_t ovf = 0;
        HB_CUDA(cudaMemcpyAsync(&ovf, ovf_flag, 4, cudaMemcpyDeviceToHost, c->stream));
        HB_CUDA(cudaStreamSynchronize(c->stream));
        if (prof) fprintf(stderr, "[hnswb200 bruteforce] optimistic pass done, overflow=%u\n", ovf);
        if (!ovf) tc_done = true;
        else {  // start over on the exact path
            HB_CUDA(cudaMemsetAsync(ovf_flag, 0, 4, c->stream));
            HB_CUDA(cudaMemsetAsync(topk.p, 0xFF, nq * k * 8, c->stream));
            HB_CUDA(cudaMemsetAsync(tau.p, 0xFF, nq * 8, c->stream));
            HB_CUDA(cudaMemsetAsync(cnt.p, 0, nq * 4, c->stream));
        }
    }
    // Exact CUDA-core path (record shapes without a tensor-core path, small bases, adversarial orders): chunks of `cap`
    // rows cannot overflow a list of `cap` entries.
    for (uint64_t done = 0; !tc_done && done < N;) {
        const uint64_t ch = std::min<uint64_t>(cap, N - done);
        HB_CUDA(launch_bf_chunk(base->d_rec, L, done, done + ch, id_offset, qrec.p, (uint32_t)nq, tau.p, buf.p, cap, cnt.p,
                                ovf_flag, c->stream));
        HB_CUDA(launch_bf_merge(topk.p, k, tau.p, buf.p, cap, cnt.p, (uint32_t)nq, c->stream));
        if (prof) fprintf(stderr, "[hnswb200 bruteforce] rows [%llu, %llu) exact (enqueued)\n", (unsigned long long)done,
                          (unsigned long long)(done + ch));
        done += ch;
    }
    HB_CUDA(launch_keys_to_out(topk.p, k, (uint32_t)nq, d_out_ids, d_out_dists, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int hnswb200_bruteforce_topk(hnswb200_ctx* c, const hnswb200_points* base, const float* queries,
                             uint64_t nq, uint32_t k, uint32_t id_offset, uint32_t* out_ids,
                             float* out_dists) {
    if (!c || !base || (nq && (!queries || !out_ids))) return fail(HNSWB200_EINVAL, "bruteforce: NULL argument");
    if (nq == 0) return 0;
    if (c->use()) return HNSWB200_ECUDA;
    const size_t qb = (nq * base->L.dim * 4 + 255) & ~(size_t)255, ib = ((size_t)nq * k * 4 + 255) & ~(size_t)255;
    if (c->ws_reserve(qb + 2 * ib)) return HNSWB200_ECUDA;
    struct { float* p; } dq{(float*)c->d_ws}, dd{(float*)((unsigned char*)c->d_ws + qb + ib)};
    struct { uint32_t* p; } dids{(uint32_t*)((unsigned char*)c->d_ws + qb)};
    HB_CUDA(cudaMemcpyAsync(dq.p, queries, nq * base->L.dim * 4, cudaMemcpyHostToDevice, c->stream));
    int rc = hnswb200_bruteforce_topk_dev(c, base, dq.p, nq, k, id_offset, dids.p, dd.p);
    if (rc) return rc;
    HB_CUDA(cudaMemcpyAsync(out_ids, dids.p, nq * k * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_dists) HB_CUDA(cudaMemcpyAsync(out_dists, dd.p, nq * k * 4, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------------------
// top-k merge
// ---------------------------------------------------------------------------
int hnswb200_topk_merge_dev(hnswb200_ctx* c, const uint32_t* d_ids, const float* d_dists, uint32_t G,
                            uint64_t nq, uint32_t k, uint32_t* d_out_ids, float* d_out_dists) {
    if (!c || !d_ids || !d_dists || !d_out_ids) return fail(HNSWB200_EINVAL, "topk_merge: NULL argument");
    if (G == 0 || k == 0 || (uint64_t)G * k > 16384) return fail(HNSWB200_EINVAL, "topk_merge: need 1 <= G*k <= 16384");
    if (c->use()) return HNSWB200_ECUDA;
    HB_CUDA(launch_topk_merge(d_ids, d_dists, G, (uint32_t)nq, k, d_out_ids, d_out_dists, c->stream));
    return 0;
}

int hnswb200_topk_merge(hnswb200_ctx* c, const uint32_t* ids, const float* dists, uint32_t G, uint64_t nq,
                        uint32_t k, uint32_t* out_ids, float* out_dists) {
    if (!c || !ids || !dists || !out_ids) return fail(HNSWB200_EINVAL, "topk_merge: NULL argument");
    if (nq == 0) return 0;
    if (c->use()) return HNSWB200_ECUDA;
    uint64_t tot = (uint64_t)G * nq * k;
    DevBuf<uint32_t> di, doi;
    DevBuf<float> dd, dod;
    HB_CUDA(di.alloc(tot));
    HB_CUDA(dd.alloc(tot));
    HB_CUDA(doi.alloc(nq * k));
    HB_CUDA(dod.alloc(nq * k));
    HB_CUDA(cudaMemcpyAsync(di.p, ids, tot * 4, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(cudaMemcpyAsync(dd.p, dists, tot * 4, cudaMemcpyHostToDevice, c->stream));
    int rc = hnswb200_topk_merge_dev(c, di.p, dd.p, G, nq, k, doi.p, dod.p);
    if (rc) return rc;
    HB_CUDA(cudaMemcpyAsync(out_ids, doi.p, nq * k * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_dists) HB_CUDA(cudaMemcpyAsync(out_dists, dod.p, nq * k * 4, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------------------
// HNSW::save / HNSW::load byte formats (all big-endian)
// ---------------------------------------------------------------------------
}  // extern "C"

namespace {
void put_u64(std::vector<uint8_t>& b, uint64_t v) { for (int i = 7; i >= 0; --i) b.push_back((uint8_t)(v >> (8 * i))); }
void put_u32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 3; i >= 0; --i) b.push_back((uint8_t)(v >> (8 * i))); }
void put_f32(std::vector<uint8_t>& b, float f) { uint32_t u; memcpy(&u, &f, 4); put_u32(b, u); }
uint64_t get_u64(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 8; ++i) v = (v << 8) | p[i]; return v; }
uint32_t get_u32(const uint8_t* p) { uint32_t v = 0; for (int i = 0; i < 4; ++i) v = (v << 8) | p[i]; return v; }
float get_f32(const uint8_t* p) { uint32_t u = get_u32(p); float f; memcpy(&f, &u, 4); return f; }
bool write_file(const std::string& path, const std::vector<uint8_t>& b) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    size_t w = b.empty() ? 0 : fwrite(b.data(), 1, b.size(), f);
    fclose(f);
    return w == b.size();
}
bool read_file(const std::string& path, std::vector<uint8_t>& b) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    b.resize((size_t)sz);
    size_t r = sz ? fread(b.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    return r == (size_t)sz;
}
}  // namespace

extern "C" {

int hnswb200_index_save_dir(hnswb200_ctx* c, const hnswb200_index* ix, const char* dir) {
    if (!c || !ix || !dir) return fail(HNSWB200_EINVAL, "save: NULL argument");
    const hnswb200_points* P = ix->points;
    const uint64_t n = P->n, dim = P->L.dim;
    if (n == 0) return fail(HNSWB200_ESTATE, "save: the index holds no points");
    std::string d(dir);
    if (mkdir(d.c_str(), 0777) != 0 && errno != EEXIST) return fail(HNSWB200_EIO, "Could not create dir " + d);
    std::vector<uint8_t> b;
    int rc;
    if (P->L.kind == HB_REC_F32) {
        // points file with VecType = FullVec: points.rs:119-131, point.rs:55-61, full.rs:55-61 (dim big-endian floats)
        std::vector<uint8_t> levels(n);
        std::vector<float> vals(n * dim);
        rc = hnswb200_points_values(c, P, vals.data(), levels.data());
        if (rc) return rc;
        b.reserve(16 + n * (1 + 4 * dim));
        put_u64(b, n);
        put_u64(b, 1 + 4 * dim);
        for (uint64_t i = 0; i < n; ++i) {
            b.push_back(levels[i]);
            for (uint64_t j = 0; j < dim; ++j) put_f32(b, vals[i * dim + j]);
        }
    } else {
        std::vector<uint8_t> codes(n * dim), levels(n);
        std::vector<float> mins(n), deltas(n);
        rc = hnswb200_points_download(c, P, codes.data(), mins.data(), deltas.data(), levels.data());
        if (rc) return rc;
        // points file: points.rs:119-131, point.rs:55-61, quant.rs:102-110 (min before delta)
        b.reserve(16 + n * (9 + dim));
        put_u64(b, n);
        put_u64(b, 9 + dim);
        for (uint64_t i = 0; i < n; ++i) {
            b.push_back(levels[i]);
            put_f32(b, mins[i]);
            put_f32(b, deltas[i]);
            b.insert(b.end(), &codes[i * dim], &codes[(i + 1) * dim]);
        }
    }
    if (!write_file(d + "/points", b)) return fail(HNSWB200_EIO, "Could not write bytes to point file");
    b.clear();  // params.rs:78-91
    put_u64(b, ix->params.m); put_u64(b, ix->params.mmax); put_u64(b, ix->params.mmax0);
    put_f32(b, ix->params.ml);
    put_u64(b, ix->params.ef_cons); put_u64(b, ix->params.dim); put_u64(b, ix->params.ep);
    if (!write_file(d + "/params", b)) return fail(HNSWB200_EIO, "Could not write bytes to params file");
    if (mkdir((d + "/layers").c_str(), 0777) != 0 && errno != EEXIST) return fail(HNSWB200_EIO, "Could not create layers dir");
    const hb::HostGraph& h = ix->graph->h;
    std::vector<uint32_t> tmp;
    for (uint32_t l = 0; l < h.n_layers(); ++l) {  // graph.rs:165-219
        // The reference pads rows to `m` words but never truncates; a node above the cap
        // (possible) would misalign the file.  Keep it self-consistent:
        // the row width in the header is max(cap, max degree).
        uint32_t width = h.cap(l);
        for (uint64_t i = 0; i < n; ++i)
            if (h.level[i] >= l) width = std::max(width, h.degree((uint32_t)i, l));
        if (width > 0xFFFF) return fail(HNSWB200_EIO, "save: degree does not fit the u16 row width");
        b.clear();
        b.push_back((uint8_t)l);
        put_u32(b, (uint32_t)h.layer_nodes[l]);
        b.push_back((uint8_t)(width >> 8));
        b.push_back((uint8_t)width);
        for (uint64_t i = 0; i < n; ++i) {
            if (h.level[i] < l) continue;
            put_u32(b, (uint32_t)i);
            h.store(l).list(h.row((uint32_t)i, l), tmp);
            for (uint32_t v : tmp) put_u32(b, v);
            for (size_t j = tmp.size(); j < width; ++j) put_u32(b, 0xFFFFFFFFu);
        }
        if (!write_file(d + "/layers/" + std::to_string(l), b)) return fail(HNSWB200_EIO, "Could not write bytes to layer file");
    }
    return 0;
}

static int load_dir_impl(hnswb200_ctx* c, const char* dir, hnswb200_index** out);

int hnswb200_index_load_dir(hnswb200_ctx* c, const char* dir, hnswb200_index** out) {
    if (!c || !dir || !out) return fail(HNSWB200_EINVAL, "load: NULL argument");
    // nothing unwinds across the C boundary: a crafted header must not turn into std::bad_alloc / length_error
    try {
        return load_dir_impl(c, dir, out);
    } catch (const std::bad_alloc&) {
        return fail(HNSWB200_ENOMEM, "load: out of host memory (corrupt header?)");
    } catch (const std::exception& e) {
        return fail(HNSWB200_EIO, std::string("load: ") + e.what());
    }
}

static int load_dir_impl(hnswb200_ctx* c, const char* dir, hnswb200_index** out) {
    std::string d(dir);
    struct stat st;
    if (stat(d.c_str(), &st) != 0) return fail(HNSWB200_EIO, "\"" + d + "\" does not exist");
    std::vector<uint8_t> b;
    if (!read_file(d + "/points", b) || b.size() < 16) return fail(HNSWB200_EIO, "Problem reading points file");
    uint64_t n = get_u64(&b[0]), psz = get_u64(&b[8]);
    // ids are < 2^31 and a point is at most 1 + 4*dim bytes: bound both before multiplying (no wrap-around)
    if (psz < 5 || psz > (1ull << 24) || n >= (1ull << 31)) return fail(HNSWB200_EIO, "points file header is not plausible");
    if ((b.size() - 16) / psz < n) return fail(HNSWB200_EIO, "points file is truncated");
    std::vector<uint8_t> pb;
    pb.swap(b);
    if (!read_file(d + "/params", b) || b.size() < 52) return fail(HNSWB200_EIO, "Problem reading params file");
    hnswb200_params prm;
    prm.m = get_u64(&b[0]); prm.mmax = get_u64(&b[8]); prm.mmax0 = get_u64(&b[16]);
    prm.ml = get_f32(&b[24]);
    prm.ef_cons = get_u64(&b[28]); prm.dim = get_u64(&b[36]); prm.ep = (uint32_t)get_u64(&b[44]);
    // The file does not name its VecType; the point size does: 1 + 8 + dim bytes for a QuantVec (quant.rs:91-93),
    // 1 + 4*dim for a FullVec (full.rs:45-47).  The two never coincide for an integer dim.
    const uint64_t dim = prm.dim;
    if (dim == 0 || dim > (1ull << 22)) return fail(HNSWB200_EIO, "params.dim is not plausible");
    const bool full = psz == 1 + 4 * dim && psz != 9 + dim;
    if (!full && psz != 9 + dim) return fail(HNSWB200_EIO, "params.dim does not match the point size");
    std::vector<uint8_t> codes, levels(n);
    std::vector<float> mins, deltas, vals;
    if (full) {
        vals.resize(n * dim);
        for (uint64_t i = 0; i < n; ++i) {
            const uint8_t* p = &pb[16 + i * psz];
            levels[i] = p[0];
            for (uint64_t j = 0; j < dim; ++j) vals[i * dim + j] = get_f32(p + 1 + 4 * j);
        }
    } else {
        codes.resize(n * dim);
        mins.resize(n);
        deltas.resize(n);
        for (uint64_t i = 0; i < n; ++i) {
            const uint8_t* p = &pb[16 + i * psz];
            levels[i] = p[0];
            mins[i] = get_f32(p + 1);
            deltas[i] = get_f32(p + 5);
            memcpy(&codes[i * dim], p + 9, dim);
        }
    }
    // layers/<idx>, sorted numerically (template.rs:103-109)
    std::vector<uint64_t> idxs;
    DIR* dd = opendir((d + "/layers").c_str());
    if (!dd) return fail(HNSWB200_EIO, "There was a problem reading layers");
    while (dirent* e = readdir(dd)) {
        if (e->d_name[0] == '.') continue;
        char* end = nullptr;
        const unsigned long long v = strtoull(e->d_name, &end, 10);
        if (end == e->d_name || *end != '\0' || v > 255) {  // the reference parses the name as a number (template.rs:103-109); a level is one byte
            closedir(dd);
            return fail(HNSWB200_EIO, std::string("layers/") + e->d_name + ": not a layer file name");
        }
        idxs.push_back(v);
    }
    closedir(dd);
    std::sort(idxs.begin(), idxs.end());
    uint32_t nl = (uint32_t)idxs.size();
    if (nl == 0) return fail(HNSWB200_EIO, "index has no layer files");
    std::vector<std::vector<uint32_t>> ids(nl), nb(nl);
    std::vector<std::vector<uint64_t>> off(nl);
    std::vector<uint32_t> caps(nl);
    std::vector<uint64_t> nn(nl);
    for (uint32_t l = 0; l < nl; ++l) {
        if (!read_file(d + "/layers/" + std::to_string(idxs[l]), b) || b.size() < 7) return fail(HNSWB200_EIO, "Problem reading layer file");
        if (b[0] != l) return fail(HNSWB200_EIO, "layer level does not match its position (template.rs:119)");
        uint32_t cnt = get_u32(&b[1]);
        uint32_t width = ((uint32_t)b[5] << 8) | b[6];
        if (b.size() != 7 + (uint64_t)cnt * 4 * (width + 1))
            return fail(HNSWB200_EIO, "layer file length != 7 + nb_nodes*4*(m+1): a row above the cap was written un-truncated by the reference (graph.rs:172-178)");
        // cap as Layers::add_level makes it (layers.rs:50); equals `width` for reference-written files
        caps[l] = std::min<uint32_t>(width, (uint32_t)(l == 0 ? 2 * prm.m : prm.m));
        nn[l] = cnt;
        std::vector<std::pair<uint32_t, std::vector<uint32_t>>> rows(cnt);
        uint64_t o = 7;
        for (uint32_t r = 0; r < cnt; ++r) {
            rows[r].first = get_u32(&b[o]);
            o += 4;
            for (uint32_t j = 0; j < width; ++j) {
                uint32_t v = get_u32(&b[o + 4 * j]);
                if (v == 0xFFFFFFFFu) break;
                rows[r].second.push_back(v);
            }
            o += 4ull * width;
        }
        off[l].push_back(0);
        for (auto& r : rows) {
            ids[l].push_back(r.first);
            nb[l].insert(nb[l].end(), r.second.begin(), r.second.end());
            off[l].push_back(nb[l].size());
        }
        if (nb[l].empty()) nb[l].push_back(0);
    }
    std::vector<const uint32_t*> pid(nl), pnb(nl);
    std::vector<const uint64_t*> poff(nl);
    for (uint32_t l = 0; l < nl; ++l) { pid[l] = ids[l].data(); pnb[l] = nb[l].data(); poff[l] = off[l].data(); }
    hnswb200_points* P = nullptr;
    hnswb200_graph* G = nullptr;
    int rc = full ? hnswb200_points_upload_f32(c, vals.data(), levels.data(), n, (uint32_t)dim, &P)
                  : hnswb200_points_upload(c, codes.data(), mins.data(), deltas.data(), levels.data(), n, (uint32_t)dim, &P);
    if (rc) return rc;
    rc = hnswb200_graph_upload(c, n, nl, caps.data(), nn.data(), pid.data(), poff.data(), pnb.data(), &G);
    if (rc) { hnswb200_points_destroy(P); return rc; }
    rc = hnswb200_index_from_parts(c, P, G, &prm, out);
    if (rc) { hnswb200_points_destroy(P); hnswb200_graph_destroy(G); }
    return rc;
}

int64_t hnswb200_load_glove(const char* path, uint64_t lim, float* out, uint64_t cap, uint64_t* dim_out) {
    FILE* f = fopen(path, "r");
    if (!f) { set_error(std::string("cannot open ") + path); return HNSWB200_EIO; }
    char* line = nullptr;
    size_t lcap = 0;
    int64_t rows = 0;
    uint64_t dim = 0, w = 0;
    while (getline(&line, &lcap, f) > 0) {
        if (lim > 0 && (uint64_t)rows >= lim) break;
        char* save = nullptr;
        char* tok = strtok_r(line, " \n\r", &save);
        if (!tok) continue;
        uint64_t dcount = 0;
        while ((tok = strtok_r(nullptr, " \n\r", &save))) {
            char* end = nullptr;
            float v = strtof(tok, &end);  // correctly rounded, like Rust's parse::<f32>()
            if (end == tok || *end != '\0') continue;
            if (out && w < cap) out[w] = v;
            ++w;
            ++dcount;
        }
        if (rows == 0) dim = dcount;
        else if (dcount != dim) {
            set_error("Line " + std::to_string(rows + 1) + ": vector is not the same size as others.");
            free(line);
            fclose(f);
            return HNSWB200_EINVAL;
        }
        ++rows;
    }
    free(line);
    fclose(f);
    if (dim_out) *dim_out = dim;
    return rows;
}

}  // extern "C"
