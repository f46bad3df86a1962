// sm_100a kernels of the HNSW query / distance path.
//   K1 quantise_kernel        <- QuantVec::new                (vectors/src/quant.rs:41-66)
//   K2 dist_*_kernel          <- distance_unrolled, dist2many, Points::distance/distance2point
//                                (quant.rs:14-37, vectors/src/lib.rs:17-22, points/src/points.rs:86-101)
//   K2f dist_full_pairs_kernel<- FullVec::distance             (vectors/src/full.rs:23-29)
//   K3 search_kernel_reg      <- HNSW::ann_by_vector + search_layer, ef <= 256: result list in registers
//      search_kernel             (search_reg.cuh); any ef: result list in shared memory (search.cuh)
//                                (hnsw/src/template.rs:306-335, template/searcher.rs:23-103)
//   K5 bf_chunk/bf_merge      <- brute_force_nns / sort_by_distance (hnsw/src/helpers/glove.rs:73-109), exact
//                                CUDA-core pass; the tensor-core filter for 128-byte records lives in bf_tc.cu
//   K6 topk_merge_kernel      <- (no reference analogue) merge of per-shard top-k lists
// All kernels of this file are byte/gather work; the only dense contraction of the path is in bf_tc.cu.
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"
#include "search.cuh"
#include "search_reg.cuh"
#include "search_params.cuh"

namespace hb {

#define HB_DISPATCH_DIM(L, ...)                                                    \
    do {                                                                           \
        if ((L).kind == HB_REC_F32) { using Q = FullQuery; __VA_ARGS__; }          \
        else if ((L).dim == 100) { using Q = RegQuery<12, 4>; __VA_ARGS__; }       \
        else if ((L).dim == 128) { using Q = RegQuery<16, 0>; __VA_ARGS__; }       \
        else if ((L).dim == 96) { using Q = RegQuery<12, 0>; __VA_ARGS__; }        \
        else if ((L).dim == 50) { using Q = RegQuery<6, 2>; __VA_ARGS__; }         \
        else { using Q = SmemQuery; __VA_ARGS__; }                                 \
    } while (0)

static inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// the record's two aux floats (layout.h: hb_aux_offset): Sum y and Sum y^2 over the warp's partial sums
__device__ __forceinline__ void write_aux(uint8_t* rp, const RecLayout& L, float sy, float sy2, int lane) {
    const uint32_t ao = hb_aux_offset(L);
    if (ao == 0xFFFFFFFFu) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sy += __shfl_xor_sync(HB_FULL, sy, o);
        sy2 += __shfl_xor_sync(HB_FULL, sy2, o);
    }
    if (lane == 0) {
        *reinterpret_cast<float*>(rp + ao) = sy;
        *reinterpret_cast<float*>(rp + ao + 4) = sy2;
    }
}

// ---------------------------------------------------------------------------
// K1: quantise rows into lane-sliced records
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) quantise_kernel(const float* __restrict__ rows, uint64_t n,
                                                       RecLayout L, uint8_t* rec, uint8_t* codes,
                                                       float* mins, float* deltas, uint32_t* nan_flag) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t r = warp; r < n; r += nwarps) {
        const float* v = rows + r * L.dim;
        float mn, dl;
        uint8_t* crow = codes ? codes + r * L.dim : nullptr;
        // pass 1: bounds + flat codes (if requested)
        bool ok = warp_quantise(v, L.dim, lane, nullptr, crow, mn, dl);
        if (!ok && lane == 0 && nan_flag) atomicOr(nan_flag, 1u);
        if (rec) {
            uint8_t* rp = rec + r * L.stride;
            float sy = 0.0f, sy2 = 0.0f;  // aux floats (hb_aux_offset): filter inputs, any summation order will do
            for (uint32_t i = lane; i < L.dim; i += 32) {
                float b = __fadd_rn(__fdiv_rn(__fsub_rn(v[i], mn), dl), 0.5f);
                float f = fminf(fmaxf(floorf(b), 0.0f), 255.0f);
                rp[hb_code_offset(L, i)] = (uint8_t)(uint32_t)f;
                const float y = __fadd_rn(__fmul_rn(f, dl), mn);
                sy += y;
                sy2 += y * y;
            }
            write_aux(rp, L, sy, sy2, lane);
            if (lane == 0) {
                *reinterpret_cast<float*>(rp + hb_min_offset(L)) = mn;
                *reinterpret_cast<float*>(rp + hb_delta_offset(L)) = dl;
            }
        }
        if (lane == 0) {
            if (mins) mins[r] = mn;
            if (deltas) deltas[r] = dl;
        }
    }
}

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ codes,
                                                   const float* __restrict__ mins,
                                                   const float* __restrict__ deltas, uint64_t n,
                                                   RecLayout L, uint8_t* rec) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t r = warp; r < n; r += nwarps) {
        uint8_t* rp = rec + r * L.stride;
        const float mn = mins[r], dl = deltas[r];
        float sy = 0.0f, sy2 = 0.0f;
        for (uint32_t i = lane; i < L.dim; i += 32) {
            const uint8_t c = codes[r * L.dim + i];
            rp[hb_code_offset(L, i)] = c;
            const float y = __fadd_rn(__fmul_rn((float)c, dl), mn);
            sy += y;
            sy2 += y * y;
        }
        write_aux(rp, L, sy, sy2, lane);
        if (lane == 0) {
            *reinterpret_cast<float*>(rp + hb_min_offset(L)) = mn;
            *reinterpret_cast<float*>(rp + hb_delta_offset(L)) = dl;
        }
    }
}

__global__ void __launch_bounds__(256) unpack_kernel(const uint8_t* __restrict__ rec, uint64_t n,
                                                     RecLayout L, uint8_t* codes, float* mins,
                                                     float* deltas) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t r = warp; r < n; r += nwarps) {
        const uint8_t* rp = rec + r * L.stride;
        if (codes)
            for (uint32_t i = lane; i < L.dim; i += 32) codes[r * L.dim + i] = rp[hb_code_offset(L, i)];
        if (lane == 0) {
            if (mins) mins[r] = *reinterpret_cast<const float*>(rp + hb_min_offset(L));
            if (deltas) deltas[r] = *reinterpret_cast<const float*>(rp + hb_delta_offset(L));
        }
    }
}

// FullVec::new (vectors/src/full.rs:18-22): the row itself, here zero-padded to whole 16-float chunks
__global__ void __launch_bounds__(256) pack_f32_kernel(const float* __restrict__ rows, uint64_t n, RecLayout L,
                                                       uint8_t* rec, uint32_t* bad_flag) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t r = warp; r < n; r += nwarps) {
        float* rp = reinterpret_cast<float*>(rec + r * L.stride);
        bool bad = false;
        for (uint32_t i = lane; i < 16 * L.W; i += 32) {
            const float x = i < L.dim ? rows[r * L.dim + i] : 0.0f;
            bad |= !(fabsf(x) <= 3.4028234664e38f);
            rp[i] = x;
        }
        if (bad && bad_flag) atomicOr(bad_flag, 1u);
    }
}

// VecBase::get_vals (vectors/src/lib.rs:24-26): the values of a stored vector, natural order
__global__ void __launch_bounds__(256) record_values_kernel(const uint8_t* __restrict__ rec, uint64_t n, RecLayout L,
                                                            float* rows) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t r = warp; r < n; r += nwarps) {
        const uint8_t* rp = rec + r * L.stride;
        if (L.kind == HB_REC_F32) {
            for (uint32_t i = lane; i < L.dim; i += 32) rows[r * L.dim + i] = reinterpret_cast<const float*>(rp)[i];
        } else {
            const float mn = *reinterpret_cast<const float*>(rp + hb_min_offset(L));
            const float dl = *reinterpret_cast<const float*>(rp + hb_delta_offset(L));
            for (uint32_t i = lane; i < L.dim; i += 32)
                rows[r * L.dim + i] = __fadd_rn(__fmul_rn((float)rp[hb_code_offset(L, i)], dl), mn);
        }
    }
}

static inline int grid_for_warps(uint64_t nwarps_needed, int warps_per_block, int cap_blocks = 148 * 16) {
    uint64_t b = (nwarps_needed + warps_per_block - 1) / warps_per_block;
    if (b < 1) b = 1;
    if (b > (uint64_t)cap_blocks) b = cap_blocks;
    return (int)b;
}

cudaError_t launch_quantise(const float* rows, uint64_t n, const RecLayout& L, uint8_t* rec,
                            uint8_t* codes, float* mins, float* deltas, uint32_t* nan_flag,
                            cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    quantise_kernel<<<grid_for_warps(n, 8), 256, 0, st>>>(rows, n, L, rec, codes, mins, deltas, nan_flag);
    return cudaGetLastError();
}
cudaError_t launch_pack(const uint8_t* codes, const float* mins, const float* deltas, uint64_t n,
                        const RecLayout& L, uint8_t* rec, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    pack_kernel<<<grid_for_warps(n, 8), 256, 0, st>>>(codes, mins, deltas, n, L, rec);
    return cudaGetLastError();
}
cudaError_t launch_pack_f32(const float* rows, uint64_t n, const RecLayout& L, uint8_t* rec, uint32_t* bad_flag,
                            cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    pack_f32_kernel<<<grid_for_warps(n, 8), 256, 0, st>>>(rows, n, L, rec, bad_flag);
    return cudaGetLastError();
}
cudaError_t launch_record_values(const uint8_t* rec, uint64_t n, const RecLayout& L, float* rows, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    record_values_kernel<<<grid_for_warps(n, 8), 256, 0, st>>>(rec, n, L, rows);
    return cudaGetLastError();
}
cudaError_t launch_unpack(const uint8_t* rec, uint64_t n, const RecLayout& L, uint8_t* codes,
                          float* mins, float* deltas, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    unpack_kernel<<<grid_for_warps(n, 8), 256, 0, st>>>(rec, n, L, codes, mins, deltas);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K0: L2-normalise rows (cosine as L2 over unit vectors: |a-b|^2 = 2 - 2 cos).  No reference analogue (the reference has
// no cosine); arithmetic in the style of FullVec::distance: one strictly sequential f32 sum of squares,
// correctly rounded sqrt and division.  A zero row stays zero.  In place is allowed.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) normalise_rows_kernel(const float* __restrict__ rows, uint64_t n, uint32_t dim, float* out) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float* v = rows + r * dim;
    float s = 0.0f;
    for (uint32_t i = 0; i < dim; ++i) s = __fadd_rn(s, __fmul_rn(v[i], v[i]));
    const float nrm = __fsqrt_rn(s);
    float* o = out + r * dim;
    for (uint32_t i = 0; i < dim; ++i) o[i] = nrm > 0.0f ? __fdiv_rn(v[i], nrm) : v[i];
}

cudaError_t launch_normalise(const float* rows, uint64_t n, uint32_t dim, float* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    normalise_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(rows, n, dim, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K2: batched distances
// ---------------------------------------------------------------------------
// one f32 query vs ids[n].  distance2point(point, idx) = point.dist2other(points[idx])
template <class Q>
__global__ void __launch_bounds__(128) dist_query_many_kernel(const uint8_t* __restrict__ rec, RecLayout L,
                                                              const float* __restrict__ query,
                                                              const uint32_t* __restrict__ ids, uint64_t n,
                                                              float* out, uint32_t* nan_flag) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    const uint32_t qd_cap = (L.dim + 7) / 8 * 8 + 8;
    float* qd = reinterpret_cast<float*>(smem) + (size_t)wib * qd_cap;
    bool ok = warp_prepare_query(L, query, qd, lane);
    if (!ok && lane == 0 && nan_flag) atomicOr(nan_flag, 1u);
    __syncwarp();
    Q q;
    q.init(L, qd, gl);
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    // two rounds of 8 candidates per iteration: the loads of the second are in flight while the first is evaluated
    for (uint64_t base = warp * 16; base < n; base += nwarps * 16) {
        const uint64_t i0 = base + grp, i1 = base + 8 + grp;
        const bool a0 = i0 < n, a1 = i1 < n;
        const uint32_t id0 = __ldg(ids + (a0 ? i0 : base)), id1 = __ldg(ids + (a1 ? i1 : base));
        const typename Q::Rec r0 = Q::load(rec + (size_t)id0 * L.stride, gl);
        const typename Q::Rec r1 = Q::load(rec + (size_t)id1 * L.stride, gl);
        const float d0 = q.dist(r0, gl, gbase);
        const float d1 = q.dist(r1, gl, gbase);
        if (a0 && gl == 0) out[i0] = d0;
        if (a1 && gl == 0) out[i1] = d1;
    }
}

cudaError_t launch_dist_query_many(const uint8_t* rec, const RecLayout& L, const float* query,
                                   const uint32_t* ids, uint64_t n, float* out, uint32_t* nan_flag,
                                   cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const uint32_t qd_cap = (L.dim + 7) / 8 * 8 + 8;
    size_t smem = (size_t)4 * qd_cap * 4;
    int grid = grid_for_warps((n + 15) / 16, 4, 148 * 8);
    HB_DISPATCH_DIM(L, {
        cudaFuncSetAttribute(dist_query_many_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dist_query_many_kernel<Q><<<grid, 128, smem, st>>>(rec, L, query, ids, n, out, nan_flag);
    });
    return cudaGetLastError();
}

__global__ void __launch_bounds__(128) dist_pairs_kernel(const uint8_t* __restrict__ rec, RecLayout L,
                                                         const uint32_t* __restrict__ a,
                                                         const uint32_t* __restrict__ b, uint64_t n,
                                                         float* out) {
    const int lane = threadIdx.x & 31;
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t base = warp * 8; base < n; base += nwarps * 8) {
        uint64_t i = base + grp;
        bool act = i < n;
        uint64_t j = act ? i : base;
        float d = rec_rec_dist(L, rec + (size_t)__ldg(a + j) * L.stride, rec + (size_t)__ldg(b + j) * L.stride,
                               gl, gbase);
        if (act && gl == 0) out[i] = d;
    }
}

cudaError_t launch_dist_pairs(const uint8_t* rec, const RecLayout& L, const uint32_t* a,
                              const uint32_t* b, uint64_t n, float* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    dist_pairs_kernel<<<grid_for_warps((n + 7) / 8, 4, 148 * 8), 128, 0, st>>>(rec, L, a, b, n, out);
    return cudaGetLastError();
}

// jobs: out[j] = distance(src[job], ids[j]), j in [off[job], off[job+1]); one warp per job.
// Used by the build for the prune distances (hnsw/src/template.rs:228-230: x = to_prune, y = n).
template <class Q>
__global__ void __launch_bounds__(128) dist_one_to_many_kernel(const uint8_t* __restrict__ rec, RecLayout L,
                                                               const uint32_t* __restrict__ src,
                                                               const uint32_t* __restrict__ off,
                                                               const uint32_t* __restrict__ ids,
                                                               uint32_t njobs, float* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    const uint32_t qd_cap = (L.dim + 7) / 8 * 8 + 8;
    float* qd = reinterpret_cast<float*>(smem) + (size_t)wib * qd_cap;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + wib;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t job = warp; job < njobs; job += nwarps) {
        __syncwarp();
        warp_dequant_record(L, rec + (size_t)__ldg(src + job) * L.stride, lane, qd);
        __syncwarp();
        Q q;
        q.init(L, qd, gl);
        const uint32_t lo = __ldg(off + job), hi = __ldg(off + job + 1);
        for (uint32_t base = lo; base < hi; base += 8) {
            uint32_t i = base + grp;
            bool act = i < hi;
            uint32_t id = __ldg(ids + (act ? i : base));
            float d = q.dist(rec + (size_t)id * L.stride, gl, gbase);
            if (act && gl == 0) out[i] = d;
        }
    }
}

cudaError_t launch_dist_one_to_many(const uint8_t* rec, const RecLayout& L, const uint32_t* src,
                                    const uint32_t* off, const uint32_t* ids, uint32_t njobs,
                                    float* out, cudaStream_t st) {
    if (njobs == 0) return cudaSuccess;
    const uint32_t qd_cap = (L.dim + 7) / 8 * 8 + 8;
    size_t smem = (size_t)4 * qd_cap * 4;
    int grid = grid_for_warps(njobs, 4, 148 * 8);
    HB_DISPATCH_DIM(L, {
        cudaFuncSetAttribute(dist_one_to_many_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dist_one_to_many_kernel<Q><<<grid, 128, smem, st>>>(rec, L, src, off, ids, njobs, out);
    });
    return cudaGetLastError();
}

// K2f: FullVec::distance, one thread per pair, one strictly sequential chain
__global__ void __launch_bounds__(128) dist_full_pairs_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ y, uint64_t n,
                                                              uint32_t dim, float* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = x + i * dim;
    const float* b = y + i * dim;
    float s = 0.0f;
    for (uint32_t k = 0; k < dim; ++k) {
        float t = __fsub_rn(a[k], b[k]);
        s = __fadd_rn(s, __fmul_rn(t, t));
    }
    out[i] = __fsqrt_rn(s);
}

cudaError_t launch_dist_full_pairs(const float* x, const float* y, uint64_t n, uint32_t dim,
                                   float* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    dist_full_pairs_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(x, y, n, dim, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K3: HNSW search, one warp per query, persistent over the batch
// ---------------------------------------------------------------------------


template <class VIS>
__host__ __device__ inline size_t search_warp_smem(uint32_t kpl, uint32_t tbits, uint32_t qd_cap) {
    return (size_t)32 * kpl * 8 + VIS::bytes(tbits) + 128 + (size_t)qd_cap * 4;
}

__device__ __forceinline__ void make_vis(Vis16& v, unsigned char* mem, const SearchParams& p) {
    v.words = reinterpret_cast<uint32_t*>(mem);
    v.tbits = p.tbits;
    v.bbits = p.bbits;
}
__device__ __forceinline__ void make_vis(Vis16N& v, unsigned char* mem, const SearchParams& p) {
    v.words = reinterpret_cast<uint32_t*>(mem);
    v.bbits = p.bbits;
}
__device__ __forceinline__ void make_vis(Vis32& v, unsigned char* mem, const SearchParams& p) {
    v.tab = reinterpret_cast<uint32_t*>(mem);
    v.tbits = p.tbits;
}

template <class Q, class VIS, int KPL>
__global__ void __launch_bounds__(SEARCH_WPB * 32, 5) search_kernel(SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gl = lane & 3, gbase = lane & ~3;
    unsigned char* wsm = smem + (size_t)wib * search_warp_smem<VIS>(p.kpl, p.tbits, p.qd_cap);
    KeyList<KPL> L;
    L.list = reinterpret_cast<u64*>(wsm);
    L.kpl = (int)p.kpl;
    VIS vis;
    make_vis(vis, wsm + (size_t)32 * p.kpl * 8, p);
    uint32_t* newbuf = reinterpret_cast<uint32_t*>(wsm + (size_t)32 * p.kpl * 8 + VIS::bytes(p.tbits));
    float* qd = reinterpret_cast<float*>(newbuf + 32);

    while (true) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(p.work_counter, 1u);
        qi = __shfl_sync(HB_FULL, qi, 0);
        if (qi >= p.nq) break;
        __syncwarp();
        // Point::new(vector): the query becomes a point exactly like a stored one (template.rs:313)
        // one read of the f32 query (it may live in pinned host memory: see hnswb200_search), then quantise in place
        {
            const float* src = (qi < p.split ? p.queries : p.queries_tail) + (size_t)qi * p.L.dim;
            for (uint32_t i = lane; i < p.L.dim; i += 32) qd[i] = src[i];
        }
        __syncwarp();
        bool ok = warp_prepare_query(p.L, qd, qd, lane);
        __syncwarp();
        uint32_t* oid = p.out_ids + (size_t)qi * p.topn;
        float* od = p.out_dists ? p.out_dists + (size_t)qi * p.topn : nullptr;
        if (!ok) {  // NaN in query: the reference panics; report through flags and an empty result
            for (uint32_t j = lane; j < p.topn; j += 32) put_result(p, oid, od, qi, j, EMPTY_ID, INFINITY);
            if (lane == 0) {
                if (p.out_counts) p.out_counts[qi] = 0;
                if (p.out_hops) p.out_hops[qi] = 0;
                if (p.out_evals) p.out_evals[qi] = 0;
                if (p.out_flags) p.out_flags[qi] = 1u;
                if (p.out_nbrs) p.out_nbrs[qi] = 0;
                if (p.nan_any) *reinterpret_cast<volatile uint32_t*>(p.nan_any) = 1u;
            }
            continue;
        }
        Q q;
        q.init(p.L, qd, gl);
        SearchCounters cnt{0u, 1u, 0u, 0u};
        // selected <- {Dist(ep, distance2point(point, ep))}   (template.rs:316-319)
        float d0 = q.dist(p.rec + (size_t)p.ep * p.L.stride, gl, gbase);
        L.reset(lane);
        if (lane == 0) L.list[0] = make_key(d0, p.ep);
        __syncwarp();
        for (uint32_t layer = p.n_layers - 1; layer >= 1; --layer)  // template.rs:322-324
            search_layer<Q, VIS, KPL>(q, p.rec, p.L.stride, p.g, layer, L, vis, newbuf, 1, lane, cnt);
        search_layer<Q, VIS, KPL>(q, p.rec, p.L.stride, p.g, 0u, L, vis, newbuf, (int)p.ef, lane, cnt);  // :326
        // get_top_selected(n)   (results.rs:59-61)
        uint32_t got = 0;
        for (uint32_t j0 = 0; j0 < p.topn; j0 += 32) {
            uint32_t j = j0 + lane;
            u64 k = (j < p.topn && j < p.ef) ? L.list[j] : SENTINEL;
            bool real = k != SENTINEL;
            if (j < p.topn) {
                put_result(p, oid, od, qi, j, real ? (uint32_t)k : EMPTY_ID, real ? __uint_as_float((uint32_t)(k >> 32)) : INFINITY);
            }
            got += __popc(__ballot_sync(HB_FULL, real));
        }
        if (lane == 0) {
            if (p.out_counts) p.out_counts[qi] = got;
            if (p.out_hops) p.out_hops[qi] = cnt.hops;
            if (p.out_evals) p.out_evals[qi] = cnt.evals;
            if (p.out_flags) p.out_flags[qi] = cnt.overflow ? 2u : 0u;
            if (p.out_nbrs) p.out_nbrs[qi] = cnt.nbrs;
        }
    }
}

// The same query path on the register-resident list (ef <= 32*KPL): no shared-memory list, no
// speculation; latency is hidden by warps (6 blocks of 4 warps per SM) instead.
// per warp: visited table | 32 candidate ids | 32 admitted keys | merge buffer (32*KPL keys) | dequantised query.
// A register-resident query (RegQuery) needs its shared-memory copy only until init(): the merge buffer reuses it.
template <class VIS, class Q, int KPL>
__host__ __device__ inline size_t search_reg_warp_smem(uint32_t tbits, uint32_t qd_cap) {
    const size_t mb = (size_t)32 * KPL * 8, qb = ((size_t)qd_cap * 4 + 15) & ~(size_t)15;
    return VIS::bytes(tbits) + 128 + 256 + (Q::kKeepsSmem ? mb + qb : (mb > qb ? mb : qb));
}

#ifndef HB_REG_MINB2
#define HB_REG_MINB2 6  // resident blocks per SM the KPL=2 kernel is compiled for (register budget)
#endif
constexpr int reg_min_blocks(int kpl) { return kpl <= 2 ? HB_REG_MINB2 : kpl <= 4 ? 5 : 4; }
// the 3584-entry visited table exists to make room for a seventh block per SM (72 registers)
// (a query that stays in shared memory takes the room of the seventh block: six then, one more than with 4096 entries)
template <class VIS, class Q> constexpr int reg_min_blocks_for(int kpl) { return reg_min_blocks(kpl); }
template <> constexpr int reg_min_blocks_for<Vis16N, RegQuery<12, 4>>(int kpl) { return kpl <= 2 ? 7 : reg_min_blocks(kpl); }
template <> constexpr int reg_min_blocks_for<Vis16N, RegQuery<12, 0>>(int kpl) { return kpl <= 2 ? 7 : reg_min_blocks(kpl); }
template <> constexpr int reg_min_blocks_for<Vis16N, RegQuery<16, 0>>(int kpl) { return kpl <= 2 ? 7 : reg_min_blocks(kpl); }
template <> constexpr int reg_min_blocks_for<Vis16N, RegQuery<6, 2>>(int kpl) { return kpl <= 2 ? 7 : reg_min_blocks(kpl); }

template <class Q, class VIS, int KPL, bool STATS>
__global__ void __launch_bounds__(SEARCH_WPB * 32, reg_min_blocks_for<VIS, Q>(KPL)) search_kernel_reg(SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gl = lane & 3, gbase = lane & ~3;
    unsigned char* wsm = smem + (size_t)wib * search_reg_warp_smem<VIS, Q, KPL>(p.tbits, p.qd_cap);
    VIS vis;
    make_vis(vis, wsm, p);
    uint32_t* newbuf = reinterpret_cast<uint32_t*>(wsm + VIS::bytes(p.tbits));
    u64* kbuf = reinterpret_cast<u64*>(newbuf + 32);
    u64* mbuf = kbuf + 32;
    float* qd = reinterpret_cast<float*>(Q::kKeepsSmem ? mbuf + 32 * KPL : mbuf);
    // A following search (launched as programmatic dependent) reads nothing this grid writes: let its blocks
    // take over the SMs as soon as this grid's blocks retire, instead of waiting for the last long query.
    asm volatile("griddepcontrol.launch_dependents;");

    while (true) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(p.work_counter, 1u);
        qi = __shfl_sync(HB_FULL, qi, 0);
        if (qi >= p.nq) break;
        __syncwarp();
        // Point::new(vector): the query becomes a point exactly like a stored one (template.rs:313)
        // one read of the f32 query (it may live in pinned host memory: see hnswb200_search), then quantise in place
        {
            const float* src = (qi < p.split ? p.queries : p.queries_tail) + (size_t)qi * p.L.dim;
            for (uint32_t i = lane; i < p.L.dim; i += 32) qd[i] = src[i];
        }
        __syncwarp();
        bool ok = warp_prepare_query(p.L, qd, qd, lane);
        __syncwarp();
        uint32_t* oid = p.out_ids + (size_t)qi * p.topn;
        float* od = p.out_dists ? p.out_dists + (size_t)qi * p.topn : nullptr;
        if (!ok) {  // NaN in query: the reference panics; report through flags and an empty result
            for (uint32_t j = lane; j < p.topn; j += 32) put_result(p, oid, od, qi, j, EMPTY_ID, INFINITY);
            if (lane == 0) {
                if (p.out_counts) p.out_counts[qi] = 0;
                if (p.out_hops) p.out_hops[qi] = 0;
                if (p.out_evals) p.out_evals[qi] = 0;
                if (p.out_flags) p.out_flags[qi] = 1u;
                if (p.out_nbrs) p.out_nbrs[qi] = 0;
                if (p.nan_any) *reinterpret_cast<volatile uint32_t*>(p.nan_any) = 1u;
            }
            continue;
        }
        Q q;
        q.init(p.L, qd, gl);
        SearchCounters cnt{0u, 0u, 0u, 0u};
        RegList<KPL> L;
        __syncwarp();  // every lane has copied its part of qd before the merge buffer may overwrite it
        search_query_reg<Q, VIS, KPL, STATS>(q, p.rec, p.L.stride, p.g, p.n_layers, p.ep, L, vis, newbuf, kbuf, mbuf, (int)p.ef, lane, cnt);
        // get_top_selected(n)   (results.rs:59-61): position lane*KPL + s
        uint32_t mine = 0;
#pragma unroll
        for (int s = 0; s < KPL; ++s) {
            const uint32_t j = (uint32_t)(lane * KPL + s);
            if (j < p.topn) {
                const u64 k = L.v[s];
                const bool real = (j < p.ef) && (k != RSENT);
                put_result(p, oid, od, qi, j, real ? rkey_id(k) : EMPTY_ID, real ? __uint_as_float((uint32_t)(k >> 32)) : INFINITY);
                mine += real ? 1u : 0u;
            }
        }
        for (uint32_t j = 32 * KPL + lane; j < p.topn; j += 32) put_result(p, oid, od, qi, j, EMPTY_ID, INFINITY);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(HB_FULL, mine, o);
        if (lane == 0) {
            if (p.out_counts) p.out_counts[qi] = mine;
            if (p.out_hops) p.out_hops[qi] = cnt.hops;
            if (p.out_evals) p.out_evals[qi] = cnt.evals;
            if (p.out_flags) p.out_flags[qi] = cnt.overflow ? 2u : 0u;
            if (p.out_nbrs) p.out_nbrs[qi] = cnt.nbrs;
        }
    }
}

template <class Q, class VIS, int KPL, bool STATS>
static cudaError_t launch_search_reg_s(const SearchParams& p, int num_sms, cudaStream_t st, bool overlap_previous) {
    size_t smem = search_reg_warp_smem<VIS, Q, KPL>(p.tbits, p.qd_cap) * SEARCH_WPB;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    // function attributes and occupancy are per device (a process may drive several GPUs)
    static int occ_cache_d[64] = {};
    static size_t occ_smem_d[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int& occ_cache = occ_cache_d[dev & 63];
    size_t& occ_smem = occ_smem_d[dev & 63];
    cudaError_t e;
    if (occ_cache == 0 || occ_smem != smem) {
        e = cudaFuncSetAttribute(search_kernel_reg<Q, VIS, KPL, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, search_kernel_reg<Q, VIS, KPL, STATS>, SEARCH_WPB * 32, smem);
        if (e != cudaSuccess) return e;
        occ_cache = occ < 1 ? 1 : occ;
        occ_smem = smem;
        if (getenv("HNSWB200_DEBUG_LAUNCH"))
            fprintf(stderr, "[hnswb200 search] KPL=%d visited bytes/warp=%zu smem/block=%zu blocks/SM=%d\n", KPL,
                    VIS::bytes(p.tbits), smem, occ_cache);
    }
    uint64_t want = ((uint64_t)p.nq + SEARCH_WPB - 1) / SEARCH_WPB;
    int occ = occ_cache;
    if (const char* ev = getenv("HNSWB200_SEARCH_BLOCKS_PER_SM")) {  // experiment knob
        int v = atoi(ev);
        if (v >= 1 && v < occ) occ = v;
    }
    uint64_t cap = (uint64_t)num_sms * occ;
    int grid = (int)(want < cap ? want : cap);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(SEARCH_WPB * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = overlap_previous ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, search_kernel_reg<Q, VIS, KPL, STATS>, p);
}

template <class Q, class VIS, int KPL>
static cudaError_t launch_search_t(const SearchParams& p, int num_sms, cudaStream_t st) {
    size_t smem = search_warp_smem<VIS>(p.kpl, p.tbits, p.qd_cap) * SEARCH_WPB;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    // function attributes and occupancy are per device (a process may drive several GPUs)
    static int occ_cache_d[64] = {};
    static size_t occ_smem_d[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int& occ_cache = occ_cache_d[dev & 63];
    size_t& occ_smem = occ_smem_d[dev & 63];
    cudaError_t e;
    if (occ_cache == 0 || occ_smem != smem) {
        e = cudaFuncSetAttribute(search_kernel<Q, VIS, KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, search_kernel<Q, VIS, KPL>, SEARCH_WPB * 32, smem);
        if (e != cudaSuccess) return e;
        occ_cache = occ < 1 ? 1 : occ;
        occ_smem = smem;
    }
    uint64_t want = ((uint64_t)p.nq + SEARCH_WPB - 1) / SEARCH_WPB;
    uint64_t cap = (uint64_t)num_sms * occ_cache;
    int grid = (int)(want < cap ? want : cap);
    search_kernel<Q, VIS, KPL><<<grid, SEARCH_WPB * 32, smem, st>>>(p);
    return cudaGetLastError();
}

// the counters (hops, evaluations, neighbour ids, overflow flag) are optional outputs: a caller that asks for none of
// them runs the variant that does not keep them
template <class Q, class VIS, int KPL>
static cudaError_t launch_search_reg_t(const SearchParams& p, int num_sms, cudaStream_t st, bool overlap_previous) {
    if (p.out_hops || p.out_evals || p.out_flags || p.out_nbrs)
        return launch_search_reg_s<Q, VIS, KPL, true>(p, num_sms, st, overlap_previous);
    return launch_search_reg_s<Q, VIS, KPL, false>(p, num_sms, st, overlap_previous);
}

// visited-table geometry shared by the query and the build kernels
void choose_visited(uint32_t ef, uint32_t S0, uint64_t n_points, uint32_t* tbits, uint32_t* bbits, bool* use16) {
    uint32_t bb = 1;
    while (bb < 31 && (1ull << bb) < n_points) ++bb;
    // observed on the C2 workload (S0 = 32): ~390 evaluations per query at ef = 10, ~830 at ef = 64,
    // ~980 at ef = 100 (p99 about 1.7x the mean); aim at a mean load factor around 0.25
    uint64_t want = (uint64_t)ef * (S0 ? S0 : 32) * 3 / 4 + 1024;
    uint32_t tb = 9;
    while ((1ull << tb) < want && tb < 16) ++tb;
    if (const char* ev = getenv("HNSWB200_VIS_SLOTS")) {  // test knob: force the overflow fallback
        uint32_t v = (uint32_t)strtoul(ev, nullptr, 10);
        if (v >= 64 && (v & (v - 1)) == 0) { tb = 0; while ((1u << tb) < v) ++tb; }
    }
    bool u16 = !getenv("HNSWB200_VIS32");
    if (u16) {
        if (bb < tb) bb = tb;
        if (bb > tb + 12) {
            if (bb - 12 <= 14) tb = bb - 12;  // a larger table makes the remainder fit 12 bits
            else u16 = false;
        }
    }
    *tbits = tb;
    *bbits = bb;
    *use16 = u16;
}

cudaError_t launch_search(const SearchLaunch& a, int num_sms, cudaStream_t st) {
    if (a.nq == 0) return cudaSuccess;
    SearchParams p;
    p.rec = a.rec;
    p.L = a.L;
    p.g.adj0 = a.g.adj0; p.g.S0 = a.g.S0; p.g.upper_off = a.g.upper_off; p.g.upper_adj = a.g.upper_adj; p.g.SU = a.g.SU;
    p.n_layers = a.g.n_layers; p.ep = a.ep;
    p.queries = a.queries; p.nq = a.nq; p.topn = a.topn; p.ef = a.ef;
    p.queries_tail = a.queries_tail ? a.queries_tail : a.queries;
    p.split = a.queries_tail ? a.split : a.nq;
    p.qd_cap = round_up(a.L.dim, 8) + 8;
    p.out_ids = a.out_ids; p.out_dists = a.out_dists; p.out_counts = a.out_counts;
    p.out_hops = a.out_hops; p.out_evals = a.out_evals; p.out_flags = a.out_flags; p.out_nbrs = a.out_nbrs;
    p.work_counter = a.work_counter;
    p.nan_any = a.nan_any;
    p.n_peers = a.n_peers < HB_MAX_PEERS ? a.n_peers : HB_MAX_PEERS;
    for (uint32_t g = 0; g < HB_MAX_PEERS; ++g) {
        p.peer_ids[g] = g < p.n_peers ? a.peer_ids[g] : nullptr;
        p.peer_dists[g] = g < p.n_peers ? a.peer_dists[g] : nullptr;
    }
    p.peer_row0 = a.peer_row0;
    p.id_offset = a.id_offset;
    bool use16;
    choose_visited(a.ef, a.g.S0, a.n_points, &p.tbits, &p.bbits, &use16);
    const bool generic_list = a.ef > 256 || getenv("HNSWB200_GENERAL_PATH");  // env: test knob
    p.kpl = round_up((a.ef + 31) / 32, 2);
    if (!generic_list && search_fast_supported(a.L, a.ef, a.n_points)) {
        // the second-generation kernel (csrc/search_fast.cuh): bucketed visited set, one reduction per batch
        if (!a.counter_is_fresh) {
            cudaError_t e0 = cudaMemsetAsync(a.work_counter, 0, sizeof(uint32_t), st);
            if (e0 != cudaSuccess) return e0;
        }
        return launch_search_fast(p, a.n_points, num_sms, st, a.overlap_previous, a.spill_ws, a.spill_cap, a.spill_warps);
    }
    {
        char nm[160];
        const char* qn = a.L.kind == HB_REC_F32 ? "FullQuery" : (a.L.dim == 100 || a.L.dim == 128 || a.L.dim == 96 || a.L.dim == 50) ? "RegQuery" : "SmemQuery";
        const bool u16n = use16 && a.ef <= 64 && p.tbits == 12 && Vis16N::fits(p.bbits) && !getenv("HNSWB200_VIS_POW2");
        snprintf(nm, sizeof nm, "hb::%s<%s dim %u,%s,ef<=%u>", generic_list ? "search_kernel" : "search_kernel_reg", qn, a.L.dim,
                 !use16 ? "Vis32" : (u16n && !generic_list) ? "Vis16N" : "Vis16", generic_list ? a.ef : (a.ef <= 64 ? 64u : a.ef <= 128 ? 128u : 256u));
        set_search_variant(nm);
    }
    if (!generic_list) {
        // register-resident list: ef <= 64 / 128 / 256 -> 2 / 4 / 8 keys per lane
        const uint32_t rkpl = a.ef <= 64 ? 2 : a.ef <= 128 ? 4 : 8;
        auto rbytes = [&](uint32_t tb) {  // upper bound over the query classes
            return ((use16 ? Vis16::bytes(tb) : Vis32::bytes(tb)) + 128 + 256 + (size_t)32 * rkpl * 8 + (size_t)p.qd_cap * 4 + 16) * SEARCH_WPB;
        };
        while (rbytes(p.tbits) > 200 * 1024 && p.tbits > 9 && (!use16 || p.bbits <= p.tbits - 1 + 12)) --p.tbits;
        if (!a.counter_is_fresh) {
            cudaError_t e0 = cudaMemsetAsync(a.work_counter, 0, sizeof(uint32_t), st);
            if (e0 != cudaSuccess) return e0;
        }
        // ef <= 64 with the default 4096-entry table: the 3584-entry table and a seventh block per SM instead (Vis16N).
        // A query that lives in registers gets 7 blocks per SM, one that stays in shared memory 6 (5 with 4096 entries).
        const bool use16n = use16 && a.ef <= 64 && p.tbits == 12 && Vis16N::fits(p.bbits) && !getenv("HNSWB200_VIS_POW2");
        HB_DISPATCH_DIM(a.L, {
            if (use16n) return launch_search_reg_t<Q, Vis16N, 2>(p, num_sms, st, a.overlap_previous);
            if (use16) {
                if (a.ef <= 64) return launch_search_reg_t<Q, Vis16, 2>(p, num_sms, st, a.overlap_previous);
                if (a.ef <= 128) return launch_search_reg_t<Q, Vis16, 4>(p, num_sms, st, a.overlap_previous);
                return launch_search_reg_t<Q, Vis16, 8>(p, num_sms, st, a.overlap_previous);
            }
            if (a.ef <= 64) return launch_search_reg_t<Q, Vis32, 2>(p, num_sms, st, a.overlap_previous);
            if (a.ef <= 128) return launch_search_reg_t<Q, Vis32, 4>(p, num_sms, st, a.overlap_previous);
            return launch_search_reg_t<Q, Vis32, 8>(p, num_sms, st, a.overlap_previous);
        });
        return cudaErrorUnknown;
    }
    // shrink the visited table if one block would not fit (the overflow fallback keeps results exact)
    auto bytes = [&](uint32_t tb) {
        return (use16 ? search_warp_smem<Vis16>(p.kpl, tb, p.qd_cap) : search_warp_smem<Vis32>(p.kpl, tb, p.qd_cap)) * SEARCH_WPB;
    };
    while (bytes(p.tbits) > 200 * 1024 && p.tbits > 9 && (!use16 || p.bbits <= p.tbits - 1 + 12)) --p.tbits;
    if (!a.counter_is_fresh) {
        cudaError_t e = cudaMemsetAsync(a.work_counter, 0, sizeof(uint32_t), st);
        if (e != cudaSuccess) return e;
    }
    // ef > 256: sorted list in shared memory, runtime width
    HB_DISPATCH_DIM(a.L, {
        if (use16) return launch_search_t<Q, Vis16, 0>(p, num_sms, st);
        return launch_search_t<Q, Vis32, 0>(p, num_sms, st);
    });
    return cudaErrorUnknown;
}

// ---------------------------------------------------------------------------
// K5: exact brute-force top-k under the quantised metric (CUDA-core version)
// Base-stationary: a warp keeps one base record dequantised in registers and
// streams the (L2-resident) query records past it, 8 per round.
// Orientation as the reference: x = query (self), y = base (glove.rs:99-101).
// ---------------------------------------------------------------------------
template <class Q>
__global__ void __launch_bounds__(128) bf_chunk_kernel(const uint8_t* __restrict__ base_rec, RecLayout L,
                                                       uint64_t b0, uint64_t b1, uint32_t id_offset,
                                                       const uint8_t* __restrict__ qrec, uint32_t nq,
                                                       const uint64_t* __restrict__ tau, uint64_t* buf,
                                                       uint32_t cap, uint32_t* cnt, uint32_t* overflow) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    const uint32_t qd_cap = (L.dim + 7) / 8 * 8 + 8;
    float* qd = reinterpret_cast<float*>(smem) + (size_t)wib * qd_cap;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t b = b0 + warp; b < b1; b += nwarps) {
        __syncwarp();
        warp_dequant_record(L, base_rec + b * L.stride, lane, qd);
        __syncwarp();
        Q q;
        q.init(L, qd, gl);
        const uint32_t gid = (uint32_t)b + id_offset;
        for (uint32_t r0 = 0; r0 < nq; r0 += 8) {
            uint32_t qi = r0 + grp;
            bool act = qi < nq;
            float d = q.dist(qrec + (size_t)(act ? qi : r0) * L.stride, gl, gbase);
            if (act && gl == 0) {
                uint64_t key = make_key(d, gid);
                if (key < __ldg(tau + qi)) {
                    uint32_t pos = atomicAdd(cnt + qi, 1u);
                    if (pos < cap) buf[(size_t)qi * cap + pos] = key;
                    else atomicOr(overflow, 1u);
                }
            }
        }
    }
}

__device__ __forceinline__ void block_bitonic_sort(uint64_t* a, uint32_t P) {
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    bool up = (i & k) == 0;
                    uint64_t x = a[i], y = a[ixj];
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// per query: top-k <- k smallest of (top-k U buffer); tau <- k-th key (or +inf); cnt <- 0
__global__ void __launch_bounds__(256) bf_merge_kernel(uint64_t* topk, uint32_t k, uint64_t* tau,
                                                       uint64_t* buf, uint32_t cap, uint32_t* cnt,
                                                       uint32_t P) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t* a = reinterpret_cast<uint64_t*>(smem);
    const uint32_t qi = blockIdx.x;
    uint32_t m = min(cnt[qi], cap);
    // sort only as wide as this query needs (P is the allocation: next_pow2(k + cap))
    uint32_t Pq = 1;
    while (Pq < k + m) Pq <<= 1;
    P = Pq;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        uint64_t v = ~0ull;
        if (i < k) v = topk[(size_t)qi * k + i];
        else if (i - k < m) v = buf[(size_t)qi * cap + (i - k)];
        a[i] = v;
    }
    __syncthreads();
    block_bitonic_sort(a, P);
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) topk[(size_t)qi * k + i] = a[i];
    if (threadIdx.x == 0) {
        tau[qi] = a[k - 1];  // ~0 (accept everything) while fewer than k keys are known
        cnt[qi] = 0;
    }
}

__global__ void keys_to_out_kernel(const uint64_t* __restrict__ topk, uint64_t total, uint32_t* ids,
                                   float* dists) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint64_t k = topk[i];
    if (k == ~0ull) {
        ids[i] = EMPTY_ID;
        if (dists) dists[i] = INFINITY;
    } else {
        ids[i] = (uint32_t)k;
        if (dists) dists[i] = __uint_as_float((uint32_t)(k >> 32));
    }
}

cudaError_t launch_bf_chunk(const uint8_t* base_rec, const RecLayout& L, uint64_t b0, uint64_t b1,
                            uint32_t id_offset, const uint8_t* qrec, uint32_t nq, const uint64_t* tau,
                            uint64_t* buf, uint32_t cap, uint32_t* cnt, uint32_t* overflow,
                            cudaStream_t st) {
    if (b1 <= b0 || nq == 0) return cudaSuccess;
    const uint32_t qd_cap = (L.dim + 7) / 8 * 8 + 8;
    size_t smem = (size_t)4 * qd_cap * 4;
    int grid = grid_for_warps(b1 - b0, 4, 148 * 12);
    HB_DISPATCH_DIM(L, {
        cudaFuncSetAttribute(bf_chunk_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        bf_chunk_kernel<Q><<<grid, 128, smem, st>>>(base_rec, L, b0, b1, id_offset, qrec, nq, tau, buf, cap, cnt, overflow);
    });
    return cudaGetLastError();
}

static inline uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

cudaError_t launch_bf_merge(uint64_t* topk, uint32_t k, uint64_t* tau, uint64_t* buf, uint32_t cap,
                            uint32_t* cnt, uint32_t nq, cudaStream_t st) {
    if (nq == 0) return cudaSuccess;
    uint32_t P = next_pow2(k + cap);
    size_t smem = (size_t)P * 8;
    cudaError_t e = cudaFuncSetAttribute(bf_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bf_merge_kernel<<<nq, 256, smem, st>>>(topk, k, tau, buf, cap, cnt, P);
    return cudaGetLastError();
}

cudaError_t launch_keys_to_out(const uint64_t* topk, uint32_t k, uint32_t nq, uint32_t* ids,
                               float* dists, cudaStream_t st) {
    uint64_t total = (uint64_t)k * nq;
    if (total == 0) return cudaSuccess;
    keys_to_out_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(topk, total, ids, dists);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K6: merge G per-shard sorted top-k lists per query under (dist, id) order
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) topk_merge_kernel(const uint32_t* __restrict__ ids,
                                                         const float* __restrict__ dists, uint32_t G,
                                                         uint32_t nq, uint32_t k, uint32_t P,
                                                         uint32_t* out_ids, float* out_dists) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t* a = reinterpret_cast<uint64_t*>(smem);
    const uint32_t qi = blockIdx.x;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        uint64_t v = ~0ull;
        if (i < G * k) {
            uint32_t g = i / k, j = i % k;
            size_t src = ((size_t)g * nq + qi) * k + j;
            uint32_t id = ids[src];
            if (id != EMPTY_ID) v = make_key(dists[src], id);
        }
        a[i] = v;
    }
    __syncthreads();
    block_bitonic_sort(a, P);
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
        uint64_t v = a[i];
        out_ids[(size_t)qi * k + i] = (v == ~0ull) ? EMPTY_ID : (uint32_t)v;
        if (out_dists) out_dists[(size_t)qi * k + i] = (v == ~0ull) ? INFINITY : __uint_as_float((uint32_t)(v >> 32));
    }
}

cudaError_t launch_topk_merge(const uint32_t* ids, const float* dists, uint32_t G, uint32_t nq,
                              uint32_t k, uint32_t* out_ids, float* out_dists, cudaStream_t st) {
    if (nq == 0 || k == 0) return cudaSuccess;
    uint32_t P = next_pow2(G * k);
    size_t smem = (size_t)P * 8;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    topk_merge_kernel<<<nq, 128, smem, st>>>(ids, dists, G, nq, k, P, out_ids, out_dists);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Peer exchange over NVLink / NVSwitch without a collective library: a rank copies its result rows into the
// gather buffers of the other ranks (peer memory mapped with hnswb200_ipc_open), then raises a flag word in each of
// them; the consumer waits for the flags of all ranks before it merges.  Stream order makes the rows visible before
// the flag: the signal kernel starts after the kernel that stored them has completed.
// ---------------------------------------------------------------------------
struct PeerPtrs { void* p[HB_MAX_PEERS]; };

__global__ void __launch_bounds__(256) peer_put_kernel(const uint4* __restrict__ src, uint64_t n16, PeerPtrs dst, uint32_t n_peers) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 v = __ldg(src + i);
        for (uint32_t g = 0; g < n_peers; ++g) reinterpret_cast<uint4*>(dst.p[g])[i] = v;
    }
}
__global__ void peer_signal_kernel(PeerPtrs flags, uint32_t n_peers, uint32_t slot, uint32_t epoch) {
    if (threadIdx.x < n_peers) {
        __threadfence_system();
        uint32_t* f = reinterpret_cast<uint32_t*>(flags.p[threadIdx.x]) + slot;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
    }
}
// flags[0..n) >= epoch (wrap-safe), or give up after ~10 s and raise *status = 2
__global__ void peer_wait_kernel(const uint32_t* flags, uint32_t n, uint32_t epoch, uint32_t* status) {
    if (threadIdx.x >= n) return;
    const long long t0 = clock64();
    while (true) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if (clock64() - t0 > 20000000000ll) {
            if (status) *reinterpret_cast<volatile uint32_t*>(status) = 2u;
            break;
        }
        __nanosleep(200);
    }
}

cudaError_t launch_peer_put(const void* src, uint64_t bytes, uint32_t n_peers, void* const* dst, int num_sms, cudaStream_t st) {
    if (!bytes || !n_peers) return cudaSuccess;
    PeerPtrs pp{};
    for (uint32_t g = 0; g < n_peers; ++g) pp.p[g] = dst[g];
    const uint64_t n16 = bytes / 16;
    uint64_t blocks = (n16 + 255) / 256;
    if (blocks > (uint64_t)num_sms * 4) blocks = (uint64_t)num_sms * 4;
    peer_put_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), n16, pp, n_peers);
    return cudaGetLastError();
}
cudaError_t launch_peer_signal(uint32_t n_peers, uint32_t* const* flags, uint32_t slot, uint32_t epoch, cudaStream_t st) {
    if (!n_peers) return cudaSuccess;
    PeerPtrs pp{};
    for (uint32_t g = 0; g < n_peers; ++g) pp.p[g] = flags[g];
    peer_signal_kernel<<<1, 32, 0, st>>>(pp, n_peers, slot, epoch);
    return cudaGetLastError();
}
cudaError_t launch_peer_wait(const uint32_t* flags, uint32_t n, uint32_t epoch, uint32_t* status, cudaStream_t st) {
    if (!n) return cudaSuccess;
    peer_wait_kernel<<<1, 32, 0, st>>>(flags, n, epoch, status);
    return cudaGetLastError();
}

}  // namespace hb
