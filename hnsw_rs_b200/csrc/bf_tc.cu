// K5 on the tensor cores: brute_force_nns (hnsw/src/helpers/glove.rs:73-109) as a tcgen05 / TMA
// dense contraction with an exact re-rank.
//
// The reference ranks every (query, base) pair by the quantised distance of vectors/src/quant.rs:14-37.
// With x_i = cq_i*dq + mq and y_i = cb_i*db + mb the squared distance is, algebraically,
//     d^2 = Sum x^2 + Sum y^2 - 2 (dq*db * <cq, cb> + mq * (db*Sum cb + dim*mb) + (dq*Sum cq) * mb)
// and the only O(dim) term is the integer dot product <cq, cb> of the u8 codes, which
// tcgen05.mma kind::i8 computes EXACTLY (s32 accumulators in TMEM; 128 * 255^2 < 2^23).
// The operands are the 128-byte lane-sliced records themselves (csrc/layout.h): a dot product does not
// care about the order of its terms, so the permuted code bytes need no second copy of the base; only
// the query side is copied once with its non-code bytes (min, delta, padding) zeroed.
//
// (Experiment, compiled with -DHB_TC_CENTRED=1, off by default -- see the note at HB_TC_CENTRED below.)
// Both sides are CENTRED at code 128: A holds (cq - 128) as s8 (code ^ 0x80), B the raw u8 record bytes, so the tensor
// core delivers dotm = Sum (cq-128)*cb, and the epilogue subtracts the per-query integer 128 * Sum (cq-128) to get
// dot'' = Sum (cq-128)*(cb-128) exactly.  With x_i = (cq_i-128)*dq + xmid, y_i = (cb_i-128)*db + ymid (xmid, ymid: the
// values of code 128, near the vectors' means)
//     d^2 = Sum x^2 + Sum y^2 - 2 (dq*db*dot'' + xmid * P + ymid * Sum x),      P = db * Sum (cb-128)
// every term next to the dot product is small (no large offsets that cancel against each other), so a whole group of
// 32 base columns can be rejected with ONE integer maximum against a per-(query, group) bound -- the float estimate
// runs only for groups that may hold a survivor (round 1 ran it for every pair: 7 instructions per pair, tensor pipe
// 14 % busy).
//
// The algebraic value differs from the reference's separately rounded chain by a few 1e-7 relative to
// Sum x^2 + Sum y^2, so it is used as a FILTER: a pair survives if its estimate is within a safety margin
// of the query's current k-th exact distance; survivors (about k per query per doubling of the base) are
// re-evaluated with the exact arithmetic (csrc/dist.cuh) and merged under (dist, id) by bf_merge_kernel.
// Top-k ids and distances are therefore bit-identical to the CUDA-core path and to the oracle.
//
// Kernel shape (one CTA per SM, 2 + 16 warps): a tile of 256 base records stays in shared memory (TMA,
// SWIZZLE_128B, K-major); tiles of 128 queries stream through a 3-stage TMA ring; one thread issues
// 4 x tcgen05.mma (M=128, N=256, K=32) per query tile into one of two 256-column TMEM accumulators;
// 16 epilogue warps read the accumulator with tcgen05.ld (32 lanes x 32 columns), form the estimate with
// packed f32x2 FMAs and append survivors to the per-query candidate lists.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "kernels.h"
#include "search.cuh"

namespace hb {

constexpr int TC_M = 128;       // queries per tile (TMEM lanes)
constexpr int TC_N = 256;       // base records per tile (TMEM columns)
constexpr int TC_K = 128;       // bytes per record = K extent
constexpr int TC_STAGES = 3;    // query-tile ring
#ifndef HB_TC_EPI_WARPS
#define HB_TC_EPI_WARPS 16
#endif
constexpr int TC_EPI_WARPS = HB_TC_EPI_WARPS;     // 4 TMEM lane quarters x (TC_EPI_WARPS / 4) column groups
constexpr int TC_COLS_PER_WARP = TC_N / (TC_EPI_WARPS / 4);
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr uint32_t TC_A_BYTES = TC_M * TC_K;  // 16 KB
constexpr uint32_t TC_B_BYTES = TC_N * TC_K;  // 32 KB
constexpr int TC_TRANSPOSE_MAX = 8;  // up to this many failing rows of a (warp, column group) are served one by one
constexpr float TC_EPS = 1e-5f;    // slack relative to Sum x^2 + Sum y^2
constexpr float TC_DELTA = 1e-4f;  // slack relative to the squared threshold

struct TcSmem {
    uint8_t b[TC_B_BYTES];              // 1024-aligned (swizzle atom)
    uint8_t a[TC_STAGES][TC_A_BYTES];
    float cu[TC_N], cv[TC_N], cw[TC_N], cb[TC_N];  // per-column constants
    // per group of 32 columns: min Bc', [min, max] of v and w, max u (quick reject of the whole group)
    float g_bmin[TC_N / 32], g_vmin[TC_N / 32], g_vmax[TC_N / 32], g_wmin[TC_N / 32], g_wmax[TC_N / 32], g_umax[TC_N / 32];
    int stage[TC_EPI_WARPS][32];  // one failing query row's 32 accumulators, read back one column per lane
    unsigned long long b_full, a_full[TC_STAGES], a_empty[TC_STAGES], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// wait of the single-thread roles (TMA producer, MMA issuer): they are far ahead of the epilogue, so they
// sleep between polls instead of taking issue slots from the epilogue warps of their scheduler
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(256);
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, unsigned long long* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
        "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// UMMA shared-memory descriptor: K-major, SWIZZLE_128B, rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t tc_smem_desc(const void* p) {
    const uint32_t lo = ((smem_u32(p) & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t hi = 64u | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// instruction descriptor: dense, D = s32 (bits 4-5 = 2), A = s8 (bits 7-9 = 1: the centred query codes), B = u8 (bits 10-12 = 0),
// both K-major, N = 256, M = 128
// HB_TC_CENTRED = 1 builds the centred operands and the per-(query, 32-column group) integer quick reject described in the
// header comment.  MEASURED AND NOT ADOPTED (profiles/r02_k5_experiments.txt): on the C4 workload the bound rejects 75-90 %
// of the (query, group) pairs, but a branch costs the whole warp and all 32 query rows of a warp pass together only ~6 % of
// the time; serving the failing rows one by one (one lane per column) makes the epilogue wait-bound instead.  Both variants
// are bit-identical to the oracle (the full brute-force suite passes with either), and both are slower than the plain
// estimate: 11.2 ms / 12.2 ms against 10.8 ms for C4 on one GPU.
#ifndef HB_TC_CENTRED
#define HB_TC_CENTRED 0
#endif
constexpr uint32_t TC_IDESC = (2u << 4) | (HB_TC_CENTRED ? (1u << 7) : 0u) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---- the filter kernel ---------------------------------------------------------------------------
struct TcFilterParams {
    uint64_t row0, row_end;      // base rows of this chunk
    const float4* bconst;        // [n_base] (u = db, v = P, w = ymid, Bc')
    const float4* qconst;        // [nq_tiles * 128] (a0, a1 = -2 xmid, a2 = -2 Sum x, a3 = -2 dq); a0 = +inf on padding rows
    const int* qshift;           // [nq_tiles * 128] 128 * Sum (cq - 128): dot'' = accumulator - qshift
    uint32_t nq, nq_tiles;
    u64* cand;                   // [nq][cap]: local base row of a survivor
    uint32_t cap;
    uint32_t* cnt;               // [nq]
    uint32_t* overflow;
};

__global__ void __launch_bounds__(TC_THREADS, 1) bf_tc_filter_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                      const __grid_constant__ CUtensorMap tmB,
                                                                      TcFilterParams p) {
    // 1024-byte alignment (swizzle atom) is requested from the launch; no pointer arithmetic here, so the
    // compiler keeps the shared address space (LDS instead of generic loads)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t tile_row0 = p.row0 + (uint64_t)blockIdx.x * TC_N;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        mbar_init(&S.b_full, 1);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&S.a_full[s], 1); mbar_init(&S.a_empty[s], 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(&S.acc_full[t], 1); mbar_init(&S.acc_empty[t], TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: both accumulator stages = all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {  // per-column constants of the stationary base tile
        for (int c = threadIdx.x - 64; c < TC_N; c += 32 * TC_EPI_WARPS) {
            const uint64_t row = tile_row0 + c;
            float4 k = make_float4(0.f, 0.f, 0.f, INFINITY);  // rows past the chunk never survive
            if (row < p.row_end) k = __ldg(p.bconst + row);
            S.cu[c] = k.x; S.cv[c] = k.y; S.cw[c] = k.z; S.cb[c] = k.w;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp >= 2 && warp < 2 + TC_N / 32) {  // aggregates of column group (warp - 2)
        const int gidx = warp - 2, c = gidx * 32 + lane;
        float bmin = S.cb[c], vmin = S.cv[c], vmax = vmin, wmin = S.cw[c], wmax = wmin, umax = S.cu[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bmin = fminf(bmin, __shfl_xor_sync(HB_FULL, bmin, o));
            vmin = fminf(vmin, __shfl_xor_sync(HB_FULL, vmin, o));
            vmax = fmaxf(vmax, __shfl_xor_sync(HB_FULL, vmax, o));
            wmin = fminf(wmin, __shfl_xor_sync(HB_FULL, wmin, o));
            wmax = fmaxf(wmax, __shfl_xor_sync(HB_FULL, wmax, o));
            umax = fmaxf(umax, __shfl_xor_sync(HB_FULL, umax, o));
        }
        if (lane == 0) {
            S.g_bmin[gidx] = bmin; S.g_vmin[gidx] = vmin; S.g_vmax[gidx] = vmax;
            S.g_wmin[gidx] = wmin; S.g_wmax[gidx] = wmax; S.g_umax[gidx] = umax;
        }
    }
    __syncthreads();
    const uint32_t tmem = S.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(&S.b_full, TC_B_BYTES);
            tma_load_2d(S.b, &tmB, &S.b_full, 0, (int)tile_row0);
            uint32_t it = 0;
            for (uint32_t t = blockIdx.y; t < p.nq_tiles; t += gridDim.y, ++it) {
                const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                mbar_wait_backoff(&S.a_empty[s], ph ^ 1);
                mbar_expect_tx(&S.a_full[s], TC_A_BYTES);
                tma_load_2d(S.a[s], &tmA, &S.a_full[s], 0, (int)(t * TC_M));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            mbar_wait(&S.b_full, 0);
            const uint64_t bdesc = tc_smem_desc(S.b);
            uint32_t it = 0;
            for (uint32_t t = blockIdx.y; t < p.nq_tiles; t += gridDim.y, ++it) {
                const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                const uint32_t acc = it & 1, aph = (it >> 1) & 1;
                mbar_wait_backoff(&S.acc_empty[acc], aph ^ 1);
                mbar_wait_backoff(&S.a_full[s], ph);
                tc_fence_after();
                const uint64_t adesc = tc_smem_desc(S.a[s]);
#pragma unroll
                for (int k = 0; k < TC_K / 32; ++k)  // 32 bytes of K per instruction: +2 in the 16-byte address field
                    tc_mma_i8(tmem + acc * TC_N, adesc + 2 * k, bdesc + 2 * k, TC_IDESC, k > 0 ? 1u : 0u);
                tc_commit(&S.a_empty[s]);     // the smem stage is free once these MMAs have read it
                tc_commit(&S.acc_full[acc]);  // ... and the accumulator is complete
            }
        }
    } else {
        // ===== epilogue: warp w reads TMEM lanes 32*(w%4).., columns TC_COLS_PER_WARP*cgroup.. =====
        const int quarter = warp & 3, cgroup = (warp - 2) >> 2;
        uint32_t it = 0;
        for (uint32_t t = blockIdx.y; t < p.nq_tiles; t += gridDim.y, ++it) {
            const uint32_t acc = it & 1, aph = (it >> 1) & 1;
            const uint32_t q = t * TC_M + quarter * 32 + lane;
            const float4 qc = __ldg(p.qconst + q);
            const int qsh = HB_TC_CENTRED ? __ldg(p.qshift + q) : 0;
            const u64 a0 = pk(qc.x, qc.x), a1 = pk(qc.y, qc.y), a2 = pk(qc.z, qc.z), a3 = pk(qc.w, qc.w);
            mbar_wait(&S.acc_full[acc], aph);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < TC_COLS_PER_WARP / 32; ++ch) {
                const int col0 = cgroup * TC_COLS_PER_WARP + ch * 32;
                uint32_t v[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(tmem + ((uint32_t)(quarter * 32) << 16) + acc * TC_N + col0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#if HB_TC_CENTRED
                {
                    // Quick reject of the whole group: val_c = base_c + (a3*u_c)*dot_c with a3*u_c <= 0, so
                    // val_c >= base_lb - smax * max(dot_c, 0) >= base_lb - smax * max(maxdot, 0) for every column c of the
                    // group, where base_lb bounds base_c = a0 + Bc' + a1*v + a2*w from below over the group's ranges and
                    // smax = -a3 * umax.  A group whose bound stays clearly positive holds no survivor.
                    int mx = (int)v[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) mx = max(mx, (int)v[j]);
                    mx -= qsh;
                    const int gi = col0 >> 5;
                    const float t1 = fminf(qc.y * S.g_vmin[gi], qc.y * S.g_vmax[gi]);
                    const float t2 = fminf(qc.z * S.g_wmin[gi], qc.z * S.g_wmax[gi]);
                    const float base_lb = qc.x + S.g_bmin[gi] + t1 + t2;
                    const float dterm = (-qc.w * S.g_umax[gi]) * (float)max(mx, 0);
                    // slack: far above the rounding of these few operations, far below the margins of real rejections
                    const float slack = 1e-4f * (fabsf(qc.x) + fabsf(S.g_bmin[gi]) + fabsf(t1) + fabsf(t2) + dterm);
                    const bool pass = base_lb - dterm > slack;  // false for NaN / -inf (non-finite parameters): those go on
                    // The test is per lane (= per query row) but a branch costs the whole warp: with ~9 % of the lanes
                    // failing, all 32 pass only ~6 % of the time.  So the few failing lanes are served one after the other
                    // by the whole warp, one lane per COLUMN: the failing lane stages its 32 accumulators in shared memory
                    // and broadcasts its query constants; the estimate is the same sequence of operations as below.
                    unsigned fm = __ballot_sync(HB_FULL, !pass);
                    if (fm == 0u) continue;
                    if (__popc(fm) <= TC_TRANSPOSE_MAX) {
                        int* stage = S.stage[warp - 2];
                        const int c = col0 + lane;
                        const float cu = S.cu[c], cv = S.cv[c], cw = S.cw[c], cb = S.cb[c];
#pragma unroll 1
                        while (fm) {
                            const int f = __ffs(fm) - 1;
                            fm &= fm - 1;
                            if (lane == f) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4)
                                    *reinterpret_cast<int4*>(stage + j) = make_int4((int)v[j], (int)v[j + 1], (int)v[j + 2], (int)v[j + 3]);
                            }
                            __syncwarp();
                            const float fa0 = __shfl_sync(HB_FULL, qc.x, f), fa1 = __shfl_sync(HB_FULL, qc.y, f),
                                        fa2 = __shfl_sync(HB_FULL, qc.z, f), fa3 = __shfl_sync(HB_FULL, qc.w, f);
                            const int fsh = __shfl_sync(HB_FULL, qsh, f);
                            const float dotf = __int2float_rn(stage[lane] - fsh);
                            const float base = __fadd_rn(__fmaf_rn(fa2, cw, __fmaf_rn(fa1, cv, cb)), fa0);
                            const float ev = __fmaf_rn(__fmul_rn(fa3, cu), dotf, base);
                            const unsigned sm = __ballot_sync(HB_FULL, ev <= 0.0f);
                            if (sm) {
                                const uint32_t fq = t * TC_M + quarter * 32 + f;
                                uint32_t pos0 = 0;
                                if (lane == 0) pos0 = atomicAdd(p.cnt + fq, (uint32_t)__popc(sm));
                                pos0 = __shfl_sync(HB_FULL, pos0, 0);
                                if (ev <= 0.0f) {
                                    const uint32_t pos = pos0 + __popc(sm & ((1u << lane) - 1u));
                                    if (pos < p.cap) p.cand[(size_t)fq * p.cap + pos] = (u64)(tile_row0 + c);
                                    else atomicOr(p.overflow, 1u);
                                }
                            }
                            __syncwarp();
                        }
                        continue;
                    }
                }
#endif
                float e[32];
                float lo = INFINITY;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int c = col0 + j;
                    const float4 u4 = *reinterpret_cast<const float4*>(&S.cu[c]);
                    const float4 v4 = *reinterpret_cast<const float4*>(&S.cv[c]);
                    const float4 w4 = *reinterpret_cast<const float4*>(&S.cw[c]);
                    const float4 b4 = *reinterpret_cast<const float4*>(&S.cb[c]);
                    // val = (a0 + Bc' + a1*v + a2*w) + (a3*u) * dot ; survivor iff val <= 0
                    const u64 base01 = add2(fma2(a2, pk(w4.x, w4.y), fma2(a1, pk(v4.x, v4.y), pk(b4.x, b4.y))), a0);
                    const u64 base23 = add2(fma2(a2, pk(w4.z, w4.w), fma2(a1, pk(v4.z, v4.w), pk(b4.z, b4.w))), a0);
                    const u64 dot01 = pk(__int2float_rn((int)v[j] - qsh), __int2float_rn((int)v[j + 1] - qsh));
                    const u64 dot23 = pk(__int2float_rn((int)v[j + 2] - qsh), __int2float_rn((int)v[j + 3] - qsh));
                    up(fma2(mul2(a3, pk(u4.x, u4.y)), dot01, base01), e[j], e[j + 1]);
                    up(fma2(mul2(a3, pk(u4.z, u4.w)), dot23, base23), e[j + 2], e[j + 3]);
                    lo = fminf(lo, fminf(fminf(e[j], e[j + 1]), fminf(e[j + 2], e[j + 3])));
                }
                if (lo <= 0.0f) {  // rare: about k survivors per query per doubling of the base
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (e[j] <= 0.0f) {
                            const uint32_t pos = atomicAdd(p.cnt + q, 1u);
                            if (pos < p.cap) p.cand[(size_t)q * p.cap + pos] = (u64)(tile_row0 + col0 + j);
                            else atomicOr(p.overflow, 1u);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.acc_empty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ---- operand preparation -------------------------------------------------------------------------
struct ByteMask { uint32_t w[32]; };  // 0xFF for every record byte that holds a code

// one warp per record: Sum c and Sum c^2 over the code bytes (exact integers), min and delta
__device__ __forceinline__ void record_stats(const uint8_t* rec, const RecLayout& L, const ByteMask& m, int lane,
                                             uint32_t& word_masked, float& s1, float& s2, float& mn, float& dl) {
    const uint32_t wd = __ldg(reinterpret_cast<const uint32_t*>(rec) + lane) & m.w[lane];
    uint32_t a = 0, b = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t c = (wd >> (8 * i)) & 0xFFu;
        a += c;
        b += c * c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(HB_FULL, a, o);
        b += __shfl_xor_sync(HB_FULL, b, o);
    }
    word_masked = wd;
    s1 = (float)a;  // <= 128 * 255, exact
    s2 = (float)b;  // <= 128 * 255^2 < 2^24, exact
    mn = __ldg(reinterpret_cast<const float*>(rec + hb_min_offset(L)));
    dl = __ldg(reinterpret_cast<const float*>(rec + hb_delta_offset(L)));
}

// (u, v, w, Bc') per base record.  Non-finite parameters make the record survive every filter.
__global__ void __launch_bounds__(256) bf_tc_base_consts_kernel(const uint8_t* __restrict__ rec, uint64_t n, RecLayout L,
                                                                ByteMask m, float4* out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t r = warp; r < n; r += nwarps) {
        uint32_t wd;
        float s1, s2, mn, dl;
        record_stats(rec + r * L.stride, L, m, lane, wd, s1, s2, mn, dl);
        const float sb = dl * s1;
        const float bc = dl * dl * s2 + 2.0f * mn * sb + (float)L.dim * mn * mn;  // Sum y^2
#if HB_TC_CENTRED
        float4 k = make_float4(dl, dl * (s1 - 128.0f * (float)L.dim), mn + 128.0f * dl, bc * (1.0f - TC_EPS));  // (db, P, ymid, Bc')
#else
        float4 k = make_float4(dl, sb + (float)L.dim * mn, mn, bc * (1.0f - TC_EPS));
#endif
        if (!(isfinite(k.x) && isfinite(k.y) && isfinite(k.z) && isfinite(k.w))) k = make_float4(0.f, 0.f, 0.f, -INFINITY);
        if (lane == 0) out[r] = k;
    }
}

// masked, centred copy of the query records (the A operand) + (dq, xmid, Sum x, Sum x^2) and the integer shift per query
__global__ void __launch_bounds__(256) bf_tc_query_prep_kernel(const uint8_t* __restrict__ qrec, uint32_t nq, RecLayout L,
                                                               ByteMask m, uint8_t* amask, float4* qstat, int* qshift) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t q = warp; q < nq; q += nwarps) {
        uint32_t wd;
        float s1, s2, mn, dl;
        record_stats(qrec + (size_t)q * L.stride, L, m, lane, wd, s1, s2, mn, dl);
        const float sq = dl * s1;
        const float qc = dl * dl * s2 + 2.0f * mn * sq + (float)L.dim * mn * mn;  // Sum x^2
#if HB_TC_CENTRED
        // A = (code - 128) as s8 on the code bytes, 0 elsewhere; x_i = (c_i - 128)*dl + xmid
        reinterpret_cast<uint32_t*>(amask + (size_t)q * TC_K)[lane] = (wd ^ 0x80808080u) & m.w[lane];
        // (dq, xmid, Sum x, Sum x^2); the integer 128 * Sum (cq - 128) goes to qshift
        if (lane == 0) {
            qstat[q] = make_float4(dl, mn + 128.0f * dl, sq + (float)L.dim * mn, qc);
            qshift[q] = 128 * ((int)s1 - 128 * (int)L.dim);
        }
#else
        reinterpret_cast<uint32_t*>(amask + (size_t)q * TC_K)[lane] = wd;
        if (lane == 0) qstat[q] = make_float4(dl, mn, sq, qc);
#endif
    }
}

// per chunk: (a0, a1, a2, a3) from the current k-th exact key of every query
__global__ void bf_tc_thresholds_kernel(const float4* __restrict__ qstat, const u64* __restrict__ tau, uint32_t nq,
                                        uint32_t nq_pad, float4* qconst) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq_pad) return;
    if (q >= nq) { qconst[q] = make_float4(INFINITY, 0.f, 0.f, 0.f); return; }
    const float4 s = qstat[q];
    const u64 key = tau[q];
    float T = INFINITY;  // fewer than k keys known yet: everything survives
    if (key != ~0ull) {
        const float t = __uint_as_float((uint32_t)(key >> 32));
        T = t * t * (1.0f + TC_DELTA);
    }
    float a0 = s.w * (1.0f - TC_EPS) - T;
    if (!isfinite(s.x) || !isfinite(s.y) || !isfinite(s.z) || !isfinite(s.w)) a0 = -INFINITY;
    qconst[q] = make_float4(a0, -2.0f * s.y, -2.0f * s.z, -2.0f * s.x);
}

// exact re-rank of the survivors: cand[q][i] (local base row) -> (dist, id) key, in place
template <class Q>
__global__ void __launch_bounds__(128) bf_rerank_kernel(const uint8_t* __restrict__ base_rec, RecLayout L, uint32_t id_offset,
                                                        const uint8_t* __restrict__ qrec, uint32_t nq, u64* cand, uint32_t cap,
                                                        const uint32_t* __restrict__ cnt) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    const uint32_t qd_cap = (L.dim + 7) / 8 * 8 + 8;
    float* qd = reinterpret_cast<float*>(smem) + (size_t)wib * qd_cap;
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + wib;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t q = warp; q < nq; q += nwarps) {
        const uint32_t n = min(__ldg(cnt + q), cap);
        if (n == 0) continue;
        __syncwarp();
        warp_dequant_record(L, qrec + (size_t)q * L.stride, lane, qd);
        __syncwarp();
        Q qq;
        qq.init(L, qd, gl);
        u64* row = cand + (size_t)q * cap;
        for (uint32_t i0 = 0; i0 < n; i0 += 8) {
            const uint32_t i = i0 + grp;
            const bool act = i < n;
            const uint32_t b = (uint32_t)row[act ? i : i0];
            const float d = qq.dist(base_rec + (size_t)b * L.stride, gl, gbase);
            if (act && gl == 0) row[i] = make_key(d, b + id_offset);
        }
    }
}

// the first rows of the base: every (query, row) pair is a "survivor"
__global__ void __launch_bounds__(256) bf_fill_first_kernel(u64* cand, uint32_t cap, uint32_t* cnt, uint32_t nq, uint32_t first) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)nq * first) return;
    const uint32_t q = (uint32_t)(i / first), r = (uint32_t)(i % first);
    cand[(size_t)q * cap + r] = r;
    if (r == 0) cnt[q] = first;
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// rows of 128 operand bytes `pitch` bytes apart (128: dim 96 / 100 records and the query copy; 144: dim 128 records, whose
// 128 code bytes come first and whose min / delta live in the 16-byte tail), box = box_rows x 128 bytes, 128-byte swizzle;
// rows past the end read as zero
static bool make_map(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t box_rows, uint32_t pitch = TC_K) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)TC_K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)TC_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool bf_tc_supported(const RecLayout& L) {
    // the K extent is the first 128 bytes of a record: every code byte must lie there (dim 96 / 100: 128-byte records;
    // dim 128: 144-byte records = 128 code bytes + a 16-byte tail)
    const bool shape = L.stride == (uint32_t)TC_K || (L.stride == 144u && L.dim == 128u);
    return L.kind == HB_REC_QUANT && shape && get_encode() != nullptr && !getenv("HNSWB200_BF_TC_128_ONLY");
}

static ByteMask code_mask(const RecLayout& L) {
    ByteMask m;
    for (int i = 0; i < 32; ++i) m.w[i] = 0;
    for (uint32_t i = 0; i < L.dim; ++i) {
        const uint32_t off = hb_code_offset(L, i);
        m.w[off / 4] |= 0xFFu << (8 * (off % 4));
    }
    return m;
}

cudaError_t bf_tc_prepare(const uint8_t* base_rec, uint64_t n, const RecLayout& L, const uint8_t* qrec, uint32_t nq,
                          float4* bconst, uint8_t* amask, float4* qstat, int* qshift, cudaStream_t st) {
    const ByteMask m = code_mask(L);
    if (n) bf_tc_base_consts_kernel<<<(unsigned)std::min<uint64_t>((n + 7) / 8, 148 * 16), 256, 0, st>>>(base_rec, n, L, m, bconst);
    if (nq) bf_tc_query_prep_kernel<<<(unsigned)std::min<uint32_t>((nq + 7) / 8, 148 * 16), 256, 0, st>>>(qrec, nq, L, m, amask, qstat, qshift);
    return cudaGetLastError();
}



#define HB_DISPATCH_DIM_T(L, ...)                                                  \
    do {                                                                           \
        if ((L).dim == 100) { using Q = RegQuery<12, 4>; __VA_ARGS__; }            \
        else if ((L).dim == 96) { using Q = RegQuery<12, 0>; __VA_ARGS__; }        \
        else if ((L).dim == 128) { using Q = RegQuery<16, 0>; __VA_ARGS__; }       \
        else { using Q = SmemQuery; __VA_ARGS__; }                                 \
    } while (0)

// Rows [0, first) of the base ranked exactly against every query (no threshold exists yet): `cand` receives the keys.
cudaError_t bf_tc_first(const uint8_t* base_rec, const RecLayout& L, uint32_t first, uint32_t id_offset, const uint8_t* qrec,
                        uint32_t nq, u64* cand, uint32_t cap, uint32_t* cnt, cudaStream_t st) {
    if (first == 0 || nq == 0) return cudaSuccess;
    if (first > cap) return cudaErrorInvalidValue;
    const uint64_t tot = (uint64_t)nq * first;
    bf_fill_first_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(cand, cap, cnt, nq, first);
    const uint32_t qd_cap = (L.dim + 7) / 8 * 8 + 8;
    const size_t rsm = (size_t)4 * qd_cap * 4;
    HB_DISPATCH_DIM_T(L, {
        cudaFuncSetAttribute(bf_rerank_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm);
        bf_rerank_kernel<Q><<<std::min<uint32_t>((nq + 3) / 4, 148 * 8), 128, rsm, st>>>(base_rec, L, id_offset, qrec, nq, cand, cap, cnt);
    });
    return cudaGetLastError();
}

// One chunk [row0, row_end) of the base: thresholds from tau, tensor-core filter, exact re-rank of the
// survivors into `cand` (as keys).  cnt must be zero on entry; *overflow is raised if a list overflowed.
cudaError_t bf_tc_chunk(const uint8_t* base_rec, uint64_t n_base, const RecLayout& L, uint64_t row0, uint64_t row_end,
                        uint32_t id_offset, const uint8_t* qrec, const uint8_t* amask, const float4* qstat,
                        const float4* bconst, float4* qconst, const int* qshift, uint32_t nq, const u64* tau, u64* cand,
                        uint32_t cap, uint32_t* cnt, uint32_t* overflow, int num_sms, cudaStream_t st) {
    if (row_end <= row0 || nq == 0) return cudaSuccess;
    const uint32_t nq_tiles = (nq + TC_M - 1) / TC_M, nq_pad = nq_tiles * TC_M;
    CUtensorMap tmA, tmB;
    if (!make_map(&tmA, amask, nq_pad, TC_M) || !make_map(&tmB, base_rec, n_base, TC_N, L.stride)) return cudaErrorNotSupported;
    bf_tc_thresholds_kernel<<<(nq_pad + 255) / 256, 256, 0, st>>>(qstat, tau, nq, nq_pad, qconst);
    TcFilterParams p;
    p.row0 = row0; p.row_end = row_end; p.bconst = bconst; p.qconst = qconst; p.qshift = qshift; p.nq = nq; p.nq_tiles = nq_tiles;
    p.cand = cand; p.cap = cap; p.cnt = cnt; p.overflow = overflow;
    const uint32_t btiles = (uint32_t)((row_end - row0 + TC_N - 1) / TC_N);
    // split the query tiles of one base tile over several CTAs when there are few base tiles
    uint32_t qsplit = 1;
    while (btiles * qsplit < (uint32_t)num_sms * 2 && qsplit * 2 <= nq_tiles) qsplit *= 2;
    const size_t smem = sizeof(TcSmem);
    static bool attr_d[64] = {};  // per device
    int dev = 0;
    cudaGetDevice(&dev);
    bool& attr = attr_d[dev & 63];
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(bf_tc_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    bf_tc_filter_kernel<<<dim3(btiles, qsplit), TC_THREADS, smem, st>>>(tmA, tmB, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const uint32_t qd_cap = (L.dim + 7) / 8 * 8 + 8;
    const size_t rsm = (size_t)4 * qd_cap * 4;
    HB_DISPATCH_DIM_T(L, {
        cudaFuncSetAttribute(bf_rerank_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm);
        bf_rerank_kernel<Q><<<std::min<uint32_t>((nq + 3) / 4, 148 * 8), 128, rsm, st>>>(base_rec, L, id_offset, qrec, nq, cand, cap, cnt);
    });
    return cudaGetLastError();
}

}  // namespace hb
