// Bit-exact quantised L2 distance primitives (device side).
//
// Arithmetic contract (reference vectors/src/quant.rs:14-37, 41-66):
//   deq(c)  = fadd(fmul((f32)c, delta), min)            two roundings, never an FMA
//   acc[j] += fmul(t, t), t = fsub(x, y)                 per residue j = i mod 8, increasing i
//   remainder elements -> acc[0], after the chunks
//   s = ((((((a0+a1)+a2)+a3)+a4)+a5)+a6)+a7 ; d = sqrt_rn(s)
// A group of 4 lanes evaluates one record; lane l owns (acc[2l], acc[2l+1]) and
// works on them as one packed f32x2 value (FADD2 on sm_100a).  Multiplications are
// issued as scalar FMULs: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2
// even with explicit rounding modifiers and --fmad=false, which would break the
// contract; scalar mul.rn.f32 is never contracted.
//
// Fast path (RegQuery): the two multiplications are issued as packed FFMA2s that are
// exact restatements of a single rounded product:
//   rn(c * delta)  = fma(2^23 + c, delta, -(2^23 * delta))   the magic-number form of the u8
//                    code enters directly; 2^23*delta is a power-of-two scaling (exact), the
//                    fma forms (2^23+c)*delta - 2^23*delta = c*delta exactly and rounds once
//   rn(t * t)      = fma(t, t, -0.0)                          adding -0 never changes a product
// The -0.0 addend is read from constant memory so that ptxas cannot see it and turn the fma
// back into a multiply it might then contract with the following add.  Records whose delta
// is not below 2^100 (2^23*delta could overflow) take the scalar-multiply path.
// tools/check_sass.py verifies the instruction mix; the GPU parity tests compare bit patterns.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "layout.h"

typedef unsigned long long u64;
#define HB_FULL 0xffffffffu

namespace hb {

__device__ __forceinline__ u64 pk(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void up(u64 v, float& a, float& b) {
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// pk(-0.0f, -0.0f); opaque to ptxas (see the header comment)
static __constant__ u64 hb_negzero2 = 0x8000000080000000ull;
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// exact u8 -> f32 for byte `b` (0..3) of word r: 0x4B0000bb = 2^23 + bb, minus 2^23 later
__device__ __forceinline__ float magic_byte(uint32_t r, int b) {
    return __uint_as_float(__byte_perm(r, 0x4B000000u, 0x7540u | (uint32_t)b));
}
__device__ __forceinline__ uint32_t word32(const uint4& w, int i) {
    return i == 0 ? w.x : i == 1 ? w.y : i == 2 ? w.z : w.w;
}

// one chunk: record bytes (b0, b0+1) of register rr against the query pair q
__device__ __forceinline__ u64 chunk_acc(u64 acc, uint32_t rr, int b0, float dl, u64 mn2, u64 q) {
    const u64 magic = pk(8388608.0f, 8388608.0f);
    u64 cf = sub2(pk(magic_byte(rr, b0), magic_byte(rr, b0 + 1)), magic);  // exact codes
    float c0, c1;
    up(cf, c0, c1);
    u64 x = add2(pk(__fmul_rn(c0, dl), __fmul_rn(c1, dl)), mn2);
    u64 t = sub2(x, q);
    float t0, t1;
    up(t, t0, t1);
    return add2(acc, pk(__fmul_rn(t0, t0), __fmul_rn(t1, t1)));
}
// fast form of chunk_acc: m2 = pk(2^23 + c0, 2^23 + c1); dl2 = pk(delta, delta);
// nmd2 = pk(-(2^23*delta), ..); nz = pk(-0, -0)
__device__ __forceinline__ u64 chunk_sq(u64 m2, u64 dl2, u64 nmd2, u64 mn2, u64 q, u64 nz) {
    u64 x = add2(fma2(m2, dl2, nmd2), mn2);
    u64 t = sub2(x, q);
    return fma2(t, t, nz);
}
__device__ __forceinline__ float rem_acc(float a0, float fm, float dl, float mn, float q) {
    float c = __fadd_rn(fm, -8388608.0f);
    float x = __fadd_rn(__fmul_rn(c, dl), mn);
    float t = __fsub_rn(x, q);
    return __fadd_rn(a0, __fmul_rn(t, t));
}

// sequential 8-way sum across the 4 lanes of a group; result broadcast to the group
__device__ __forceinline__ float group_sum_sqrt(u64 acc, int gl, int gbase) {
    float a0, a1;
    up(acc, a0, a1);
    float p = __fadd_rn(a0, a1);  // lane 0: (0 + a0) + a1 with 0 + a0 == a0
    float t = __shfl_sync(HB_FULL, p, gbase + 0);
    if (gl == 1) p = __fadd_rn(__fadd_rn(t, a0), a1);
    t = __shfl_sync(HB_FULL, p, gbase + 1);
    if (gl == 2) p = __fadd_rn(__fadd_rn(t, a0), a1);
    t = __shfl_sync(HB_FULL, p, gbase + 2);
    if (gl == 3) p = __fadd_rn(__fadd_rn(t, a0), a1);
    p = __shfl_sync(HB_FULL, p, gbase + 3);
    return __fsqrt_rn(p);
}

// ---------------------------------------------------------------------------
// Query held in registers, dimension known at compile time (dim = 8*NCH + REM).
// All 32 lanes of the warp must call dist() together (it shuffles).
// ---------------------------------------------------------------------------
template <int NCH, int REM>
struct RegQuery {
    static constexpr bool kKeepsSmem = false;  // the dequantised values are copied to registers by init()
    static constexpr bool kPrefetchBeforeVisited = true;  // records of one or two lines: see search_reg.cuh
    static constexpr int W = (int)hb_layout_W(NCH);
    static constexpr int TAIL = (int)hb_layout_tail(NCH, REM);
    static constexpr int RP = (REM + 1) / 2;  // remainder pairs
    u64 q[NCH ? NCH : 1];
    u64 qr[RP ? RP : 1];
    u64 nz;

    // qd: dequantised query values in shared memory, natural element order
    __device__ __forceinline__ void init(const RecLayout&, const float* qd, int gl) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            float2 v = *reinterpret_cast<const float2*>(qd + 8 * k + 2 * gl);
            q[k] = pk(v.x, v.y);
        }
#pragma unroll
        for (int r = 0; r < RP; ++r)
            qr[r] = pk(qd[8 * NCH + 2 * r], (2 * r + 1 < REM) ? qd[8 * NCH + 2 * r + 1] : 0.0f);
        nz = hb_negzero2;
    }

    // magic-number form (2^23 + code) of slice position pz of this lane
    __device__ __forceinline__ static float slice_magic(const uint4 (&w)[W], int pz) {
        return magic_byte(word32(w[pz / 16], (pz % 16) / 4), pz % 4);
    }

    // the bytes of one record that this lane needs, in registers (lets a caller issue the loads of several
    // records before it evaluates the first)
    struct Rec {
        uint4 w[W];
        uint4 tw;
    };
    __device__ __forceinline__ static Rec load(const uint8_t* __restrict__ rec, int gl) {
        const uint4* p = reinterpret_cast<const uint4*>(rec);
        Rec r;
#pragma unroll
        for (int j = 0; j < W; ++j) r.w[j] = __ldg(p + 4 * j + gl);
        r.tw = make_uint4(0, 0, 0, 0);
        if (TAIL) r.tw = __ldg(p + 4 * W);
        return r;
    }
    __device__ __forceinline__ float dist(const uint8_t* __restrict__ rec, int gl, int gbase) const {
        return dist(load(rec, gl), gl, gbase);
    }
    __device__ __forceinline__ float dist(const Rec& R, int gl, int gbase) const {
        const uint4 (&w)[W] = R.w;
        const uint4 tw = R.tw;
        float mn, dl;
        if (TAIL) {
            mn = __uint_as_float(tw.x);
            dl = __uint_as_float(tw.y);
        } else {
            uint32_t last = w[W - 1].w;
            mn = __uint_as_float(__shfl_sync(HB_FULL, last, gbase + 1));
            dl = __uint_as_float(__shfl_sync(HB_FULL, last, gbase + 2));
        }
        const u64 mn2 = pk(mn, mn);
        u64 acc = pk(0.0f, 0.0f);
        if (__any_sync(HB_FULL, !(dl < 1.2676506e30f))) {
            // delta >= 2^100, inf or NaN somewhere in this round: separately rounded multiplies
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int j = k / 8, c = k % 8;
                acc = chunk_acc(acc, word32(w[j], c / 2), (c & 1) * 2, dl, mn2, q[k]);
            }
            if (REM > 0) {
                float a0, a1;
                up(acc, a0, a1);
                float n0 = a0;
#pragma unroll
                for (int r = 0; r < REM; ++r) {
                    float fm = TAIL ? magic_byte(word32(tw, (8 + r) / 4), (8 + r) % 4) : slice_magic(w, 2 * NCH + r);
                    float qa, qb;
                    up(qr[r / 2], qa, qb);
                    n0 = rem_acc(n0, fm, dl, mn, (r & 1) ? qb : qa);
                }
                acc = pk(gl == 0 ? n0 : a0, a1);
            }
            return group_sum_sqrt(acc, gl, gbase);
        }
        const u64 dl2 = pk(dl, dl);
        const float nmd = __fmul_rn(-8388608.0f, dl);
        const u64 nmd2 = pk(nmd, nmd);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            u64 m2 = pk(slice_magic(w, 2 * k), slice_magic(w, 2 * k + 1));
            acc = add2(acc, chunk_sq(m2, dl2, nmd2, mn2, q[k], nz));
        }
        if (REM > 0) {
            // remainder elements: squares in pairs, accumulated one by one into acc[0] (lane 0 of the group)
            float a0, a1;
            up(acc, a0, a1);
            float n0 = a0;
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                float f0, f1;
                if (TAIL) {
                    f0 = magic_byte(word32(tw, (8 + 2 * r) / 4), (8 + 2 * r) % 4);
                    f1 = (2 * r + 1 < REM) ? magic_byte(word32(tw, (9 + 2 * r) / 4), (9 + 2 * r) % 4) : 8388608.0f;
                } else {
                    f0 = slice_magic(w, 2 * NCH + 2 * r);
                    f1 = (2 * r + 1 < REM) ? slice_magic(w, 2 * NCH + 2 * r + 1) : 8388608.0f;
                }
                float s0, s1;
                up(chunk_sq(pk(f0, f1), dl2, nmd2, mn2, qr[r], nz), s0, s1);
                n0 = __fadd_rn(n0, s0);
                if (2 * r + 1 < REM) n0 = __fadd_rn(n0, s1);
            }
            acc = pk(gl == 0 ? n0 : a0, a1);
        }
        return group_sum_sqrt(acc, gl, gbase);
    }
};

// ---------------------------------------------------------------------------
// Query held in shared memory, runtime dimension (any dim).
// ---------------------------------------------------------------------------
struct SmemQuery {
    static constexpr bool kKeepsSmem = true;  // dist() reads the dequantised values from shared memory
    static constexpr bool kPrefetchBeforeVisited = true;
    const float* qd;
    RecLayout L;
    __device__ __forceinline__ void init(const RecLayout& l, const float* q, int) {
        qd = q;
        L = l;
    }
    struct Rec { const uint8_t* p; };  // runtime dimension: nothing is preloaded
    __device__ __forceinline__ static Rec load(const uint8_t* __restrict__ rec, int) { return Rec{rec}; }
    __device__ __forceinline__ float dist(const Rec& r, int gl, int gbase) const { return dist(r.p, gl, gbase); }
    __device__ __forceinline__ float dist(const uint8_t* __restrict__ rec, int gl, int gbase) const {
        const uint4* p = reinterpret_cast<const uint4*>(rec);
        const float mn = __ldg(reinterpret_cast<const float*>(rec + hb_min_offset(L)));
        const float dl = __ldg(reinterpret_cast<const float*>(rec + hb_delta_offset(L)));
        const u64 mn2 = pk(mn, mn);
        u64 acc = pk(0.0f, 0.0f);
        for (uint32_t j = 0; j < L.W; ++j) {
            uint4 w = __ldg(p + 4 * j + gl);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t k = 8 * j + c;
                if (k < L.nch) {
                    float2 v = *reinterpret_cast<const float2*>(qd + 8 * k + 2 * gl);
                    acc = chunk_acc(acc, word32(w, c / 2), (c & 1) * 2, dl, mn2, pk(v.x, v.y));
                }
            }
        }
        if (L.rem) {
            float a0, a1;
            up(acc, a0, a1);
            float n0 = a0;
            for (uint32_t r = 0; r < L.rem; ++r) {
                uint32_t byte = __ldg(rec + hb_code_offset(L, 8 * L.nch + r));
                float fm = __uint_as_float(0x4B000000u | byte);
                n0 = rem_acc(n0, fm, dl, mn, qd[8 * L.nch + r]);
            }
            acc = pk(gl == 0 ? n0 : a0, a1);
        }
        return group_sum_sqrt(acc, gl, gbase);
    }
};

// ---------------------------------------------------------------------------
// f32 records (the reference's FullVec, vectors/src/full.rs:23-29):
//   d = sqrt( sum_i (x_i - y_i)^2 ), ONE strictly sequential f32 sum in element order, no FMA.
// (x - y)^2 == (y - x)^2 bit for bit, so the orientation does not matter.  The 4 lanes of a group
// take one 16-float chunk at a time (lane l: floats 4l..4l+3, one 16-byte load) and pass the
// running sum from lane to lane: every lane adds its four squares to the sum it was handed, and the
// group keeps the result of the lane whose turn it is.  Zero padding adds +0 to a non-negative sum.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float full_chain16(float s, const uint4& x, const float4& y, u64 nz, int gbase) {
    const u64 t01 = sub2(pk(__uint_as_float(x.x), __uint_as_float(x.y)), pk(y.x, y.y));
    const u64 t23 = sub2(pk(__uint_as_float(x.z), __uint_as_float(x.w)), pk(y.z, y.w));
    float q0, q1, q2, q3;
    up(fma2(t01, t01, nz), q0, q1);  // rn(t*t): adding -0 never changes a product (header comment)
    up(fma2(t23, t23, nz), q2, q3);
#pragma unroll
    for (int st = 0; st < 4; ++st) {
        const float t = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s, q0), q1), q2), q3);
        s = __shfl_sync(HB_FULL, t, gbase + st);
    }
    return s;
}

struct FullQuery {
    static constexpr bool kKeepsSmem = true;  // dist() reads the query from shared memory
    static constexpr bool kPrefetchBeforeVisited = false;  // 4*dim bytes per record: only records that will be evaluated
    const float* qd;  // natural order, zero-padded to 16*W floats, 16-byte aligned
    uint32_t W;
    u64 nz;
    __device__ __forceinline__ void init(const RecLayout& l, const float* q, int) {
        qd = q;
        W = l.W;
        nz = hb_negzero2;
    }
    struct Rec { const uint8_t* p; };  // runtime dimension: nothing is preloaded
    __device__ __forceinline__ static Rec load(const uint8_t* __restrict__ rec, int) { return Rec{rec}; }
    __device__ __forceinline__ float dist(const Rec& r, int gl, int gbase) const { return dist(r.p, gl, gbase); }
    __device__ __forceinline__ float dist(const uint8_t* __restrict__ rec, int gl, int gbase) const {
        const uint4* p = reinterpret_cast<const uint4*>(rec) + gl;
        const float4* q = reinterpret_cast<const float4*>(qd) + gl;
        float s = 0.0f;
        // two chunks in flight while one chunk's chain runs (a chain is ~160 cycles, an L2 hit ~300)
        uint4 w0 = __ldg(p);
        uint4 w1 = W > 1 ? __ldg(p + 4) : w0;
        for (uint32_t j = 0; j < W; ++j) {
            const uint4 cur = w0;
            w0 = w1;
            if (j + 2 < W) w1 = __ldg(p + 4 * (j + 2));
            s = full_chain16(s, cur, q[4 * j], nz, gbase);
        }
        return __fsqrt_rn(s);
    }
};

__device__ __forceinline__ float full_rec_rec_dist(const RecLayout& L, const uint8_t* __restrict__ ra,
                                                   const uint8_t* __restrict__ rb, int gl, int gbase) {
    const uint4* pa = reinterpret_cast<const uint4*>(ra) + gl;
    const uint4* pb = reinterpret_cast<const uint4*>(rb) + gl;
    const u64 nz = hb_negzero2;
    float s = 0.0f;
    for (uint32_t j = 0; j < L.W; ++j) {
        const uint4 x = __ldg(pa + 4 * j);
        const uint4 yb = __ldg(pb + 4 * j);
        const float4 y = make_float4(__uint_as_float(yb.x), __uint_as_float(yb.y), __uint_as_float(yb.z), __uint_as_float(yb.w));
        s = full_chain16(s, x, y, nz, gbase);
    }
    return __fsqrt_rn(s);
}

// record <-> record distance, both streamed from memory (Points::distance(a,b),
// points/src/points.rs:86-93: x = a, y = b).  Runtime dimension.
__device__ __forceinline__ float rec_rec_dist(const RecLayout& L, const uint8_t* __restrict__ ra,
                                              const uint8_t* __restrict__ rb, int gl, int gbase) {
    if (L.kind == HB_REC_F32) return full_rec_rec_dist(L, ra, rb, gl, gbase);
    const uint4* pa = reinterpret_cast<const uint4*>(ra);
    const uint4* pb = reinterpret_cast<const uint4*>(rb);
    const float mna = __ldg(reinterpret_cast<const float*>(ra + hb_min_offset(L)));
    const float dla = __ldg(reinterpret_cast<const float*>(ra + hb_delta_offset(L)));
    const float mnb = __ldg(reinterpret_cast<const float*>(rb + hb_min_offset(L)));
    const float dlb = __ldg(reinterpret_cast<const float*>(rb + hb_delta_offset(L)));
    const u64 mna2 = pk(mna, mna);
    const u64 magic = pk(8388608.0f, 8388608.0f);
    u64 acc = pk(0.0f, 0.0f);
    for (uint32_t j = 0; j < L.W; ++j) {
        uint4 wa = __ldg(pa + 4 * j + gl);
        uint4 wb = __ldg(pb + 4 * j + gl);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t k = 8 * j + c;
            if (k < L.nch) {
                uint32_t rr = word32(wb, c / 2);
                int b0 = (c & 1) * 2;
                u64 cf = sub2(pk(magic_byte(rr, b0), magic_byte(rr, b0 + 1)), magic);
                float c0, c1;
                up(cf, c0, c1);
                u64 y = add2(pk(__fmul_rn(c0, dlb), __fmul_rn(c1, dlb)), pk(mnb, mnb));
                // t = x_a - y_b: negate the roles in chunk_acc by feeding y as the "query"
                acc = chunk_acc(acc, word32(wa, c / 2), b0, dla, mna2, y);
            }
        }
    }
    if (L.rem) {
        float a0, a1;
        up(acc, a0, a1);
        float n0 = a0;
        for (uint32_t r = 0; r < L.rem; ++r) {
            uint32_t off = hb_code_offset(L, 8 * L.nch + r);
            float y = __fadd_rn(__fmul_rn((float)__ldg(rb + off), dlb), mnb);
            float fm = __uint_as_float(0x4B000000u | (uint32_t)__ldg(ra + off));
            n0 = rem_acc(n0, fm, dla, mna, y);
        }
        acc = pk(gl == 0 ? n0 : a0, a1);
    }
    return group_sum_sqrt(acc, gl, gbase);
}

// ---------------------------------------------------------------------------
// Warp-level quantiser (QuantVec::new, vectors/src/quant.rs:41-66).
// v: f32[dim] in global memory.  Writes the dequantised values to qd (shared or
// global, natural order), optionally the codes, and returns min/delta.
// Returns false if the vector holds a NaN (the reference panics).
// max_by keeps the LAST of equal maxima, min_by the FIRST of equal minima.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool warp_quantise(const float* __restrict__ v, uint32_t dim, int lane,
                                              float* qd, uint8_t* codes, float& mn_out,
                                              float& dl_out) {
    float mxv = -INFINITY, mnv = INFINITY;
    int mxi = -1, mni = 0x7fffffff;
    bool nan = false;
    for (uint32_t i = lane; i < dim; i += 32) {
        float x = v[i];
        nan |= (x != x);
        if (!(mxv > x)) { mxv = x; mxi = (int)i; }
        if (x < mnv) { mnv = x; mni = (int)i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(HB_FULL, mxv, o);
        int oi = __shfl_xor_sync(HB_FULL, mxi, o);
        if (ov > mxv || (ov == mxv && oi > mxi)) { mxv = ov; mxi = oi; }
        ov = __shfl_xor_sync(HB_FULL, mnv, o);
        oi = __shfl_xor_sync(HB_FULL, mni, o);
        if (ov < mnv || (ov == mnv && oi < mni)) { mnv = ov; mni = oi; }
    }
    nan = __any_sync(HB_FULL, nan);
    const float lb = mnv;
    const float dl = __fdiv_rn(__fsub_rn(mxv, lb), 255.0f);
    for (uint32_t i = lane; i < dim; i += 32) {
        float b = __fadd_rn(__fdiv_rn(__fsub_rn(v[i], lb), dl), 0.5f);
        float f = floorf(b);
        f = fminf(fmaxf(f, 0.0f), 255.0f);  // Rust `as u8`: saturating, NaN -> 0
        uint32_t c = (uint32_t)f;
        if (codes) codes[i] = (uint8_t)c;
        if (qd) qd[i] = __fadd_rn(__fmul_rn((float)c, dl), lb);
    }
    mn_out = lb;
    dl_out = dl;
    return !nan;
}

// Point::new(vector) for a query (points/src/point.rs:24-30): the values the distance arithmetic sees, written to
// qd in natural order (src may be qd itself).  QuantVec: quantise, then dequantise (the codes are not kept);
// FullVec: the vector itself, zero-padded to whole 16-float chunks.  Returns false for a vector the reference would
// panic on: a NaN (QuantVec), any non-finite value (FullVec: inf - inf would make a NaN distance).
__device__ __forceinline__ bool warp_prepare_query(const RecLayout& L, const float* src, float* qd, int lane) {
    if (L.kind == HB_REC_F32) {
        bool bad = false;
        for (uint32_t i = lane; i < 16 * L.W; i += 32) {
            const float x = i < L.dim ? src[i] : 0.0f;
            bad |= !(fabsf(x) <= 3.4028234664e38f);
            qd[i] = x;
        }
        return !__any_sync(HB_FULL, bad);
    }
    float mn, dl;
    return warp_quantise(src, L.dim, lane, qd, nullptr, mn, dl);
}

// dequantise a stored record into natural element order (qd[dim])
__device__ __forceinline__ void warp_dequant_record(const RecLayout& L, const uint8_t* __restrict__ rec,
                                                    int lane, float* qd) {
    if (L.kind == HB_REC_F32) {  // FullVec: the stored values themselves, padding included
        const float* r = reinterpret_cast<const float*>(rec);
        for (uint32_t i = lane; i < 16 * L.W; i += 32) qd[i] = __ldg(r + i);
        return;
    }
    const float mn = __ldg(reinterpret_cast<const float*>(rec + hb_min_offset(L)));
    const float dl = __ldg(reinterpret_cast<const float*>(rec + hb_delta_offset(L)));
    for (uint32_t i = lane; i < L.dim; i += 32) {
        uint32_t c = __ldg(rec + hb_code_offset(L, i));
        qd[i] = __fadd_rn(__fmul_rn((float)c, dl), mn);
    }
}

}  // namespace hb
