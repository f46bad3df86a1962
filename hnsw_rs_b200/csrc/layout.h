// Record layout of a quantised point in HBM ("lane-sliced" record).
//
// A QuantVec (vectors/src/quant.rs:6-11: delta, min, codes[dim]) is evaluated by a
// group of 4 lanes.  The reference accumulates element i into acc[i mod 8] for the
// first 8*floor(dim/8) elements and every remainder element into acc[0]
// (quant.rs:14-37).  Lane l of the group owns the accumulator pair
// (acc[2l], acc[2l+1]) and therefore exactly the bytes code[8k+2l], code[8k+2l+1]
// for every full chunk k, in increasing k -- the reference's summation order.
//
// Memory order is word-major so that the 4 lanes of a group read 64 contiguous
// bytes with one 16-byte load each:   offset(word j, lane l) = 16 * (4 j + l)
//   lane slice position p = 2k + {0,1}  ->  word j = p / 16, byte p % 16
// compact form (tail == 0):
//   remainder bytes: lane 0, slice positions 2*nch .. 2*nch+rem-1
//   min   : last 4 bytes of lane 1's last word
//   delta : last 4 bytes of lane 2's last word
// tail form (tail == 1), used when the slices have no spare room:
//   one extra 16-byte word at offset 64*W: [min f32][delta f32][rem bytes][pad]
// stride = 64*W + 16*tail bytes: 128 B for dim 96/100 (one line per candidate),
// 144 B for dim 128, 64 B for dim 50.
//
// f32 records (kind == HB_REC_F32; the reference's other VecType, FullVec, vectors/src/full.rs:3-6):
// the dim floats in natural order, zero-padded to a multiple of 16 floats.  FullVec::distance sums
// strictly sequentially (full.rs:23-29), so the 4 lanes of a group take the 16-float chunks of the
// record in order, lane l holding floats 4l..4l+3 of a chunk (one 16-byte load each, 64 contiguous
// bytes per group), and hand the running sum from lane to lane.  A padding element contributes
// (0 - 0)^2 = +0 to a non-negative sum, which leaves it unchanged bit for bit.
// stride = 64 * ceil(dim/16) bytes: 448 B for dim 100, 512 B for dim 128.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HB_HD __host__ __device__ __forceinline__
#else
#define HB_HD inline
#endif

enum : uint32_t { HB_REC_QUANT = 0, HB_REC_F32 = 1 };

struct RecLayout {
    uint32_t kind;    // HB_REC_QUANT (QuantVec) or HB_REC_F32 (FullVec)
    uint32_t dim;     // 8*nch + rem
    uint32_t nch;     // full 8-element chunks
    uint32_t rem;     // remainder elements (all accumulate into acc[0])
    uint32_t W;       // 16-byte words per lane slice
    uint32_t tail;    // 1: trailing [min][delta][rem..] word
    uint32_t stride;  // bytes per record
};

HB_HD constexpr uint32_t hb_layout_W(uint32_t nch) {
    return (2 * nch + 15) / 16 == 0 ? 1u : (2 * nch + 15) / 16;
}
HB_HD constexpr uint32_t hb_layout_tail(uint32_t nch, uint32_t rem) {
    return (16 * hb_layout_W(nch) - 2 * nch >= 4 && 16 * hb_layout_W(nch) - 2 * nch >= rem) ? 0u : 1u;
}

HB_HD RecLayout hb_make_layout(uint32_t dim) {
    RecLayout L;
    L.kind = HB_REC_QUANT;
    L.dim = dim;
    L.nch = dim / 8;
    L.rem = dim % 8;
    L.W = hb_layout_W(L.nch);
    L.tail = hb_layout_tail(L.nch, L.rem);
    L.stride = 64 * L.W + 16 * L.tail;
    return L;
}

// f32 records: W = number of 16-float chunks (one 16-byte word per lane per chunk); nch/rem/tail unused
HB_HD RecLayout hb_make_layout_f32(uint32_t dim) {
    RecLayout L;
    L.kind = HB_REC_F32;
    L.dim = dim;
    L.nch = dim / 8;
    L.rem = dim % 8;
    L.W = (dim + 15) / 16 == 0 ? 1u : (dim + 15) / 16;
    L.tail = 0;
    L.stride = 64 * L.W;
    return L;
}
HB_HD RecLayout hb_make_layout_kind(uint32_t dim, uint32_t kind) {
    return kind == HB_REC_F32 ? hb_make_layout_f32(dim) : hb_make_layout(dim);
}

// byte offset of element i of the vector inside its record (quantised records)
HB_HD uint32_t hb_code_offset(const RecLayout& L, uint32_t i) {
    if (i < 8 * L.nch) {
        uint32_t k = i / 8, r = i % 8, l = r / 2, p = 2 * k + (r & 1);
        return 16 * (4 * (p / 16) + l) + (p % 16);
    }
    uint32_t r = i - 8 * L.nch;
    if (L.tail) return 64 * L.W + 8 + r;
    uint32_t p = 2 * L.nch + r;
    return 16 * (4 * (p / 16) + 0) + (p % 16);
}
// Two spare floats of a quantised record, Sum y_i and Sum y_i^2 of its dequantised values, for the search kernel's
// pre-filter (csrc/search_fast.cuh).  Compact form: bytes 8..15 of the last word of lane 3's slice (free when
// 16*W - 2*nch >= 8); tail form: bytes 8..15 of the tail word (free when rem == 0).  0xFFFFFFFF: no room in this layout
// (e.g. dim 50), the search then evaluates every candidate exactly.  Written by quantise_kernel / pack_kernel; nothing
// on the exact-arithmetic path reads them.
HB_HD uint32_t hb_aux_offset(const RecLayout& L) {
    if (L.kind != HB_REC_QUANT) return 0xFFFFFFFFu;
    if (L.tail) return L.rem == 0 ? 64 * L.W + 8 : 0xFFFFFFFFu;
    return (16 * L.W >= 2 * L.nch + 8) ? 16 * (4 * (L.W - 1) + 3) + 8 : 0xFFFFFFFFu;
}
HB_HD uint32_t hb_min_offset(const RecLayout& L) {
    return L.tail ? 64 * L.W : 16 * (4 * (L.W - 1) + 1) + 12;
}
HB_HD uint32_t hb_delta_offset(const RecLayout& L) {
    return L.tail ? 64 * L.W + 4 : 16 * (4 * (L.W - 1) + 2) + 12;
}
