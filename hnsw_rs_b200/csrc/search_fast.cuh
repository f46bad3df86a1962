// HNSW::ann_by_vector, second generation of the one-warp-per-query kernel (ef <= 128, quantised records of a
// compile-time dimension, ids < 2^21) -- the path bench.py measures.
//
// Same algorithm, results and counters as search_query_reg (csrc/search_reg.cuh), which restates
// Searcher::search_layer (hnsw/src/template/searcher.rs:23-103) driven by HNSW::ann_by_vector
// (hnsw/src/template.rs:306-335) over the Results sets (hnsw/src/template/results.rs:26-33).  The round-1 kernel
// spent 61 % of its 46 k warp instructions per query on bookkeeping; what changed, by share of that:
//
//  * visited (results.rs:101-103): buckets of FOUR 15-bit entries (8 bytes, one LDS.64).  The home bucket of an id is
//    read once; "is it there" is one zero-halfword test over the two words, "where does it go" is a population count
//    of the empty marks (entries fill a bucket front to back and are never removed).  One claim + read-back step
//    instead of a lock-step probe loop (4.2 steps x 23 instructions per batch in round 1).  A full home bucket or a
//    lost claim (two new ids of one batch, same bucket) goes to a continuation loop (vis_slow) over the following buckets
//    (displacement <= 7 stored in the entry), and an id that finds 8 full buckets goes to a small exact spill list --
//    so membership, and with it the evaluation counter, stays exact (the round-1 table over-counted after an overflow).
//    (bucket, entry) names the id exactly: h = id * odd mod 2^B is a bijection on B-bit ids, bucket = floor(h * NB / 2^B),
//    and the entry keeps bits [s, B) of (h * NB) mod 2^B with 2^s < NB, which differ between two ids of one bucket because
//    those values are NB apart (s = 9 for 512 < NB <= 1024, 8 for 256 < NB <= 512; the displacement gets the bits that
//    are left of the 15: three, or two when B - s = 13).
//  * distance (vectors/src/quant.rs:14-37): the four lanes of a group no longer pass the 8-way sum from lane to lane
//    (3 dependent shuffles + sqrt + key + admission test per ROUND of 8 candidates).  Every lane stores its accumulator
//    pair to shared memory; after the last round lane i sums the eight accumulators of candidate i in the reference's
//    order, takes the square root and tests admission -- once per BATCH of up to 32 candidates.  The remainder elements
//    (dim mod 8, all added to acc[0] in order, quant.rs:31-35) are squared by one lane each instead of by every lane.
//  * the fused all-gather stores one id per lane for all peers at once.
//
// Arithmetic contract: exactly csrc/dist.cuh (packed FFMA2 restatements of single rounded products, explicit
// rounding everywhere, sequential sums in the reference's order); tools/check_sass.py covers this kernel.
#pragma once
#include "search_params.cuh"
#include "search_reg.cuh"
#include "vis_geometry.h"

#ifndef HB_FAST_PREFETCH_ALL
#define HB_FAST_PREFETCH_ALL 1  // request the record of every neighbour before the visited test (see search_reg.cuh)
#endif

// (measured and removed, profiles/r02_ab_variants.txt calls U-W: handing out the free entries of a home bucket by lane rank with one
// match.any instead of claim + read-back: -2.4 %; L1 / L2 eviction hints on the record and row loads: +-0.2 %, except
// L1::no_allocate on the pre-filter's record loads: -24 %, the four 16-byte loads of a record no longer merge into one request per line)
#ifndef HB_FAST_SLOWINL
#define HB_FAST_SLOWINL 1  // 1: vis_slow is inlined at its one call site (no call, no packing of the predicates into registers:
                           // 29 % of the batches go there; +0.95 %, call U)
#endif
#ifndef HB_FAST_PRMTPOP
#define HB_FAST_PRMTPOP 1  // 1: the free marks of a bucket are counted on the four high bytes gathered by one byte permute (call V)
#endif
#if HB_FAST_SLOWINL
#define HB_VIS_SLOW_ATTR __forceinline__
#else
#define HB_VIS_SLOW_ATTR __noinline__
#endif

namespace hb {

constexpr int FAST_SPILL = 16;        // words of the shared-memory spill area: 12 ids, their number, the number of ids in the
                                      // global set, the global slice this query holds (+1; 0 = none), one spare
constexpr int FAST_SPILL_IDS = 12;
constexpr int FAST_SPILL_N1 = 12, FAST_SPILL_N2 = 13, FAST_SPILL_SLICE = 14;
constexpr int FAST_ACC_STRIDE = 12;   // floats per candidate in the accumulator buffer: acc[0..8), remainder squares [8..12)
constexpr uint32_t FAST_SCRATCH_BYTES = 32 * FAST_ACC_STRIDE * 4;  // 1536: accumulators | admitted keys + merge buffer | query

__device__ __forceinline__ void lds_v2(uint32_t a, uint32_t& x, uint32_t& y) {
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// number of free entries (0xFFFF; valid entries are < 0x8000) of a bucket
__device__ __forceinline__ int vis_free(uint32_t w0, uint32_t w1) {
#if HB_FAST_PRMTPOP
    return __popc(__byte_perm(w0, w1, 0x7531) & 0x80808080u);
#else
    return __popc(w0 & 0x80008000u) + __popc(w1 & 0x80008000u);
#endif
}

// One probe of bucket address `a` for entry value `mine`: is it there; if not, and `want_claim`, claim the first free
// entry and read it back.  Returns found; won = this lane's claim stuck; full = no free entry.  All lanes call together.
__device__ __forceinline__ bool vis_probe(uint32_t a, uint32_t mine, bool active, bool& won, bool& full) {
    uint32_t w0, w1;
    lds_v2(a, w0, w1);
    const uint32_t pat = mine * 0x10001u;  // mine < 2^15: both halves
    const uint32_t x0 = w0 ^ pat, x1 = w1 ^ pat;
    // a zero halfword in x0 or x1 (a borrow can only flag the high half falsely when the low half already matched)
    const uint32_t z = (((x0 - 0x00010001u) & ~x0) | ((x1 - 0x00010001u) & ~x1)) & 0x80008000u;
    const bool found = z != 0u;
    // valid entries are < 0x8000, a free entry is 0xFFFF: the free ones are the last `e` of the bucket
    const int e = vis_free(w0, w1);
    const bool claim = active && !found && e > 0;
    const uint32_t sa = claim ? a + 8u - 2u * (uint32_t)e : a;
    if (claim) sts_u16(sa, mine);
    __syncwarp();
    // two lanes that claim the same entry in this step wrote different values (distinct ids of one batch, and
    // (bucket, entry) names the id): exactly the lane whose value stuck has won
    won = claim && lds_u16(sa) == mine;
    full = active && !found && e == 0;
    return found;
}

// Continuation for the lanes the single-step insert could not settle (29 % of the batches on C2; inlined at its one
// call site): lost claims retry the same bucket, full buckets move on (displacement + 1).  bit0 of the result: id is new (recorded now); bit1: 8 full buckets in a
// row, the caller consults the spill list.
__device__ HB_VIS_SLOW_ATTR uint32_t vis_slow(uint32_t sbase, uint32_t nb, uint32_t home, uint32_t mine0, bool pending,
                                              bool home_full, uint32_t dmax) {
    // a lane whose home bucket was full continues behind it; a lane that lost a claim looks at its home again
    uint32_t b = home_full ? (home + 1u == nb ? 0u : home + 1u) : home, d = home_full ? 1u : 0u;
    bool isnew = false, ovf = false;
    while (__any_sync(HB_FULL, pending)) {
        bool won, full;
        const bool found = vis_probe(sbase + b * 8u, mine0 + d, pending, won, full);
        isnew = isnew || won;
        if (full) {
            ++d;
            b = b + 1u == nb ? 0u : b + 1u;
            if (d > dmax) ovf = true;
        }
        pending = pending && !found && !won && !ovf;
    }
    return (isnew ? 1u : 0u) | (ovf ? 2u : 0u);
}

struct VisB4 {
    uint32_t sbase;   // shared-space byte address of the table: nb buckets of 8 bytes
    uint32_t nb;      // buckets, 512 < nb <= 1024
    uint32_t mul;     // odd << (32 - B): h32 = id * mul holds the B-bit bijection value top-aligned
    uint32_t rsh;     // (h32 * nb) >> rsh, low displacement bits cleared = entry with displacement 0 (see fast_vis_geometry)
    uint32_t dmax;    // largest displacement an entry can hold (7 or 3)
    uint32_t* spill;  // ids that found 8 full buckets: [0, 14) ids, [14] their number, [15] number of ids in the global continuation

    __device__ __forceinline__ void clear(int lane) const {
        const uint32_t n16 = nb / 2;  // 16-byte chunks
        for (uint32_t i = lane; i < n16; i += 32)
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(sbase + i * 16u), "r"(0xFFFFFFFFu) : "memory");
        if (lane == 0) { spill[FAST_SPILL_N1] = 0u; spill[FAST_SPILL_N2] = 0u; }
        __syncwarp();
    }
    // results.insert_visited for all 32 lanes (want: this lane holds an id).  Returns "id was not yet visited" per lane.
    // Where the table cannot decide for a lane (ovf: 8 full buckets in a row, see vis_slow) the caller's exact spill set
    // does: on_ovf(isnew, ovf) -> isnew, called by all lanes when any lane overflowed.  That call sits inside the slow-path
    // branch -- an id can only overflow there -- and not behind a vote of its own in every batch (+0.65 %,
    // profiles/r02_ab_variants.txt call U).
    template <class OVF>
    __device__ __forceinline__ bool insert_warp(uint32_t id, bool want, OVF&& on_ovf) const {
        uint32_t home, mine0;
        fast_vis_slot(mul, rsh, dmax, nb, id, home, mine0);
        bool won, full;
        const bool found = vis_probe(sbase + home * 8u, mine0, want, won, full);
        const bool pending = want && !found && !won;
        if (__any_sync(HB_FULL, pending)) {
            const uint32_t r = vis_slow(sbase, nb, home, mine0, pending, full, dmax);
            won = won || (r & 1u);
            const bool ovf = (r & 2u) != 0u;
            if (__any_sync(HB_FULL, ovf)) won = on_ovf(won, ovf);  // rare
        }
        return won;
    }
};

// Global continuation of the spill set: a small pool of slices (hash sets of `cap` 32-bit ids, 0xFFFFFFFF = free) that a
// query borrows when its 12 shared-memory entries are used up and returns, wiped, when it is done.  pool[0..npool) are
// the owner words (0 = free), the slices follow.  Overlapping launches of one context (PDL) share the pool safely:
// a slice has one owner at a time.
struct SpillPool {
    uint32_t* base;
    uint32_t npool, cap;  // cap: a power of two >= 32
    __device__ __forceinline__ uint32_t* slice(uint32_t i) const { return base + npool + (size_t)i * cap; }
};

// Out-of-line: the lanes in `om` hold ids that found 8 full buckets in a row.  The exact spill set decides whether such an
// id was seen before: up to 12 ids in shared memory; once those are used up, a borrowed global slice used as an open-
// addressing hash set in which ALL the lanes concerned probe at once (compare-and-swap on 32-bit entries; the ids of a batch
// are distinct).  When no slice can be had or a probe sequence runs too long, list membership keeps the answers exact and
// only the evaluation counter may over-count (flag bit 1).
// Returns bit0: this lane's "id is new"; bits 1-2: the flag bits to raise.
template <int KPL>
__device__ __noinline__ uint32_t vis_spill(uint32_t* spill, SpillPool pool, unsigned om, uint32_t nb, bool isnew,
                                           const RegList<KPL> L, int lane) {
    uint32_t fl = 4u;
    const bool mine = (om >> lane) & 1u;
    const uint32_t n1 = spill[FAST_SPILL_N1];
    bool found = false;
    for (uint32_t i = 0; i < n1; ++i) found = found || spill[i] == nb;  // broadcast reads
    bool pending = mine && !found;
    unsigned pm = __ballot_sync(HB_FULL, pending);
    if (pm == 0u) return fl | (isnew ? 1u : 0u);
    uint32_t sl = spill[FAST_SPILL_SLICE];  // slice index + 1
    if (sl == 0u && n1 + (uint32_t)__popc(pm) <= (uint32_t)FAST_SPILL_IDS) {  // room in the shared-memory list
        if (pending) spill[n1 + __popc(pm & ((1u << lane) - 1u))] = nb;
        if (lane == 0) spill[FAST_SPILL_N1] = n1 + (uint32_t)__popc(pm);
        __syncwarp();
        return fl | ((isnew || pending) ? 1u : 0u);
    }
    if (sl == 0u && pool.base) {  // borrow a slice
        if (lane == 0) {
            const uint32_t start = (blockIdx.x * SEARCH_WPB + (threadIdx.x >> 5)) % pool.npool;
            for (uint32_t k = 0; k < pool.npool && sl == 0u; ++k) {
                const uint32_t i = start + k < pool.npool ? start + k : start + k - pool.npool;
                if (atomicCAS(pool.base + i, 0u, 1u) == 0u) sl = i + 1u;
            }
            __threadfence();  // the previous owner's wipe is visible before this query reads the slice
            spill[FAST_SPILL_SLICE] = sl;
        }
        sl = __shfl_sync(HB_FULL, sl, 0);
    }
    bool fresh = false, over = pending && sl == 0u;
    if (sl != 0u) {
        uint32_t* g = pool.slice(sl - 1u);
        const uint32_t mask = pool.cap - 1u;
        uint32_t h = ((nb * 0x9E3779B1u) >> 7) & mask;
        int probes = 0;
        while (__any_sync(HB_FULL, pending)) {
            uint32_t old = nb;
            if (pending) old = atomicCAS(g + h, 0xFFFFFFFFu, nb);
            const bool won = pending && old == 0xFFFFFFFFu;
            const bool step = pending && !won && old != nb;
            fresh = fresh || won;
            h = step ? ((h + 1u) & mask) : h;
            probes += step ? 1 : 0;
            over = over || (step && probes >= 96);  // the set is too full for this id
            pending = step && !over;
        }
        const unsigned wm = __ballot_sync(HB_FULL, fresh);
        if (lane == 0 && wm) spill[FAST_SPILL_N2] = spill[FAST_SPILL_N2] + (uint32_t)__popc(wm);
    }
    unsigned ovm = __ballot_sync(HB_FULL, over);
    if (ovm) {
        fl |= 2u;
        while (ovm) {
            const int src = __ffs(ovm) - 1;
            ovm &= ovm - 1;
            const bool in_list = L.holds_id(__shfl_sync(HB_FULL, nb, src));
            if (lane == src) fresh = !in_list;
        }
    }
    __syncwarp();
    return fl | ((isnew || fresh) ? 1u : 0u);
}

// A query that borrowed a global slice wipes it and gives it back (at the end of every layer search that used it).
__device__ __noinline__ void spill_release(uint32_t* spill, SpillPool pool, int lane) {
    const uint32_t sl = spill[FAST_SPILL_SLICE];
    if (sl == 0u) return;
    uint32_t* g = pool.slice(sl - 1u);
    for (uint32_t i = lane * 4u; i < pool.cap; i += 128u)
        *reinterpret_cast<uint4*>(g + i) = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        atomicExch(pool.base + (sl - 1u), 0u);
        spill[FAST_SPILL_SLICE] = 0u;
    }
    __syncwarp();
}

// Query of a compile-time dimension (8*NCH + REM, REM <= 4) in registers; evaluates one record per group of 4 lanes
// and leaves the accumulators to the caller (no cross-lane sum here).
#ifndef HB_FAST_QSMEM
#define HB_FAST_QSMEM 1  // 1: the query's value pairs stay in shared memory (one LDS.64 per chunk) instead of 2*NCH registers:
                         // -2 % at equal occupancy, but 64 registers and an eighth block per SM: +5 % (profiles/r02_ab_variants.txt)
#endif
#ifndef HB_FAST_PIPE
#define HB_FAST_PIPE 0   // 1: the record of the next round is loaded before the current one is evaluated
#endif
#ifndef HB_FAST_ROWPF
#define HB_FAST_ROWPF 1  // 1: the adjacency row of every admitted key is requested (L2) at admission
#endif
#ifndef HB_FAST_FILTER
#define HB_FAST_FILTER 1  // 1: integer pre-filter (dp4a) before the exact evaluation, see FastQuery::prefilter
#endif
constexpr uint32_t FAST_QS_BYTES = HB_FAST_QSMEM ? 16 * 32 + 0 : 0;  // up to 16 chunks x 4 lanes x 8 bytes
// pre-filter operands: the query's codes in record layout (128 bytes), (dq, dq * Sum cq, mq, Sum x^2), 32 surviving ids
constexpr uint32_t FAST_QC_BYTES = HB_FAST_FILTER ? 128 + 16 + 128 : 0;

template <int NCH, int REM>
struct FastQuery {
    static_assert(REM <= 4, "one remainder element per lane of a group");
    static_assert(NCH <= 16, "shared-memory query area holds 16 chunks");
    using RQ = RegQuery<NCH, REM>;
    static constexpr int W = RQ::W;
    static constexpr int TAIL = RQ::TAIL;
    using Rec = typename RQ::Rec;
    static constexpr int kRem = REM;
    static constexpr uint32_t kStride = 64u * W + 16u * TAIL;  // bytes per record (csrc/layout.h)
    // the record carries Sum y and Sum y^2 (layout.h: hb_aux_offset) -> the integer pre-filter can run
    static constexpr bool kAux = HB_FAST_FILTER && (TAIL ? (REM == 0) : (16 * W >= 2 * NCH + 8));
    // request the lines of one record (L2)
    __device__ __forceinline__ static void prefetch(const uint8_t* rp) {
        prefetch_l2(rp);
        if (kStride > 128u || (kStride % 128u) != 0u) prefetch_l2(rp + kStride - 1);  // a record that can straddle lines
    }
    // remainder bytes: tail form -> bytes 8..11 of the tail word; compact form -> lane 0's slice positions 2*NCH + r
    static constexpr int RP0 = TAIL ? 8 : 2 * NCH;             // slice position (or tail byte) of remainder element 0
    static constexpr bool STRADDLE = REM > 0 && (RP0 % 4) + REM > 4;  // the REM bytes span two 32-bit words
#if HB_FAST_QSMEM
    static constexpr int QROW = (NCH + 1) / 2 * 2;  // chunks per lane of a group, padded to whole 16-byte loads
    const u64* qs;    // this lane's row of the [lane of group][chunk] table in shared memory (16-byte aligned)
    __device__ __forceinline__ u64 qk(int k) const { return qs[k]; }
    // two chunks per 16-byte load
    __device__ __forceinline__ void qk2(int k, u64& a, u64& b) const {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(qs + k);
        a = v.x;
        b = v.y;
    }
#else
    u64 q[NCH ? NCH : 1];
    __device__ __forceinline__ u64 qk(int k) const { return q[k]; }
#endif
    float qrem;       // query value of remainder element gl (lanes gl >= REM: 0)
    u64 nz;
    const uint4* qcode;   // pre-filter: this lane's slice of the query's codes in record layout (words gl, gl + 4, ...)
    const float4* qaux;   // pre-filter: (dq, dq * Sum cq, mq, Sum x^2)

    // qd: the dequantised query in shared memory (natural order); qtab: FAST_QS_BYTES of shared memory that stay valid
    // for the whole query (only used when the query lives in shared memory)
    __device__ __forceinline__ void init(const RecLayout&, const float* qd, int gl, u64* qtab, int lane) {
#if HB_FAST_QSMEM
        for (int i = lane; i < 4 * NCH; i += 32) {  // entry (l, k) = values 8k+2l, 8k+2l+1
            const int k = i >> 2, l = i & 3;
            float2 v = *reinterpret_cast<const float2*>(qd + 8 * k + 2 * l);
            qtab[l * QROW + k] = pk(v.x, v.y);
        }
        qs = qtab + gl * QROW;
#else
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            float2 v = *reinterpret_cast<const float2*>(qd + 8 * k + 2 * gl);
            q[k] = pk(v.x, v.y);
        }
#endif
        qrem = (REM > 0 && gl < REM) ? qd[8 * NCH + gl] : 0.0f;
        nz = hb_negzero2;
        qcode = nullptr;
        qaux = nullptr;
    }
    // Pre-filter operands.  ctmp: the query's codes in natural order (dim bytes); qc: 128 + 16 bytes of shared memory that
    // stay valid for the whole query; mn / dl: the query's own quantiser.  Call after init (all 32 lanes).
    __device__ __forceinline__ void init_filter(const RecLayout& L, const float* qd, const uint8_t* ctmp, float mn, float dl,
                                                unsigned char* qc, int gl, int lane) {
        uint32_t* w = reinterpret_cast<uint32_t*>(qc);
        w[lane] = 0u;  // 128 bytes
        __syncwarp();
        uint32_t sc = 0;
        float qq = 0.0f;
        for (uint32_t i = lane; i < L.dim; i += 32) {
            const uint32_t c = ctmp[i];
            qc[hb_code_offset(L, i)] = (unsigned char)c;
            sc += c;
            qq += qd[i] * qd[i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sc += __shfl_xor_sync(HB_FULL, sc, o);
            qq += __shfl_xor_sync(HB_FULL, qq, o);
        }
        if (lane == 0) *reinterpret_cast<float4*>(qc + 128) = make_float4(dl, dl * (float)sc, mn, qq);
        qcode = reinterpret_cast<const uint4*>(qc) + gl;
        qaux = reinterpret_cast<const float4*>(qc + 128);
    }

    // Integer pre-filter (the algebra of the tensor-core brute-force filter, csrc/bf_tc.cu, on CUDA cores): with
    // x_i = cq_i*dq + mq and y_i = cb_i*db + mb,
    //     d^2 = Sum x^2 + Sum y^2 - 2 (dq*db*<cq, cb> + (dq * Sum cq) * mb + mq * Sum y),
    // and <cq, cb> is 8 dp4a per lane on the record bytes the lane holds anyway.  The value differs from the reference's
    // separately rounded chain by a few 1e-7 of Sum x^2 + Sum y^2, so it only ever REJECTS: a candidate whose estimate
    // exceeds T = worst^2 * (1 + 1e-4) by more than 1e-5 * (Sum x^2 + Sum y^2) has an exact key above the admission bound,
    // the reference would evaluate it and drop it (searcher.rs:74-94), and so it is counted as evaluated and skipped.
    // Everything else (NaN / inf included) goes to the exact arithmetic.  Returns "may be admitted"; meaningful in the
    // lanes that hold the aux floats: lane 3 of a group (compact records), every lane (tail records).
    __device__ __forceinline__ bool prefilter(const Rec& R, int gl, int gbase, float T) const {
        const uint4 (&w)[W] = R.w;
        float mn, dl, sy, bc;
        if (TAIL) {
            mn = __uint_as_float(R.tw.x);
            dl = __uint_as_float(R.tw.y);
            sy = __uint_as_float(R.tw.z);
            bc = __uint_as_float(R.tw.w);
        } else {
            const uint32_t last = w[W - 1].w;
            mn = __uint_as_float(__shfl_sync(HB_FULL, last, gbase + 1));
            dl = __uint_as_float(__shfl_sync(HB_FULL, last, gbase + 2));
            sy = __uint_as_float(w[W - 1].z);  // lane 3's own words
            bc = __uint_as_float(last);
        }
        unsigned dot = 0;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const uint4 q4 = qcode[4 * j];
            dot = __dp4a(w[j].x, q4.x, dot);
            dot = __dp4a(w[j].y, q4.y, dot);
            dot = __dp4a(w[j].z, q4.z, dot);
            dot = __dp4a(w[j].w, q4.w, dot);
        }
        dot += __shfl_xor_sync(HB_FULL, dot, 1);
        dot += __shfl_xor_sync(HB_FULL, dot, 2);
        const float4 a = *qaux;
        // (separate multiplies and adds: tools/check_sass.py admits no scalar FFMA in this kernel, and three instructions
        // per round are a fair price for keeping that guard simple)
        const float t = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(a.x, dl), (float)dot), __fmul_rn(a.y, mn)), __fmul_rn(a.z, sy));
        const float nrm = __fadd_rn(a.w, bc);
        const float est = __fadd_rn(nrm, __fmul_rn(-2.0f, t));
        return !(est > __fadd_rn(T, __fmul_rn(1e-5f, nrm)));
    }
    __device__ __forceinline__ static Rec load(const uint8_t* __restrict__ rec, int gl) { return RQ::load(rec, gl); }

    // The pre-filter with TWO lanes per record (16 candidates per round instead of 8): lane h = lane & 1 of a pair holds
    // slices h and 2 + h of the record (w[2j + e] = 16-byte word j of slice 2e + h), so that each load instruction reads one
    // whole 32-byte sector per pair.  The dot product is an exact integer and the float expression is the one above, so
    // the survivors are the same candidates.  Meaningful in lane h == 1 (it holds min, Sum y, Sum y^2; delta comes from
    // lane h == 0) or in every lane (tail records).
    struct Rec2 {
        uint4 w[2 * W];
        uint4 tw;
    };
    __device__ __forceinline__ static Rec2 load2(const uint8_t* __restrict__ rec, int h, bool act) {
        const uint4* p = reinterpret_cast<const uint4*>(rec) + h;
        Rec2 r;
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < W; ++j) {
            r.w[2 * j] = act ? __ldg(p + 4 * j) : z;
            r.w[2 * j + 1] = act ? __ldg(p + 4 * j + 2) : z;
        }
        r.tw = z;
        if (TAIL) r.tw = act ? __ldg(p - h + 4 * W) : z;
        return r;
    }
    // a group without a candidate in this round loads nothing
    __device__ __forceinline__ static Rec load_if(const uint8_t* __restrict__ rec, int gl, bool act) {
        const uint4* p = reinterpret_cast<const uint4*>(rec);
        Rec r;
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < W; ++j) r.w[j] = act ? __ldg(p + 4 * j + gl) : z;
        r.tw = z;
        if (TAIL) r.tw = act ? __ldg(p + 4 * W) : z;
        return r;
    }
    __device__ __forceinline__ bool prefilter2(const Rec2& R, int h, int lane, float T) const {
        float mn, dl, sy, bc;
        if (TAIL) {
            mn = __uint_as_float(R.tw.x);
            dl = __uint_as_float(R.tw.y);
            sy = __uint_as_float(R.tw.z);
            bc = __uint_as_float(R.tw.w);
        } else {
            mn = __uint_as_float(R.w[2 * W - 2].w);                                         // slice 1 (h == 1, e == 0)
            dl = __uint_as_float(__shfl_sync(HB_FULL, R.w[2 * W - 1].w, lane & ~1));         // slice 2 (h == 0, e == 1)
            sy = __uint_as_float(R.w[2 * W - 1].z);                                         // slice 3 (h == 1, e == 1)
            bc = __uint_as_float(R.w[2 * W - 1].w);
        }
        const uint4* qc2 = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(qaux) - 128) + h;
        unsigned dot = 0;
#pragma unroll
        for (int j = 0; j < 2 * W; ++j) {
            const uint4 q4 = qc2[2 * j];  // word j/2 of slice 2*(j%2) + h: 16-byte index 4*(j/2) + 2*(j%2) + h
            dot = __dp4a(R.w[j].x, q4.x, dot);
            dot = __dp4a(R.w[j].y, q4.y, dot);
            dot = __dp4a(R.w[j].z, q4.z, dot);
            dot = __dp4a(R.w[j].w, q4.w, dot);
        }
        dot += __shfl_xor_sync(HB_FULL, dot, 1);
        const float4 a = *qaux;
        const float t = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(a.x, dl), (float)dot), __fmul_rn(a.y, mn)), __fmul_rn(a.z, sy));
        const float nrm = __fadd_rn(a.w, bc);
        const float est = __fadd_rn(nrm, __fmul_rn(-2.0f, t));
        return !(est > __fadd_rn(T, __fmul_rn(1e-5f, nrm)));
    }

    // acc: this lane's accumulator pair (acc[2gl], acc[2gl+1]) after the full chunks; rsq: the square of remainder
    // element gl (meaningful for gl < REM).  All 32 lanes call together.
    __device__ __forceinline__ void partial(const Rec& R, int gl, int gbase, u64& acc_out, float& rsq) const {
        const uint4 (&w)[W] = R.w;
        float mn, dl;
        uint32_t rw = 0;  // the word holding this lane's remainder byte
        if (TAIL) {
            mn = __uint_as_float(R.tw.x);
            dl = __uint_as_float(R.tw.y);
            if (REM > 0) rw = R.tw.z;
        } else {
            const uint32_t last = w[W - 1].w;
            mn = __uint_as_float(__shfl_sync(HB_FULL, last, gbase + 1));
            dl = __uint_as_float(__shfl_sync(HB_FULL, last, gbase + 2));
            if (REM > 0) {
                constexpr int P = 2 * NCH;  // lane 0's slice position of remainder element 0
                const uint32_t a = word32(w[P / 16], (P % 16) / 4);
                rw = __shfl_sync(HB_FULL, a, gbase);
                if (STRADDLE) {
                    constexpr int P1 = 2 * NCH + REM - 1;
                    const uint32_t b = __shfl_sync(HB_FULL, word32(w[P1 / 16], (P1 % 16) / 4), gbase);
                    rw = ((P % 4) + gl >= 4) ? b : rw;
                }
            }
        }
        const u64 mn2 = pk(mn, mn);
        u64 acc = pk(0.0f, 0.0f);
        rsq = 0.0f;
        // PRMT selector: byte of remainder element gl, then 0x00 0x00 0x4B (the magic-number form, dist.cuh)
        const uint32_t remsel = 0x7540u | (uint32_t)((RP0 + gl) & 3);
        if (__any_sync(HB_FULL, !(dl < 1.2676506e30f))) {
            // delta >= 2^100, inf or NaN somewhere in this round: separately rounded multiplies (dist.cuh)
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int j = k / 8, c = k % 8;
                acc = chunk_acc(acc, word32(w[j], c / 2), (c & 1) * 2, dl, mn2, qk(k));
            }
            if (REM > 0) {
                const float fm = __uint_as_float(__byte_perm(rw, 0x4B000000u, remsel));
                rsq = rem_acc(0.0f, fm, dl, mn, qrem);  // 0 + t*t == t*t
            }
        } else {
            const u64 dl2 = pk(dl, dl);
            const float nmd = __fmul_rn(-8388608.0f, dl);
            const u64 nmd2 = pk(nmd, nmd);
#if HB_FAST_QSMEM
#pragma unroll
            for (int k = 0; k + 1 < NCH; k += 2) {
                u64 qa, qb;
                qk2(k, qa, qb);
                const u64 ma = pk(RQ::slice_magic(w, 2 * k), RQ::slice_magic(w, 2 * k + 1));
                acc = add2(acc, chunk_sq(ma, dl2, nmd2, mn2, qa, nz));
                const u64 mb = pk(RQ::slice_magic(w, 2 * k + 2), RQ::slice_magic(w, 2 * k + 3));
                acc = add2(acc, chunk_sq(mb, dl2, nmd2, mn2, qb, nz));
            }
            if (NCH & 1) {
                constexpr int k = NCH - 1;
                const u64 m2 = pk(RQ::slice_magic(w, 2 * k), RQ::slice_magic(w, 2 * k + 1));
                acc = add2(acc, chunk_sq(m2, dl2, nmd2, mn2, qk(k), nz));
            }
#else
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const u64 m2 = pk(RQ::slice_magic(w, 2 * k), RQ::slice_magic(w, 2 * k + 1));
                acc = add2(acc, chunk_sq(m2, dl2, nmd2, mn2, qk(k), nz));
            }
#endif
            if (REM > 0) {
                // one remainder element per lane; the second half of the packed pair is a dummy (code 0 against 0)
                const float fm = __uint_as_float(__byte_perm(rw, 0x4B000000u, remsel));
                float s0, s1;
                up(chunk_sq(pk(fm, 8388608.0f), dl2, nmd2, mn2, pk(qrem, 0.0f), nz), s0, s1);
                rsq = s0;
            }
        }
        acc_out = acc;
    }
};

// per warp (wsm, 16-byte aligned): 32 candidate ids | spill list | worst key | scratch | visited buckets
constexpr uint32_t FAST_OFF_SPILL = 128, FAST_OFF_WORST = FAST_OFF_SPILL + FAST_SPILL * 4, FAST_OFF_SCRATCH = FAST_OFF_WORST + 16,
                   FAST_OFF_QTAB = FAST_OFF_SCRATCH + FAST_SCRATCH_BYTES, FAST_OFF_QCODE = FAST_OFF_QTAB + FAST_QS_BYTES,
                   FAST_OFF_QAUX = FAST_OFF_QCODE + 128, FAST_OFF_SURV = FAST_OFF_QAUX + 16,
                   FAST_OFF_TABLE = FAST_OFF_QCODE + FAST_QC_BYTES;

#ifndef HB_FAST_SPEC
#define HB_FAST_SPEC 0  // (measured: -5 %) load the adjacency row of the probable next expansion one hop ahead and prefetch its records
#endif
#ifndef HB_FAST_FILTER2
#define HB_FAST_FILTER2 1  // the pre-filter runs with two lanes per record (16 candidates per round)
#endif
#ifndef HB_FAST_PREDLD
#define HB_FAST_PREDLD 0  // 1: groups / pairs without a candidate in a round do not load a record (measured: -2 %, the selects cost more than the requests)
#endif
#ifndef HB_FAST_INS1
#define HB_FAST_INS1 3  // a batch that admits at most this many keys inserts them in registers (no shared-memory merge); 0: off
#endif
#ifndef HB_FAST_MERGE
#define HB_FAST_MERGE 1  // 1: merge_ranked below; 0: RegList::merge (round 1)
#endif

// candidates.first() without popping: the id of the first entry whose "expanded" bit is clear
template <int KPL>
__device__ __forceinline__ bool list_peek(const RegList<KPL>& L, uint32_t& cid) {
    uint32_t pick = 0xFFFFFFFFu;
#pragma unroll
    for (int s = KPL - 1; s >= 0; --s) {
        const uint32_t lo = (uint32_t)L.v[s];
        pick = (lo & 1u) ? pick : lo;
    }
    const unsigned m = __ballot_sync(HB_FULL, !(pick & 1u));
    if (!m) return false;
    cid = __shfl_sync(HB_FULL, pick, __ffs(m) - 1) >> 1;
    return true;
}

// selected <- ef smallest of (selected U new keys): the same result as RegList::merge (and as the reference's one-by-one
// admission, searcher.rs:74-94), with the work arranged so that nothing but the rank among the NEW keys loops over them:
//   * the sorted list is dumped to shared memory once (one 16-byte store per lane); every lane that holds a new key
//     finds its rank in it by a branch-free binary search (log2(32*KPL) steps);
//   * rank among the new keys: one pass over the m compacted new keys (the only O(m) part, 7 instructions per key);
//   * a new key's final position is the sum of its two ranks; an OR-reduction (REDUX) of those positions gives the
//     occupancy mask of the merged list, and every list slot then knows whether it takes a new key (the c-th smallest,
//     c = occupied positions below it) or the old key c places to its left.
// key/want: this lane's new key; am = ballot(want); obuf: 32*KPL keys, kbuf and sbuf: 32 keys each.
template <int KPL>
__device__ __forceinline__ void merge_ranked(RegList<KPL>& L, u64 key, bool want, unsigned am, u64* obuf, u64* kbuf, u64* sbuf,
                                             int& len, int ef, int lane, u64* worst_p) {
    constexpr int C = 32 * KPL;
    const int m = __popc(am);
#if HB_FAST_INS1
    if (m <= HB_FAST_INS1) {
        // one new key (or a few): insertions in registers, one after the other.  p = number of list keys below the key;
        // slot q keeps its key (q < p), takes the new one (q == p) or the key one place to its left (q > p); a key that is
        // no longer below the bound by the time its turn comes has p == ef and falls off the end, as in the merge.
#pragma unroll 1
        for (unsigned todo = am; todo; todo &= todo - 1u) {
            const u64 k1 = __shfl_sync(HB_FULL, key, __ffs(todo) - 1);
            int p = 0;
#pragma unroll
            for (int s = 0; s < KPL; ++s) p += __popc(__ballot_sync(HB_FULL, L.v[s] < k1));
            const u64 from_left = __shfl_up_sync(HB_FULL, L.v[KPL - 1], 1);
            const int newlen = min(ef, len + 1);
#pragma unroll
            for (int s = KPL - 1; s >= 0; --s) {  // downwards: L.v[s - 1] is still the old key when slot s reads it
                const int q = lane * KPL + s;
                const u64 left = s ? L.v[s ? s - 1 : 0] : from_left;
                u64 v = q < p ? L.v[s] : (q == p ? k1 : left);
                v = q < newlen ? v : RSENT;
                L.v[s] = v;
                if (q == ef - 1 && newlen == ef) *worst_p = v;
            }
            len = newlen;
        }
        __syncwarp();
        return;
    }
#endif
    __syncwarp();  // every lane has read its accumulators: the merge scratch may overwrite them
    // old list -> shared memory (position lane*KPL + s), compacted new keys -> kbuf
#pragma unroll
    for (int s = 0; s < KPL; s += 2)
        *reinterpret_cast<ulonglong2*>(obuf + lane * KPL + s) = make_ulonglong2(L.v[s], L.v[s + 1]);
    if (want) kbuf[__popc(am & ((1u << lane) - 1u))] = key;
    __syncwarp();
    // rank in the old list: number of list keys below `key` (the list ends in sentinels or the key is below list[ef-1],
    // so the rank is at most C - 1 and C/2 + ... + 1 steps reach it)
    int rl = 0;
#pragma unroll
    for (int st = C / 2; st >= 1; st >>= 1) rl += (obuf[rl + st - 1] < key) ? st : 0;
    // rank among the new keys
    int rn = 0;
    {
        int j = 0;
#pragma unroll 1
        for (; j + 1 < m; j += 2) {  // two keys per 16-byte (broadcast) load
            const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(kbuf + j);
            rn += (kk.x < key) ? 1 : 0;
            rn += (kk.y < key) ? 1 : 0;
        }
        if (j < m) rn += (kbuf[j] < key) ? 1 : 0;
    }
    const int pos = rl + rn;
    if (want) sbuf[rn] = key;  // the new keys in ascending order
    // occupancy of the merged list by new keys, one word per 32 positions
    uint32_t occ[KPL];
#pragma unroll
    for (int w = 0; w < KPL; ++w)
        occ[w] = __reduce_or_sync(HB_FULL, (want && (pos >> 5) == w) ? (1u << (pos & 31)) : 0u);
    __syncwarp();
    const int newlen = min(ef, len + m);
    // this lane's KPL slots lie in one word of the mask
    const int q0 = lane * KPL, wq = q0 >> 5;
    uint32_t mine = occ[0];
    int below = 0;
#pragma unroll
    for (int w = 1; w < KPL; ++w) {
        below += (w <= wq) ? __popc(occ[w - 1]) : 0;
        mine = (w == wq) ? occ[w] : mine;
    }
    int c = below + __popc(mine & ((1u << (q0 & 31)) - 1u));
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
        const int q = q0 + s;
        const bool isn = (mine >> (q & 31)) & 1u;
        u64 v = isn ? sbuf[c] : obuf[q - c];
        v = q < newlen ? v : RSENT;
        L.v[s] = v;
        if (q == ef - 1 && newlen == ef) *worst_p = v;  // the new admission bound (it stays the sentinel while |selected| < ef)
        c += isn ? 1 : 0;
    }
    len = newlen;
    __syncwarp();
}

// One whole query.  On exit L holds the <= ef nearest evaluated nodes of layer 0, sorted.
// pool: the global continuation of the spill set (base may be null).
template <class Q, int KPL, bool STATS>
__device__ __forceinline__ void search_query_fast(const Q& query, const uint8_t* __restrict__ rec, uint32_t rec_stride,
                                                  const GraphView& g, uint32_t n_layers, uint32_t ep, RegList<KPL>& L,
                                                  const VisB4& vis, unsigned char* wsm, const SpillPool pool, int ef, int lane,
                                                  SearchCounters& cnt) {
    uint32_t* newbuf = reinterpret_cast<uint32_t*>(wsm);
    float* scratch = reinterpret_cast<float*>(wsm + FAST_OFF_SCRATCH);
    // key at position ef_l - 1 of the list (the sentinel, i.e. the maximum, while |selected| < ef_l): kept in shared
    // memory, it is read once per batch
    u64* worst_p = reinterpret_cast<u64*>(wsm + FAST_OFF_WORST);
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    const unsigned lt = (1u << lane) - 1u;
    float* accbuf = scratch;
    // merge scratch (aliases the accumulators): old list | compacted new keys | sorted new keys
    u64* obuf = reinterpret_cast<u64*>(scratch);
    u64* kbuf = obuf + 32 * KPL;
    u64* sbuf = kbuf + 32;
    static_assert((32 * KPL + 64) * 8 <= FAST_SCRATCH_BYTES, "merge scratch does not fit");
    uint32_t layer = n_layers - 1;
    int len = 0;  // |selected|
    L.reset();
    if (lane == 0) {
        *worst_p = RSENT;
        vis.spill[FAST_SPILL_SLICE] = 0u;
    }
    vis.clear(lane);
    // boot batch: the entry point (selected <- {Dist(ep, d)}, template.rs:316-319)
    uint32_t nb = lane == 0 ? ep : EMPTY_ID;
    bool seed = false;
    bool fresh_row = false;  // nb was just loaded for a popped candidate and its records have not been requested yet
    uint32_t row = EMPTY_ID, next = EMPTY_ID, b0 = 0;
    // speculation: spec_nb = this lane's slot of the adjacency row of spec_id, the probable next expansion
    uint32_t spec_id = EMPTY_ID, spec_nb = EMPTY_ID;
    bool spec_pending = false;  // the records of spec_nb's row are still to be requested
#pragma unroll 1
    while (true) {
        const int ef_l = layer ? 1 : ef;
        // ---- one batch of up to 32 ids: results.insert_visited(node) (results.rs:101-103) ----
        const bool valid = !(nb & CHAIN_BIT);  // EMPTY_ID and chain markers carry bit 31
#if HB_FAST_PREFETCH_ALL
        // request the record of every neighbour before the visited test (unless that happened a hop ago: speculation hit)
        if (fresh_row && valid) Q::prefetch(rec + (size_t)nb * Q::kStride);
#endif
        fresh_row = false;
        bool isnew = vis.insert_warp(nb, valid, [&](bool isnew_in, bool ovf) -> bool {
            const uint32_t r = vis_spill<KPL>(vis.spill, pool, __ballot_sync(HB_FULL, ovf), nb, isnew_in, L, lane);
            if (STATS) cnt.overflow |= r & 6u;
            return (r & 1u) != 0u;
        });
        isnew = isnew && !seed;
        const unsigned nm = __ballot_sync(HB_FULL, isnew);
        const int ncnt = __popc(nm);
#if HB_FAST_SPEC == 1
        if (spec_pending) {
            // the row of the probable next expansion has arrived by now (it was requested at the pop): request the records
            // of its neighbours a whole hop before they are evaluated
            if (!(spec_nb & CHAIN_BIT)) Q::prefetch(rec + (size_t)spec_nb * Q::kStride);
            spec_pending = false;
        }
#endif
        if (ncnt) {
            if (STATS) cnt.evals += ncnt;
            if (isnew) {
                const int my = __popc(nm & lt);
                newbuf[my] = nb;
#if !HB_FAST_PREFETCH_ALL
                Q::prefetch(rec + (size_t)nb * Q::kStride);
#endif
            }
            __syncwarp();
            // ---- which of the ncnt candidates reach the exact evaluation ----
            const uint32_t* cbuf = newbuf;  // their ids
            int ecnt = ncnt;
            if (Q::kAux) {
                const u64 wk = *worst_p;
                if (wk != RSENT) {  // |selected| == ef: a key is admitted only below the bound
                    const float wd = __uint_as_float((uint32_t)(wk >> 32));
                    const float T = __fmul_rn(__fmul_rn(wd, wd), 1.0001f);
                    uint32_t* surv = reinterpret_cast<uint32_t*>(wsm + FAST_OFF_SURV);
                    int sc = 0;
#if HB_FAST_FILTER2
#pragma unroll 1
                    for (int r0 = 0; r0 < ncnt; r0 += 16) {
                        const int idx = r0 + (lane >> 1), h = lane & 1;
                        const uint32_t cand = newbuf[idx < ncnt ? idx : 0];
                        const bool keep = query.prefilter2(Q::load2(rec + (size_t)cand * Q::kStride, h, HB_FAST_PREDLD ? idx < ncnt : true), h, lane, T);
                        const bool sv = keep && h == 1 && idx < ncnt;
                        const unsigned sm = __ballot_sync(HB_FULL, sv);
                        if (sv) surv[sc + __popc(sm & lt)] = cand;
                        sc += __popc(sm);
                    }
#else
#pragma unroll 1
                    for (int r0 = 0; r0 < ncnt; r0 += 8) {
                        const int idx = r0 + grp;
                        const uint32_t cand = newbuf[idx < ncnt ? idx : 0];
                        const bool keep = query.prefilter(Q::load(rec + (size_t)cand * Q::kStride, gl), gl, gbase, T);
                        const bool sv = keep && gl == 3 && idx < ncnt;
                        const unsigned sm = __ballot_sync(HB_FULL, sv);
                        if (sv) surv[sc + __popc(sm & lt)] = cand;
                        sc += __popc(sm);
                    }
#endif
                    __syncwarp();
                    cbuf = surv;
                    ecnt = sc;
                }
            }
            if (ecnt == 0) goto next_batch;  // every candidate of the batch is above the admission bound
#if HB_FAST_PIPE
            {
                // two record buffers: the loads of round r+1 are in flight while round r is evaluated
                auto rec_of = [&](int r0) {
                    const int idx = r0 + grp;
                    return Q::load(rec + (size_t)cbuf[idx < ecnt ? idx : 0] * Q::kStride, gl);
                };
                auto eval = [&](const typename Q::Rec& R, int r0) {
                    const int idx = r0 + grp;
                    u64 acc;
                    float rsq;
                    query.partial(R, gl, gbase, acc, rsq);
                    *reinterpret_cast<u64*>(accbuf + idx * FAST_ACC_STRIDE + 2 * gl) = acc;
                    if (Q::kRem > 0) accbuf[idx * FAST_ACC_STRIDE + 8 + gl] = rsq;
                };
                typename Q::Rec A = rec_of(0), B;
#pragma unroll 1
                for (int r0 = 0;; r0 += 16) {
                    const bool moreB = r0 + 8 < ecnt;
                    if (moreB) B = rec_of(r0 + 8);
                    eval(A, r0);
                    if (!moreB) break;
                    const bool moreA = r0 + 16 < ecnt;
                    if (moreA) A = rec_of(r0 + 16);
                    eval(B, r0 + 8);
                    if (!moreA) break;
                }
            }
#else
#pragma unroll 1
            for (int r0 = 0; r0 < ecnt; r0 += 8) {
                const int idx = r0 + grp;
                const uint32_t cand = cbuf[idx < ecnt ? idx : 0];
                // index.get_point(node).dist2other(point)  (searcher.rs:66-69), the sum left to the lanes below
                u64 acc;
                float rsq;
                query.partial(Q::load_if(rec + (size_t)cand * Q::kStride, gl, HB_FAST_PREDLD ? idx < ecnt : true), gl, gbase, acc, rsq);
                *reinterpret_cast<u64*>(accbuf + idx * FAST_ACC_STRIDE + 2 * gl) = acc;
                if (Q::kRem > 0) accbuf[idx * FAST_ACC_STRIDE + 8 + gl] = rsq;
            }
#endif
            __syncwarp();
            // lane i: candidate i.  acc[0] takes the remainder squares in order (quant.rs:31-35), then
            // acc.iter().sum() left to right (quant.rs:36) and the square root
            const float4 a03 = *reinterpret_cast<const float4*>(accbuf + lane * FAST_ACC_STRIDE);
            const float4 a47 = *reinterpret_cast<const float4*>(accbuf + lane * FAST_ACC_STRIDE + 4);
            float s = a03.x;
            if (Q::kRem > 0) {
                const float4 rq = *reinterpret_cast<const float4*>(accbuf + lane * FAST_ACC_STRIDE + 8);
                constexpr int R = Q::kRem;
                if (R > 0) s = __fadd_rn(s, rq.x);
                if (R > 1) s = __fadd_rn(s, rq.y);
                if (R > 2) s = __fadd_rn(s, rq.z);
                if (R > 3) s = __fadd_rn(s, rq.w);
            }
            s = __fadd_rn(s, a03.y);
            s = __fadd_rn(s, a03.z);
            s = __fadd_rn(s, a03.w);
            s = __fadd_rn(s, a47.x);
            s = __fadd_rn(s, a47.y);
            s = __fadd_rn(s, a47.z);
            s = __fadd_rn(s, a47.w);
            // lanes without a candidate hold stale scratch bytes: give them a value on the square root's fast path, or the
            // whole warp waits for their slow path (denormal / NaN inputs) in four batches out of five
            const float d = __fsqrt_rn(lane < ecnt ? s : 1.0f);
            const u64 key = make_rkey(d, cbuf[lane]);
            // admission (searcher.rs:74-94): key < list[ef-1] covers |selected| < ef and strict <.  `worst` is the
            // batch's starting value: a key admitted against it may still fall off the end in the merge, exactly as a
            // later, nearer key would have evicted it one by one.
            const bool want = lane < ecnt && key < *worst_p;
            const unsigned am = __ballot_sync(HB_FULL, want);
#if HB_FAST_ROWPF
            // an admitted key may be expanded a few hops from now: request its adjacency row (layer 0) today
            if (want && layer == 0) prefetch_l2(g.adj0 + (size_t)cbuf[lane] * g.S0);
#endif
            if (am) {
#if HB_FAST_MERGE
                merge_ranked<KPL>(L, key, want, am, obuf, kbuf, sbuf, len, ef_l, lane, worst_p);
#else
                __syncwarp();  // every lane has read its accumulators: the merge scratch may overwrite them
                const int kcnt = __popc(am);
                if (want) kbuf[__popc(am & lt)] = key;
                __syncwarp();
                L.merge(kbuf, kcnt, obuf, len, ef_l, lane);
                if (lane == 0) *worst_p = len == ef_l ? obuf[ef_l - 1] : RSENT;
                __syncwarp();
#endif
            }
        }
    next_batch:
        // ---- next batch ----
        const uint32_t S = layer ? g.SU : g.S0;
        if (row != EMPTY_ID) {  // more of the current adjacency row (rows wider than 32, continuation rows)
            b0 += 32;
            if (b0 >= S) { row = next; next = EMPTY_ID; b0 = 0; }
        }
        seed = false;
        bool have_nb = false;
        if (row == EMPTY_ID) {
            uint32_t cid;
            if (!L.pop(cid, lane)) {
                // this layer is finished: selected survives as the entry set of the next one
                L.clear_flags();
                if (layer == 0) {
                    if (vis.spill[FAST_SPILL_SLICE] != 0u) spill_release(vis.spill, pool, lane);
                    break;
                }
                --layer;
                {
                    const int efn = layer ? 1 : ef;
                    const u64 w = len == efn ? L.get(efn - 1) : RSENT;
                    if (lane == 0) *worst_p = w;
                }
                if (vis.spill[FAST_SPILL_SLICE] != 0u) spill_release(vis.spill, pool, lane);
                vis.clear(lane);
                // seed batch: visited <- ids(selected); the upper layers ran with ef = 1, so the entry set is the
                // single key at position 0
                nb = (lane == 0 && L.v[0] != RSENT) ? rkey_id(L.v[0]) : EMPTY_ID;
                seed = true;
                continue;
            }
            if (STATS) cnt.hops++;
            // layer.neighbors_vec(cid)  (graph/src/graph.rs:103-113) as fixed-stride rows
            row = layer ? __ldg(g.upper_off + cid) + (layer - 1) : cid;
#if HB_FAST_SPEC
            if (layer == 0) {
                if (cid == spec_id) {  // the guess of the previous hop was right: its row is already here
                    nb = spec_nb;
                    have_nb = true;
                    fresh_row = HB_FAST_SPEC != 1;  // (mode 1 requested the records of this row a hop ago)
                }
                // guess the next expansion: the best entry that is still unexpanded.  The reference pops exactly that one
                // next unless this batch admits a nearer key (searcher.rs:36-44); a wrong guess costs one row load.
                uint32_t c2;
                spec_id = EMPTY_ID;
                if (list_peek<KPL>(L, c2)) {
                    spec_id = c2;
                    spec_nb = ((uint32_t)lane < g.S0) ? __ldg(g.adj0 + (size_t)c2 * g.S0 + lane) : EMPTY_ID;
                    spec_pending = HB_FAST_SPEC == 1;
                }
            }
#endif
        }
        if (!have_nb) {
            const uint32_t* adj = layer ? g.upper_adj : g.adj0;
            const uint32_t* rp = adj + (size_t)row * S;
            const uint32_t i = b0 + lane;
            nb = (i < S) ? __ldg(rp + i) : EMPTY_ID;
            fresh_row = true;
        }
        {
            const bool ok = !(nb & CHAIN_BIT);
            const unsigned mk = __ballot_sync(HB_FULL, !ok && nb != EMPTY_ID);
            if (mk) next = __shfl_sync(HB_FULL, nb, __ffs(mk) - 1) & ~CHAIN_BIT;
            if (STATS) cnt.nbrs += __popc(__ballot_sync(HB_FULL, ok));
        }
    }
}

}  // namespace hb
