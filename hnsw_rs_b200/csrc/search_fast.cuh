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
//    lost claim (two new ids of one batch, same bucket) goes to an out-of-line loop over the following buckets
//    (displacement <= 7 stored in the entry), and an id that finds 8 full buckets goes to a small exact spill list --
//    so membership, and with it the evaluation counter, stays exact (the round-1 table over-counted after an overflow).
//    (bucket, entry) names the id exactly: h = id * odd mod 2^B is a bijection on B-bit ids, bucket = floor(h * NB / 2^B),
//    and the entry keeps bits [9, B) of (h * NB) mod 2^B, which differ between two ids of one bucket because those
//    values are NB > 512 apart.
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

#ifndef HB_FAST_PREFETCH_ALL
#define HB_FAST_PREFETCH_ALL 1  // request the record of every neighbour before the visited test (see search_reg.cuh)
#endif

namespace hb {

constexpr int FAST_SPILL = 16;        // words of the shared-memory spill area: 14 ids + the two list lengths
constexpr int FAST_SPILL_IDS = FAST_SPILL - 2;
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
    const int e = __popc(w0 & 0x80008000u) + __popc(w1 & 0x80008000u);
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

// Out-of-line continuation for the lanes the single-step insert could not settle: lost claims retry the same bucket,
// full buckets move on (displacement + 1).  bit0 of the result: id is new (recorded now); bit1: 8 full buckets in a
// row, the caller consults the spill list.
__device__ __noinline__ uint32_t vis_slow(uint32_t sbase, uint32_t nb, uint32_t home, uint32_t mine0, bool pending,
                                          bool home_full) {
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
            if (d > 7u) ovf = true;
        }
        pending = pending && !found && !won && !ovf;
    }
    return (isnew ? 1u : 0u) | (ovf ? 2u : 0u);
}

struct VisB4 {
    uint32_t sbase;   // shared-space byte address of the table: nb buckets of 8 bytes
    uint32_t nb;      // buckets, 512 < nb <= 1024
    uint32_t mul;     // odd << (32 - B): h32 = id * mul holds the B-bit bijection value top-aligned
    uint32_t rsh;     // 38 - B: (h32 * nb) >> rsh, low three bits cleared = entry with displacement 0
    uint32_t* spill;  // ids that found 8 full buckets: [0, 14) ids, [14] their number, [15] number of ids in the global continuation

    __device__ __forceinline__ void clear(int lane) const {
        const uint32_t n16 = nb / 2;  // 16-byte chunks
        for (uint32_t i = lane; i < n16; i += 32)
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(sbase + i * 16u), "r"(0xFFFFFFFFu) : "memory");
        if (lane == 0) { spill[FAST_SPILL_IDS] = 0u; spill[FAST_SPILL_IDS + 1] = 0u; }
        __syncwarp();
    }
    // results.insert_visited for all 32 lanes (want: this lane holds an id).  Returns "id was not yet visited" per
    // lane; ovf: the table could not decide for this lane (see vis_slow), isnew is then false and the caller decides.
    __device__ __forceinline__ bool insert_warp(uint32_t id, bool want, bool& ovf) const {
        const uint32_t h32 = id * mul;
        const uint32_t home = __umulhi(h32, nb);
        const uint32_t mine0 = ((h32 * nb) >> rsh) & 0x7FF8u;
        bool won, full;
        const bool found = vis_probe(sbase + home * 8u, mine0, want, won, full);
        const bool pending = want && !found && !won;
        ovf = false;
        if (__any_sync(HB_FULL, pending)) {
            const uint32_t r = vis_slow(sbase, nb, home, mine0, pending, full);
            won = won || (r & 1u);
            ovf = (r & 2u) != 0u;
        }
        return won;
    }
};

// Out-of-line: the lanes in `om` hold ids that found 8 full buckets in a row.  The exact spill list decides whether
// such an id was seen before (shared memory first, then this warp's slice of a global workspace); when both are
// full, list membership keeps the answers exact and only the evaluation counter may over-count (*flags bit 1).
// Returns bit0: this lane's "id is new"; bits 1-2: the flag bits to raise.
template <int KPL>
__device__ __noinline__ uint32_t vis_spill(uint32_t* spill, uint32_t* gspill, uint32_t gcap, unsigned om, uint32_t nb,
                                           bool isnew, const RegList<KPL> L, int lane) {
    uint32_t fl = 4u;
    while (om) {
        const int src = __ffs(om) - 1;
        om &= om - 1;
        const uint32_t id = __shfl_sync(HB_FULL, nb, src);
        uint32_t n1 = spill[FAST_SPILL_IDS], n2 = spill[FAST_SPILL_IDS + 1];
        bool hit = (uint32_t)lane < n1 && spill[lane] == id;
        for (uint32_t base = 0; base < n2; base += 32)
            hit = hit || (base + lane < n2 && *reinterpret_cast<volatile uint32_t*>(gspill + base + lane) == id);
        bool fresh;
        if (__any_sync(HB_FULL, hit)) fresh = false;
        else if (n1 < (uint32_t)FAST_SPILL_IDS) {
            if (lane == 0) { spill[n1] = id; spill[FAST_SPILL_IDS] = n1 + 1u; }
            fresh = true;
        } else if (n2 < gcap) {
            if (lane == 0) {
                *reinterpret_cast<volatile uint32_t*>(gspill + n2) = id;
                spill[FAST_SPILL_IDS + 1] = n2 + 1u;
            }
            fresh = true;
        } else {
            fl |= 2u;
            fresh = !L.holds_id(id);
        }
        __syncwarp();
        if (lane == src) isnew = fresh;
    }
    return fl | (isnew ? 1u : 0u);
}

// Query of a compile-time dimension (8*NCH + REM, REM <= 4) in registers; evaluates one record per group of 4 lanes
// and leaves the accumulators to the caller (no cross-lane sum here).
template <int NCH, int REM>
struct FastQuery {
    static_assert(REM <= 4, "one remainder element per lane of a group");
    using RQ = RegQuery<NCH, REM>;
    static constexpr int W = RQ::W;
    static constexpr int TAIL = RQ::TAIL;
    using Rec = typename RQ::Rec;
    static constexpr int kRem = REM;
    // remainder bytes: tail form -> bytes 8..11 of the tail word; compact form -> lane 0's slice positions 2*NCH + r
    static constexpr int RP0 = TAIL ? 8 : 2 * NCH;             // slice position (or tail byte) of remainder element 0
    static constexpr bool STRADDLE = REM > 0 && (RP0 % 4) + REM > 4;  // the REM bytes span two 32-bit words
    u64 q[NCH ? NCH : 1];
    float qrem;       // query value of remainder element gl (lanes gl >= REM: 0)
    u64 nz;

    __device__ __forceinline__ void init(const RecLayout&, const float* qd, int gl) {
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            float2 v = *reinterpret_cast<const float2*>(qd + 8 * k + 2 * gl);
            q[k] = pk(v.x, v.y);
        }
        qrem = (REM > 0 && gl < REM) ? qd[8 * NCH + gl] : 0.0f;
        nz = hb_negzero2;
    }
    __device__ __forceinline__ static Rec load(const uint8_t* __restrict__ rec, int gl) { return RQ::load(rec, gl); }

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
                acc = chunk_acc(acc, word32(w[j], c / 2), (c & 1) * 2, dl, mn2, q[k]);
            }
            if (REM > 0) {
                const float fm = __uint_as_float(__byte_perm(rw, 0x4B000000u, remsel));
                rsq = rem_acc(0.0f, fm, dl, mn, qrem);  // 0 + t*t == t*t
            }
        } else {
            const u64 dl2 = pk(dl, dl);
            const float nmd = __fmul_rn(-8388608.0f, dl);
            const u64 nmd2 = pk(nmd, nmd);
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const u64 m2 = pk(RQ::slice_magic(w, 2 * k), RQ::slice_magic(w, 2 * k + 1));
                acc = add2(acc, chunk_sq(m2, dl2, nmd2, mn2, q[k], nz));
            }
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
                   FAST_OFF_TABLE = FAST_OFF_SCRATCH + FAST_SCRATCH_BYTES;

// One whole query.  On exit L holds the <= ef nearest evaluated nodes of layer 0, sorted.
// gspill / gcap: this warp's slice of the global continuation of the spill list (may be null / 0).
template <class Q, int KPL, bool STATS>
__device__ __forceinline__ void search_query_fast(const Q& query, const uint8_t* __restrict__ rec, uint32_t rec_stride,
                                                  const GraphView& g, uint32_t n_layers, uint32_t ep, RegList<KPL>& L,
                                                  const VisB4& vis, unsigned char* wsm, uint32_t* gspill_ws, uint32_t gcap,
                                                  int ef, int lane, SearchCounters& cnt) {
    uint32_t* newbuf = reinterpret_cast<uint32_t*>(wsm);
    float* scratch = reinterpret_cast<float*>(wsm + FAST_OFF_SCRATCH);
    // key at position ef_l - 1 of the list (the sentinel, i.e. the maximum, while |selected| < ef_l): kept in shared
    // memory, it is read once per batch
    u64* worst_p = reinterpret_cast<u64*>(wsm + FAST_OFF_WORST);
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    const unsigned lt = (1u << lane) - 1u;
    float* accbuf = scratch;
    u64* kbuf = reinterpret_cast<u64*>(scratch);
    u64* mbuf = kbuf + 32;
    uint32_t layer = n_layers - 1;
    int ef_l = layer ? 1 : ef;
    int len = 0;  // |selected|
    L.reset();
    if (lane == 0) *worst_p = RSENT;
    vis.clear(lane);
    // boot batch: the entry point (selected <- {Dist(ep, d)}, template.rs:316-319)
    uint32_t nb = lane == 0 ? ep : EMPTY_ID;
    bool seed = false;
    uint32_t row = EMPTY_ID, next = EMPTY_ID, b0 = 0;
    uint32_t S = layer ? g.SU : g.S0;
    const uint32_t* adj = layer ? g.upper_adj : g.adj0;
#pragma unroll 1
    while (true) {
        // ---- one batch of up to 32 ids: results.insert_visited(node) (results.rs:101-103) ----
        const bool valid = !(nb & CHAIN_BIT);  // EMPTY_ID and chain markers carry bit 31
#if HB_FAST_PREFETCH_ALL
        if (valid) prefetch_record(rec + (size_t)nb * rec_stride, rec_stride);
#endif
        bool ovf;
        bool isnew = vis.insert_warp(nb, valid, ovf);
        if (__any_sync(HB_FULL, ovf)) {  // rare: 8 full buckets in a row
            uint32_t* gsp = gspill_ws ? gspill_ws + ((size_t)blockIdx.x * SEARCH_WPB + (threadIdx.x >> 5)) * gcap : nullptr;
            const uint32_t r = vis_spill<KPL>(vis.spill, gsp, gsp ? gcap : 0u, __ballot_sync(HB_FULL, ovf), nb, isnew, L, lane);
            isnew = (r & 1u) != 0u;
            if (STATS) cnt.overflow |= r & 6u;
        }
        isnew = isnew && !seed;
        const unsigned nm = __ballot_sync(HB_FULL, isnew);
        const int ncnt = __popc(nm);
        if (ncnt) {
            if (STATS) cnt.evals += ncnt;
            if (isnew) {
                const int my = __popc(nm & lt);
                newbuf[my] = nb;
#if !HB_FAST_PREFETCH_ALL
                prefetch_record(rec + (size_t)nb * rec_stride, rec_stride);
#endif
            }
            __syncwarp();
#pragma unroll 1
            for (int r0 = 0; r0 < ncnt; r0 += 8) {
                const int idx = r0 + grp;
                const uint32_t cand = newbuf[idx < ncnt ? idx : 0];
                // index.get_point(node).dist2other(point)  (searcher.rs:66-69), the sum left to the lanes below
                u64 acc;
                float rsq;
                query.partial(Q::load(rec + (size_t)cand * rec_stride, gl), gl, gbase, acc, rsq);
                *reinterpret_cast<u64*>(accbuf + idx * FAST_ACC_STRIDE + 2 * gl) = acc;
                if (Q::kRem > 0) accbuf[idx * FAST_ACC_STRIDE + 8 + gl] = rsq;
            }
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
            const float d = __fsqrt_rn(s);
            const u64 key = make_rkey(d, newbuf[lane]);
            // admission (searcher.rs:74-94): key < list[ef-1] covers |selected| < ef and strict <.  `worst` is the
            // batch's starting value: a key admitted against it may still fall off the end in the merge, exactly as a
            // later, nearer key would have evicted it one by one.
            const bool want = lane < ncnt && key < *worst_p;
            const unsigned am = __ballot_sync(HB_FULL, want);
            if (am) {
                const int kcnt = __popc(am);
                __syncwarp();  // every lane has read its accumulators: the key buffer may overwrite them
                if (want) kbuf[__popc(am & lt)] = key;
                __syncwarp();
                L.merge(kbuf, kcnt, mbuf, len, ef_l, lane);
                if (lane == 0) *worst_p = len == ef_l ? mbuf[ef_l - 1] : RSENT;
                __syncwarp();
            }
        }
        // ---- next batch ----
        if (row != EMPTY_ID) {  // more of the current adjacency row (rows wider than 32, continuation rows)
            b0 += 32;
            if (b0 >= S) { row = next; next = EMPTY_ID; b0 = 0; }
        }
        seed = false;
        if (row == EMPTY_ID) {
            uint32_t cid;
            if (!L.pop(cid, lane)) {
                // this layer is finished: selected survives as the entry set of the next one
                L.clear_flags();
                if (layer == 0) break;
                --layer;
                S = layer ? g.SU : g.S0;
                adj = layer ? g.upper_adj : g.adj0;
                ef_l = layer ? 1 : ef;
                {
                    const u64 w = len == ef_l ? L.get(ef_l - 1) : RSENT;
                    if (lane == 0) *worst_p = w;
                }
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
        }
        {
            const uint32_t* rp = adj + (size_t)row * S;
            const uint32_t i = b0 + lane;
            nb = (i < S) ? __ldg(rp + i) : EMPTY_ID;
            const bool ok = !(nb & CHAIN_BIT);
            const unsigned mk = __ballot_sync(HB_FULL, !ok && nb != EMPTY_ID);
            if (mk) next = __shfl_sync(HB_FULL, nb, __ffs(mk) - 1) & ~CHAIN_BIT;
            if (STATS) cnt.nbrs += __popc(__ballot_sync(HB_FULL, ok));
        }
    }
}

}  // namespace hb
