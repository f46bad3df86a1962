// Fast path of the query search: the result list lives in REGISTERS (KPL keys per lane,
// lane-major, sorted), the visited set is a 16-bit-entry exact hash table.  Same
// semantics as search.cuh (one expansion per iteration, (dist,id) keys, counters),
// roughly half the instructions per hop and half the shared memory per query, which
// doubles the number of queries resident per SM.
//
// List: position p lives in lane p / KPL, register p % KPL.  Unused positions hold the
// sentinel ~0 (whose "expanded" bit is set, so it is never picked and never beaten).
// selected.len() < ef is therefore the same test as key < list[ef-1].
//
// Visited set (ids < 2^B, table of T = 2^t 16-bit entries): h = (id * odd) mod 2^B is a
// bijection on B-bit ids; home = top t bits of h, rem = low B-t bits.  An entry stores
// (rem << 4 | displacement) with displacement <= 14, so (slot, entry) determines the id
// exactly: no false positives, no false negatives.  Needs B - t <= 12.
#pragma once
#include "search.cuh"

namespace hb {

constexpr u64 SENTINEL = ~0ull;

struct Vis16 {
    uint32_t* words;   // T/2 32-bit words holding two entries each
    uint32_t tbits;    // log2(T)
    uint32_t bbits;    // B
    uint32_t bmask;    // 2^B - 1
};

__device__ __forceinline__ void vis16_clear(const Vis16& v, int lane) {
    uint4* p = reinterpret_cast<uint4*>(v.words);
    const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    const uint32_t n16 = (1u << v.tbits) / 8;  // 8 entries per uint4
    for (uint32_t i = lane; i < n16; i += 32) p[i] = e;
    __syncwarp();
}

// true if id was not in the set (and is now recorded).  *ovf: probe window exhausted,
// the id is reported new but not recorded (caller falls back to a list scan).
__device__ __forceinline__ bool vis16_insert(const Vis16& v, uint32_t id, bool* ovf) {
    const uint32_t h = (id * 0x9E3779B1u) & v.bmask;
    const uint32_t rbits = v.bbits - v.tbits;
    const uint32_t home = h >> rbits;
    const uint32_t rem = h & ((1u << rbits) - 1u);
    const uint32_t tmask = (1u << v.tbits) - 1u;
#pragma unroll 1
    for (uint32_t d = 0; d < 15; ++d) {
        const uint32_t slot = (home + d) & tmask;
        const uint32_t mine = (rem << 4) | d;
        uint32_t* wp = v.words + (slot >> 1);
        const uint32_t sh = (slot & 1u) * 16u;
        uint32_t w = *reinterpret_cast<volatile uint32_t*>(wp);
        while (true) {
            uint32_t e = (w >> sh) & 0xFFFFu;
            if (e == mine) return false;
            if (e != 0xFFFFu) break;  // occupied by another id: next displacement
            uint32_t nw = (w & ~(0xFFFFu << sh)) | (mine << sh);
            uint32_t old = atomicCAS(wp, w, nw);
            if (old == w) return true;
            w = old;  // the word changed under us (neighbouring entry or this one): re-examine
        }
    }
    *ovf = true;
    return true;
}

template <int KPL>
struct RegList {
    u64 k[KPL];

    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int s = 0; s < KPL; ++s) k[s] = SENTINEL;
    }
    __device__ __forceinline__ u64 pick(int slot) const {
        u64 v = k[0];
#pragma unroll
        for (int s = 1; s < KPL; ++s) v = (slot == s) ? k[s] : v;
        return v;
    }
    // key at position p, broadcast to the warp
    __device__ __forceinline__ u64 at(int p) const {
        u64 v = pick(p % KPL);
        return __shfl_sync(HB_FULL, v, p / KPL);
    }
    // number of entries (masked) strictly smaller than key
    __device__ __forceinline__ int lower_bound(u64 key) const {
        int c = 0;
#pragma unroll
        for (int s = 0; s < KPL; ++s) c += ((k[s] & KEY_MASK) < key) ? 1 : 0;
        int full = __popc(__ballot_sync(HB_FULL, c == KPL));
        int cb = __shfl_sync(HB_FULL, c, full & 31);
        return full * KPL + (full < 32 ? cb : 0);
    }
    // insert at pos (pos < ef <= 32*KPL); position ef falls off
    __device__ __forceinline__ void insert(int pos, u64 key, int ef, int lane) {
        const int lp = pos / KPL, sp = pos % KPL;
        const u64 carry = __shfl_up_sync(HB_FULL, k[KPL - 1], 1);
        const bool after = lane > lp, here = lane == lp;
#pragma unroll
        for (int s = KPL - 1; s >= 1; --s) {
            bool mv = after || (here && s > sp);
            k[s] = mv ? k[s - 1] : k[s];
        }
        k[0] = after ? carry : k[0];
#pragma unroll
        for (int s = 0; s < KPL; ++s)
            if (here && s == sp) k[s] = key;
        if (ef < 32 * KPL) {
            const int le = ef / KPL, se = ef % KPL;
#pragma unroll
            for (int s = 0; s < KPL; ++s)
                if (lane == le && s == se) k[s] = SENTINEL;
        }
    }
    __device__ __forceinline__ int count() const {
        int c = 0;
#pragma unroll
        for (int s = 0; s < KPL; ++s) c += (k[s] != SENTINEL) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(HB_FULL, c, o);
        return c;
    }
};

struct WarpScratch16 {
    Vis16 vis;
    uint32_t* newbuf;  // [32]
    float* qd;
};

template <class Q, int KPL>
__device__ __forceinline__ void search_layer_reg(const Q& query, const uint8_t* __restrict__ rec,
                                                 uint32_t rec_stride, const GraphView& g, uint32_t layer,
                                                 const WarpScratch16& s, RegList<KPL>& L, int ef, int lane,
                                                 SearchCounters& cnt) {
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    vis16_clear(s.vis, lane);
    {   // visited <- ids(selected)   (results.rs:159-168)
        bool ovf = false;
#pragma unroll
        for (int t = 0; t < KPL; ++t)
            if (L.k[t] != SENTINEL) vis16_insert(s.vis, (uint32_t)L.k[t], &ovf);
        if (__any_sync(HB_FULL, ovf)) cnt.overflow = 1;
        __syncwarp();
    }
    u64 worst = L.at(ef - 1) & KEY_MASK;  // sentinel (max) while |selected| < ef
    while (true) {
        // candidates.pop_first(): first entry whose "expanded" bit is clear
        int myslot = -1;
#pragma unroll
        for (int t = KPL - 1; t >= 0; --t)
            if (!(L.k[t] & EXP_FLAG)) myslot = t;
        unsigned um = __ballot_sync(HB_FULL, myslot >= 0);
        if (!um) break;
        const int owner = __ffs(um) - 1;
        u64 ck = (myslot >= 0) ? L.pick(myslot) : 0ull;
        ck = __shfl_sync(HB_FULL, ck, owner);
        if (lane == owner) {
#pragma unroll
            for (int t = 0; t < KPL; ++t)
                if (t == myslot) L.k[t] |= EXP_FLAG;
        }
        const uint32_t cid = (uint32_t)ck;
        cnt.hops++;

        const uint32_t* base;
        uint32_t S, row;
        if (layer == 0) { base = g.adj0; S = g.S0; row = cid; }
        else { base = g.upper_adj; S = g.SU; row = __ldg(g.upper_off + cid) + (layer - 1); }
        while (row != EMPTY_ID) {
            const uint32_t* rp = base + (size_t)row * S;
            uint32_t next = EMPTY_ID;
            for (uint32_t b0 = 0; b0 < S; b0 += 32) {
                uint32_t i = b0 + lane;
                uint32_t nb = (i < S) ? __ldg(rp + i) : EMPTY_ID;
                bool marker = (nb != EMPTY_ID) && (nb & CHAIN_BIT);
                unsigned mk = __ballot_sync(HB_FULL, marker);
                if (mk) next = __shfl_sync(HB_FULL, nb, __ffs(mk) - 1) & ~CHAIN_BIT;
                bool valid = (nb != EMPTY_ID) && !marker;
                cnt.nbrs += __popc(__ballot_sync(HB_FULL, valid));
                bool ovf = false;
                bool isnew = valid && vis16_insert(s.vis, nb, &ovf);
                if (__any_sync(HB_FULL, ovf)) {
                    // rare: probe window exhausted.  Exactness is kept by testing list membership.
                    cnt.overflow = 1;
                    unsigned om = __ballot_sync(HB_FULL, ovf);
                    while (om) {
                        int src = __ffs(om) - 1;
                        om &= om - 1;
                        uint32_t id = __shfl_sync(HB_FULL, nb, src);
                        bool hit = false;
#pragma unroll
                        for (int t = 0; t < KPL; ++t)
                            hit |= (L.k[t] != SENTINEL) && ((uint32_t)(L.k[t] & ~EXP_FLAG) == id);
                        if (__any_sync(HB_FULL, hit) && lane == src) isnew = false;
                    }
                }
                unsigned nm = __ballot_sync(HB_FULL, isnew);
                int ncnt = __popc(nm);
                if (ncnt == 0) continue;
                cnt.evals += ncnt;
                if (isnew) s.newbuf[__popc(nm & ((1u << lane) - 1))] = nb;
                __syncwarp();
                for (int r0 = 0; r0 < ncnt; r0 += 8) {
                    int idx = r0 + grp;
                    bool act = idx < ncnt;
                    uint32_t cand = s.newbuf[act ? idx : 0];
                    float d = query.dist(rec + (size_t)cand * rec_stride, gl, gbase);
                    u64 key = make_key(d, cand);
                    // admission (searcher.rs:74-94): key < list[ef-1] covers both |selected| < ef and strict <
                    bool want = act && gl == 0 && key < worst;
                    unsigned am = __ballot_sync(HB_FULL, want);
                    while (am) {
                        int src = __ffs(am) - 1;
                        am &= am - 1;
                        u64 k = __shfl_sync(HB_FULL, key, src);
                        if (k < worst) {
                            int pos = L.lower_bound(k);
                            L.insert(pos, k, ef, lane);
                            worst = L.at(ef - 1) & KEY_MASK;
                        }
                    }
                }
                __syncwarp();
            }
            row = next;
        }
    }
    // clear_candidates: drop the expanded marks for the next layer (sentinels keep theirs)
#pragma unroll
    for (int t = 0; t < KPL; ++t)
        if (L.k[t] != SENTINEL) L.k[t] &= KEY_MASK;
}

}  // namespace hb
