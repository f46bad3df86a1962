// HNSW::ann_by_vector on a register-resident result list (ef <= 32*KPL), one warp per query.
//
// Same algorithm and the same results, counters included, as search_layer in search.cuh
// (Searcher::search_layer, hnsw/src/template/searcher.rs:23-103, driven by HNSW::ann_by_vector,
// hnsw/src/template.rs:306-335).  What differs is the shape of the code: the kernel is bound by
// instruction issue and by the instruction cache (24 warps per SM, each at its own program
// counter), so the whole query -- entry point, greedy upper layers, layer 0 -- runs through ONE
// loop with one instance each of the visited-set insert, the distance evaluation and the list
// insertion.  One iteration handles one batch of up to 32 neighbour ids:
//   boot batch   the entry point alone (selected <- {Dist(ep, d)}, template.rs:316-319)
//   seed batch   the surviving entry of the layer above, recorded as visited but not evaluated
//                (results.rs:148-168 at the start of every search_layer)
//   row batch    32 slots of the adjacency row of the popped candidate
//
// Key layout: (f32 bits of dist << 32) | (id << 1) | expanded.  ids are < 2^31, so id << 1 fits;
// keys of different ids compare like Dist::cmp (graph/src/dist.rs:30-37) whatever their flag bits,
// and a new key (flag clear) is never < an entry with the same (dist, id).  Position p of the
// sorted list lives in lane p / KPL, slot p % KPL.  The keys admitted while one batch of neighbours is
// evaluated are merged in one step (RegList::merge): every key computes its final position by counting.
#pragma once
#include "search.cuh"

#ifndef HB_PREFETCH_ALL
#define HB_PREFETCH_ALL 1  // 0: only the records of rounds >= 2 are prefetched, after the visited test
#endif


namespace hb {

constexpr u64 RSENT = ~0ull;

__device__ __forceinline__ u64 make_rkey(float d, uint32_t id) {
    return ((u64)__float_as_uint(d) << 32) | (u64)(id << 1);
}
__device__ __forceinline__ uint32_t rkey_id(u64 k) { return (uint32_t)k >> 1; }
__device__ __forceinline__ u64 sel64(bool c, u64 a, u64 b) { return c ? a : b; }

template <int KPL>
struct RegList {
    u64 v[KPL];

    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int s = 0; s < KPL; ++s) v[s] = RSENT;
    }
    // key at position p (warp-uniform), broadcast to all lanes
    __device__ __forceinline__ u64 get(int p) const {
        const int s = p % KPL;
        u64 x = v[0];
#pragma unroll
        for (int t = 1; t < KPL; ++t) x = sel64(s == t, v[t], x);
        return __shfl_sync(HB_FULL, x, p / KPL);
    }
    // Merge the m (<= 32) new keys kbuf[0..m) -- all different from each other and from the list's keys -- into
    // the sorted list and keep the ef smallest: selected <- ef smallest of (selected U new), which is what the
    // reference's one-by-one admission (searcher.rs:74-94) ends with, whatever the order.  Every key learns its
    // final position by counting: a list key moves up by the number of new keys below it, a new key lands at
    // (list keys below it) + (new keys below it); the permutation itself goes through mbuf (32*KPL slots of
    // shared memory).  len = number of real keys, updated.
    __device__ __forceinline__ void merge(const u64* kbuf, int m, u64* mbuf, int& len, int ef, int lane) {
        const u64 nk = lane < m ? kbuf[lane] : RSENT;
        // shift[s]: number of new keys below this lane's list key s; own: number of new keys below this lane's own
        // new key; myr: number of list keys below this lane's own new key
        int shift[KPL];
#pragma unroll
        for (int s = 0; s < KPL; ++s) shift[s] = 0;
        int own = 0, myr = 0;
#pragma unroll 1
        for (int j = 0; j < m; ++j) {
            const u64 kj = kbuf[j];  // broadcast read
            int r = 0;
#pragma unroll
            for (int s = 0; s < KPL; ++s) {
                // lt = kj < v[s]: new key below this list key (keys are distinct; true for sentinels);
                // shift[s] += lt; r += number of lanes with !lt.  One predicate feeds the add and the vote
                // (written out: the compiler evaluated the comparison twice and copied the counters around).
                uint32_t bal;
                asm volatile("{\n .reg .pred p;\n setp.lt.u64 p, %2, %3;\n @p add.s32 %0, %0, 1;\n"
                             " vote.sync.ballot.b32 %1, !p, 0xffffffff;\n}"
                             : "+r"(shift[s]), "=r"(bal) : "l"(kj), "l"(v[s]));
                r += __popc(bal);
            }
            asm("{\n .reg .pred p;\n setp.lt.u64 p, %1, %2;\n @p add.s32 %0, %0, 1;\n}" : "+r"(own) : "l"(kj), "l"(nk));
            myr = lane == j ? r : myr;
        }
#pragma unroll
        for (int s = 0; s < KPL; ++s) {
            const int p = lane * KPL + s + shift[s];
            if (v[s] != RSENT && p < ef) mbuf[p] = v[s];
        }
        if (lane < m) {
            const int p = myr + own;
            if (p < ef) mbuf[p] = nk;
        }
        __syncwarp();
        len = min(ef, len + m);
#pragma unroll
        for (int s = 0; s < KPL; ++s) {
            const int p = lane * KPL + s;
            v[s] = p < len ? mbuf[p] : RSENT;
        }
        __syncwarp();
    }
    // candidates.pop_first(): the first entry whose "expanded" bit is clear; marks it expanded and returns its id.
    // Returns false when there is none.  Only the low words are looked at: the flag and the id live there, and the
    // sentinel's low bit is set, so an empty slot reads as expanded.
    __device__ __forceinline__ bool pop(uint32_t& cid, int lane) {
        uint32_t lo[KPL], hi[KPL];
#pragma unroll
        for (int s = 0; s < KPL; ++s) asm("mov.b64 {%0,%1}, %2;" : "=r"(lo[s]), "=r"(hi[s]) : "l"(v[s]));
        uint32_t pick = 0xFFFFFFFFu;  // low bit set: nothing to expand in this lane
#pragma unroll
        for (int s = KPL - 1; s >= 0; --s) pick = (lo[s] & 1u) ? pick : lo[s];
        const unsigned m = __ballot_sync(HB_FULL, !(pick & 1u));
        if (!m) return false;
        const uint32_t got = __shfl_sync(HB_FULL, pick, __ffs(m) - 1);
        cid = got >> 1;
        // ids are unique in the list, so the low word `got` names its entry: set the "expanded" bit there
#pragma unroll
        for (int s = 0; s < KPL; ++s) {
            lo[s] |= (lo[s] == got) ? 1u : 0u;
            asm("mov.b64 %0, {%1,%2};" : "=l"(v[s]) : "r"(lo[s]), "r"(hi[s]));
        }
        return true;
    }
    __device__ __forceinline__ void clear_flags() {  // clear_candidates (searcher.rs:100)
#pragma unroll
        for (int s = 0; s < KPL; ++s) v[s] = sel64(v[s] != RSENT, v[s] & ~1ull, v[s]);
    }
    __device__ __forceinline__ bool holds_id(uint32_t id) const {
        bool hit = false;
#pragma unroll
        for (int s = 0; s < KPL; ++s) hit |= (v[s] != RSENT) && (rkey_id(v[s]) == id);
        return __any_sync(HB_FULL, hit);
    }
};

// One whole query.  On exit L holds the <= ef nearest evaluated nodes of layer 0, sorted.
// Layers n_layers-1 .. 1 are searched with ef = 1 (template.rs:322-324), layer 0 with ef (:326).
template <class Q, class VIS, int KPL, bool STATS>
__device__ __forceinline__ void search_query_reg(const Q& query, const uint8_t* __restrict__ rec,
                                                 uint32_t rec_stride, const GraphView& g, uint32_t n_layers,
                                                 uint32_t ep, RegList<KPL>& L, const VIS& vis,
                                                 uint32_t* newbuf, u64* kbuf, u64* mbuf, int ef, int lane,
                                                 SearchCounters& cnt) {
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    uint32_t layer = n_layers - 1;
    int ef_l = layer ? 1 : ef;
    int len = 0;  // |selected|
    L.reset();
    u64 worst = RSENT;  // key at position ef_l - 1: the sentinel (max) while |selected| < ef_l
    vis.clear(lane);
    // boot batch: the entry point
    uint32_t nb = lane == 0 ? ep : EMPTY_ID;
    bool seed = false;
    uint32_t row = EMPTY_ID, next = EMPTY_ID, b0 = 0;
    // adjacency geometry of the current layer
    uint32_t S = layer ? g.SU : g.S0;
    const uint32_t* adj = layer ? g.upper_adj : g.adj0;
#pragma unroll 1
    while (true) {
        // ---- one batch of up to 32 ids: results.insert_visited(node) (results.rs:101-103) ----
        const bool valid = !(nb & CHAIN_BIT);  // EMPTY_ID and chain markers carry bit 31
#if HB_PREFETCH_ALL
        // request the record of every neighbour before the visited test: the memory latency overlaps the hash probing
        // (about half of the neighbours turn out to be visited already: DRAM has the headroom, the issue slots do not).
        // Not for f32 records: at four lines per record the wasted half costs more DRAM time than the overlap saves.
        if (Q::kPrefetchBeforeVisited && valid) {
            const uint8_t* rp8 = rec + (size_t)nb * rec_stride;
            prefetch_record(rp8, rec_stride);
        }
#endif
        bool ovf = false;
        bool isnew = vis.insert_warp(nb, valid, &ovf);
        if (__any_sync(HB_FULL, ovf)) {
            // rare: probe window exhausted.  Exactness is kept by testing list membership.
            if (STATS) cnt.overflow = 1;
            unsigned om = __ballot_sync(HB_FULL, ovf);
#pragma unroll 1
            while (om) {
                const int src = __ffs(om) - 1;
                om &= om - 1;
                const uint32_t id = __shfl_sync(HB_FULL, nb, src);
                if (L.holds_id(id) && lane == src) isnew = false;
            }
        }
        isnew = isnew && !seed;
        const unsigned nm = __ballot_sync(HB_FULL, isnew);
        const int ncnt = __popc(nm);
        if (ncnt) {
            if (STATS) cnt.evals += ncnt;
            if (isnew) {
                const int my = __popc(nm & ((1u << lane) - 1));
                newbuf[my] = nb;
                // records of the second and later rounds are requested now, so that those rounds
                // do not pay a second memory latency
                if ((!HB_PREFETCH_ALL && my >= 8) || !Q::kPrefetchBeforeVisited) {
                    const uint8_t* rp8 = rec + (size_t)nb * rec_stride;
                    prefetch_record(rp8, rec_stride);
                }
            }
            __syncwarp();
            int kcnt = 0;
#pragma unroll 1
            for (int r0 = 0; r0 < ncnt; r0 += 8) {
                const int idx = r0 + grp;
                const bool act = idx < ncnt;
                const uint32_t cand = newbuf[act ? idx : 0];
                // index.get_point(node).dist2other(point)  (searcher.rs:66-69)
                const float d = query.dist(rec + (size_t)cand * rec_stride, gl, gbase);
                const u64 key = make_rkey(d, cand);
                // admission (searcher.rs:74-94): key < list[ef-1] covers |selected| < ef and strict <.
                // `worst` is the batch's starting value: a key admitted against it may still fall off the end
                // in the merge, exactly as a later, nearer key would have evicted it one by one.
                const bool want = act && gl == 0 && key < worst;
                const unsigned am = __ballot_sync(HB_FULL, want);
                if (want) kbuf[kcnt + __popc(am & ((1u << lane) - 1))] = key;
                kcnt += __popc(am);
            }
            __syncwarp();
            if (kcnt) {
                L.merge(kbuf, kcnt, mbuf, len, ef_l, lane);
                worst = len == ef_l ? mbuf[ef_l - 1] : RSENT;
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
                worst = len == ef_l ? L.get(ef_l - 1) : RSENT;
                vis.clear(lane);
                // seed batch: visited <- ids(selected); the upper layers ran with ef = 1, so the
                // entry set is the single key at position 0
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
