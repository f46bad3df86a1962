// Warp-resident HNSW layer search (device side), shared by the query kernel and
// the build kernel.
//
// Restates Searcher::search_layer (hnsw/src/template/searcher.rs:23-103) over the
// Results sets (hnsw/src/template/results.rs:26-33) with one 32-lane CTA-slice
// (one warp) per query:
//   selected    -> sorted array of u64 keys in shared memory, key =
//                  (f32 bits of dist << 32) | id.  Non-negative f32 bit patterns
//                  order like unsigned ints, so integer compare == Dist::cmp
//                  (graph/src/dist.rs:30-37: dist, then id).
//   candidates  -> the not-yet-expanded members of `selected` (bit 31 of the id
//                  half marks "expanded").  A candidate that has fallen out of
//                  `selected` is > the worst selected, so popping it is exactly
//                  the reference's break (searcher.rs:41-44); hence "no
//                  unexpanded entry left" == the reference's loop exit.
//   visited     -> open-addressing hash set of ids in shared memory.
// Exactly one candidate is expanded per iteration; all unvisited
// neighbours of that candidate are evaluated in parallel, 8 per round, 4 lanes
// each, which is result- and counter-identical to the reference (App. C-5).
#pragma once
#include "dist.cuh"

namespace hb {

constexpr uint32_t EMPTY_ID = 0xFFFFFFFFu;
constexpr uint32_t CHAIN_BIT = 0x80000000u;  // adjacency slot: continuation row marker
constexpr u64 EXP_FLAG = 0x80000000ull;      // list key: "expanded"
constexpr u64 KEY_MASK = ~EXP_FLAG;

struct GraphView {
    const uint32_t* adj0;       // layer 0: row r = node id (r < n_points) or chain row
    uint32_t S0;                // slots per layer-0 row
    const uint32_t* upper_off;  // [n_points] first upper row of the node, EMPTY_ID if level 0
    const uint32_t* upper_adj;  // rows for layers >= 1: row = upper_off[node] + (layer - 1)
    uint32_t SU;                // slots per upper row
};

struct WarpScratch {
    u64* list;          // [ef_cap]
    uint32_t* vis;      // [vis_slots]
    uint32_t* newbuf;   // [32]
    float* qd;          // [dim padded]
    uint32_t vis_slots; // power of two >= 64
};

struct SearchCounters {
    uint32_t hops;
    uint32_t evals;
    uint32_t overflow;  // visited set could not record an id (results stay exact)
    uint32_t nbrs;      // neighbour ids read: sum of degrees of the expanded nodes
};

__device__ __forceinline__ u64 make_key(float d, uint32_t id) {
    return ((u64)__float_as_uint(d) << 32) | (u64)id;
}

// number of list entries (masked) strictly smaller than key; list sorted ascending
__device__ __forceinline__ int list_lower_bound(const u64* list, int n, u64 key, int lane) {
    int lo = 0, len = n;
    while (len > 0) {
        int step = (len + 31) >> 5;
        int idx = lo + lane * step;
        bool less = (idx < lo + len) && ((list[idx] & KEY_MASK) < key);
        int c = __popc(__ballot_sync(HB_FULL, less));
        int nlo = c > 0 ? lo + (c - 1) * step + 1 : lo;
        int nhi = min(lo + c * step, lo + len);
        lo = nlo;
        len = nhi - nlo;
    }
    return lo;
}

// insert key at pos, keeping at most ef entries (the last one is dropped when full)
__device__ __forceinline__ void list_insert_at(u64* list, int& n, int ef, int pos, u64 key, int lane) {
    int last_src = (n < ef ? n : ef - 1) - 1;
    for (int top = last_src; top >= pos; top -= 32) {
        int j = top - lane;
        bool mv = j >= pos;
        u64 v = 0;
        if (mv) v = list[j];
        __syncwarp();
        if (mv) list[j + 1] = v;
        __syncwarp();
    }
    if (lane == 0) list[pos] = key;
    if (n < ef) ++n;
    __syncwarp();
}

__device__ __forceinline__ uint32_t vis_slot(uint32_t id, uint32_t shift) {
    return (id * 0x9E3779B1u) >> shift;
}

// returns true if id was not yet in the set (and records it).  On a full probe
// window the id is reported new but NOT recorded and *ovf is raised; the caller
// then falls back to a list-membership test so results stay exact.
__device__ __forceinline__ bool vis_insert(uint32_t* tab, uint32_t mask, uint32_t shift, uint32_t id,
                                           bool* ovf) {
    uint32_t h = vis_slot(id, shift);
    for (int probe = 0; probe < 48; ++probe) {
        uint32_t old = atomicCAS(&tab[h], EMPTY_ID, id);
        if (old == EMPTY_ID) return true;
        if (old == id) return false;
        h = (h + 1) & mask;
    }
    *ovf = true;
    return true;
}

__device__ __forceinline__ void vis_clear(uint32_t* tab, uint32_t slots, int lane) {
    uint4* p = reinterpret_cast<uint4*>(tab);
    const uint4 e = make_uint4(EMPTY_ID, EMPTY_ID, EMPTY_ID, EMPTY_ID);
    for (uint32_t i = lane; i < slots / 4; i += 32) p[i] = e;
    __syncwarp();
}

// One layer of best-first search.  On entry list[0..n) holds the entry set
// (sorted, flags clear); on exit it holds the <= ef nearest evaluated nodes
// (sorted, flags clear).  All 32 lanes execute this together.
template <class Q>
__device__ __forceinline__ void search_layer(const Q& query, const uint8_t* __restrict__ rec,
                                             uint32_t rec_stride, const GraphView& g, uint32_t layer,
                                             const WarpScratch& s, int& n, int ef, int lane,
                                             SearchCounters& cnt) {
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    // upper layers see few nodes: use (and clear) only a slice of the table
    const uint32_t slots = (layer == 0) ? s.vis_slots : min(s.vis_slots, 1024u);
    const uint32_t vmask = slots - 1, vshift = 32 - (31 - __clz(slots));
    vis_clear(s.vis, slots, lane);
    // visited <- ids(selected)   (results.rs:159-168)
    {
        bool ovf = false;
        for (int i = lane; i < n; i += 32) vis_insert(s.vis, vmask, vshift, (uint32_t)s.list[i], &ovf);
        if (__any_sync(HB_FULL, ovf)) cnt.overflow = 1;
        __syncwarp();
    }
    u64 worst = (n > 0) ? (s.list[n - 1] & KEY_MASK) : ~0ull;
    int cursor = 0;  // every entry before `cursor` is expanded
    while (true) {
        // candidates.pop_first(): first unexpanded entry
        int found = -1;
        for (int c = cursor; c < n; c += 32) {
            int i = c + lane;
            bool un = (i < n) && !(s.list[i] & EXP_FLAG);
            unsigned b = __ballot_sync(HB_FULL, un);
            if (b) { found = c + __ffs(b) - 1; break; }
        }
        if (found < 0) break;
        cursor = found;
        const u64 ck = s.list[cursor];
        __syncwarp();
        if (lane == 0) s.list[cursor] = ck | EXP_FLAG;
        __syncwarp();
        const uint32_t cid = (uint32_t)ck;
        cnt.hops++;
        int minpos = 0x7fffffff;

        // layer.neighbors_vec(cid)  (graph/src/graph.rs:103-113) as fixed-stride rows
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
                // results.insert_visited(node)  (results.rs:101-103)
                bool ovf = false;
                bool isnew = valid && vis_insert(s.vis, vmask, vshift, nb, &ovf);
                if (__any_sync(HB_FULL, ovf)) {
                    // rare: table window full.  Exactness is kept by testing list membership.
                    cnt.overflow = 1;
                    unsigned om = __ballot_sync(HB_FULL, ovf);
                    while (om) {
                        int src = __ffs(om) - 1;
                        om &= om - 1;
                        uint32_t id = __shfl_sync(HB_FULL, nb, src);
                        bool hit = false;
                        for (int i2 = lane; i2 < n; i2 += 32) hit |= ((uint32_t)(s.list[i2] & ~EXP_FLAG) == id);
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
                    // index.get_point(node).dist2other(point)  (searcher.rs:66-69)
                    float d = query.dist(rec + (size_t)cand * rec_stride, gl, gbase);
                    u64 key = make_key(d, cand);
                    // admission (searcher.rs:74-94): unconditional while |selected| < ef, else strict <
                    bool want = act && gl == 0 && (n < ef || key < worst);
                    unsigned am = __ballot_sync(HB_FULL, want);
                    while (am) {
                        int src = __ffs(am) - 1;
                        am &= am - 1;
                        u64 k = __shfl_sync(HB_FULL, key, src);
                        if (n < ef || k < worst) {
                            int pos = list_lower_bound(s.list, n, k, lane);
                            list_insert_at(s.list, n, ef, pos, k, lane);
                            minpos = min(minpos, pos);
                            worst = (n >= ef) ? (s.list[n - 1] & KEY_MASK) : ~0ull;
                        }
                    }
                }
                __syncwarp();
            }
            row = next;
        }
        cursor = min(cursor, minpos);
    }
    // clear_candidates (searcher.rs:100): drop the expanded marks for the next layer
    for (int i = lane; i < n; i += 32) s.list[i] &= KEY_MASK;
    __syncwarp();
}

}  // namespace hb
