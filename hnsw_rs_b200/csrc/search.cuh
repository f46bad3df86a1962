// Warp-resident HNSW layer search with the result list in shared memory (any ef), used by the build
// kernel (csrc/builder.cu) and by queries with ef > 256; plus the visited sets shared with the
// register-list query path (csrc/search_reg.cuh, ef <= 256 -- the path bench.py measures).
//
// Restates Searcher::search_layer (hnsw/src/template/searcher.rs:23-103) over the
// Results sets (hnsw/src/template/results.rs:26-33) with one warp per query:
//   selected    -> sorted array of u64 keys in shared memory, key =
//                  (f32 bits of dist << 32) | id.  Non-negative f32 bit patterns
//                  order like unsigned ints, so integer compare == Dist::cmp
//                  (graph/src/dist.rs:30-37: dist, then id).  The array has
//                  C = 32*KPL slots; unused slots hold the sentinel ~0, so
//                  "|selected| < ef" is the same test as key < list[ef-1].
//   candidates  -> the not-yet-expanded members of `selected` (bit 31 of the id
//                  half marks "expanded"; the sentinel has it set).  A candidate
//                  that has fallen out of `selected` is > the worst selected, so
//                  popping it is exactly the reference's break (searcher.rs:41-44);
//                  hence "no unexpanded entry left" == the reference's loop exit.
//   visited     -> exact open-addressing hash set of ids in shared memory
//                  (Vis16: 16-bit entries for ids < 2^(t+12); Vis32: 32-bit entries).
// Exactly one candidate is expanded per iteration; all unvisited
// neighbours of that candidate are evaluated in parallel, 8 per round, 4 lanes
// each, which is result- and counter-identical to the reference (App. C-5).
//
// List maintenance is lane-major: lane l owns slots [l*KPL, (l+1)*KPL).  A position
// search is one pass (each lane compares its KPL keys, one ballot), an insertion is one
// predicated store per moved key.
#pragma once
#include "dist.cuh"

namespace hb {

constexpr uint32_t EMPTY_ID = 0xFFFFFFFFu;
constexpr uint32_t CHAIN_BIT = 0x80000000u;  // adjacency slot: continuation row marker
constexpr u64 EXP_FLAG = 0x80000000ull;      // list key: "expanded"
constexpr u64 KEY_MASK = ~EXP_FLAG;
constexpr u64 SENTINEL = ~0ull;

struct GraphView {
    const uint32_t* adj0;       // layer 0: row r = node id (r < n_points) or chain row
    uint32_t S0;                // slots per layer-0 row
    const uint32_t* upper_off;  // [n_points] first upper row of the node, EMPTY_ID if level 0
    const uint32_t* upper_adj;  // rows for layers >= 1: row = upper_off[node] + (layer - 1)
    uint32_t SU;                // slots per upper row
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

// ---------------------------------------------------------------------------
// visited sets
// ---------------------------------------------------------------------------
// 32-bit entries: any id < 2^31.
struct Vis32 {
    uint32_t* tab;
    uint32_t tbits;
    static __host__ __device__ size_t bytes(uint32_t tbits) { return (size_t)4 << tbits; }
    __device__ __forceinline__ void clear(int lane) const {
        uint4* p = reinterpret_cast<uint4*>(tab);
        const uint4 e = make_uint4(EMPTY_ID, EMPTY_ID, EMPTY_ID, EMPTY_ID);
        for (uint32_t i = lane; i < (1u << tbits) / 4; i += 32) p[i] = e;
        __syncwarp();
    }
    // Insertion for all 32 lanes at once with warp-uniform control flow (want: this lane has an id to record);
    // returns "id was new" per lane.  On an exhausted probe window the id is reported new but NOT recorded and
    // *ovf is raised; the caller then falls back to a list-membership test so results stay exact.
    __device__ __forceinline__ bool insert_warp(uint32_t id, bool want, bool* ovf) const {
        const uint32_t mask = (1u << tbits) - 1u;
        uint32_t h = (id * 0x9E3779B1u) >> (32 - tbits);
        bool pending = want, isnew = false;
        int probes = 0;
#pragma unroll 1
        while (__any_sync(HB_FULL, pending)) {
            uint32_t old = id;
            if (pending) old = atomicCAS(&tab[h], EMPTY_ID, id);
            const bool won = pending && old == EMPTY_ID;
            const bool step = pending && old != EMPTY_ID && old != id;
            h = step ? ((h + 1) & mask) : h;
            probes += step ? 1 : 0;
            const bool full = step && probes >= 48;
            if (full) *ovf = true;
            isnew = isnew || won || full;
            pending = step && !full;
        }
        return isnew;
    }
};

// 16-bit entries (ids < 2^B, table of T = 2^t entries): h = (id * odd) mod 2^B is a bijection
// on B-bit ids; home = top t bits of h, rem = low B-t bits.  An entry stores
// (rem << 4 | displacement) with displacement <= 14, so (slot, entry) determines the id:
// no false positives, no false negatives.  Needs B - t <= 12.
#ifndef HB_VIS_STS
#define HB_VIS_STS 1  // 1: claim entries with a 16-bit store + read-back; 0: 32-bit compare-and-swap on the entry pair (A/B)
#endif
struct Vis16 {
    uint32_t* words;  // T/2 words holding two entries each
    uint32_t tbits, bbits;
    static __host__ __device__ size_t bytes(uint32_t tbits) { return (size_t)2 << tbits; }
    __device__ __forceinline__ void clear(int lane) const {
        uint4* p = reinterpret_cast<uint4*>(words);
        const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        for (uint32_t i = lane; i < (1u << tbits) / 8; i += 32) p[i] = e;
        __syncwarp();
    }
    // Insertion for all 32 lanes at once with warp-uniform control flow (want: this lane has an id to
    // record).  Every lane executes every iteration of the one loop, so the warp never splits into
    // separately scheduled fragments (per-lane probe loops did: independent thread scheduling never
    // re-joined them and every later collective ran once per fragment).  Returns "id was new" per lane;
    // on an exhausted displacement window the id is reported new but not recorded and *ovf is raised.
    __device__ __forceinline__ bool insert_warp(uint32_t id, bool want, bool* ovf) const {
        // h = (id * odd) mod 2^B, kept top-aligned: the home slot is its top t bits, the remainder the next B-t
        const uint32_t h = (id * 0x9E3779B1u) << (32u - bbits);
        const uint32_t tmask = (1u << tbits) - 1u;
        // (h << t) holds the remainder in its top B-t bits and zeros below: this shift leaves it at bits [4, 4+B-t)
        const uint32_t rem16 = (h << tbits) >> (28u - (bbits - tbits));
        uint32_t slot = h >> (32u - tbits), mine = rem16;  // mine = rem16 | displacement
        bool pending = want, isnew = false;
#if HB_VIS_STS
        // The table belongs to this warp alone, so an entry is claimed with a plain 16-bit store and read back:
        // two lanes that claim the same free entry in the same step write different values ((slot, entry) names the
        // id and the ids of a batch are distinct), so exactly the lane whose value stuck has won, and the loser finds
        // the entry taken by another id and moves on like any lane that met an occupied entry.  A pending lane
        // therefore advances by one displacement per step, in lockstep with the (warp-uniform) step counter: the
        // loop needs no per-lane bookkeeping beyond `pending`, and "window exhausted" is "still pending after 15 steps".
        volatile uint16_t* tab = reinterpret_cast<volatile uint16_t*>(words);
#pragma unroll 1
        for (int it = 0; it < 15 && __any_sync(HB_FULL, pending); ++it) {
            const uint32_t e = tab[slot];
            const bool claim = pending && e == 0xFFFFu;
            if (claim) tab[slot] = (uint16_t)mine;
            __syncwarp();
            const bool done = tab[slot] == mine;  // already there (e == mine) or claimed just now
            isnew = isnew || (claim && done);
            pending = pending && !done;
            slot = (slot + 1u) & tmask;
            ++mine;
        }
        if (pending) *ovf = true;
        return isnew || pending;
#else
#pragma unroll 1
        while (__any_sync(HB_FULL, pending)) {
            uint32_t* wp = words + (slot >> 1);
            const uint32_t sh = (slot & 1u) * 16u;
            const uint32_t w = *reinterpret_cast<volatile uint32_t*>(wp);
            const uint32_t e = (w >> sh) & 0xFFFFu;
            const bool hit = pending && e == mine;
            const bool free_ = pending && e == 0xFFFFu;
            bool won = false;
            if (free_) won = atomicCAS(wp, w, (w & ~(0xFFFFu << sh)) | (mine << sh)) == w;
            // occupied by another id: next displacement; a lost race re-examines the same slot
            const bool step = pending && !hit && !free_;
            slot = step ? ((slot + 1) & tmask) : slot;
            mine += step ? 1u : 0u;
            const bool full = step && (mine & 15u) == 15u;
            if (full) *ovf = true;
            isnew = isnew || won || full;
            pending = pending && !hit && !won && !full;
        }
        return isnew;
#endif
    }
};

// Vis16 with a table that is not a power of two: T = 7 * 512 = 3584 entries (7 KB).  With the 128 + 256 + 512 bytes of
// the other per-warp buffers a warp then needs 8064 bytes, so SEVEN blocks of four warps fit one SM (28 warps instead of
// the 24 that a 4096-entry table allows; the kernel is latency bound and gains more from the four extra warps than it
// loses to the higher load factor).  home = floor(hB * T / 2^B) for the B-bit bijection value hB; the entry stores
// rem = hB - ceil(home * 2^B / T), the offset of hB inside its home's range, so (slot, entry) still names the id exactly.
// Needs 12 <= B <= 23 (rem < 4096).
struct Vis16N {
    static constexpr uint32_t T = 3584;
    uint32_t* words;
    uint32_t bbits;
    static __host__ __device__ size_t bytes(uint32_t) { return (size_t)2 * T; }
    static __host__ __device__ bool fits(uint32_t bbits) { return bbits >= 12 && bbits <= 23; }
    __device__ __forceinline__ void clear(int lane) const {
        uint4* p = reinterpret_cast<uint4*>(words);
        const uint4 e = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        for (uint32_t i = lane; i < T / 8; i += 32) p[i] = e;
        __syncwarp();
    }
    // same contract and the same store-and-read-back protocol as Vis16::insert_warp
    __device__ __forceinline__ bool insert_warp(uint32_t id, bool want, bool* ovf) const {
        const uint32_t h = (id * 0x9E3779B1u) << (32u - bbits);  // hB, top-aligned
        const uint32_t home = __umulhi(h, T);                    // floor(hB * T / 2^B)
        const uint32_t base = ((home << (bbits - 9u)) + 6u) / 7u;  // ceil(home * 2^B / T), T = 7 * 2^9
        uint32_t slot = home, mine = ((h >> (32u - bbits)) - base) << 4;  // mine = rem << 4 | displacement
        bool pending = want, isnew = false;
        volatile uint16_t* tab = reinterpret_cast<volatile uint16_t*>(words);
#pragma unroll 1
        for (int it = 0; it < 15 && __any_sync(HB_FULL, pending); ++it) {
            const uint32_t e = tab[slot];
            const bool claim = pending && e == 0xFFFFu;
            if (claim) tab[slot] = (uint16_t)mine;
            __syncwarp();
            const bool done = tab[slot] == mine;
            isnew = isnew || (claim && done);
            pending = pending && !done;
            ++slot;
            slot = slot == T ? 0u : slot;
            ++mine;
        }
        if (pending) *ovf = true;
        return isnew || pending;
    }
};

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// every 128-byte line of one record (f32 records span several)
__device__ __forceinline__ void prefetch_record(const uint8_t* rp8, uint32_t rec_stride) {
    prefetch_l2(rp8);
    if (rec_stride > 128) {
        prefetch_l2(rp8 + 128);
        for (uint32_t o = 256; o < rec_stride; o += 128) prefetch_l2(rp8 + o);
    }
}

// ---------------------------------------------------------------------------
// sorted key list in shared memory, lane-major access
// ---------------------------------------------------------------------------
// KPL > 0: compile-time keys per lane (capacity 32*KPL); KPL == 0: runtime kpl.
template <int KPL>
struct KeyList {
    u64* list;
    int kpl;  // keys per lane (even)
    __device__ __forceinline__ int K() const { return KPL ? KPL : kpl; }
    __device__ __forceinline__ int cap() const { return 32 * K(); }
    __device__ __forceinline__ void reset(int lane) const {
        for (int i = lane; i < cap(); i += 32) list[i] = SENTINEL;
        __syncwarp();
    }
    // insert `key` (< list[ef-1] masked) keeping the list sorted and at most ef long.
    // Returns the position.  All lanes call together.
    __device__ __forceinline__ int insert(u64 key, int ef, int lane) const {
        const int k = K();
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(list + lane * k);
        int c = 0;
        if (KPL) {
            u64 mine[KPL ? KPL : 2];
#pragma unroll
            for (int j = 0; j < KPL / 2; ++j) {
                ulonglong2 v = p[j];
                mine[2 * j] = v.x;
                mine[2 * j + 1] = v.y;
            }
#pragma unroll
            for (int j = 0; j < KPL; ++j) c += ((mine[j] & KEY_MASK) < key) ? 1 : 0;
            const int full = __popc(__ballot_sync(HB_FULL, c == KPL));
            const int cb = __shfl_sync(HB_FULL, c, full & 31);
            const int pos = full * KPL + (full < 32 ? cb : 0);
            __syncwarp();
            // slot q >= pos moves to q + 1; the last slot of the array falls off
#pragma unroll
            for (int j = 0; j < KPL; ++j) {
                const int q = lane * KPL + j;
                if (q >= pos && q + 1 < 32 * KPL) list[q + 1] = mine[j];
            }
            __syncwarp();
            if (lane == 0) {
                list[pos] = key;
                if (ef < 32 * KPL) list[ef] = SENTINEL;
            }
            __syncwarp();
            return pos;
        } else {
            for (int j = 0; j < k / 2; ++j) {
                ulonglong2 v = p[j];
                c += ((v.x & KEY_MASK) < key) ? 1 : 0;
                c += ((v.y & KEY_MASK) < key) ? 1 : 0;
            }
            const int full = __popc(__ballot_sync(HB_FULL, c == k));
            const int cb = __shfl_sync(HB_FULL, c, full & 31);
            const int pos = full * k + (full < 32 ? cb : 0);
            // move from the top down, one 32-wide stripe at a time
            const int last_src = min(ef, 32 * k) - 2;
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
            __syncwarp();
            return pos;
        }
    }
};

// helpers on a plain sorted array (used by the build's candidate window)
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

// ---------------------------------------------------------------------------
// one layer of best-first search
// ---------------------------------------------------------------------------
// On entry the list holds the entry set (sorted, flags clear, sentinels behind it); on exit
// it holds the <= ef nearest evaluated nodes (sorted, flags clear).  newbuf: 32 u32 scratch.
template <class Q, class VIS, int KPL>
__device__ __forceinline__ void search_layer(const Q& query, const uint8_t* __restrict__ rec,
                                             uint32_t rec_stride, const GraphView& g, uint32_t layer,
                                             const KeyList<KPL>& L, const VIS& vis, uint32_t* newbuf,
                                             int ef, int lane, SearchCounters& cnt) {
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    u64* list = L.list;
    vis.clear(lane);
    {   // visited <- ids(selected)   (results.rs:159-168)
        bool ovf = false;
        for (int i0 = 0; i0 < ef; i0 += 32) {
            const int i = i0 + lane;
            const u64 k = i < ef ? list[i] : SENTINEL;
            vis.insert_warp((uint32_t)k, k != SENTINEL, &ovf);
        }
        if (__any_sync(HB_FULL, ovf)) cnt.overflow = 1;
        __syncwarp();
    }
    u64 worst = list[ef - 1] & KEY_MASK;  // masked sentinel (max) while |selected| < ef
    int cursor = 0;                       // every entry before `cursor` is expanded
    while (true) {
        // candidates.pop_first(): first entry whose "expanded" bit is clear
        int found = -1;
        for (int c = cursor; c < ef; c += 32) {
            int i = c + lane;
            bool un = (i < ef) && !(list[i] & EXP_FLAG);
            unsigned b = __ballot_sync(HB_FULL, un);
            if (b) { found = c + __ffs(b) - 1; break; }
        }
        if (found < 0) break;
        cursor = found;
        const u64 ck = list[cursor];
        __syncwarp();
        if (lane == 0) list[cursor] = ck | EXP_FLAG;
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
                bool isnew = vis.insert_warp(nb, valid, &ovf);
                if (__any_sync(HB_FULL, ovf)) {
                    // rare: probe window exhausted.  Exactness is kept by testing list membership.
                    cnt.overflow = 1;
                    unsigned om = __ballot_sync(HB_FULL, ovf);
                    while (om) {
                        int src = __ffs(om) - 1;
                        om &= om - 1;
                        uint32_t id = __shfl_sync(HB_FULL, nb, src);
                        bool hit = false;
                        for (int i2 = lane; i2 < ef; i2 += 32) {
                            u64 k2 = list[i2];
                            hit |= (k2 != SENTINEL) && ((uint32_t)(k2 & ~EXP_FLAG) == id);
                        }
                        if (__any_sync(HB_FULL, hit) && lane == src) isnew = false;
                    }
                }
                // all records of this batch are requested at once (a second round does not pay
                // a second memory latency)
                if (isnew) {
                    const uint8_t* rp8 = rec + (size_t)nb * rec_stride;
                    prefetch_record(rp8, rec_stride);
                }
                unsigned nm = __ballot_sync(HB_FULL, isnew);
                int ncnt = __popc(nm);
                if (ncnt == 0) continue;
                cnt.evals += ncnt;
                if (isnew) newbuf[__popc(nm & ((1u << lane) - 1))] = nb;
                __syncwarp();
                for (int r0 = 0; r0 < ncnt; r0 += 8) {
                    int idx = r0 + grp;
                    bool act = idx < ncnt;
                    uint32_t cand = newbuf[act ? idx : 0];
                    // index.get_point(node).dist2other(point)  (searcher.rs:66-69)
                    float d = query.dist(rec + (size_t)cand * rec_stride, gl, gbase);
                    u64 key = make_key(d, cand);
                    // admission (searcher.rs:74-94): key < list[ef-1] covers |selected| < ef and strict <
                    bool want = act && gl == 0 && key < worst;
                    unsigned am = __ballot_sync(HB_FULL, want);
                    while (am) {
                        int src = __ffs(am) - 1;
                        am &= am - 1;
                        u64 k = __shfl_sync(HB_FULL, key, src);
                        if (k < worst) {
                            int pos = L.insert(k, ef, lane);
                            minpos = min(minpos, pos);
                            worst = list[ef - 1] & KEY_MASK;
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
    for (int i = lane; i < ef; i += 32) {
        u64 k = list[i];
        if (k != SENTINEL) list[i] = k & KEY_MASK;
    }
    __syncwarp();
}

// number of real entries among the first ef slots
__device__ __forceinline__ int list_count(const u64* list, int ef, int lane) {
    int c = 0;
    for (int i = lane; i < ef; i += 32) c += (list[i] != SENTINEL) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(HB_FULL, c, o);
    return c;
}

}  // namespace hb
