// Index build: HNSW::insert_bulk / insert (hnsw/src/template.rs:177-251, 388-444) with
// the per-point search + neighbour-selection heuristic on the device and the edge
// commit (make_connections / prune_connections / make_pruned_connections) on the host.
//
// Division of labour (north-star item 3): everything that evaluates distances runs in
// build_kernel below -- Inserter::build_insertion_results (template/inserter.rs:40-126):
// greedy descent, search_layer(ef_cons) and select_heuristic (template/searcher.rs:109-153,
// template/results.rs:105-146,69-77) for a BATCH of new points against one frozen
// snapshot of the graph.  The host then commits the chosen edges point by point in the
// reference's order; the prune step needs only distances that the device already
// produced (every edge carries its length), so the commit never waits for the GPU.
//
// batch == 1 reproduces the reference's single-thread build exactly (each point sees
// every earlier point's edges).  batch > 1 is the device analogue of the reference's
// nb_threads > 1 mode, where concurrently inserted points do not see each other's
// edges either (template.rs:401-440) -- but deterministic.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "commit.h"
#include "engine.h"
#include "search.cuh"

namespace hb {

#define HB_DISPATCH_DIM_B(L, ...)                                                  \
    do {                                                                           \
        if ((L).kind == HB_REC_F32) { using Q = FullQuery; __VA_ARGS__; }          \
        else if ((L).dim == 100) { using Q = RegQuery<12, 4>; __VA_ARGS__; }       \
        else if ((L).dim == 128) { using Q = RegQuery<16, 0>; __VA_ARGS__; }       \
        else if ((L).dim == 96) { using Q = RegQuery<12, 0>; __VA_ARGS__; }        \
        else if ((L).dim == 50) { using Q = RegQuery<6, 2>; __VA_ARGS__; }         \
        else { using Q = SmemQuery; __VA_ARGS__; }                                 \
    } while (0)

// ---------------------------------------------------------------------------
// scatter of touched adjacency rows into the device mirror
// ---------------------------------------------------------------------------
__global__ void scatter_rows_kernel(uint32_t* dst, uint32_t S, const uint32_t* __restrict__ rows,
                                    const uint32_t* __restrict__ data, uint32_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)n * S) return;
    uint32_t r = (uint32_t)(i / S), c = (uint32_t)(i % S);
    dst[(size_t)rows[r] * S + c] = data[i];
}
cudaError_t launch_scatter_rows(uint32_t* dst, uint32_t S, const uint32_t* rows, const uint32_t* data,
                                uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    uint64_t tot = (uint64_t)n * S;
    scatter_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(dst, S, rows, data, n);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// build kernel
// ---------------------------------------------------------------------------
struct BuildParams {
    const uint8_t* rec;
    RecLayout L;
    GraphView g;
    uint32_t n_layers, ep;
    const uint32_t* job_ids;    // [njobs] point ids to insert
    const uint8_t* job_levels;  // [njobs]
    uint32_t njobs;
    uint32_t ef_cons, m;
    uint32_t kpl, tbits, bbits, cand_cap, m_cap, qd_cap;
    uint32_t* out_ids;   // [njobs][n_layers][m]
    float* out_dists;    // [njobs][n_layers][m]
    uint32_t* out_cnt;   // [njobs][n_layers]
    uint32_t* out_evals; // [njobs]
    uint32_t* work_counter;
};

constexpr int BUILD_WPB = 2;

template <class VIS>
__host__ __device__ inline size_t build_warp_smem(const BuildParams& p) {
    return (size_t)32 * p.kpl * 8 + (size_t)p.cand_cap * 8 + (size_t)p.m_cap * 16 + VIS::bytes(p.tbits) + 128 +
           (size_t)p.qd_cap * 8;
}
__device__ __forceinline__ void make_vis(Vis16& v, unsigned char* mem, const BuildParams& p) {
    v.words = reinterpret_cast<uint32_t*>(mem);
    v.tbits = p.tbits;
    v.bbits = p.bbits;
}
__device__ __forceinline__ void make_vis(Vis32& v, unsigned char* mem, const BuildParams& p) {
    v.tab = reinterpret_cast<uint32_t*>(mem);
    v.tbits = p.tbits;
}

// One pass of extend_candidates_with_neighbors (results.rs:122-146) restricted to keys
// > lower: candidates <- {selected} U {Dist(n, d(point, n)) : s in selected, n in N(s)},
// kept as the cand_cap smallest in `cand` (sorted).  Returns true if something was dropped.
template <class Q, class VIS>
__device__ __forceinline__ bool heuristic_fill(const Q& query, const uint8_t* __restrict__ rec,
                                               uint32_t rec_stride, const GraphView& g, uint32_t layer,
                                               const u64* list, int n, u64* cand, int& cn, int cand_cap,
                                               u64 lower, bool have_lower, const VIS& vis,
                                               uint32_t* newbuf, int lane, uint32_t& evals) {
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    bool dropped = false;
    cn = 0;
    vis.clear(lane);
    // the old selected set itself (select_setup, results.rs:105-111)
    for (int i = 0; i < n; ++i) {
        u64 k = list[i] & KEY_MASK;
        if (have_lower && k <= lower) continue;
        if (cn < cand_cap || k < (cand[cn - 1] & KEY_MASK)) {
            if (cn == cand_cap) dropped = true;
            int pos = list_lower_bound(cand, cn, k, lane);
            list_insert_at(cand, cn, cand_cap, pos, k, lane);
        } else dropped = true;
    }
    {
        bool ovf = false;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            vis.insert_warp(i < n ? (uint32_t)(list[i] & KEY_MASK) : 0u, i < n, &ovf);
        }
        __syncwarp();
    }
    for (int si = 0; si < n; ++si) {
        const uint32_t sid = (uint32_t)(list[si] & KEY_MASK);
        const uint32_t* base;
        uint32_t S, row;
        if (layer == 0) { base = g.adj0; S = g.S0; row = sid; }
        else { base = g.upper_adj; S = g.SU; row = __ldg(g.upper_off + sid) + (layer - 1); }
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
                // The reference evaluates every occurrence and lets the ordered set collapse
                // duplicates; evaluating each distinct id once gives the same set.  If the
                // table window is full the id is simply evaluated again (duplicate keys are
                // skipped by the consumer).
                bool ovf = false;
                bool isnew = vis.insert_warp(nb, valid, &ovf);
                unsigned nm = __ballot_sync(HB_FULL, isnew);
                int ncnt = __popc(nm);
                if (ncnt == 0) continue;
                evals += ncnt;
                if (isnew) newbuf[__popc(nm & ((1u << lane) - 1))] = nb;
                __syncwarp();
                for (int r0 = 0; r0 < ncnt; r0 += 8) {
                    int idx = r0 + grp;
                    bool act = idx < ncnt;
                    uint32_t c = newbuf[act ? idx : 0];
                    // points.distance(point.id, neighbor)  (results.rs:141-144)
                    float d = query.dist(rec + (size_t)c * rec_stride, gl, gbase);
                    u64 key = make_key(d, c);
                    bool cons = act && gl == 0 && !(have_lower && key <= lower);
                    unsigned am = __ballot_sync(HB_FULL, cons);
                    while (am) {
                        int src = __ffs(am) - 1;
                        am &= am - 1;
                        u64 k = __shfl_sync(HB_FULL, key, src);
                        if (cn < cand_cap || k < (cand[cn - 1] & KEY_MASK)) {
                            if (cn == cand_cap) dropped = true;
                            int pos = list_lower_bound(cand, cn, k, lane);
                            // an id evaluated twice (table overflow) yields an identical key: skip it
                            bool dup = pos < cn && (cand[pos] & KEY_MASK) == k;
                            if (!dup) list_insert_at(cand, cn, cand_cap, pos, k, lane);
                        } else dropped = true;
                    }
                }
                __syncwarp();
            }
            row = next;
        }
    }
    return dropped;
}

// Searcher::select_heuristic(m, extend_cands = true, keep_pruned = true)
// (searcher.rs:109-153).  On exit list[0..n) is the new selected set, sorted.
template <class Q, class VIS>
__device__ __forceinline__ void select_heuristic(const Q& query, const BuildParams& p, uint32_t layer,
                                                 u64* list, int& n, u64* cand, u64* sel, u64* rej,
                                                 const VIS& vis, uint32_t* newbuf, float* qd2, int lane,
                                                 uint32_t& evals) {
    const int gl = lane & 3, gbase = lane & ~3, grp = lane >> 2;
    const int m = (int)p.m;
    int cn = 0, ci = 0, sn = 0, rn = 0;
    bool dropped = heuristic_fill(query, p.rec, p.L.stride, p.g, layer, list, n, cand, cn, (int)p.cand_cap,
                                  0ull, false, vis, newbuf, lane, evals);
    u64 last = 0;
    bool first = true;
    while (sn < m) {
        if (ci >= cn) {
            if (!dropped) break;  // candidates exhausted
            // the bounded window ran dry: refill it with the next smallest keys > last
            dropped = heuristic_fill(query, p.rec, p.L.stride, p.g, layer, list, n, cand, cn, (int)p.cand_cap,
                                     last, true, vis, newbuf, lane, evals);
            ci = 0;
            if (cn == 0) break;
        }
        const u64 e = cand[ci++] & KEY_MASK;
        last = e;
        if (first) {  // the nearest candidate is always accepted (searcher.rs:124-125)
            if (lane == 0) sel[sn] = e;
            ++sn;
            first = false;
            __syncwarp();
            continue;
        }
        // get_nearest_from_selected(e): min_s Dist(s.id, d(e, s))  (results.rs:69-77)
        __syncwarp();
        warp_dequant_record(p.L, p.rec + (size_t)(uint32_t)e * p.L.stride, lane, qd2);
        __syncwarp();
        Q eq;
        eq.init(p.L, qd2, gl);
        u64 nearest = ~0ull;
        for (int r0 = 0; r0 < sn; r0 += 8) {
            int idx = r0 + grp;
            bool act = idx < sn;
            uint32_t sid = (uint32_t)sel[act ? idx : 0];
            float d = eq.dist(p.rec + (size_t)sid * p.L.stride, gl, gbase);
            u64 k = act ? make_key(d, sid) : ~0ull;
            nearest = min(nearest, k);
        }
        evals += sn;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nearest = min(nearest, __shfl_xor_sync(HB_FULL, nearest, o));
        if (e < nearest) {
            if (lane == 0) sel[sn] = e;
            ++sn;
        } else if (rn < m) {  // keep_pruned: only the first m rejected can ever be used
            if (lane == 0) rej[rn] = e;
            ++rn;
        }
        __syncwarp();
    }
    // fill from visited_h in ascending order (searcher.rs:139-144), then emit the union sorted
    int take = min(rn, m - sn);
    // merge sel[0..sn) and rej[0..take), both ascending
    __syncwarp();
    if (lane == 0) {
        int a = 0, b = 0, o = 0;
        while (a < sn || b < take) {
            bool ta = (b >= take) || (a < sn && sel[a] < rej[b]);
            list[o++] = ta ? sel[a++] : rej[b++];
        }
    }
    n = sn + take;
    __syncwarp();
}

template <class Q, class VIS>
__global__ void __launch_bounds__(BUILD_WPB * 32) build_kernel(BuildParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gl = lane & 3, gbase = lane & ~3;
    unsigned char* wsm = smem + (size_t)wib * build_warp_smem<VIS>(p);
    KeyList<0> L;
    L.list = reinterpret_cast<u64*>(wsm);
    L.kpl = (int)p.kpl;
    u64* list = L.list;
    u64* cand = list + 32 * p.kpl;
    u64* sel = cand + p.cand_cap;
    u64* rej = sel + p.m_cap;
    unsigned char* vmem = reinterpret_cast<unsigned char*>(rej + p.m_cap);
    VIS vis;
    make_vis(vis, vmem, p);
    uint32_t* newbuf = reinterpret_cast<uint32_t*>(vmem + VIS::bytes(p.tbits));
    float* qd = reinterpret_cast<float*>(newbuf + 32);
    float* qd2 = qd + p.qd_cap;

    while (true) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(p.work_counter, 1u);
        j = __shfl_sync(HB_FULL, j, 0);
        if (j >= p.njobs) break;
        const uint32_t pid = __ldg(p.job_ids + j);
        const uint32_t level = __ldg(p.job_levels + j);
        uint32_t* ocnt = p.out_cnt + (size_t)j * p.n_layers;
        for (uint32_t l = lane; l < p.n_layers; l += 32) ocnt[l] = 0;
        if (pid == p.ep) {  // inserter.rs:42-45: the entry point is never inserted
            if (lane == 0) p.out_evals[j] = 0;
            continue;
        }
        __syncwarp();
        warp_dequant_record(p.L, p.rec + (size_t)pid * p.L.stride, lane, qd);
        __syncwarp();
        Q q;
        q.init(p.L, qd, gl);
        SearchCounters cnt{0u, 1u, 0u, 0u};
        // setup_insert (inserter.rs:53-68): selected <- {Dist(ep, distance(ep, id))}
        float d0 = q.dist(p.rec + (size_t)p.ep * p.L.stride, gl, gbase);
        L.reset(lane);
        if (lane == 0) list[0] = make_key(d0, p.ep);
        __syncwarp();
        // traverse_layers_above (inserter.rs:70-89)
        for (uint32_t layer = p.n_layers - 1; layer > level; --layer)
            search_layer<Q, VIS, 0>(q, p.rec, p.L.stride, p.g, layer, L, vis, newbuf, 1, lane, cnt);
        // traverse_layers_below (inserter.rs:91-126)
        uint32_t bound = min(level, p.n_layers - 1);
        for (uint32_t layer = bound + 1; layer-- > 0;) {
            search_layer<Q, VIS, 0>(q, p.rec, p.L.stride, p.g, layer, L, vis, newbuf, (int)p.ef_cons, lane, cnt);
            int n = list_count(list, (int)p.ef_cons, lane);
            select_heuristic<Q, VIS>(q, p, layer, list, n, cand, sel, rej, vis, newbuf, qd2, lane, cnt.evals);
            // the heuristic's picks are the entry set of the next layer: restore the sentinel tail
            for (int i = n + lane; i < 32 * (int)p.kpl; i += 32) list[i] = SENTINEL;
            // save_layer_results (results.rs:79-84)
            uint32_t* oi = p.out_ids + ((size_t)j * p.n_layers + layer) * p.m;
            float* od = p.out_dists + ((size_t)j * p.n_layers + layer) * p.m;
            for (int i = lane; i < n; i += 32) {
                u64 k = list[i];
                oi[i] = (uint32_t)k;
                od[i] = __uint_as_float((uint32_t)(k >> 32));
            }
            if (lane == 0) ocnt[layer] = (uint32_t)n;
            __syncwarp();
        }
        if (lane == 0) p.out_evals[j] = cnt.evals;
    }
}

// ---------------------------------------------------------------------------
// rand 0.8.5 StdRng (ChaCha12, seed_from_u64 via PCG32) level draw, as
// SimplePoints::new does (points/src/points.rs:39-48,148-160).  Third-party
// algorithm restated from its published definition; not pinned by any
// reference test (DESIGN.md "unpinned").
// ---------------------------------------------------------------------------
struct ChaCha12 {
    uint32_t key[8];
    uint64_t counter = 0;
    uint32_t buf[64];
    int idx = 64;
    static uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
    void seed_from_u64(uint64_t state) {
        for (int i = 0; i < 8; ++i) {
            state = state * 6364136223846793005ull + 11634580027462260723ull;
            uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
        }
        counter = 0;
        idx = 64;
    }
    void block(uint64_t ctr, uint32_t* out) {
        uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                          key[4], key[5], key[6], key[7], (uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
        uint32_t x[16];
        memcpy(x, s, sizeof(x));
        auto qr = [&](int a, int b, int c, int d) {
            x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
            x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
            x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
            x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
        };
        for (int r = 0; r < 6; ++r) {
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
        }
        for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
    }
    uint32_t next_u32() {
        if (idx >= 64) {
            for (int b = 0; b < 4; ++b) block(counter + b, buf + 16 * b);
            counter += 4;
            idx = 0;
        }
        return buf[idx++];
    }
    float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
};

void draw_levels(uint64_t m, uint64_t n, uint8_t* out) {
    ChaCha12 rng;
    rng.seed_from_u64(0);
    const float ml = 1.0f / logf((float)m);
    for (uint64_t i = 0; i < n; ++i) {
        float r = 0.0f;
        while (r == 0.0f || r == 1.0f) r = rng.gen_f32();
        out[i] = (uint8_t)(size_t)floorf(-logf(r) * ml);
    }
}

// edge lengths for a graph that was imported without them
static int ensure_weights(hnswb200_ctx* c, hnswb200_index* ix) {
    HostGraph& h = ix->graph->h;
    if (h.weights_valid) return 0;
    for (int which = 0; which < 2; ++which) {
        AdjStore& s = which ? h.au : h.a0;
        std::vector<uint32_t> src, off, ids, rows;
        off.push_back(0);
        for (uint64_t node = 0; node < h.n_points(); ++node) {
            uint32_t lo = which ? 1 : 0, hi = which ? h.level[node] : 0;
            for (uint32_t l = lo; l <= hi; ++l) {
                uint32_t row = h.row((uint32_t)node, l);
                if (s.deg[row] == 0) continue;
                src.push_back((uint32_t)node);
                rows.push_back(row);
                for (uint32_t i = 0; i < s.deg[row]; ++i) ids.push_back(s.get(row, i));
                off.push_back((uint32_t)ids.size());
            }
        }
        if (src.empty()) continue;
        DevBuf<uint32_t> dsrc, doff, dids;
        DevBuf<float> dout;
        HB_CUDA(dsrc.alloc(src.size()));
        HB_CUDA(doff.alloc(off.size()));
        HB_CUDA(dids.alloc(ids.size()));
        HB_CUDA(dout.alloc(ids.size()));
        HB_CUDA(cudaMemcpyAsync(dsrc.p, src.data(), src.size() * 4, cudaMemcpyHostToDevice, c->stream));
        HB_CUDA(cudaMemcpyAsync(doff.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, c->stream));
        HB_CUDA(cudaMemcpyAsync(dids.p, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, c->stream));
        HB_CUDA(launch_dist_one_to_many(ix->points->d_rec, ix->points->L, dsrc.p, doff.p, dids.p, (uint32_t)src.size(),
                                        dout.p, c->stream));
        std::vector<float> out(ids.size());
        HB_CUDA(cudaMemcpyAsync(out.data(), dout.p, ids.size() * 4, cudaMemcpyDeviceToHost, c->stream));
        HB_CUDA(cudaStreamSynchronize(c->stream));
        for (size_t j = 0; j < rows.size(); ++j)
            for (uint32_t i = 0; i < s.deg[rows[j]]; ++i) s.setw(rows[j], i, out[off[j] + i]);
    }
    h.weights_valid = true;
    return 0;
}

// Insert the (already stored) points new_ids: level classes from the top, ascending id
// inside a class (template.rs:403-416 with the oracle's iteration convention).
// Everything build_insert can refuse for reasons known up front (parameters, shared-memory working set), checked
// BEFORE store_points mutates the index: a refused insert leaves the index exactly as it was.
static int build_validate(const hnswb200_index* ix, uint64_t n_points_after) {
    const hnswb200_params& prm = ix->params;
    if (prm.ef_cons < prm.m)
        return (set_error("build: ef_cons < m is not supported by the device build"), HNSWB200_EINVAL);
    if (prm.m < 2) return (set_error("build: m must be >= 2"), HNSWB200_EINVAL);
    BuildParams p{};
    p.L = ix->points->L;
    p.ef_cons = (uint32_t)prm.ef_cons;
    p.m = (uint32_t)prm.m;
    p.kpl = ((std::max<uint32_t>(p.ef_cons, p.m) + 31) / 32 + 1) / 2 * 2;
    p.cand_cap = std::max<uint32_t>(256, 32 * p.kpl);
    p.m_cap = (p.m + 1) / 2 * 2;
    p.qd_cap = (p.L.dim + 7) / 8 * 8 + 8;
    bool use16;
    choose_visited(p.ef_cons, ix->graph->h.a0.S, n_points_after, &p.tbits, &p.bbits, &use16);
    auto bytes = [&]() { return (use16 ? build_warp_smem<Vis16>(p) : build_warp_smem<Vis32>(p)) * BUILD_WPB; };
    while (bytes() > 200 * 1024 && p.tbits > 9 && (!use16 || p.bbits <= p.tbits - 1 + 12)) --p.tbits;
    if (bytes() > 227 * 1024) return (set_error("build: ef_cons too large for the shared-memory working set"), HNSWB200_EINVAL);
    return 0;
}

int build_insert(hnswb200_ctx* c, hnswb200_index* ix, const std::vector<uint32_t>& new_ids, uint32_t batch) {
    if (new_ids.empty()) return 0;
    if (c->use()) return HNSWB200_ECUDA;
    hnswb200_graph* G = ix->graph;
    HostGraph& h = G->h;
    const hnswb200_params& prm = ix->params;
    int rc = build_validate(ix, h.n_points());
    if (rc) return rc;
    rc = ensure_weights(c, ix);
    if (rc) return rc;
    std::vector<uint32_t> d0, du;
    rc = G->sync_new_nodes(*std::min_element(new_ids.begin(), new_ids.end()));
    if (rc) return rc;
    const uint32_t nl = h.n_layers();
    const uint32_t m = (uint32_t)prm.m;
    const uint32_t maxb = batch ? batch : 4096;

    BuildParams p;
    p.rec = ix->points->d_rec;
    p.L = ix->points->L;
    p.n_layers = nl;
    p.ep = prm.ep;
    p.ef_cons = (uint32_t)prm.ef_cons;
    p.m = m;
    p.kpl = ((std::max<uint32_t>(p.ef_cons, m) + 31) / 32 + 1) / 2 * 2;
    p.cand_cap = std::max<uint32_t>(256, 32 * p.kpl);
    p.m_cap = (m + 1) / 2 * 2;
    p.qd_cap = (p.L.dim + 7) / 8 * 8 + 8;
    bool use16;
    choose_visited(p.ef_cons, h.a0.S, h.n_points(), &p.tbits, &p.bbits, &use16);
    auto bytes = [&]() { return (use16 ? build_warp_smem<Vis16>(p) : build_warp_smem<Vis32>(p)) * BUILD_WPB; };
    while (bytes() > 200 * 1024 && p.tbits > 9 && (!use16 || p.bbits <= p.tbits - 1 + 12)) --p.tbits;
    const size_t smem = bytes();
    if (smem > 227 * 1024) return (set_error("build: ef_cons too large for the shared-memory working set"), HNSWB200_EINVAL);

    DevBuf<uint32_t> d_jobs, d_oids, d_ocnt, d_oev;
    DevBuf<uint8_t> d_jlv;
    DevBuf<float> d_od;
    HB_CUDA(d_jobs.alloc(maxb));
    HB_CUDA(d_jlv.alloc(maxb));
    HB_CUDA(d_oids.alloc((size_t)maxb * nl * m));
    HB_CUDA(d_od.alloc((size_t)maxb * nl * m));
    HB_CUDA(d_ocnt.alloc((size_t)maxb * nl));
    HB_CUDA(d_oev.alloc(maxb));
    p.job_ids = d_jobs.p;
    p.job_levels = d_jlv.p;
    p.out_ids = d_oids.p;
    p.out_dists = d_od.p;
    p.out_cnt = d_ocnt.p;
    p.out_evals = d_oev.p;
    p.work_counter = c->d_scratch;
    PinBuf<uint32_t> o_ids, o_cnt;  // page-locked: the copies back run at link speed and do not block the host
    PinBuf<float> o_d;
    HB_CUDA(o_ids.alloc((size_t)maxb * nl * m));
    HB_CUDA(o_cnt.alloc((size_t)maxb * nl));
    HB_CUDA(o_d.alloc((size_t)maxb * nl * m));
    std::vector<uint8_t> jl(maxb);

    int grid_cap = 0;
    HB_DISPATCH_DIM_B(p.L, {
        int occ = 0;
        if (use16) {
            HB_CUDA(cudaFuncSetAttribute(build_kernel<Q, Vis16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, build_kernel<Q, Vis16>, BUILD_WPB * 32, smem));
        } else {
            HB_CUDA(cudaFuncSetAttribute(build_kernel<Q, Vis32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, build_kernel<Q, Vis32>, BUILD_WPB * 32, smem));
        }
        grid_cap = c->num_sms * (occ < 1 ? 1 : occ);
    });

    // order: level classes from the top; within a class ascending id
    std::vector<uint32_t> order(new_ids);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return h.level[a] != h.level[b] ? h.level[a] > h.level[b] : a < b;
    });
    // points already linked before this call count towards the batch ramp
    uint64_t linked = h.n_points() - new_ids.size();
    size_t pos = 0;
    std::vector<LayerSel> res;
    CommitScratch cs;
    const bool prof = getenv("HNSWB200_BUILD_PROFILE") != nullptr;
    double t_kernel = 0, t_commit = 0, t_upload = 0;
    uint64_t n_batches = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    while (pos < order.size()) {
        double t0 = now();
        ++n_batches;
        // ramp: a batch never exceeds a quarter of what is already linked, so early
        // points (which shape the upper layers) are inserted (almost) one by one
        uint64_t lim = batch == 1 ? 1 : std::max<uint64_t>(1, linked / 4);
        uint32_t nb = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(maxb, lim), order.size() - pos);
        for (uint32_t j = 0; j < nb; ++j) jl[j] = h.level[order[pos + j]];
        p.njobs = nb;
        p.g.adj0 = G->d_adj0; p.g.S0 = h.a0.S; p.g.upper_off = G->d_upper_off; p.g.upper_adj = G->d_adju; p.g.SU = h.au.S;
        HB_CUDA(cudaMemcpyAsync(d_jobs.p, &order[pos], nb * 4, cudaMemcpyHostToDevice, c->stream));
        HB_CUDA(cudaMemcpyAsync(d_jlv.p, jl.data(), nb, cudaMemcpyHostToDevice, c->stream));
        HB_CUDA(cudaMemsetAsync(c->d_scratch, 0, 4, c->stream));
        int grid = std::min<int>(grid_cap, (int)((nb + BUILD_WPB - 1) / BUILD_WPB));
        HB_DISPATCH_DIM_B(p.L, {
            if (use16) build_kernel<Q, Vis16><<<grid, BUILD_WPB * 32, smem, c->stream>>>(p);
            else build_kernel<Q, Vis32><<<grid, BUILD_WPB * 32, smem, c->stream>>>(p);
        });
        HB_CUDA(cudaGetLastError());
        HB_CUDA(cudaMemcpyAsync(o_ids.data(), d_oids.p, (size_t)nb * nl * m * 4, cudaMemcpyDeviceToHost, c->stream));
        HB_CUDA(cudaMemcpyAsync(o_d.data(), d_od.p, (size_t)nb * nl * m * 4, cudaMemcpyDeviceToHost, c->stream));
        HB_CUDA(cudaMemcpyAsync(o_cnt.data(), d_ocnt.p, (size_t)nb * nl * 4, cudaMemcpyDeviceToHost, c->stream));
        HB_CUDA(cudaStreamSynchronize(c->stream));
        double t1 = now();
        // the commit is a chain of dependent cache misses on the rows of the selected neighbours:
        // touch the rows of a job a few jobs ahead of the one being committed
        auto touch_rows = [&](uint32_t j) {
            for (uint32_t l = 0; l < nl; ++l) {
                uint32_t cnt = o_cnt[(size_t)j * nl + l];
                if (cnt == 0) continue;
                const AdjStore& s = h.store(l);
                const uint32_t* oi = &o_ids[((size_t)j * nl + l) * m];
                for (uint32_t i = 0; i < cnt; ++i) {
                    uint32_t row = h.row(oi[i], l);
                    const char* a = (const char*)&s.data[(size_t)row * s.S];
                    const char* w = (const char*)&s.w[(size_t)row * s.S];
                    for (uint32_t b = 0; b < s.S * 4; b += 64) {
                        __builtin_prefetch(a + b, 1);
                        __builtin_prefetch(w + b, 1);
                    }
                    __builtin_prefetch(&s.deg[row], 1);
                }
            }
        };
        constexpr uint32_t COMMIT_AHEAD = 4;
        for (uint32_t j = 0; j < std::min(COMMIT_AHEAD, nb); ++j) touch_rows(j);
        for (uint32_t j = 0; j < nb; ++j) {
            if (j + COMMIT_AHEAD < nb) touch_rows(j + COMMIT_AHEAD);
            uint32_t pid = order[pos + j];
            res.clear();
            for (uint32_t l = 0; l < nl; ++l) {
                uint32_t cnt = o_cnt[(size_t)j * nl + l];
                if (cnt == 0) continue;
                res.push_back(LayerSel{l, cnt, &o_ids[((size_t)j * nl + l) * m], &o_d[((size_t)j * nl + l) * m]});
            }
            const char* cerr = nullptr;
            if (commit_point(h, pid, res, d0, du, cs, &cerr)) return (set_error(cerr), HNSWB200_ESTATE);
        }
        double t2 = now();
        rc = G->upload_rows(d0, du);
        if (rc) return rc;
        pos += nb;
        linked += nb;
        double t3 = now();
        t_kernel += t1 - t0; t_commit += t2 - t1; t_upload += t3 - t2;
    }
    if (prof)
        fprintf(stderr, "[hnswb200 build] points=%zu batches=%llu kernel+copy=%.3fs commit=%.3fs upload=%.3fs (stage %.3fs = dedupe %.3f + alloc %.3f + gather %.3f + wide rows; %llu rows, %llu wide; transfer %.3fs, %llu full uploads) smem/block=%zu grid_cap=%d\n",
                order.size(), (unsigned long long)n_batches, t_kernel, t_commit, t_upload, G->t_stage, G->t_up[0], G->t_up[1], G->t_up[2], (unsigned long long)G->n_up_rows, (unsigned long long)G->n_up_big, G->t_xfer, (unsigned long long)G->n_full_uploads, smem, grid_cap);
    return 0;
}

}  // namespace hb

using namespace hb;

extern "C" {

// store_points (template.rs:269-293): quantise, assign levels, add to layers, pick the entry point
static int store_points(hnswb200_ctx* c, hnswb200_index* ix, const float* rows, uint64_t n, uint32_t dim,
                        const uint8_t* levels, std::vector<uint32_t>& ids) {
    if (n == 0) return (set_error("insert: no vectors given"), HNSWB200_EINVAL);
    if (dim != ix->params.dim) {  // check_points_dim (template.rs:253-262): the reference panics
        set_error("The current index dimension is " + std::to_string(ix->params.dim) +
                  ", but tried inserting points of dimension " + std::to_string(dim));
        return HNSWB200_EINVAL;
    }
    int vrc = build_validate(ix, ix->graph->h.n_points() + n);
    if (vrc) return vrc;
    std::vector<uint8_t> lv(n);
    if (levels) memcpy(lv.data(), levels, n);
    else draw_levels(ix->params.m, n, lv.data());  // re-seeded for every batch (points.rs:40)
    // append to the device-resident SimplePoints
    int rc = points_append_f32(c, ix->points, rows, n, lv.data());
    if (rc) return rc;
    HostGraph& h = ix->graph->h;
    for (uint64_t i = 0; i < n; ++i) ids.push_back(h.add_node(lv[i]));
    // ep = first key of the top layer's map (template.rs:283-290); convention: smallest id
    uint32_t top = h.n_layers() - 1;
    for (uint64_t i = 0; i < h.n_points(); ++i)
        if (h.level[i] == top) { ix->params.ep = (uint32_t)i; break; }
    return 0;
}

int hnswb200_index_insert_bulk(hnswb200_ctx* c, hnswb200_index* ix, const float* rows, uint64_t n,
                               uint32_t dim, const uint8_t* levels, uint32_t batch) {
    if (!c || !ix || !rows) return (set_error("insert_bulk: NULL argument"), HNSWB200_EINVAL);
    if (c->use()) return HNSWB200_ECUDA;
    std::vector<uint32_t> ids;
    int rc = store_points(c, ix, rows, n, dim, levels, ids);
    if (rc) return rc;
    return build_insert(c, ix, ids, batch);
}

int hnswb200_index_insert_vec(hnswb200_ctx* c, hnswb200_index* ix, const float* row, uint32_t dim,
                              uint32_t* id_out) {
    if (!c || !ix || !row) return (set_error("insert_vec: NULL argument"), HNSWB200_EINVAL);
    if (c->use()) return HNSWB200_ECUDA;
    std::vector<uint32_t> ids;
    int rc = store_points(c, ix, row, 1, dim, nullptr, ids);
    if (rc) return rc;
    if (id_out) *id_out = ids[0];
    return build_insert(c, ix, ids, 1);
}

int hnswb200_build(hnswb200_ctx* c, const float* rows, uint64_t n, uint32_t dim, const hnswb200_params* params,
                   const uint8_t* levels, uint32_t batch, hnswb200_index** out) {
    if (!c || !params || !out || (n && !rows)) return (set_error("build: NULL argument"), HNSWB200_EINVAL);
    if (params->dim != dim) return (set_error("build: params.dim != dim"), HNSWB200_EINVAL);
    if (params->m < 2) return (set_error("build: m must be >= 2"), HNSWB200_EINVAL);
    if (c->use()) return HNSWB200_ECUDA;
    hnswb200_points* P = nullptr;
    int rc = hnswb200_points_from_f32(c, rows, 0, dim, nullptr, &P);  // HNSW::new: empty index
    if (rc) return rc;
    hnswb200_graph* G = new hnswb200_graph();
    G->ctx = c;
    G->h.init((uint32_t)params->m, (uint32_t)(2 * params->m), (uint32_t)params->m);
    hnswb200_index* ix = new hnswb200_index();
    ix->ctx = c;
    ix->points = P;
    ix->graph = G;
    ix->params = *params;
    ix->params.ep = 0;
    if (n) {
        rc = hnswb200_index_insert_bulk(c, ix, rows, n, dim, levels, batch);
        if (rc) { hnswb200_index_destroy(ix); return rc; }
    }
    *out = ix;
    return 0;
}

}  // extern "C"
