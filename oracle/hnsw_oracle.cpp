// oracle/hnsw_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (C++17) of the hot path of Gumo-A/hnsw_rs, written from a
// reading of the reference's Rust sources.  It exists only so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// can check and time the CUDA path against the reference's algorithm.  Nothing
// under hnsw_rs_b200/ may include, link or call this file.
//
// Why a restatement and not the reference itself: the reference is a Rust
// workspace and there is no cargo/rustc in this image or on the GPU box, so
// oracle/_ref cannot be built.  Pinning: the reference's own known-answer
// tests and fixture test are reproduced in tests/test_oracle_pins.py
//   * vectors/src/quant.rs:143-202  (quantised distance KATs, exact equality)
//   * vectors/src/full.rs:88-147    (f32 distance KATs)
//   * vectors/tests/full_lvq_tests.rs:3-27 (quantisation error < 1 %)
//   * hnsw/src/template/results.rs:223-231 + graph/src/dist.rs:30-37 (tie rule)
//   * graph/src/graph.rs:299-486    (graph invariants / replace_neighbors)
//   * hnsw/src/template.rs:518-572  (test-data recall@10 > 0.99, min degree > 0)
//   * hnsw/src/template.rs:574-611  (save/load round trip)
// UNPINNED (no reference test or golden vector fixes them; documented in
// DESIGN.md): the rand 0.8.5 StdRng level sequence (restated below from the
// published ChaCha12 / PCG32 seed-expansion algorithm, third-party crates
// rand 0.8.5, rand_chacha 0.3.1, rand_core 0.6.4 per Cargo.lock:685-706), and
// every place where the reference iterates a hashbrown map (entry point choice,
// insertion order inside a level class, order of prune application).  The
// oracle's convention there is ascending node id.
//
// Float contract: every f32 operation is a separately rounded
// IEEE-754 binary32 op.  Build with -ffp-contract=off and no -ffast-math.

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <map>
#include <set>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;

// ---------------------------------------------------------------------------
// vectors crate
// ---------------------------------------------------------------------------

// vectors/src/quant.rs:6-11
struct QuantVec {
    float delta = 0.f;
    float min = 0.f;
    std::vector<uint8_t> codes;
    // vectors/src/full.rs:3-6: when the index was created with VecType = FullVec (points/src/point.rs:4 flipped) a
    // point carries its f32 values instead of the three fields above
    std::vector<float> full;
};

// Rust `f32 as u8`: saturating, NaN -> 0.
inline uint8_t sat_u8(float x) {
    if (!(x == x)) return 0;
    if (x <= 0.f) return 0;
    if (x >= 255.f) return 255;
    return (uint8_t)x;
}

// vectors/src/quant.rs:41-66 (QuantVec::new)
// max_by(partial_cmp) keeps the LAST of equal maxima, min_by the FIRST of equal
// minima (only observable through the sign of a zero bound).
bool quantise(const float* v, size_t d, QuantVec& out) {
    if (d == 0) { g_err = "cannot quantise an empty vector"; return false; }
    float ub = v[0], lb = v[0];
    for (size_t i = 0; i < d; ++i)
        if (v[i] != v[i]) { g_err = "NaN in vector"; return false; }
    for (size_t i = 1; i < d; ++i) {
        if (!(ub > v[i])) ub = v[i];   // Ordering != Greater -> take the later one
        if (v[i] < lb) lb = v[i];      // strictly less -> keep the first
    }
    float range = ub - lb;
    float delta = range / 255.0f;      // 2^8 - 1
    out.delta = delta;
    out.min = lb;
    out.codes.resize(d);
    for (size_t i = 0; i < d; ++i) {
        float t = v[i] - lb;
        float b = t / delta;
        float c = b + 0.5f;
        out.codes[i] = sat_u8(std::floor(c));
    }
    return true;
}

inline float deq(uint8_t c, float delta, float mn) {
    float p = (float)c * delta;
    return p + mn;
}

// vectors/src/quant.rs:14-37 (distance_unrolled == dist2other)
float dist_unrolled(const uint8_t* xc, float xd, float xm, const uint8_t* yc, float yd,
                    float ym, size_t d) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t nfull = d / 8;
    for (size_t k = 0; k < nfull; ++k) {
        for (size_t j = 0; j < 8; ++j) {
            float x = deq(xc[8 * k + j], xd, xm);
            float y = deq(yc[8 * k + j], yd, ym);
            float t = x - y;
            float sq = t * t;
            acc[j] = acc[j] + sq;
        }
    }
    for (size_t i = 8 * nfull; i < d; ++i) {
        float x = deq(xc[i], xd, xm);
        float y = deq(yc[i], yd, ym);
        float t = x - y;
        float sq = t * t;
        acc[0] = acc[0] + sq;
    }
    float s = 0.f;
    for (int j = 0; j < 8; ++j) s = s + acc[j];
    return std::sqrt(s);
}

// vectors/src/full.rs:23-29 and quant.rs:67-73 (generic `distance`): one
// strictly sequential chain over zip(iter_vals) then sqrt.
float dist_sequential(const float* x, const float* y, size_t d) {
    float s = 0.f;
    for (size_t i = 0; i < d; ++i) {
        float t = x[i] - y[i];
        float sq = t * t;
        s = s + sq;
    }
    return std::sqrt(s);
}

// ---------------------------------------------------------------------------
// graph crate
// ---------------------------------------------------------------------------

// graph/src/dist.rs:4-37
struct Dist {
    uint32_t id;
    float dist;
};
inline bool operator<(const Dist& a, const Dist& b) {
    if (a.dist < b.dist) return true;
    if (a.dist > b.dist) return false;
    return a.id < b.id;
}
inline bool operator>(const Dist& a, const Dist& b) { return b < a; }

// graph/src/graph.rs:11-163.  Node map = dense row index; neighbour set kept as
// an insertion-ordered vector with set semantics (results never depend on the
// iteration order).
struct Graph {
    size_t level = 0;
    size_t m = 0;
    std::vector<int32_t> row_of;            // node id -> row, -1 if absent
    std::vector<uint32_t> node_ids;         // row -> node id
    std::vector<std::vector<uint32_t>> nb;  // row -> neighbour set

    bool contains(uint32_t id) const { return id < row_of.size() && row_of[id] >= 0; }
    void add_node(uint32_t id) {  // graph.rs:31-35
        if (id >= row_of.size()) row_of.resize((size_t)id + 1, -1);
        if (row_of[id] >= 0) return;
        row_of[id] = (int32_t)node_ids.size();
        node_ids.push_back(id);
        nb.emplace_back();
    }
    std::vector<uint32_t>* nbrs(uint32_t id) {
        return contains(id) ? &nb[row_of[id]] : nullptr;
    }
    const std::vector<uint32_t>* nbrs(uint32_t id) const {
        return contains(id) ? &nb[row_of[id]] : nullptr;
    }
    static void set_insert(std::vector<uint32_t>& s, uint32_t v) {
        if (std::find(s.begin(), s.end(), v) == s.end()) s.push_back(v);
    }
    static void set_remove(std::vector<uint32_t>& s, uint32_t v) {
        auto it = std::find(s.begin(), s.end(), v);
        if (it != s.end()) s.erase(it);
    }
    // graph.rs:37-52; returns 0 ok, 1 self connection, 2 node missing
    int add_edge(uint32_t a, uint32_t b) {
        if (a == b) return 1;
        auto* na = nbrs(a);
        auto* nbb = nbrs(b);
        if (!na || !nbb) return 2;
        set_insert(*na, b);
        set_insert(*nbb, a);
        return 0;
    }
    int remove_edge(uint32_t a, uint32_t b) {  // graph.rs:72-83
        auto* na = nbrs(a);
        auto* nbb = nbrs(b);
        if (!na || !nbb) return 2;
        set_remove(*na, b);
        set_remove(*nbb, a);
        return 0;
    }
    size_t degree(uint32_t id) const { return nbrs(id)->size(); }
    // graph.rs:85-94: never cut the edge to a neighbour whose degree is 1
    void isolate_node(uint32_t node) {
        std::vector<uint32_t> snap = *nbrs(node);
        for (uint32_t n : snap) {
            if (degree(n) == 1) continue;
            remove_edge(node, n);
        }
    }
    int add_neighbors(uint32_t node, const std::vector<uint32_t>& nn) {  // graph.rs:139-148
        for (uint32_t n : nn) {
            int r = add_edge(node, n);
            if (r) return r;
        }
        return 0;
    }
    int replace_neighbors(uint32_t node, const std::vector<uint32_t>& nn) {  // graph.rs:128-137
        if (!contains(node)) return 2;
        isolate_node(node);
        return add_neighbors(node, nn);
    }
    size_t nb_nodes() const { return node_ids.size(); }
};

// graph/src/layers.rs:7-71
struct Layers {
    std::vector<Graph> levels;
    size_t m = 0;
    void add_level(size_t level) {  // layers.rs:48-60: layer 0 cap 2m, others m
        while (levels.size() <= level) {
            Graph g;
            g.level = levels.size();
            g.m = levels.empty() ? m * 2 : m;
            levels.push_back(std::move(g));
        }
    }
    void add_node(uint32_t id, size_t level) {  // layers.rs:62-70
        add_level(level);
        for (size_t l = 0; l <= level; ++l) levels[l].add_node(id);
    }
};

// ---------------------------------------------------------------------------
// rand 0.8.5 StdRng (ChaCha12) restatement -- UNPINNED, see header.
// ---------------------------------------------------------------------------
struct ChaChaRng {
    uint32_t key[8];
    uint64_t counter = 0;
    uint32_t buf[64];
    int idx = 64;
    int rounds = 12;

    static inline uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
    static void block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) {
        uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                          key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                          (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
        uint32_t x[16];
        memcpy(x, s, sizeof(x));
#define QR(a, b, c, d)                                   \
    x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl(x[d], 16);   \
    x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl(x[b], 12);   \
    x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl(x[d], 8);    \
    x[c] += x[d]; x[b] ^= x[c]; x[b] = rotl(x[b], 7);
        for (int r = 0; r < rounds; r += 2) {
            QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
            QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
        }
#undef QR
        for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
    }
    // rand_core 0.6 SeedableRng::seed_from_u64: PCG32 expansion of the u64
    void seed_from_u64(uint64_t state) {
        const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
        for (int i = 0; i < 8; ++i) {
            state = state * MUL + INC;
            uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
            uint32_t rot = (uint32_t)(state >> 59);
            key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        }
        counter = 0;
        idx = 64;
    }
    uint32_t next_u32() {
        if (idx >= 64) {  // rand_chacha refills 4 blocks at a time
            for (int b = 0; b < 4; ++b) block(key, counter + b, rounds, buf + 16 * b);
            counter += 4;
            idx = 0;
        }
        return buf[idx++];
    }
    // rand 0.8 Standard for f32: 24 random bits scaled into [0,1)
    float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
};

// points/src/points.rs:148-160
size_t new_layer(float ml, ChaChaRng& rng) {
    float r = 0.0f;
    while (r == 0.0f || r == 1.0f) r = rng.gen_f32();
    float l = -std::log(r) * ml;
    return (size_t)std::floor(l);
}

// ---------------------------------------------------------------------------
// points crate (SimplePoints with VecType = QuantVec), flat storage
// ---------------------------------------------------------------------------
struct Points {
    size_t dim = 0;
    bool full = false;           // VecType = FullVec: `vals` is the store, FullVec::distance the metric
    std::vector<uint8_t> codes;  // n * dim
    std::vector<float> mins, deltas;
    std::vector<float> vals;     // n * dim (full mode)
    std::vector<uint8_t> levels;
    size_t len() const { return levels.size(); }
    // Point::new (points/src/point.rs:24-30): VecType::new(vector)
    bool make_point(const float* v, size_t d, QuantVec& out) const;
    void push(const QuantVec& q, uint8_t level) {
        if (full) {
            vals.insert(vals.end(), q.full.begin(), q.full.end());
        } else {
            codes.insert(codes.end(), q.codes.begin(), q.codes.end());
            mins.push_back(q.min);
            deltas.push_back(q.delta);
        }
        levels.push_back(level);
    }
    // the stored point `id` as a value (get_point(id).clone())
    void get(uint32_t id, QuantVec& out) const {
        if (full) {
            out.full.assign(&vals[(size_t)id * dim], &vals[(size_t)(id + 1) * dim]);
        } else {
            out.delta = deltas[id];
            out.min = mins[id];
            out.codes.assign(&codes[(size_t)id * dim], &codes[(size_t)(id + 1) * dim]);
        }
    }
    // points.rs:86-93  distance(a,b) = a.dist2other(b)
    float distance(uint32_t a, uint32_t b) const {
        if (full) return dist_sequential(&vals[(size_t)a * dim], &vals[(size_t)b * dim], dim);
        return dist_unrolled(&codes[(size_t)a * dim], deltas[a], mins[a],
                             &codes[(size_t)b * dim], deltas[b], mins[b], dim);
    }
    // points.rs:95-101  distance2point(point, idx) = point.dist2other(points[idx])
    float distance2point(const QuantVec& p, uint32_t b) const {
        if (full) return dist_sequential(p.full.data(), &vals[(size_t)b * dim], dim);
        return dist_unrolled(p.codes.data(), p.delta, p.min, &codes[(size_t)b * dim], deltas[b],
                             mins[b], dim);
    }
    // point.rs:35-37 via searcher.rs:66-69: index.get_point(node).dist2other(point)
    float node2point(uint32_t a, const QuantVec& p) const {
        if (full) return dist_sequential(&vals[(size_t)a * dim], p.full.data(), dim);
        return dist_unrolled(&codes[(size_t)a * dim], deltas[a], mins[a], p.codes.data(), p.delta,
                             p.min, dim);
    }
};

bool Points::make_point(const float* v, size_t d, QuantVec& out) const {
    if (!full) return quantise(v, d, out);
    // FullVec::new clones the vector (full.rs:18-22).  A non-finite value can make a NaN distance, on which the
    // reference panics (graph/src/dist.rs:32): refused here, like a NaN in quantise().
    for (size_t i = 0; i < d; ++i)
        if (!std::isfinite(v[i])) { g_err = "non-finite value in vector"; return false; }
    out.full.assign(v, v + d);
    return true;
}

// ---------------------------------------------------------------------------
// hnsw crate
// ---------------------------------------------------------------------------

// hnsw/src/params.rs:4-62
struct Params {
    uint32_t ep = 0;
    size_t m = 0, mmax = 0, mmax0 = 0;
    float ml = 0.f;
    size_t ef_cons = 0, dim = 0;
};
inline float default_ml(size_t m) { return 1.0f / std::log((float)m); }  // params.rs:15-17

typedef std::set<Dist> OrderedDists;

// hnsw/src/template/results.rs:26-33
struct Results {
    OrderedDists selected, candidates, visited_h;
    // visited: epoch-stamped dense array (same membership semantics as IntSet)
    std::vector<uint32_t> stamp;
    uint32_t epoch = 1;
    std::map<size_t, std::map<uint32_t, OrderedDists>> insertion_results, prune_results;
    uint64_t hops = 0, evals = 0;

    void ensure(size_t n) {
        if (stamp.size() < n) stamp.resize(n, 0);
    }
    bool insert_visited(uint32_t id) {  // results.rs:101-103
        if (stamp[id] == epoch) return false;
        stamp[id] = epoch;
        return true;
    }
    void clear_visited() {  // results.rs:178-180
        if (++epoch == 0) {
            std::fill(stamp.begin(), stamp.end(), 0);
            epoch = 1;
        }
    }
    void clear_all() {  // results.rs:182-190
        selected.clear();
        candidates.clear();
        clear_visited();
        visited_h.clear();
        insertion_results.clear();
        prune_results.clear();
    }
};

struct Index {
    Params params;
    Layers layers;
    Points points;
};

// hnsw/src/template/searcher.rs:23-103
bool search_layer(Results& r, const Graph& layer, const QuantVec& point, const Index& index,
                  size_t ef) {
    r.ensure(index.points.len());
    for (const Dist& d : r.selected) r.candidates.insert(d);   // results.rs:148-157
    for (const Dist& d : r.selected) r.insert_visited(d.id);   // results.rs:159-168
    std::vector<Dist> fresh;
    while (!r.candidates.empty()) {
        Dist cand = *r.candidates.begin();
        r.candidates.erase(r.candidates.begin());
        Dist furthest = *r.selected.rbegin();
        if (cand > furthest) break;
        const std::vector<uint32_t>* neigh = layer.nbrs(cand.id);
        if (!neigh) {
            g_err = "Error in search_layer: " + std::to_string(cand.id) + " not in Graph";
            return false;
        }
        r.hops++;
        fresh.clear();
        for (uint32_t node : *neigh) {
            if (!r.insert_visited(node)) continue;
            fresh.push_back(Dist{node, index.points.node2point(node, point)});
        }
        r.evals += fresh.size();
        for (const Dist& t : fresh) {
            Dist f = *r.selected.rbegin();
            if (r.selected.size() < ef) {
                r.selected.insert(t);
                r.candidates.insert(t);
                continue;
            }
            if (t < f) {
                r.selected.insert(t);
                r.candidates.insert(t);
                if (r.selected.size() > ef) r.selected.erase(std::prev(r.selected.end()));
            }
        }
    }
    r.candidates.clear();
    r.clear_visited();
    return true;
}

// hnsw/src/template/searcher.rs:109-153 with results.rs:105-146,69-77
bool select_heuristic(Results& r, const Graph& layer, uint32_t point_id, const Points& points,
                      size_t m, bool extend_cands, bool keep_pruned) {
    // select_setup, results.rs:105-111
    r.visited_h.clear();
    r.candidates.clear();
    r.candidates.insert(r.selected.begin(), r.selected.end());
    r.selected.clear();
    if (extend_cands) {  // results.rs:122-146 (no dedup before evaluation)
        std::vector<uint32_t> neighbors;
        for (const Dist& node : r.candidates) {
            const std::vector<uint32_t>* nn = layer.nbrs(node.id);
            if (!nn) { g_err = "Node is not in the Graph"; return false; }
            neighbors.insert(neighbors.end(), nn->begin(), nn->end());
        }
        for (uint32_t n : neighbors) {
            r.candidates.insert(Dist{n, points.distance(point_id, n)});
            r.evals++;
        }
    }
    if (r.candidates.empty()) { g_err = "select_heuristic: no candidates"; return false; }
    Dist e = *r.candidates.begin();
    r.candidates.erase(r.candidates.begin());
    r.selected.insert(e);
    while (!r.candidates.empty() && r.selected.size() < m) {
        e = *r.candidates.begin();
        r.candidates.erase(r.candidates.begin());
        // get_nearest_from_selected, results.rs:69-77: min over s of Dist(s.id, d(e, s))
        bool have = false;
        Dist nearest{0, 0.f};
        for (const Dist& s : r.selected) {
            Dist c{s.id, points.distance(e.id, s.id)};
            r.evals++;
            if (!have || c < nearest) { nearest = c; have = true; }
        }
        if (e < nearest) r.selected.insert(e);
        else if (keep_pruned) r.visited_h.insert(e);
    }
    if (keep_pruned) {
        while (!r.visited_h.empty() && r.selected.size() < m) {
            r.selected.insert(*r.visited_h.begin());
            r.visited_h.erase(r.visited_h.begin());
        }
    }
    return true;
}

// hnsw/src/template/inserter.rs:40-126
bool build_insertion_results(Results& r, Index& ix, uint32_t id) {
    if (id == ix.params.ep) return true;  // inserter.rs:42-45 (results NOT cleared)
    r.clear_all();                        // setup_insert, inserter.rs:53-68
    r.ensure(ix.points.len());
    QuantVec point;
    ix.points.get(id, point);
    size_t level = ix.points.levels[id];
    r.selected.insert(Dist{ix.params.ep, ix.points.distance(ix.params.ep, id)});
    r.evals++;
    size_t L = ix.layers.levels.size();
    for (size_t l = L; l-- > level + 1;)  // traverse_layers_above
        if (!search_layer(r, ix.layers.levels[l], point, ix, 1)) return false;
    size_t bound = std::min(level, L - 1);
    for (size_t l = bound + 1; l-- > 0;) {  // traverse_layers_below
        if (!search_layer(r, ix.layers.levels[l], point, ix, ix.params.ef_cons)) return false;
        if (!select_heuristic(r, ix.layers.levels[l], id, ix.points, ix.params.m, true, true))
            return false;
        r.insertion_results[l][id] = r.selected;  // save_layer_results, results.rs:79-84
    }
    return true;
}

// hnsw/src/template.rs:177-251 + select_simple template.rs:614-621
bool insert(Index& ix, uint32_t id, Results& r) {
    if (!build_insertion_results(r, ix, id)) return false;
    // make_connections, template.rs:196-207
    for (auto& lr : r.insertion_results) {
        Graph& layer = ix.layers.levels[lr.first];
        for (auto& nd : lr.second) {
            std::vector<uint32_t> ids;
            for (const Dist& d : nd.second) ids.push_back(d.id);
            int rc = layer.add_neighbors(nd.first, ids);
            if (rc) { g_err = "make_connections: add_edge failed"; return false; }
        }
    }
    // prune_connections, template.rs:209-238
    r.prune_results.clear();
    for (auto& lr : r.insertion_results) {
        Graph& layer = ix.layers.levels[lr.first];
        for (auto& nd : lr.second) {
            for (const Dist& x : nd.second) {
                if (!(layer.degree(x.id) > layer.m)) continue;
                std::vector<Dist> cands;
                for (uint32_t n : *layer.nbrs(x.id)) {
                    cands.push_back(Dist{n, ix.points.distance(x.id, n)});
                    r.evals++;
                }
                std::sort(cands.begin(), cands.end());
                OrderedDists nearest;
                for (size_t i = 0; i < cands.size() && i < layer.m; ++i) nearest.insert(cands[i]);
                r.prune_results[lr.first][x.id] = nearest;
            }
        }
    }
    // make_pruned_connections, template.rs:240-251
    for (auto& lr : r.prune_results) {
        Graph& layer = ix.layers.levels[lr.first];
        for (auto& nd : lr.second) {
            std::vector<uint32_t> ids;
            for (const Dist& d : nd.second) ids.push_back(d.id);
            int rc = layer.replace_neighbors(nd.first, ids);
            if (rc) { g_err = "make_pruned_connections: replace_neighbors failed"; return false; }
        }
    }
    return true;
}

// hnsw/src/template.rs:269-293 (store_points) with points.rs:39-48
bool store_points(Index& ix, const float* rows, size_t n, size_t dim,
                  std::vector<uint32_t>& ids, const uint8_t* forced_levels) {
    if (n == 0) { g_err = "store_points: no vectors"; return false; }
    if (dim != ix.params.dim) {  // template.rs:253-262 (reference panics)
        g_err = "The current index dimension is " + std::to_string(ix.params.dim) +
                ", but tried inserting points of dimension " + std::to_string(dim);
        return false;
    }
    ChaChaRng rng;
    rng.seed_from_u64(0);  // points.rs:40: re-seeded for every batch
    float ml = default_ml(ix.params.m);
    ix.points.dim = dim;
    QuantVec q;
    for (size_t i = 0; i < n; ++i) {
        size_t level = forced_levels ? forced_levels[i] : new_layer(ml, rng);
        if (!ix.points.make_point(rows + i * dim, dim, q)) return false;
        uint32_t id = (uint32_t)ix.points.len();
        ix.points.push(q, (uint8_t)level);
        ids.push_back(id);
    }
    for (uint32_t id : ids) ix.layers.add_node(id, ix.points.levels[id]);
    // template.rs:283-290: ep = first key of the top layer's map (oracle: smallest id)
    const Graph& top = ix.layers.levels.back();
    ix.params.ep = *std::min_element(top.node_ids.begin(), top.node_ids.end());
    return true;
}

// hnsw/src/template.rs:388-444, nb_threads = 1 semantics (ascending id per level class)
bool insert_bulk(Index& ix, const float* rows, size_t n, size_t dim, const uint8_t* forced_levels,
                 uint64_t* evals_out) {
    std::vector<uint32_t> ids;
    if (!store_points(ix, rows, n, dim, ids, forced_levels)) return false;
    std::vector<uint8_t> stored(ix.points.len(), 0);
    for (uint32_t id : ids) stored[id] = 1;
    Results r;
    uint64_t evals = 0;
    for (size_t l = ix.layers.levels.size(); l-- > 0;) {
        std::vector<uint32_t> todo;
        for (uint32_t id : ix.layers.levels[l].node_ids)
            if (stored[id] && ix.points.levels[id] == (uint8_t)l) todo.push_back(id);
        std::sort(todo.begin(), todo.end());
        for (uint32_t id : todo) {
            r.evals = 0;
            if (!insert(ix, id, r)) return false;
            evals += r.evals;
        }
    }
    if (evals_out) *evals_out = evals;
    return true;
}

// hnsw/src/template.rs:306-335
bool ann_by_vector(const Index& ix, const float* v, size_t n, size_t ef, Results& r,
                   std::vector<Dist>& out) {
    QuantVec point;
    if (!ix.points.make_point(v, ix.params.dim, point)) return false;
    r.selected.clear();
    r.candidates.clear();
    r.hops = 0;
    r.evals = 1;
    r.ensure(ix.points.len());
    r.selected.insert(Dist{ix.params.ep, ix.points.distance2point(point, ix.params.ep)});
    size_t L = ix.layers.levels.size();
    for (size_t l = L; l-- > 1;)
        if (!search_layer(r, ix.layers.levels[l], point, ix, 1)) return false;
    if (!search_layer(r, ix.layers.levels[0], point, ix, ef)) return false;
    out.clear();
    for (const Dist& d : r.selected) {  // get_top_selected, results.rs:59-61
        if (out.size() >= n) break;
        out.push_back(d);
    }
    return true;
}

// hnsw/src/helpers/glove.rs:73-109 and template.rs:531-541: all N distances,
// full sort under Dist order, take k.  (partial_sort is order-equivalent.)
void brute_force_one(const Index& ix, const QuantVec& q, size_t k, std::vector<Dist>& out) {
    size_t n = ix.points.len();
    std::vector<Dist> d(n);
    for (size_t i = 0; i < n; ++i) d[i] = Dist{(uint32_t)i, ix.points.distance2point(q, (uint32_t)i)};
    size_t kk = std::min(k, n);
    std::partial_sort(d.begin(), d.begin() + kk, d.end());
    out.assign(d.begin(), d.begin() + kk);
}

// ---------------------------------------------------------------------------
// byte formats (all big-endian)
// ---------------------------------------------------------------------------
void put_u64(std::vector<uint8_t>& b, uint64_t v) { for (int i = 7; i >= 0; --i) b.push_back((uint8_t)(v >> (8 * i))); }
void put_u32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 3; i >= 0; --i) b.push_back((uint8_t)(v >> (8 * i))); }
void put_u16(std::vector<uint8_t>& b, uint16_t v) { b.push_back((uint8_t)(v >> 8)); b.push_back((uint8_t)v); }
void put_f32(std::vector<uint8_t>& b, float f) { uint32_t u; memcpy(&u, &f, 4); put_u32(b, u); }
uint64_t get_u64(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 8; ++i) v = (v << 8) | p[i]; return v; }
uint32_t get_u32(const uint8_t* p) { uint32_t v = 0; for (int i = 0; i < 4; ++i) v = (v << 8) | p[i]; return v; }
float get_f32(const uint8_t* p) { uint32_t u = get_u32(p); float f; memcpy(&f, &u, 4); return f; }

bool write_file(const std::string& path, const std::vector<uint8_t>& b) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { g_err = "Could not create " + path; return false; }
    size_t w = b.empty() ? 0 : fwrite(b.data(), 1, b.size(), f);
    fclose(f);
    if (w != b.size()) { g_err = "short write " + path; return false; }
    return true;
}
bool read_file(const std::string& path, std::vector<uint8_t>& b) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { g_err = "Problem reading " + path; return false; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    b.resize((size_t)sz);
    size_t r = sz ? fread(b.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    return r == (size_t)sz;
}

// template.rs:43-73; params.rs:78-91; points.rs:119-131; graph.rs:165-219
bool save_dir(const Index& ix, const std::string& dir) {
    mkdir(dir.c_str(), 0777);
    std::vector<uint8_t> b;
    const Points& P = ix.points;
    put_u64(b, P.len());
    put_u64(b, P.full ? 1 + 4 * P.dim : 9 + P.dim);  // 1 + VecType::size(): full.rs:45-47 / quant.rs:91-93
    for (size_t i = 0; i < P.len(); ++i) {
        b.push_back(P.levels[i]);
        if (P.full) {
            for (size_t j = 0; j < P.dim; ++j) put_f32(b, P.vals[i * P.dim + j]);  // full.rs:55-61
            continue;
        }
        put_f32(b, P.mins[i]);
        put_f32(b, P.deltas[i]);
        b.insert(b.end(), &P.codes[i * P.dim], &P.codes[(i + 1) * P.dim]);
    }
    if (!write_file(dir + "/points", b)) return false;
    b.clear();
    put_u64(b, ix.params.m); put_u64(b, ix.params.mmax); put_u64(b, ix.params.mmax0);
    put_f32(b, ix.params.ml);
    put_u64(b, ix.params.ef_cons); put_u64(b, ix.params.dim); put_u64(b, ix.params.ep);
    if (!write_file(dir + "/params", b)) return false;
    if (mkdir((dir + "/layers").c_str(), 0777) != 0 && errno != EEXIST) {
        g_err = "Could not create layers dir";
        return false;
    }
    for (size_t l = 0; l < ix.layers.levels.size(); ++l) {
        const Graph& g = ix.layers.levels[l];
        b.clear();
        // Format hazard: the reference pads rows to `m` words but never
        // truncates, so a node whose degree exceeds the cap (possible, App. C-6) would
        // misalign every later row.  We keep the file self-consistent instead: the row
        // width written in the header is max(cap, max degree); readers that follow
        // graph.rs:226-252 load it correctly.
        size_t width = g.m;
        for (auto& s : g.nb) width = std::max(width, s.size());
        b.push_back((uint8_t)g.level);
        put_u32(b, (uint32_t)g.nb_nodes());
        put_u16(b, (uint16_t)width);
        for (size_t r = 0; r < g.nb_nodes(); ++r) {
            put_u32(b, g.node_ids[r]);
            size_t c = 0;
            for (uint32_t n : g.nb[r]) { put_u32(b, n); ++c; }
            for (; c < width; ++c) put_u32(b, 0xFFFFFFFFu);
        }
        if (!write_file(dir + "/layers/" + std::to_string(l), b)) return false;
    }
    return true;
}

// template.rs:75-131; params.rs:93-114; points.rs:133-145; graph.rs:226-252
bool load_dir(Index& ix, const std::string& dir) {
    struct stat st;
    if (stat(dir.c_str(), &st) != 0) { g_err = dir + " does not exist"; return false; }
    std::vector<uint8_t> b;
    if (!read_file(dir + "/points", b) || b.size() < 16) { g_err = "Problem reading points file"; return false; }
    size_t len = get_u64(&b[0]), psz = get_u64(&b[8]);
    if (psz < 5 || b.size() < 16 + len * psz) { g_err = "points file truncated"; return false; }
    std::vector<uint8_t> pb;
    pb.swap(b);
    if (!read_file(dir + "/params", b) || b.size() < 52) { g_err = "Problem reading params file"; return false; }
    // the VecType the file was written with is not recorded; the point size tells: 1 + 4*dim (FullVec) or 9 + dim
    const size_t pdim = get_u64(&b[36]);
    Points& P = ix.points;
    P = Points();
    P.dim = pdim;
    P.full = psz == 1 + 4 * pdim && psz != 9 + pdim;
    if (!P.full && psz != 9 + pdim) { g_err = "params.dim does not match the point size"; return false; }
    for (size_t i = 0; i < len; ++i) {
        const uint8_t* p = &pb[16 + i * psz];
        P.levels.push_back(p[0]);
        if (P.full) {
            for (size_t j = 0; j < pdim; ++j) P.vals.push_back(get_f32(p + 1 + 4 * j));
            continue;
        }
        P.mins.push_back(get_f32(p + 1));
        P.deltas.push_back(get_f32(p + 5));
        P.codes.insert(P.codes.end(), p + 9, p + psz);
    }
    ix.params.m = get_u64(&b[0]); ix.params.mmax = get_u64(&b[8]); ix.params.mmax0 = get_u64(&b[16]);
    ix.params.ml = get_f32(&b[24]);
    ix.params.ef_cons = get_u64(&b[28]); ix.params.dim = get_u64(&b[36]);
    ix.params.ep = (uint32_t)get_u64(&b[44]);
    ix.layers = Layers();
    ix.layers.m = ix.params.m;
    std::vector<size_t> idxs;
    DIR* d = opendir((dir + "/layers").c_str());
    if (!d) { g_err = "There was a problem reading layers"; return false; }
    while (dirent* e = readdir(d)) {
        if (e->d_name[0] == '.') continue;
        idxs.push_back((size_t)strtoull(e->d_name, nullptr, 10));
    }
    closedir(d);
    std::sort(idxs.begin(), idxs.end());
    for (size_t li : idxs) {
        if (!read_file(dir + "/layers/" + std::to_string(li), b) || b.size() < 7) { g_err = "Problem reading layer file"; return false; }
        Graph g;
        g.level = b[0];
        uint32_t nn = get_u32(&b[1]);
        size_t width = ((size_t)b[5] << 8) | b[6];
        g.m = width;
        if (b.size() != 7 + (size_t)nn * 4 * (width + 1)) {
            g_err = "layer file length does not match nb_nodes*(m+1) rows (over-full row written by the reference?)";
            return false;
        }
        if (g.level != ix.layers.levels.size()) { g_err = "layer level mismatch"; return false; }
        size_t off = 7;
        for (uint32_t r = 0; r < nn; ++r) {
            uint32_t node = get_u32(&b[off]);
            off += 4;
            g.add_node(node);
            std::vector<uint32_t>& s = g.nb[g.row_of[node]];
            for (size_t j = 0; j < width; ++j) {
                uint32_t v = get_u32(&b[off + 4 * j]);
                if (v == 0xFFFFFFFFu) break;
                Graph::set_insert(s, v);
            }
            off += 4 * width;
        }
        // cap as Layers::add_level would have made it (layers.rs:50); equals the header
        // value for every file the reference itself can write consistently
        g.m = std::min(width, g.level == 0 ? 2 * ix.params.m : ix.params.m);
        ix.layers.levels.push_back(std::move(g));
    }
    return true;
}

}  // namespace

// ---------------------------------------------------------------------------
// C API used by tests/ and bench.py through ctypes
// ---------------------------------------------------------------------------
extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }

int oracle_quantise(const float* v, uint64_t dim, uint8_t* codes, float* mn, float* delta) {
    QuantVec q;
    if (!quantise(v, dim, q)) return -1;
    memcpy(codes, q.codes.data(), dim);
    *mn = q.min;
    *delta = q.delta;
    return 0;
}

float oracle_dist_quant(const uint8_t* xc, float xd, float xm, const uint8_t* yc, float yd, float ym,
                        uint64_t dim) {
    return dist_unrolled(xc, xd, xm, yc, yd, ym, dim);
}

float oracle_dist_full(const float* x, const float* y, uint64_t dim) { return dist_sequential(x, y, dim); }

// quant.rs:67-73 generic distance between two quantised vectors (iter_vals zip)
float oracle_dist_quant_generic(const uint8_t* xc, float xd, float xm, const uint8_t* yc, float yd,
                                float ym, uint64_t dim) {
    std::vector<float> x(dim), y(dim);
    for (uint64_t i = 0; i < dim; ++i) { x[i] = deq(xc[i], xd, xm); y[i] = deq(yc[i], yd, ym); }
    return dist_sequential(x.data(), y.data(), dim);
}

void oracle_dequantise(const uint8_t* c, float delta, float mn, uint64_t dim, float* out) {
    for (uint64_t i = 0; i < dim; ++i) out[i] = deq(c[i], delta, mn);
}

// Dist::cmp: -1, 0, 1
int oracle_dist_cmp(uint32_t ida, float da, uint32_t idb, float db) {
    Dist a{ida, da}, b{idb, db};
    if (a < b) return -1;
    if (b < a) return 1;
    return 0;
}

void oracle_levels(uint64_t m, uint64_t n, uint8_t* out) {
    ChaChaRng rng;
    rng.seed_from_u64(0);
    float ml = default_ml(m);
    for (uint64_t i = 0; i < n; ++i) out[i] = (uint8_t)new_layer(ml, rng);
}

// raw ChaCha block for validating the core against the RFC 7539 vector
void oracle_chacha_block(const uint32_t* key, uint64_t counter, uint32_t nonce0, uint32_t nonce1,
                         int rounds, uint32_t* out16) {
    (void)nonce0; (void)nonce1;
    ChaChaRng::block(key, counter, rounds, out16);
}

void* oracle_index_new(uint64_t m, int64_t ef_cons, uint64_t dim) {  // template.rs:133-144
    Index* ix = new Index();
    ix->params.m = m;
    ix->params.mmax = m;
    ix->params.mmax0 = 2 * m;
    ix->params.ml = default_ml(m);
    ix->params.ef_cons = ef_cons >= 0 ? (size_t)ef_cons : 2 * m;
    ix->params.dim = dim;
    ix->params.ep = 0;
    ix->layers.m = m;
    ix->points.dim = dim;
    return ix;
}
void oracle_index_free(void* h) { delete (Index*)h; }
// VecType of an EMPTY index: 0 = QuantVec (the reference as committed), 1 = FullVec (points/src/point.rs:4 flipped)
int oracle_set_vec_type(void* h, int full) {
    Index& ix = *(Index*)h;
    if (ix.points.len() != 0) { g_err = "the vector type can only be chosen while the index is empty"; return -1; }
    ix.points.full = full != 0;
    return 0;
}
int oracle_vec_type(void* h) { return ((Index*)h)->points.full ? 1 : 0; }
void oracle_export_values(void* h, float* vals) {  // FullVec store, n*dim
    const Points& P = ((Index*)h)->points;
    if (vals && P.full) memcpy(vals, P.vals.data(), 4 * P.vals.size());
}

int oracle_insert_bulk(void* h, const float* rows, uint64_t n, uint64_t dim, const uint8_t* levels,
                       uint64_t* evals) {
    return insert_bulk(*(Index*)h, rows, n, dim, levels, evals) ? 0 : -1;
}

int64_t oracle_insert_vec(void* h, const float* row, uint64_t dim) {  // template.rs:165-173
    Index& ix = *(Index*)h;
    std::vector<uint32_t> ids;
    if (!store_points(ix, row, 1, dim, ids, nullptr)) return -1;
    ix.layers.add_node(ids[0], ix.points.levels[ids[0]]);
    Results r;
    if (!insert(ix, ids[0], r)) return -1;
    return ids[0];
}

uint64_t oracle_len(void* h) { return ((Index*)h)->points.len(); }
uint64_t oracle_nb_layers(void* h) { return ((Index*)h)->layers.levels.size(); }
uint32_t oracle_ep(void* h) { return ((Index*)h)->params.ep; }
void oracle_set_ep(void* h, uint32_t ep) { ((Index*)h)->params.ep = ep; }
uint64_t oracle_dim(void* h) { return ((Index*)h)->params.dim; }
void oracle_params(void* h, uint64_t* out6, float* ml) {
    const Params& p = ((Index*)h)->params;
    out6[0] = p.m; out6[1] = p.mmax; out6[2] = p.mmax0; out6[3] = p.ef_cons; out6[4] = p.dim; out6[5] = p.ep;
    *ml = p.ml;
}
float oracle_distance(void* h, uint32_t a, uint32_t b) {  // template.rs:150-152; NaN if missing
    Index& ix = *(Index*)h;
    if (a >= ix.points.len() || b >= ix.points.len()) return NAN;
    return ix.points.distance(a, b);
}

// flat export: codes[n*dim], mins[n], deltas[n], levels[n]
void oracle_export_points(void* h, uint8_t* codes, float* mins, float* deltas, uint8_t* levels) {
    const Points& P = ((Index*)h)->points;
    if (codes) memcpy(codes, P.codes.data(), P.codes.size());
    if (mins) memcpy(mins, P.mins.data(), 4 * P.len());
    if (deltas) memcpy(deltas, P.deltas.data(), 4 * P.len());
    if (levels) memcpy(levels, P.levels.data(), P.len());
}
uint64_t oracle_layer_nb_nodes(void* h, uint64_t l) { return ((Index*)h)->layers.levels[l].nb_nodes(); }
uint64_t oracle_layer_nb_edges(void* h, uint64_t l) {
    uint64_t e = 0;
    for (auto& s : ((Index*)h)->layers.levels[l].nb) e += s.size();
    return e;
}
uint64_t oracle_layer_cap(void* h, uint64_t l) { return ((Index*)h)->layers.levels[l].m; }
// CSR export of one layer, rows in ascending node id, neighbours ascending
void oracle_export_layer(void* h, uint64_t l, uint32_t* node_ids, uint64_t* offsets, uint32_t* nbrs) {
    const Graph& g = ((Index*)h)->layers.levels[l];
    std::vector<uint32_t> order(g.node_ids);
    std::sort(order.begin(), order.end());
    uint64_t off = 0;
    for (size_t r = 0; r < order.size(); ++r) {
        node_ids[r] = order[r];
        offsets[r] = off;
        std::vector<uint32_t> s = g.nb[g.row_of[order[r]]];
        std::sort(s.begin(), s.end());
        for (uint32_t v : s) nbrs[off++] = v;
    }
    offsets[order.size()] = off;
}

// import an index from flat parts (so the oracle can search a GPU-built graph)
void* oracle_index_from_parts(uint64_t m, uint64_t ef_cons, uint64_t dim, uint32_t ep, uint64_t n,
                              const uint8_t* codes, const float* mins, const float* deltas,
                              const uint8_t* levels, uint64_t n_layers, const uint64_t* n_nodes,
                              const uint32_t* const* node_ids, const uint64_t* const* offsets,
                              const uint32_t* const* nbrs) {
    Index* ix = (Index*)oracle_index_new(m, (int64_t)ef_cons, dim);
    ix->params.ep = ep;
    if (!mins && !deltas) {  // FullVec parts: `codes` points at n*dim f32 values
        ix->points.full = true;
        const float* v = reinterpret_cast<const float*>(codes);
        ix->points.vals.assign(v, v + n * dim);
    } else {
        ix->points.codes.assign(codes, codes + n * dim);
        ix->points.mins.assign(mins, mins + n);
        ix->points.deltas.assign(deltas, deltas + n);
    }
    ix->points.levels.assign(levels, levels + n);
    ix->layers.add_level(n_layers ? n_layers - 1 : 0);
    for (uint64_t l = 0; l < n_layers; ++l) {
        Graph& g = ix->layers.levels[l];
        for (uint64_t r = 0; r < n_nodes[l]; ++r) g.add_node(node_ids[l][r]);
        for (uint64_t r = 0; r < n_nodes[l]; ++r) {
            std::vector<uint32_t>& s = g.nb[g.row_of[node_ids[l][r]]];
            s.assign(nbrs[l] + offsets[l][r], nbrs[l] + offsets[l][r + 1]);
        }
    }
    return ix;
}

int oracle_save(void* h, const char* dir) { return save_dir(*(Index*)h, dir) ? 0 : -1; }
void* oracle_load(const char* dir) {
    Index* ix = new Index();
    if (!load_dir(*ix, dir)) { delete ix; return nullptr; }
    return ix;
}

// ann_by_vector for a batch of queries, statically partitioned over `threads`
// std::threads sharing the read-only index (the reference itself is one query
// per call on one thread; threads > 1 is the "all host cores" baseline).
// out_ids/out_dists: q*n (padded with 0xFFFFFFFF / +inf), counts[q], hops[q], evals[q].
int oracle_search_batch(void* h, const float* queries, uint64_t q, uint64_t n, uint64_t ef,
                        uint32_t threads, uint32_t* out_ids, float* out_dists, uint32_t* counts,
                        uint32_t* hops, uint32_t* evals) {
    const Index& ix = *(Index*)h;
    if (ix.points.len() == 0) { g_err = "empty index"; return -1; }
    if (threads == 0) threads = 1;
    std::atomic<int> failed{0};
    std::string err;
    auto work = [&](uint64_t lo, uint64_t hi) {
        Results r;
        std::vector<Dist> out;
        for (uint64_t i = lo; i < hi; ++i) {
            if (!ann_by_vector(ix, queries + i * ix.params.dim, n, ef, r, out)) {
                if (!failed.exchange(1)) err = g_err;
                return;
            }
            for (uint64_t j = 0; j < n; ++j) {
                if (out_ids) out_ids[i * n + j] = j < out.size() ? out[j].id : 0xFFFFFFFFu;
                if (out_dists) out_dists[i * n + j] = j < out.size() ? out[j].dist : INFINITY;
            }
            if (counts) counts[i] = (uint32_t)out.size();
            if (hops) hops[i] = (uint32_t)r.hops;
            if (evals) evals[i] = (uint32_t)r.evals;
        }
    };
    if (threads == 1) {
        work(0, q);
    } else {
        std::vector<std::thread> th;
        uint64_t per = (q + threads - 1) / threads;
        for (uint32_t t = 0; t < threads; ++t) {
            uint64_t lo = std::min<uint64_t>(q, t * per), hi = std::min<uint64_t>(q, lo + per);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto& t : th) t.join();
    }
    if (failed) { g_err = err; return -1; }
    return 0;
}

// exact top-k under the quantised metric with (dist,id) order, batch of f32 queries
int oracle_bruteforce(void* h, const float* queries, uint64_t q, uint64_t k, uint32_t threads,
                      uint32_t* out_ids, float* out_dists) {
    const Index& ix = *(Index*)h;
    if (threads == 0) threads = 1;
    std::atomic<int> failed{0};
    auto work = [&](uint64_t lo, uint64_t hi) {
        std::vector<Dist> out;
        QuantVec qv;
        for (uint64_t i = lo; i < hi; ++i) {
            if (!ix.points.make_point(queries + i * ix.params.dim, ix.params.dim, qv)) { failed = 1; return; }
            brute_force_one(ix, qv, k, out);
            for (uint64_t j = 0; j < k; ++j) {
                if (out_ids) out_ids[i * k + j] = j < out.size() ? out[j].id : 0xFFFFFFFFu;
                if (out_dists) out_dists[i * k + j] = j < out.size() ? out[j].dist : INFINITY;
            }
        }
    };
    std::vector<std::thread> th;
    uint64_t per = (q + threads - 1) / threads;
    for (uint32_t t = 0; t < threads; ++t) {
        uint64_t lo = std::min<uint64_t>(q, t * per), hi = std::min<uint64_t>(q, lo + per);
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto& t : th) t.join();
    return failed ? -1 : 0;
}

// distances from one f32 query (quantised first) to a list of ids: dist2many
int oracle_dist_query_many(void* h, const float* query, const uint32_t* ids, uint64_t n, float* out) {
    const Index& ix = *(Index*)h;
    QuantVec qv;
    if (!ix.points.make_point(query, ix.params.dim, qv)) return -1;
    for (uint64_t i = 0; i < n; ++i) out[i] = ix.points.distance2point(qv, ids[i]);
    return 0;
}

// reference test helper: per-layer min / max degree (template.rs:556-571)
void oracle_layer_degree_range(void* h, uint64_t l, uint64_t* mn, uint64_t* mx) {
    const Graph& g = ((Index*)h)->layers.levels[l];
    uint64_t a = UINT64_MAX, b = 0;
    for (auto& s : g.nb) { a = std::min<uint64_t>(a, s.size()); b = std::max<uint64_t>(b, s.size()); }
    *mn = a; *mx = b;
}

// ---- bare Graph API for the graph.rs unit-test pins ----
void* oracle_graph_new(uint64_t level, uint64_t m) { Graph* g = new Graph(); g->level = level; g->m = m; return g; }
void oracle_graph_free(void* g) { delete (Graph*)g; }
void oracle_graph_add_node(void* g, uint32_t id) { ((Graph*)g)->add_node(id); }
int oracle_graph_add_edge(void* g, uint32_t a, uint32_t b) { return ((Graph*)g)->add_edge(a, b); }
int oracle_graph_remove_edge(void* g, uint32_t a, uint32_t b) { return ((Graph*)g)->remove_edge(a, b); }
int oracle_graph_contains(void* g, uint32_t id) { return ((Graph*)g)->contains(id) ? 1 : 0; }
int64_t oracle_graph_degree(void* g, uint32_t id) { return ((Graph*)g)->contains(id) ? (int64_t)((Graph*)g)->degree(id) : -1; }
int64_t oracle_graph_neighbors(void* g, uint32_t id, uint32_t* out, uint64_t cap) {
    const std::vector<uint32_t>* s = ((Graph*)g)->nbrs(id);
    if (!s) return -1;
    for (size_t i = 0; i < s->size() && i < cap; ++i) out[i] = (*s)[i];
    return (int64_t)s->size();
}
int oracle_graph_replace_neighbors(void* g, uint32_t id, const uint32_t* nn, uint64_t n) {
    return ((Graph*)g)->replace_neighbors(id, std::vector<uint32_t>(nn, nn + n));
}
uint64_t oracle_graph_nb_nodes(void* g) { return ((Graph*)g)->nb_nodes(); }

// GloVe text loader (helpers/glove.rs:14-71): `word v1 .. vd` per line, values
// parsed straight to f32 (strtof is correctly rounded, like Rust's parse::<f32>).
// Returns rows read; writes at most cap floats; *dim_out = values per row.
int64_t oracle_load_glove(const char* path, uint64_t lim, float* out, uint64_t cap, uint64_t* dim_out) {
    FILE* f = fopen(path, "r");
    if (!f) { g_err = std::string("cannot open ") + path; return -1; }
    char* line = nullptr;
    size_t lcap = 0;
    int64_t rows = 0;
    uint64_t dim = 0, w = 0;
    while (getline(&line, &lcap, f) > 0) {
        if (lim > 0 && (uint64_t)rows >= lim) break;
        char* save = nullptr;
        char* tok = strtok_r(line, " \n\r", &save);  // the word
        if (!tok) continue;
        uint64_t d = 0;
        while ((tok = strtok_r(nullptr, " \n\r", &save))) {
            char* end = nullptr;
            float v = strtof(tok, &end);
            if (end == tok || *end != '\0') continue;  // non-numeric token joins the word
            if (out && w < cap) out[w] = v;
            ++w;
            ++d;
        }
        if (rows == 0) dim = d;
        else if (d != dim) { g_err = "vector is not the same size as others"; free(line); fclose(f); return -1; }
        ++rows;
    }
    free(line);
    fclose(f);
    if (dim_out) *dim_out = dim;
    return rows;
}

}  // extern "C"
