// Host-side mirror of the layered graph (graph/src/graph.rs, graph/src/layers.rs) in
// the flat fixed-stride form the device kernels read.  The reference keeps, per layer,
// a hash map node -> mutex-guarded hash set of neighbours; here layer 0 is one row per
// point and all upper layers share one row pool (row = upper_off[node] + layer - 1).
// Edge mutation semantics (add_edge / remove_edge / isolate_node / replace_neighbors,
// graph.rs:37-148) are restated on these rows; the build commits edges here and ships
// only the touched rows to the device.
#pragma once
#include <stdint.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <unordered_map>
#include <vector>

namespace hb {

constexpr uint32_t H_EMPTY = 0xFFFFFFFFu;

// neighbour sets with set semantics; S inline slots per row, one more slot per row in `xdata`
// (a full row holds cap + 1 neighbours between make_connections and prune_connections, which
// is the steady state of the build), rare overflow beyond that in `spill`.
// Every edge also carries its length d(a,b) (bit-identical in both directions, since
// (x-y)^2 is exactly symmetric): the build's prune step (hnsw/src/template.rs:209-238)
// needs d(x, n) for every neighbour n of x, and all of those were already produced by
// the device when the edge was selected.
struct AdjStore {
    uint32_t cap = 0;  // Graph.m: the degree cap the reference prunes to
    uint32_t S = 0;    // slots per row (cap rounded up to a multiple of 4)
    std::vector<uint32_t> data;
    std::vector<float> w;
    std::vector<uint32_t> deg;
    std::vector<uint32_t> xdata;  // slot S of every row
    std::vector<float> xw;
    std::unordered_map<uint32_t, std::vector<uint32_t>> spill;  // slots S+1...
    std::unordered_map<uint32_t, std::vector<float>> wspill;

    void init(uint32_t cap_) {
        cap = cap_;
        S = (cap_ + 3) / 4 * 4;
        if (S < 4) S = 4;
    }
    uint64_t rows() const { return deg.size(); }
    void add_rows(uint64_t n) {
        data.resize(data.size() + n * S, H_EMPTY);
        w.resize(w.size() + n * S, 0.0f);
        deg.resize(deg.size() + n, 0);
        xdata.resize(xdata.size() + n, H_EMPTY);
        xw.resize(xw.size() + n, 0.0f);
    }
    uint32_t get(uint32_t row, uint32_t i) const {
        if (i < S) return data[(size_t)row * S + i];
        if (i == S) return xdata[row];
        return spill.find(row)->second[i - S - 1];
    }
    float getw(uint32_t row, uint32_t i) const {
        if (i < S) return w[(size_t)row * S + i];
        if (i == S) return xw[row];
        return wspill.find(row)->second[i - S - 1];
    }
    void set(uint32_t row, uint32_t i, uint32_t v, float wt) {
        if (i < S) {
            data[(size_t)row * S + i] = v;
            w[(size_t)row * S + i] = wt;
        } else if (i == S) {
            xdata[row] = v;
            xw[row] = wt;
        } else {
            auto& sp = spill[row];
            auto& ws = wspill[row];
            if (sp.size() <= i - S - 1) { sp.resize(i - S); ws.resize(i - S); }
            sp[i - S - 1] = v;
            ws[i - S - 1] = wt;
        }
    }
    void setw(uint32_t row, uint32_t i, float wt) {
        if (i < S) w[(size_t)row * S + i] = wt;
        else if (i == S) xw[row] = wt;
        else wspill[row][i - S - 1] = wt;
    }
    int find(uint32_t row, uint32_t v) const {
        uint32_t d = deg[row];
        const uint32_t* p = &data[(size_t)row * S];
        uint32_t lim = d < S ? d : S;
#if defined(__SSE2__)
        // S is a multiple of 4 and unused slots hold H_EMPTY (never a node id): compare whole quads
        const __m128i vv = _mm_set1_epi32((int)v);
        for (uint32_t i = 0; i < lim; i += 4) {
            int hit = _mm_movemask_ps(_mm_castsi128_ps(_mm_cmpeq_epi32(_mm_loadu_si128((const __m128i*)(p + i)), vv)));
            if (hit) return (int)(i + (uint32_t)__builtin_ctz((unsigned)hit));
        }
#else
        for (uint32_t i = 0; i < lim; ++i)
            if (p[i] == v) return (int)i;
#endif
        if (d > S) {
            if (xdata[row] == v) return (int)S;
            if (d > S + 1) {
                const auto& sp = spill.find(row)->second;
                for (uint32_t i = 0; i < d - S - 1; ++i)
                    if (sp[i] == v) return (int)(S + 1 + i);
            }
        }
        return -1;
    }
    bool insert(uint32_t row, uint32_t v, float wt = 0.0f) {  // IntSet::insert
        if (find(row, v) >= 0) return false;
        set(row, deg[row], v, wt);
        deg[row]++;
        return true;
    }
    bool remove(uint32_t row, uint32_t v) {  // IntSet::remove (order is irrelevant)
        int i = find(row, v);
        if (i < 0) return false;
        uint32_t last = deg[row] - 1;
        set(row, (uint32_t)i, get(row, last), getw(row, last));
        if (last < S) data[(size_t)row * S + last] = H_EMPTY;
        else if (last == S) xdata[row] = H_EMPTY;
        else {
            auto it = spill.find(row);
            it->second.pop_back();
            wspill.find(row)->second.pop_back();
            if (it->second.empty()) { spill.erase(it); wspill.erase(row); }
        }
        deg[row] = last;
        return true;
    }
    void list(uint32_t row, std::vector<uint32_t>& out) const {
        out.clear();
        for (uint32_t i = 0; i < deg[row]; ++i) out.push_back(get(row, i));
    }
};

struct HostGraph {
    uint32_t m = 0;  // Layers.m: upper-layer cap; layer 0 cap is 2m (layers.rs:50)
    std::vector<uint8_t> level;       // per node: highest layer it belongs to
    std::vector<uint32_t> upper_off;  // per node: first row in `au`, H_EMPTY if level 0
    AdjStore a0, au;
    std::vector<uint64_t> layer_nodes;  // nodes per layer (Graph::nb_nodes)
    bool weights_valid = true;          // false for imported graphs until the build recomputes them

    void init(uint32_t m_, uint32_t cap0, uint32_t capu) {
        m = m_;
        a0.init(cap0);
        au.init(capu);
    }
    uint64_t n_points() const { return level.size(); }
    uint32_t n_layers() const { return (uint32_t)layer_nodes.size(); }
    // Layers::add_node(id, level) for id == n_points()  (layers.rs:62-70)
    uint32_t add_node(uint32_t lvl) {
        uint32_t id = (uint32_t)level.size();
        level.push_back((uint8_t)lvl);
        a0.add_rows(1);
        if (lvl > 0) {
            upper_off.push_back((uint32_t)au.rows());
            au.add_rows(lvl);
        } else {
            upper_off.push_back(H_EMPTY);
        }
        while (layer_nodes.size() <= lvl) layer_nodes.push_back(0);
        for (uint32_t l = 0; l <= lvl; ++l) layer_nodes[l]++;
        return id;
    }
    bool in_layer(uint32_t node, uint32_t layer) const { return node < level.size() && level[node] >= layer; }
    AdjStore& store(uint32_t layer) { return layer == 0 ? a0 : au; }
    const AdjStore& store(uint32_t layer) const { return layer == 0 ? a0 : au; }
    uint32_t row(uint32_t node, uint32_t layer) const { return layer == 0 ? node : upper_off[node] + layer - 1; }
    uint32_t cap(uint32_t layer) const { return layer == 0 ? a0.cap : au.cap; }
    uint32_t degree(uint32_t node, uint32_t layer) const { return store(layer).deg[row(node, layer)]; }

    // graph.rs:37-52; 0 ok, 1 self connection, 2 node not in graph
    int add_edge(uint32_t layer, uint32_t a, uint32_t b, float wt, std::vector<uint32_t>* dirty) {
        if (a == b) return 1;
        if (!in_layer(a, layer) || !in_layer(b, layer)) return 2;
        AdjStore& s = store(layer);
        uint32_t ra = row(a, layer), rb = row(b, layer);
        if (s.insert(ra, b, wt) && dirty) dirty->push_back(ra);
        if (s.insert(rb, a, wt) && dirty) dirty->push_back(rb);
        return 0;
    }
    int remove_edge(uint32_t layer, uint32_t a, uint32_t b, std::vector<uint32_t>* dirty) {  // graph.rs:72-83
        if (!in_layer(a, layer) || !in_layer(b, layer)) return 2;
        AdjStore& s = store(layer);
        uint32_t ra = row(a, layer), rb = row(b, layer);
        if (s.remove(ra, b) && dirty) dirty->push_back(ra);
        if (s.remove(rb, a) && dirty) dirty->push_back(rb);
        return 0;
    }
    // graph.rs:128-137 with isolate_node graph.rs:85-94.  The reference removes every edge of
    // `node` except those to degree-1 neighbours and then re-adds `nn`; for members of `nn`
    // that is a remove + re-add with no net effect, so only the neighbours outside `nn`
    // change: they lose the edge unless their degree is 1.  nn must be a subset of the
    // current neighbour set or new ids with known weights (wts parallel to nn).
    int replace_neighbors(uint32_t layer, uint32_t node, const std::vector<uint32_t>& nn,
                          const std::vector<float>& wts, std::vector<uint32_t>* dirty) {
        if (!in_layer(node, layer)) return 2;
        std::vector<uint32_t> snap;
        store(layer).list(row(node, layer), snap);
        for (uint32_t nb : snap) {
            if (std::find(nn.begin(), nn.end(), nb) != nn.end()) continue;
            if (degree(nb, layer) == 1) continue;
            remove_edge(layer, node, nb, dirty);
        }
        for (size_t i = 0; i < nn.size(); ++i) {
            int r = add_edge(layer, node, nn[i], i < wts.size() ? wts[i] : 0.0f, dirty);
            if (r) return r;
        }
        return 0;
    }
};

}  // namespace hb
