// Host commit of one point's insertion results (hnsw/src/template.rs:196-251): make_connections,
// prune_connections, make_pruned_connections on the host mirror of the graph.  Host only (no CUDA),
// so the CPU tests and tools/dev/commit_replay.cpp can drive it without a device.
#pragma once
#include <string.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "hostgraph.h"

namespace hb {

typedef unsigned long long commit_u64;

// one layer's selection for one point: a view into the batch's result arrays
struct LayerSel {
    uint32_t layer;
    uint32_t n;
    const uint32_t* ids;
    const float* dists;
};

struct CommitScratch {
    struct Prune { uint32_t layer, node, drop_off, drop_n; };
    struct Lost { uint32_t node, other; float w; };  // `node` lost its edge to `other` while another row was pruned
    std::vector<Prune> prunes;
    std::vector<uint32_t> drop_ids;
    std::vector<Lost> lost;
    std::vector<std::pair<commit_u64, uint32_t>> keyed;
};

inline commit_u64 commit_key(float w, uint32_t id) {  // Dist order (graph/src/dist.rs:16-37) for distances >= 0
    uint32_t bits;
    memcpy(&bits, &w, 4);
    return ((commit_u64)bits << 32) | id;
}

// 0, or 1 with *err set
inline int commit_point(HostGraph& h, uint32_t pid, const std::vector<LayerSel>& res, std::vector<uint32_t>& dirty0,
                        std::vector<uint32_t>& dirtyu, CommitScratch& cs, const char** err) {
    // make_connections: every layer first (ascending layer, ascending Dist)
    for (const LayerSel& ls : res) {
        std::vector<uint32_t>* dirty = ls.layer == 0 ? &dirty0 : &dirtyu;
        for (uint32_t i = 0; i < ls.n; ++i) {
            int r = h.add_edge(ls.layer, pid, ls.ids[i], ls.dists[i], dirty);
            if (r) { *err = "make_connections: add_edge failed (self connection or node not in graph)"; return 1; }
        }
    }
    // prune_connections: every new neighbour x above the layer cap keeps its cap nearest (select_simple,
    // template.rs:614-621).  All selections are made before any is applied.  Only what x drops is written
    // down: the kept set is "the row as it is now, minus the dropped".
    cs.prunes.clear();
    cs.drop_ids.clear();
    for (const LayerSel& ls : res) {
        const AdjStore& s = h.store(ls.layer);
        const uint32_t cap = h.cap(ls.layer);
        for (uint32_t xi = 0; xi < ls.n; ++xi) {
            const uint32_t x = ls.ids[xi];
            const uint32_t row = h.row(x, ls.layer);
            const uint32_t d = s.deg[row];
            if (!(d > cap)) continue;
            // (prune_results is a map keyed by node; a node occurs once in one point's selection)
            CommitScratch::Prune pr{ls.layer, x, (uint32_t)cs.drop_ids.size(), d - cap};
            if (d == cap + 1) {  // the steady state of a build: exactly one too many, drop the (dist, id) maximum
                commit_u64 worst = 0;
                const uint32_t* pd = &s.data[(size_t)row * s.S];
                const float* pw = &s.w[(size_t)row * s.S];
                const uint32_t lim = std::min(d, s.S);
                for (uint32_t i = 0; i < lim; ++i) worst = std::max(worst, commit_key(pw[i], pd[i]));
                for (uint32_t i = lim; i < d; ++i) worst = std::max(worst, commit_key(s.getw(row, i), s.get(row, i)));
                cs.drop_ids.push_back((uint32_t)worst);
            } else {
                cs.keyed.clear();
                for (uint32_t i = 0; i < d; ++i) cs.keyed.push_back({commit_key(s.getw(row, i), s.get(row, i)), i});
                std::nth_element(cs.keyed.begin(), cs.keyed.begin() + cap, cs.keyed.end());
                for (uint32_t i = cap; i < d; ++i) cs.drop_ids.push_back((uint32_t)cs.keyed[i].first);
            }
            cs.prunes.push_back(pr);
        }
    }
    // make_pruned_connections: ascending layer, ascending node id (oracle convention for the
    // reference's hash-map iteration order).  replace_neighbors(x, kept) = isolate_node(x) +
    // add_neighbors(x, kept) (graph.rs:85-94,128-148): members of `kept` are removed and re-added
    // (no net change), the dropped lose the edge unless their degree is 1.  A kept edge has to be
    // re-created only if an earlier replacement of this same point cut it: those cuts are remembered
    // in `lost` with their length, and x re-adds the ones it did not itself drop.
    std::sort(cs.prunes.begin(), cs.prunes.end(), [](const CommitScratch::Prune& a, const CommitScratch::Prune& b) {
        return a.layer != b.layer ? a.layer < b.layer : a.node < b.node;
    });
    cs.lost.clear();
    uint32_t lost_layer = 0xFFFFFFFFu;
    for (const CommitScratch::Prune& pr : cs.prunes) {
        std::vector<uint32_t>* dirty = pr.layer == 0 ? &dirty0 : &dirtyu;
        if (pr.layer != lost_layer) { cs.lost.clear(); lost_layer = pr.layer; }
        AdjStore& s = h.store(pr.layer);
        const uint32_t rowx = h.row(pr.node, pr.layer);
        const size_t n_lost = cs.lost.size();  // cuts made before this replacement
        for (uint32_t i = 0; i < pr.drop_n; ++i) {
            const uint32_t nb = cs.drop_ids[pr.drop_off + i];
            const int at = s.find(rowx, nb);
            if (at < 0) continue;  // already cut earlier
            const uint32_t rownb = h.row(nb, pr.layer);
            if (s.deg[rownb] == 1) continue;
            cs.lost.push_back({nb, pr.node, s.getw(rowx, (uint32_t)at)});
            s.remove(rowx, nb);
            s.remove(rownb, pr.node);
            if (dirty) { dirty->push_back(rowx); dirty->push_back(rownb); }
        }
        for (size_t k = 0; k < n_lost; ++k) {
            if (cs.lost[k].node != pr.node) continue;
            const uint32_t y = cs.lost[k].other;
            const uint32_t* db = &cs.drop_ids[pr.drop_off];
            if (std::find(db, db + pr.drop_n, y) != db + pr.drop_n) continue;  // x drops it itself
            int r = h.add_edge(pr.layer, pr.node, y, cs.lost[k].w, dirty);
            if (r) { *err = "make_pruned_connections: replace_neighbors failed"; return 1; }
        }
    }
    return 0;
}

}  // namespace hb
