// The plain, step-by-step form of the build's commit (hnsw/src/template.rs:196-251): for every over-full
// neighbour the kept list is materialised (select_simple, template.rs:614-621) and replace_neighbors is
// applied as "cut what is not kept, re-add what is kept".  Test reference for the product's faster
// hnsw_rs_b200/csrc/commit.h, which must leave exactly the same graph.
#pragma once
#include "../../hnsw_rs_b200/csrc/commit.h"

namespace hb {

struct PlainScratch {
    struct Prune { uint32_t layer, node, kept_off, kept_n, drop_off, drop_n; };
    std::vector<Prune> prunes;
    std::vector<uint32_t> kept_ids, drop_ids, lost;
    std::vector<float> kept_w;
    std::vector<std::pair<commit_u64, uint32_t>> keyed;
};

// 0, or 1 with *err set
inline int commit_point_plain(HostGraph& h, uint32_t pid, const std::vector<LayerSel>& res, std::vector<uint32_t>& dirty0,
                        std::vector<uint32_t>& dirtyu, PlainScratch& cs, const char** err) {
    // make_connections: every layer first (ascending layer, ascending Dist)
    for (const LayerSel& ls : res) {
        std::vector<uint32_t>* dirty = ls.layer == 0 ? &dirty0 : &dirtyu;
        for (uint32_t i = 0; i < ls.n; ++i) {
            int r = h.add_edge(ls.layer, pid, ls.ids[i], ls.dists[i], dirty);
            if (r) { *err = "make_connections: add_edge failed (self connection or node not in graph)"; return 1; }
        }
    }
    // prune_connections: for every new neighbour x above the layer cap keep the cap nearest
    // (select_simple, template.rs:614-621).  All results are computed before any is applied.
    cs.prunes.clear();
    cs.kept_ids.clear();
    cs.kept_w.clear();
    cs.drop_ids.clear();
    for (const LayerSel& ls : res) {
        const AdjStore& s = h.store(ls.layer);
        const uint32_t cap = h.cap(ls.layer);
        for (uint32_t xi = 0; xi < ls.n; ++xi) {
            const uint32_t x = ls.ids[xi];
            uint32_t row = h.row(x, ls.layer);
            uint32_t d = s.deg[row];
            if (!(d > cap)) continue;
            // (prune_results is a map keyed by node; a node occurs once in one point's selection)
            cs.keyed.clear();
            for (uint32_t i = 0; i < d; ++i) {
                float w = s.getw(row, i);
                uint32_t bits;
                memcpy(&bits, &w, 4);
                cs.keyed.push_back({((commit_u64)bits << 32) | s.get(row, i), i});
            }
            // keep the `cap` smallest (dist, id) keys; almost always d == cap + 1
            std::nth_element(cs.keyed.begin(), cs.keyed.begin() + cap, cs.keyed.end());
            PlainScratch::Prune pr;
            pr.layer = ls.layer;
            pr.node = x;
            pr.kept_off = (uint32_t)cs.kept_ids.size();
            pr.kept_n = cap;
            pr.drop_off = (uint32_t)cs.drop_ids.size();
            pr.drop_n = d - cap;
            for (uint32_t i = 0; i < d; ++i) {
                if (i < cap) {
                    cs.kept_ids.push_back((uint32_t)cs.keyed[i].first);
                    cs.kept_w.push_back(s.getw(row, cs.keyed[i].second));
                } else {
                    cs.drop_ids.push_back((uint32_t)cs.keyed[i].first);
                }
            }
            cs.prunes.push_back(pr);
        }
    }
    // make_pruned_connections: ascending layer, ascending node id (oracle convention for the
    // reference's hash-map iteration order).  replace_neighbors(x, kept) = isolate_node(x) +
    // add_neighbors(x, kept) (graph.rs:85-94,128-148): members of `kept` are removed and re-added
    // (no net change), the others lose the edge unless their degree is 1.  A kept edge has to be
    // re-created only if an earlier replacement of this same point cut it, i.e. x lost an edge.
    std::sort(cs.prunes.begin(), cs.prunes.end(), [](const PlainScratch::Prune& a, const PlainScratch::Prune& b) {
        return a.layer != b.layer ? a.layer < b.layer : a.node < b.node;
    });
    cs.lost.clear();
    uint32_t lost_layer = 0xFFFFFFFFu;
    for (const PlainScratch::Prune& pr : cs.prunes) {
        std::vector<uint32_t>* dirty = pr.layer == 0 ? &dirty0 : &dirtyu;
        if (pr.layer != lost_layer) { cs.lost.clear(); lost_layer = pr.layer; }
        const bool x_lost = std::find(cs.lost.begin(), cs.lost.end(), pr.node) != cs.lost.end();
        for (uint32_t i = 0; i < pr.drop_n; ++i) {
            uint32_t nb = cs.drop_ids[pr.drop_off + i];
            if (h.store(pr.layer).find(h.row(pr.node, pr.layer), nb) < 0) continue;  // already cut earlier
            if (h.degree(nb, pr.layer) == 1) continue;
            h.remove_edge(pr.layer, pr.node, nb, dirty);
            cs.lost.push_back(nb);
        }
        if (x_lost) {
            for (uint32_t i = 0; i < pr.kept_n; ++i) {
                int r = h.add_edge(pr.layer, pr.node, cs.kept_ids[pr.kept_off + i], cs.kept_w[pr.kept_off + i], dirty);
                if (r) { *err = "make_pruned_connections: replace_neighbors failed"; return 1; }
            }
        }
    }
    return 0;
}

}  // namespace hb
