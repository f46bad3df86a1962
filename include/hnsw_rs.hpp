// hnsw_rs.hpp -- C++17 host-side mirror of the public API of the Rust workspace Gumo-A/hnsw_rs for the
// search / distance / build path, over the C ABI of libhnsw_b200.so (include/hnsw_b200.h).
//
// The reference is compiled code (Rust) and there is no Rust toolchain in this image, so the host side
// above the C ABI is written in C++ with the reference's names, argument meaning and error behaviour:
//   Result<_, String>::Err  -> throws hnsw_rs::Error        (message = the library's last_error)
//   Option::None            -> std::nullopt
//   panic!                  -> throws hnsw_rs::Panic         (e.g. check_points_dim, template.rs:253-262)
// Header only; link with -lhnsw_b200.  Nothing here computes: every distance, search and build runs in the
// CUDA library, and without a CUDA device every compute call throws.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <optional>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "hnsw_b200.h"

namespace hnsw_rs {

using NodeID = uint32_t;  // graph/src/lib.rs:1

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct Panic : std::logic_error {
    using std::logic_error::logic_error;
};
inline void check(int rc) {
    if (rc != HNSWB200_OK) throw Error(rc, hnswb200_last_error());
}

// one per GPU; shared by the objects created from it
class Context {
   public:
    explicit Context(int device = 0) { check(hnswb200_ctx_create(device, &h_)); }
    ~Context() { hnswb200_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    hnswb200_ctx* get() const { return h_; }
    static Context& global() {
        static Context c(0);
        return c;
    }

   private:
    hnswb200_ctx* h_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------
namespace vectors {

// vectors/src/full.rs
struct FullVec {
    std::vector<float> vector;
    static FullVec new_(const std::vector<float>& v) { return FullVec{v}; }
    size_t dim() const { return vector.size(); }
    std::vector<float> get_vals() const { return vector; }
    // full.rs:23-29: strictly sequential f32 sum over zip(self, other), sqrt
    float distance(const FullVec& o, Context& c = Context::global()) const {
        const size_t d = std::min(vector.size(), o.vector.size());
        float out = 0.f;
        if (d == 0) return 0.f;
        check(hnswb200_dist_full_pairs(c.get(), vector.data(), o.vector.data(), 1, (uint32_t)d, &out));
        return out;
    }
    float dist2other(const FullVec& o, Context& c = Context::global()) const { return distance(o, c); }
    std::vector<float> dist2many(const std::vector<FullVec>& others, Context& c = Context::global()) const {
        std::vector<float> out;
        for (const auto& o : others) out.push_back(dist2other(o, c));
        return out;
    }
    // Serializer, full.rs:44-70: dim x f32 big-endian
    size_t size() const { return vector.size() * 4; }
    std::vector<uint8_t> serialize() const {
        std::vector<uint8_t> b;
        for (float f : vector) {
            uint32_t u;
            std::memcpy(&u, &f, 4);
            for (int i = 3; i >= 0; --i) b.push_back((uint8_t)(u >> (8 * i)));
        }
        return b;
    }
    static FullVec deserialize(const std::vector<uint8_t>& b) {
        FullVec v;
        for (size_t i = 0; i + 4 <= b.size(); i += 4) {
            uint32_t u = ((uint32_t)b[i] << 24) | ((uint32_t)b[i + 1] << 16) | ((uint32_t)b[i + 2] << 8) | b[i + 3];
            float f;
            std::memcpy(&f, &u, 4);
            v.vector.push_back(f);
        }
        return v;
    }
};

// vectors/src/quant.rs
struct QuantVec {
    float delta = 0.f, min = 0.f;
    std::vector<uint8_t> codes;
    // quant.rs:41-66 (on the device; NaN input -> Panic like partial_cmp().unwrap())
    static QuantVec new_(const std::vector<float>& v, Context& c = Context::global()) {
        QuantVec q;
        q.codes.resize(v.size());
        int rc = hnswb200_quantise(c.get(), v.data(), 1, (uint32_t)v.size(), q.codes.data(), &q.min, &q.delta);
        if (rc == HNSWB200_EINVAL) throw Panic(hnswb200_last_error());
        check(rc);
        return q;
    }
    size_t dim() const { return codes.size(); }
    // quant.rs:79-83: (c as f32) * delta + min, two roundings
    std::vector<float> get_vals() const {
        std::vector<float> out;
        for (uint8_t cde : codes) {
            volatile float p = (float)cde * delta;  // volatile: keep the product rounded on its own
            out.push_back(p + min);
        }
        return out;
    }
    // quant.rs:75-77 -> distance_unrolled (quant.rs:14-37)
    float dist2other(const QuantVec& o, Context& c = Context::global()) const { return dist2many({o}, c)[0]; }
    std::vector<float> dist2many(const std::vector<QuantVec>& others, Context& c = Context::global()) const {
        if (others.empty()) return {};
        const size_t d = dim(), n = others.size() + 1;
        std::vector<uint8_t> cb(n * d);
        std::vector<float> mn(n), dl(n);
        std::memcpy(cb.data(), codes.data(), d);
        mn[0] = min;
        dl[0] = delta;
        for (size_t i = 0; i < others.size(); ++i) {
            if (others[i].dim() != d) throw Panic("dist2many: dimension mismatch");
            std::memcpy(cb.data() + (i + 1) * d, others[i].codes.data(), d);
            mn[i + 1] = others[i].min;
            dl[i + 1] = others[i].delta;
        }
        hnswb200_points* p = nullptr;
        check(hnswb200_points_upload(c.get(), cb.data(), mn.data(), dl.data(), nullptr, n, (uint32_t)d, &p));
        std::vector<uint32_t> a(others.size(), 0), b(others.size());
        for (size_t i = 0; i < others.size(); ++i) b[i] = (uint32_t)(i + 1);
        std::vector<float> out(others.size());
        int rc = hnswb200_dist_pairs(c.get(), p, a.data(), b.data(), others.size(), out.data());
        hnswb200_points_destroy(p);
        check(rc);
        return out;
    }
    // Serializer, quant.rs:90-125: min f32, delta f32, codes (big-endian)
    size_t size() const { return 8 + codes.size(); }
    std::vector<uint8_t> serialize() const {
        std::vector<uint8_t> b;
        for (float f : {min, delta}) {
            uint32_t u;
            std::memcpy(&u, &f, 4);
            for (int i = 3; i >= 0; --i) b.push_back((uint8_t)(u >> (8 * i)));
        }
        b.insert(b.end(), codes.begin(), codes.end());
        return b;
    }
    static QuantVec deserialize(const std::vector<uint8_t>& b) {
        auto rd = [&](size_t i) {
            uint32_t u = ((uint32_t)b[i] << 24) | ((uint32_t)b[i + 1] << 16) | ((uint32_t)b[i + 2] << 8) | b[i + 3];
            float f;
            std::memcpy(&f, &u, 4);
            return f;
        };
        QuantVec q;
        q.min = rd(0);
        q.delta = rd(4);
        q.codes.assign(b.begin() + 8, b.end());
        return q;
    }
};

}  // namespace vectors

// ---------------------------------------------------------------------------------------------------
namespace graph {

// graph/src/dist.rs:4-37
struct Dist {
    NodeID id;
    float dist;
    bool operator==(const Dist& o) const { return id == o.id && dist == o.dist; }
    bool operator<(const Dist& o) const {
        if (std::isnan(dist) || std::isnan(o.dist)) throw Panic("Dist: NaN distance (partial_cmp().unwrap())");
        return dist < o.dist || (dist == o.dist && id < o.id);
    }
};

// graph/src/graph.rs as a read-only view of one layer exported from the device-side index
struct Graph {
    std::map<NodeID, std::set<NodeID>> nodes;
    size_t level = 0, m = 0;
    size_t nb_nodes() const { return nodes.size(); }
    bool contains(NodeID n) const { return nodes.count(n) != 0; }
    std::optional<size_t> degree(NodeID n) const {
        auto it = nodes.find(n);
        if (it == nodes.end()) return std::nullopt;
        return it->second.size();
    }
    std::optional<std::vector<NodeID>> neighbors_vec(NodeID n) const {
        auto it = nodes.find(n);
        if (it == nodes.end()) return std::nullopt;
        return std::vector<NodeID>(it->second.begin(), it->second.end());
    }
};

}  // namespace graph

// ---------------------------------------------------------------------------------------------------
namespace hnsw {

// hnsw/src/params.rs:4-62
struct Params {
    NodeID ep = 0;
    size_t m = 0, mmax = 0, mmax0 = 0;
    float ml = 0.f;
    size_t ef_cons = 0, dim = 0;
    static Params from_c(const hnswb200_params& c) {
        return Params{c.ep, (size_t)c.m, (size_t)c.mmax, (size_t)c.mmax0, c.ml, (size_t)c.ef_cons, (size_t)c.dim};
    }
    static Params from_m(size_t m, size_t dim) {
        hnswb200_params c;
        hnswb200_params_default(m, -1, dim, &c);
        return from_c(c);
    }
    static Params from_m_efcons(size_t m, size_t ef_cons, size_t dim) {
        hnswb200_params c;
        hnswb200_params_default(m, (int64_t)ef_cons, dim, &c);
        return from_c(c);
    }
};

class HNSW {
   public:
    Params params;

    // template.rs:133-144.  metric: HNSWB200_METRIC_L2 (the reference) or HNSWB200_METRIC_COSINE (an addition: rows and
    // queries are L2-normalised on the device, then the reference's L2 path runs on the unit vectors)
    // vec_type: HNSWB200_VEC_QUANT (the reference as committed, `type VecType = QuantVec;`, points/src/point.rs:4) or
    // HNSWB200_VEC_FULL (the same alias flipped to FullVec: f32 vectors, FullVec::distance)
    static HNSW new_(size_t m, std::optional<size_t> ef_cons, size_t dim, Context& c = Context::global(),
                     int metric = HNSWB200_METRIC_L2, int vec_type = HNSWB200_VEC_QUANT) {
        hnswb200_params p;
        hnswb200_params_default(m, ef_cons ? (int64_t)*ef_cons : -1, dim, &p);
        hnswb200_index* ix = nullptr;
        const int prev = hnswb200_ctx_vec_type(c.get());
        check(hnswb200_ctx_set_vec_type(c.get(), vec_type));
        const int brc = hnswb200_build(c.get(), nullptr, 0, (uint32_t)dim, &p, nullptr, 0, &ix);
        hnswb200_ctx_set_vec_type(c.get(), prev);
        check(brc);
        if (metric != HNSWB200_METRIC_L2) {
            int rc = hnswb200_index_set_metric(ix, metric);
            if (rc) { hnswb200_index_destroy(ix); check(rc); }
        }
        return HNSW(&c, ix);
    }
    int metric() const { return hnswb200_index_metric(ix_); }
    int vec_type() const { return hnswb200_points_vec_type(hnswb200_index_points(ix_)); }
    HNSW(HNSW&& o) noexcept : params(o.params), ctx_(o.ctx_), ix_(o.ix_) { o.ix_ = nullptr; }
    HNSW& operator=(HNSW&& o) noexcept {
        if (this != &o) {
            if (ix_) hnswb200_index_destroy(ix_);
            params = o.params; ctx_ = o.ctx_; ix_ = o.ix_; o.ix_ = nullptr;
        }
        return *this;
    }
    HNSW(const HNSW&) = delete;
    ~HNSW() { if (ix_) hnswb200_index_destroy(ix_); }

    size_t len() const { return (size_t)hnswb200_index_len(ix_); }  // template.rs:146-148

    // template.rs:388-444: consumes the index, returns it.  nb_threads == 1 is the reference's deterministic
    // order; > 1 inserts batches concurrently against a frozen snapshot of the graph.
    HNSW insert_bulk(const std::vector<std::vector<float>>& vectors, size_t nb_threads, bool /*verbose*/) && {
        std::vector<float> flat = flatten(vectors);
        check(hnswb200_index_insert_bulk(ctx_->get(), ix_, flat.data(), vectors.size(), (uint32_t)params.dim, nullptr,
                                         nb_threads <= 1 ? 1u : 0u));
        refresh();
        return std::move(*this);
    }
    // template.rs:165-173
    NodeID insert_vec(const std::vector<float>& v) {
        if (v.size() != params.dim) dim_panic(v.size());
        uint32_t id = 0;
        check(hnswb200_index_insert_vec(ctx_->get(), ix_, v.data(), (uint32_t)v.size(), &id));
        refresh();
        return id;
    }
    // template.rs:306-335: <= n ids in ascending (dist, id) order
    std::vector<NodeID> ann_by_vector(const std::vector<float>& v, size_t n, size_t ef) const {
        return ann_batch({v}, n, ef)[0];
    }
    // many queries per call: the device boundary sits here
    std::vector<std::vector<NodeID>> ann_batch(const std::vector<std::vector<float>>& queries, size_t n, size_t ef) const {
        std::vector<float> flat = flatten(queries);
        std::vector<uint32_t> ids(queries.size() * n, HNSWB200_NO_ID), counts(queries.size());
        check(hnswb200_search(ctx_->get(), ix_, flat.data(), queries.size(), (uint32_t)params.dim, (uint32_t)n, (uint32_t)ef,
                              ids.data(), nullptr, counts.data(), nullptr));
        std::vector<std::vector<NodeID>> out(queries.size());
        for (size_t q = 0; q < queries.size(); ++q) out[q].assign(ids.begin() + q * n, ids.begin() + q * n + counts[q]);
        return out;
    }
    // template.rs:150-152
    std::optional<float> distance(NodeID a, NodeID b) const {
        if (a >= len() || b >= len()) return std::nullopt;
        float out = 0.f;
        check(hnswb200_dist_pairs(ctx_->get(), hnswb200_index_points(ix_), &a, &b, 1, &out));
        return out;
    }
    size_t nb_layers() const { return hnswb200_graph_nb_layers(hnswb200_index_graph(ix_)); }
    // template.rs:192-194
    graph::Graph get_layer(size_t layer_nb) const {
        const hnswb200_graph* g = hnswb200_index_graph(ix_);
        if (layer_nb >= nb_layers()) throw Panic("Layer " + std::to_string(layer_nb) + " not found in the structure.");  // layers.rs:28
        const uint64_t nn = hnswb200_graph_layer_nb_nodes(g, (uint32_t)layer_nb), ne = hnswb200_graph_layer_nb_edges(g, (uint32_t)layer_nb);
        std::vector<uint32_t> ids(nn), nb(ne);
        std::vector<uint64_t> off(nn + 1);
        check(hnswb200_graph_export_layer(g, (uint32_t)layer_nb, ids.data(), off.data(), nb.data()));
        graph::Graph out;
        out.level = layer_nb;
        out.m = hnswb200_graph_layer_cap(g, (uint32_t)layer_nb);
        for (uint64_t r = 0; r < nn; ++r) out.nodes[ids[r]] = std::set<NodeID>(nb.begin() + off[r], nb.begin() + off[r + 1]);
        return out;
    }
    // template.rs:341-370
    bool assert_param_compliance() const {
        bool ok = true;
        for (size_t l = 0; l < nb_layers(); ++l) {
            const size_t max_degree = l > 0 ? params.mmax : params.mmax0;
            const size_t lim = (size_t)std::ceil((float)max_degree * 1.1f);
            graph::Graph g = get_layer(l);
            for (const auto& kv : g.nodes) {
                if (kv.second.size() > lim) ok = false;
                if (kv.second.empty() && g.nb_nodes() > 1) ok = false;
            }
        }
        return ok;
    }
    // template.rs:43-131: the reference's directory layout and big-endian formats
    void save(const std::string& dir) const { check(hnswb200_index_save_dir(ctx_->get(), ix_, dir.c_str())); }
    static HNSW load(const std::string& dir, Context& c = Context::global()) {
        hnswb200_index* ix = nullptr;
        check(hnswb200_index_load_dir(c.get(), dir.c_str(), &ix));
        return HNSW(&c, ix);
    }
    const hnswb200_index* handle() const { return ix_; }
    Context& context() const { return *ctx_; }

   private:
    HNSW(Context* c, hnswb200_index* ix) : ctx_(c), ix_(ix) { refresh(); }
    void refresh() {
        hnswb200_params p;
        check(hnswb200_index_params(ix_, &p));
        params = Params::from_c(p);
    }
    [[noreturn]] void dim_panic(size_t got) const {
        throw Panic("The current index dimension is " + std::to_string(params.dim) + ", but tried inserting points of dimension " +
                    std::to_string(got));  // template.rs:257
    }
    std::vector<float> flatten(const std::vector<std::vector<float>>& rows) const {
        std::vector<float> flat;
        flat.reserve(rows.size() * params.dim);
        for (const auto& r : rows) {
            if (r.size() != params.dim) dim_panic(r.size());
            flat.insert(flat.end(), r.begin(), r.end());
        }
        return flat;
    }
    Context* ctx_;
    hnswb200_index* ix_;
};

namespace helpers {
// helpers/glove.rs:14-71
inline std::vector<std::vector<float>> load_glove_array(size_t lim, const std::string& path) {
    uint64_t dim = 0;
    int64_t rows = hnswb200_load_glove(path.c_str(), lim, nullptr, 0, &dim);
    if (rows < 0) throw Error((int)rows, hnswb200_last_error());
    std::vector<float> flat((size_t)rows * dim);
    hnswb200_load_glove(path.c_str(), lim, flat.data(), flat.size(), &dim);
    std::vector<std::vector<float>> out((size_t)rows);
    for (int64_t r = 0; r < rows; ++r) out[r].assign(flat.begin() + r * dim, flat.begin() + (r + 1) * dim);
    return out;
}
// helpers/glove.rs:73-109: exact top-k of every query under the quantised metric, (dist, id) ties
inline std::vector<std::vector<NodeID>> brute_force_nns(size_t nb_nns, const HNSW& index, const std::vector<std::vector<float>>& queries) {
    std::vector<float> flat;
    for (const auto& q : queries) flat.insert(flat.end(), q.begin(), q.end());
    std::vector<uint32_t> ids(queries.size() * nb_nns, HNSWB200_NO_ID);
    check(hnswb200_bruteforce_topk(index.context().get(), hnswb200_index_points(index.handle()), flat.data(), queries.size(),
                                   (uint32_t)nb_nns, 0, ids.data(), nullptr));
    std::vector<std::vector<NodeID>> out(queries.size());
    for (size_t q = 0; q < queries.size(); ++q)
        for (size_t j = 0; j < nb_nns; ++j)
            if (ids[q * nb_nns + j] != HNSWB200_NO_ID) out[q].push_back(ids[q * nb_nns + j]);
    return out;
}
}  // namespace helpers

}  // namespace hnsw
}  // namespace hnsw_rs
