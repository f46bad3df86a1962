// The reference's own unit / integration tests for the hot path, restated against the C++ host mirror
// (include/hnsw_rs.hpp) so that they read like the originals:
//   vectors/src/full.rs:78-147   vectors/src/quant.rs:133-202   vectors/tests/full_lvq_tests.rs:3-27
//   hnsw/src/template.rs:465-611 (hnsw_init, hnsw_build, insert one/many after build, can_not_add_different_dim,
//                                 hnsw_glove_build_eval, hnsw_serialize)
// usage: reference_suite <dir with store.txt and queries.txt> [--list]      exit code = number of failed tests
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <random>
#include <set>
#include <string>

#include "hnsw_rs.hpp"

using namespace hnsw_rs;
using hnsw_rs::hnsw::HNSW;
using hnsw_rs::vectors::FullVec;
using hnsw_rs::vectors::QuantVec;

static int g_failed = 0;
#define ASSERT(c)                                                                  \
    do {                                                                           \
        if (!(c)) { std::printf("    assertion failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); throw 1; } \
    } while (0)

static std::mt19937 g_rng(12345);
static std::vector<std::vector<float>> gen_rand_vecs(size_t dim, size_t n) {  // vectors/src/lib.rs:29-37
    std::uniform_real_distribution<float> u(0.0f, 1.0f);
    std::vector<std::vector<float>> out(n, std::vector<float>(dim));
    for (auto& v : out)
        for (auto& x : v) x = u(g_rng);
    return out;
}
static std::vector<std::vector<float>> make_rand_vectors(size_t n, size_t dim) { return gen_rand_vecs(dim, n); }  // template.rs:630-638

template <class V>
static void distance_kats() {
    std::vector<V> others;
    for (int i = 0; i < 100; ++i) others.push_back(V::new_(gen_rand_vecs(128, 1)[0]));
    V a = V::new_(gen_rand_vecs(128, 1)[0]);
    for (float d : a.dist2many(others)) ASSERT(d >= 0.0f);
    auto kat = [](std::vector<float> x, std::vector<float> y, float want) {
        V p = V::new_(x), q = V::new_(y);
        float dist = p.dist2other(q), dist2other = p.dist2other(q);
        ASSERT(dist == want);
        ASSERT(dist == dist2other);
    };
    kat({0.5f}, {0.25f}, 0.25f);
    kat({0.75f}, {0.25f}, 0.5f);
    kat({0.0f, 0.0f}, {0.0f, 1.0f}, 1.0f);
    kat({1.0f, 0.0f}, {0.0f, 1.0f}, std::sqrt(2.0f));
    kat({-1.0f, 0.0f}, {0.0f, 1.0f}, std::sqrt(2.0f));
    kat({1.0f, 0.0f}, {0.0f, -1.0f}, std::sqrt(2.0f));
    V c = V::new_(gen_rand_vecs(128, 1)[0]);
    V d = c;
    ASSERT(c.dist2other(d) == 0.0f);
}

static void full_distance() { distance_kats<FullVec>(); }    // full.rs:88-147
static void quant_distance() { distance_kats<QuantVec>(); }  // quant.rs:143-202

static void full_serialization() {  // full.rs:78-86
    FullVec a = FullVec::new_(gen_rand_vecs(128, 1)[0]);
    FullVec b = FullVec::deserialize(a.serialize());
    ASSERT(a.get_vals() == b.get_vals());
}
static void quant_serialization() {  // quant.rs:133-141
    QuantVec a = QuantVec::new_(gen_rand_vecs(128, 1)[0]);
    QuantVec b = QuantVec::deserialize(a.serialize());
    ASSERT(a.get_vals() == b.get_vals());
    ASSERT(a.size() == 8 + 128);
}
static void dist_err_lt_one_percent() {  // vectors/tests/full_lvq_tests.rs:3-27 (quant <-> quant leg)
    for (int i = 0; i < 200; ++i) {
        auto v = gen_rand_vecs(128, 2);
        float d_full = FullVec::new_(v[0]).distance(FullVec::new_(v[1]));
        float d_quant = QuantVec::new_(v[0]).dist2other(QuantVec::new_(v[1]));
        ASSERT(std::fabs(d_quant - d_full) / d_full < 0.01f);
    }
}

static const size_t DIM = 10, N = 100, M = 12;
static void hnsw_init() { HNSW index = HNSW::new_(12, std::nullopt, 128); (void)index; }
static void hnsw_build() {
    HNSW index = HNSW::new_(12, std::nullopt, DIM).insert_bulk(make_rand_vectors(N, DIM), 1, false);
    ASSERT(index.len() == N);
}
static void hnsw_insert_one_after_build() {
    HNSW index = HNSW::new_(12, std::nullopt, DIM).insert_bulk(make_rand_vectors(N, DIM), 1, false);
    ASSERT(index.len() == N);
    index.insert_vec(make_rand_vectors(1, DIM)[0]);
    ASSERT(index.len() == N + 1);
}
static void hnsw_insert_many_after_build() {
    HNSW index = HNSW::new_(12, std::nullopt, DIM).insert_bulk(make_rand_vectors(N, DIM), 1, false);
    ASSERT(index.len() == N);
    index = std::move(index).insert_bulk(make_rand_vectors(N, DIM), 1, false);
    ASSERT(index.len() == N * 2);
}
static void can_not_add_different_dim() {  // #[should_panic]
    HNSW index = HNSW::new_(12, std::nullopt, 128).insert_bulk(make_rand_vectors(10, 128), 1, false);
    bool panicked = false;
    try {
        index = std::move(index).insert_bulk(make_rand_vectors(10, 512), 1, false);
    } catch (const Panic&) {
        panicked = true;
    }
    ASSERT(panicked);
}
static std::string g_data;
static int g_vec_type = HNSWB200_VEC_QUANT;  // the suite's index tests run once per VecType (points/src/point.rs:4)
static void hnsw_glove_build_eval() {
    auto stored = hnsw_rs::hnsw::helpers::load_glove_array(0, g_data + "/store.txt");
    auto queries = hnsw_rs::hnsw::helpers::load_glove_array(0, g_data + "/queries.txt");
    ASSERT(stored.size() == 1000 && queries.size() == 100);
    HNSW index = HNSW::new_(M, std::nullopt, queries[0].size(), Context::global(), HNSWB200_METRIC_L2, g_vec_type)
                     .insert_bulk(stored, 1, false);
    ASSERT(index.vec_type() == g_vec_type);
    // ground truth: all distance2point values sorted by Dist, first 10 ids (template.rs:531-541)
    auto queries_nn = hnsw_rs::hnsw::helpers::brute_force_nns(10, index, queries);
    size_t total_hits = 0;
    for (size_t i = 0; i < queries.size(); ++i) {
        auto ann = index.ann_by_vector(queries[i], 10, 100);
        std::set<NodeID> a(ann.begin(), ann.end()), t(queries_nn[i].begin(), queries_nn[i].end());
        for (NodeID x : t) total_hits += a.count(x);
    }
    float final_acc = (float)total_hits / (float)(queries.size() * 10);
    std::printf("    Final accuracy was %f\n", final_acc);
    ASSERT(final_acc > 0.99f);
    for (size_t l = 0; l < index.nb_layers(); ++l) {
        auto layer = index.get_layer(l);
        if (layer.nb_nodes() <= 1) continue;
        size_t min_degree = SIZE_MAX;
        for (const auto& kv : layer.nodes) min_degree = std::min(min_degree, kv.second.size());
        ASSERT(min_degree > 0);
    }
    ASSERT(index.assert_param_compliance());
    ASSERT(index.ann_by_vector(queries[0], 10, 5).size() == 5);  // ef < n returns ef ids (results.rs:59-61)
}
static void hnsw_serialize() {
    for (int rep = 0; rep < 5; ++rep) {
        HNSW index = HNSW::new_(12, std::nullopt, DIM, Context::global(), HNSWB200_METRIC_L2, g_vec_type)
                         .insert_bulk(make_rand_vectors(N, DIM), 1, false);
        std::string path = "/tmp/hnsw_rs_serialization_test";
        std::string rm = "rm -rf " + path;
        (void)!std::system(rm.c_str());
        index.save(path);
        HNSW loaded = HNSW::load(path);
        (void)!std::system(rm.c_str());
        ASSERT(loaded.len() == N && loaded.vec_type() == g_vec_type);
        auto a = index.get_layer(0), b = loaded.get_layer(0);
        ASSERT(a.nodes == b.nodes);
        for (NodeID i = 0; i + 1 < (NodeID)N; ++i) ASSERT(*index.distance(i, i + 1) == *loaded.distance(i, i + 1));
        ASSERT(!index.distance(0, (NodeID)N + 7).has_value());
    }
}

int main(int argc, char** argv) {
    struct T { const char* name; std::function<void()> fn; };
    std::vector<T> tests = {
        {"vectors::full::distance", full_distance}, {"vectors::full::serialization", full_serialization},
        {"vectors::quant::distance", quant_distance}, {"vectors::quant::serialization", quant_serialization},
        {"vectors::full_lvq_tests::dist_err_lt_one_percent", dist_err_lt_one_percent},
        {"hnsw::hnsw_init", hnsw_init}, {"hnsw::hnsw_build", hnsw_build},
        {"hnsw::hnsw_insert_one_after_build", hnsw_insert_one_after_build},
        {"hnsw::hnsw_insert_many_after_build", hnsw_insert_many_after_build},
        {"hnsw::can_not_add_different_dim", can_not_add_different_dim},
        {"hnsw::hnsw_glove_build_eval", hnsw_glove_build_eval}, {"hnsw::hnsw_serialize", hnsw_serialize},
        // the same two with `type VecType = FullVec;`
        {"hnsw(FullVec)::hnsw_glove_build_eval", [] { g_vec_type = HNSWB200_VEC_FULL; hnsw_glove_build_eval(); g_vec_type = HNSWB200_VEC_QUANT; }},
        {"hnsw(FullVec)::hnsw_serialize", [] { g_vec_type = HNSWB200_VEC_FULL; hnsw_serialize(); g_vec_type = HNSWB200_VEC_QUANT; }},
    };
    if (argc > 2 && std::string(argv[2]) == "--list") {
        for (auto& t : tests) std::printf("%s\n", t.name);
        return 0;
    }
    if (argc < 2) { std::printf("usage: reference_suite <data dir> [--list]\n"); return 100; }
    g_data = argv[1];
    for (auto& t : tests) {
        std::printf("test %s ... ", t.name);
        std::fflush(stdout);
        try {
            t.fn();
            std::printf("ok\n");
        } catch (const std::exception& e) {
            std::printf("FAILED: %s\n", e.what());
            ++g_failed;
        } catch (...) {
            std::printf("FAILED\n");
            ++g_failed;
        }
    }
    std::printf("test result: %s. %zu passed; %d failed\n", g_failed ? "FAILED" : "ok", tests.size() - g_failed, g_failed);
    return g_failed;
}
