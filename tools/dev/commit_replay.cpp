// Host-only timing of the build's commit step (hnsw_rs_b200/csrc/commit.h) on a synthetic stream of
// insertion results: points on a line, each selecting its m nearest already-inserted points, so the
// row contents, fill levels and prune / cut rates resemble a real build (rows fill up, every new edge
// to a full row prunes it).   g++ -O2 -std=c++17 -o /tmp/commit_replay tools/dev/commit_replay.cpp
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>

#include "../../hnsw_rs_b200/csrc/commit.h"

using namespace hb;

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? atoi(argv[1]) : 1000000, m = argc > 2 ? atoi(argv[2]) : 16;
    HostGraph h;
    h.init(m, 2 * m, m);
    std::mt19937_64 rng(1);
    std::vector<float> x(n);
    for (auto& v : x) v = (float)(rng() >> 40) / (float)(1 << 24);
    std::vector<uint8_t> lv(n);
    const float ml = 1.0f / logf((float)m);
    for (auto& l : lv) { float r = ((rng() >> 40) + 1) / (float)((1 << 24) + 1); l = (uint8_t)floorf(-logf(r) * ml); }
    for (uint32_t i = 0; i < n; ++i) h.add_node(lv[i]);
    const uint32_t nl = h.n_layers();
    // per layer: ordered map coordinate -> id of inserted points
    std::vector<std::multimap<float, uint32_t>> line(nl);
    // selections precomputed in batches so that the map walk is outside the timed part
    const uint32_t B = 4096;
    std::vector<uint32_t> o_ids((size_t)B * nl * m), o_cnt((size_t)B * nl);
    std::vector<float> o_d((size_t)B * nl * m);
    std::vector<uint32_t> d0, du;
    std::vector<LayerSel> res;
    CommitScratch cs;
    double t_commit = 0, t_stage = 0;
    uint64_t staged = 0;
    std::vector<uint8_t> mark;
    std::vector<uint32_t> stage;
    const size_t PF = argc > 3 ? atoi(argv[3]) : 8;
    uint64_t edges = 0;
    for (uint32_t pos = 0; pos < n; pos += B) {
        uint32_t nb = std::min(B, n - pos);
        for (uint32_t j = 0; j < nb; ++j) {
            uint32_t pid = pos + j;
            for (uint32_t l = 0; l < nl; ++l) {
                uint32_t cnt = 0;
                if (l <= lv[pid]) {
                    auto hi = line[l].lower_bound(x[pid]);
                    auto lo = hi;
                    std::vector<std::pair<float, uint32_t>> c;
                    for (uint32_t k = 0; k < m && hi != line[l].end(); ++k, ++hi) c.push_back({fabsf(hi->first - x[pid]), hi->second});
                    for (uint32_t k = 0; k < m && lo != line[l].begin(); ++k) { --lo; c.push_back({fabsf(lo->first - x[pid]), lo->second}); }
                    std::sort(c.begin(), c.end());
                    for (auto& e : c) {
                        if (cnt == m) break;
                        o_ids[((size_t)j * nl + l) * m + cnt] = e.second;
                        o_d[((size_t)j * nl + l) * m + cnt] = e.first;
                        ++cnt;
                    }
                }
                o_cnt[(size_t)j * nl + l] = cnt;
            }
        }
        auto t0 = std::chrono::steady_clock::now();
        for (uint32_t j = 0; j < nb; ++j) {
            uint32_t pid = pos + j;
            res.clear();
            for (uint32_t l = 0; l < nl; ++l) {
                uint32_t cnt = o_cnt[(size_t)j * nl + l];
                if (cnt) res.push_back(LayerSel{l, cnt, &o_ids[((size_t)j * nl + l) * m], &o_d[((size_t)j * nl + l) * m]});
                edges += cnt;
            }
            const char* err = nullptr;
            if (commit_point(h, pid, res, d0, du, cs, &err)) { printf("error: %s\n", err); return 1; }
        }
        t_commit += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        {   // the staging half of hnswb200_graph::upload_rows (api.cu): de-duplicate, gather rows
            auto t1 = std::chrono::steady_clock::now();
            const AdjStore& s = h.a0;
            if (mark.size() < s.rows()) mark.resize(s.rows(), 0);
            size_t uniq = 0;
            for (uint32_t r : d0)
                if (!mark[r]) { mark[r] = 1; d0[uniq++] = r; }
            d0.resize(uniq);
            for (uint32_t r : d0) mark[r] = 0;
            if (stage.size() < (uniq + 64) * s.S) stage.resize((uniq + 64) * s.S * 2);
            size_t cnt = 0;
            for (size_t di = 0; di < d0.size(); ++di) {
                const uint32_t r = d0[di];
                if (PF && di + PF < d0.size()) {
                    const char* nx = (const char*)&s.data[(size_t)d0[di + PF] * s.S];
                    for (uint32_t b = 0; b < s.S * 4; b += 64) __builtin_prefetch(nx + b);
                    __builtin_prefetch(&s.deg[d0[di + PF]]);
                }
                if (s.deg[r] <= s.S) { memcpy(&stage[cnt * s.S], &s.data[(size_t)r * s.S], s.S * 4); ++cnt; }
            }
            staged += cnt;
            t_stage += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
        }
        d0.clear();
        du.clear();
        for (uint32_t j = 0; j < nb; ++j)
            for (uint32_t l = 0; l <= lv[pos + j]; ++l) line[l].insert({x[pos + j], pos + j});
    }
    uint64_t deg = 0, wide = 0;
    for (uint32_t i = 0; i < n; ++i) { deg += h.a0.deg[i]; wide += h.a0.deg[i] > h.a0.S; }
    printf("n=%u m=%u commit=%.3fs  %.2f us/point  edges offered=%llu  mean degree layer0=%.2f rows wider than S=%llu spill rows=%zu\n", n, m, t_commit,
           t_commit / n * 1e6, (unsigned long long)edges, (double)deg / n, (unsigned long long)wide, h.a0.spill.size());
    printf("stage=%.3fs rows staged=%llu (%.1f ns/row) prefetch distance %zu\n", t_stage, (unsigned long long)staged, t_stage / staged * 1e9, PF);
    return 0;
}
