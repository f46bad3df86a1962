// Host-only checks of the record layouts (hnsw_rs_b200/csrc/layout.h): every element of a vector has its own byte(s) inside
// the record, min / delta do not collide with a code, the lane-sliced order is the one the distance kernels assume, and the
// f32 (FullVec) layout is the natural order padded to whole 16-float chunks.
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

#include "../../hnsw_rs_b200/csrc/layout.h"

#define CHECK(c)                                                                  \
    do {                                                                          \
        if (!(c)) { std::printf("FAILED %s (dim %u) at line %d\n", #c, dim, __LINE__); return 1; } \
    } while (0)

int main() {
    for (uint32_t dim = 1; dim <= 520; ++dim) {
        // ---- quantised records ----
        const RecLayout L = hb_make_layout(dim);
        CHECK(L.kind == HB_REC_QUANT && L.dim == dim && L.nch == dim / 8 && L.rem == dim % 8);
        CHECK(L.stride == 64 * L.W + 16 * L.tail && L.stride % 16 == 0);
        std::set<uint32_t> used;
        for (uint32_t i = 0; i < dim; ++i) {
            const uint32_t o = hb_code_offset(L, i);
            CHECK(o < L.stride && used.insert(o).second);
        }
        for (uint32_t b = 0; b < 4; ++b) {
            CHECK(hb_min_offset(L) + b < L.stride && used.insert(hb_min_offset(L) + b).second);
            CHECK(hb_delta_offset(L) + b < L.stride && used.insert(hb_delta_offset(L) + b).second);
        }
        CHECK(hb_min_offset(L) % 4 == 0 && hb_delta_offset(L) % 4 == 0);
        // lane l of a group owns elements 8k + 2l, 8k + 2l + 1 of every full chunk k, consecutive in its own 16-byte words,
        // in increasing k: word j of lane l sits at 16 * (4 j + l)
        for (uint32_t k = 0; k < L.nch; ++k)
            for (uint32_t l = 0; l < 4; ++l)
                for (uint32_t e = 0; e < 2; ++e) {
                    const uint32_t p = 2 * k + e;  // position inside the lane's slice
                    CHECK(hb_code_offset(L, 8 * k + 2 * l + e) == 16 * (4 * (p / 16) + l) + p % 16);
                }
        if (dim == 100 || dim == 96) CHECK(L.stride == 128);
        if (dim == 128) CHECK(L.stride == 144);
        if (dim == 50) CHECK(L.stride == 64);
        // ---- f32 records ----
        const RecLayout F = hb_make_layout_f32(dim);
        CHECK(F.kind == HB_REC_F32 && F.dim == dim && F.W == (dim + 15) / 16 && F.stride == 64 * F.W);
        CHECK(F.stride >= 4 * dim && F.stride < 4 * dim + 64);
        CHECK(hb_make_layout_kind(dim, HB_REC_F32).stride == F.stride && hb_make_layout_kind(dim, HB_REC_QUANT).stride == L.stride);
        if (dim == 100) CHECK(F.stride == 448);
        if (dim == 128) CHECK(F.stride == 512);
        // the shared-memory query buffer of the kernels (round_up(dim, 8) + 8 floats) holds the padded vector
        CHECK((dim + 7) / 8 * 8 + 8 >= 16 * F.W);
    }
    std::printf("all checks passed\n");
    return 0;
}
