// Host-only check of the visited set's geometry (hnsw_rs_b200/csrc/vis_geometry.h, used by VisB4 in csrc/search_fast.cuh):
// for every id below 2^B the pair (home bucket, 15-bit entry) is unique, so "the entry is in the bucket" means "this id was
// visited" with no false positive and no false negative -- results.insert_visited (hnsw/src/template/results.rs:101-103)
// stays exact although the table stores 15 bits per id.  Exhaustive over the ids for every B the kernel admits (10..21)
// and the bucket counts the launcher can produce (every even count for B <= 16, the compiled and test-knob counts above).
#include <cstdio>
#include <cstdint>
#include <vector>

#include "../../hnsw_rs_b200/csrc/vis_geometry.h"

static int check(uint32_t B, uint32_t nb, std::vector<uint64_t>& seen) {
    const FastVisGeometry g = fast_vis_geometry(B, nb);
    if (g.dmax != 3u && g.dmax != 7u) { std::printf("FAILED B=%u nb=%u: dmax %u\n", B, nb, g.dmax); return 1; }
    if ((g.mul >> (32u - B)) % 2u != 1u) { std::printf("FAILED B=%u: multiplier not odd\n", B); return 1; }
    seen.assign(((size_t)nb << 15) / 64, 0ull);
    const uint32_t n = 1u << B;
    for (uint32_t id = 0; id < n; ++id) {
        uint32_t home, mine0;
        fast_vis_slot(g.mul, g.rsh, g.dmax, nb, id, home, mine0);
        // the bucket exists; the entry is a valid one (bit 15 clear: 0xFFFF marks a free entry) with room for the displacement
        if (home >= nb || (mine0 & g.dmax) != 0u || mine0 + g.dmax > 0x7FFFu) {
            std::printf("FAILED B=%u nb=%u id=%u: home %u entry %#x\n", B, nb, id, home, mine0);
            return 1;
        }
        const size_t bit = ((size_t)home << 15) | mine0;
        if (seen[bit >> 6] & (1ull << (bit & 63))) {
            std::printf("FAILED B=%u nb=%u id=%u: (bucket %u, entry %#x) already names another id\n", B, nb, id, home, mine0);
            return 1;
        }
        seen[bit >> 6] |= 1ull << (bit & 63);
    }
    return 0;
}

int main() {
    std::vector<uint64_t> seen;
    unsigned cases = 0;
    for (uint32_t B = 10; B <= 16; ++B)
        for (uint32_t nb = 258; nb <= 1024; nb += 2, ++cases)
            if (check(B, nb, seen)) return 1;
    // compiled bucket counts (460: 9 blocks/SM, 380 / 564 / 576 / 704 / 768 / 896: the A/B builds, 1024: ef <= 128), the test
    // knob's values (258, 300, 514) and the edges of the two entry formats
    const uint32_t nbs[] = {258, 260, 300, 380, 460, 510, 512, 514, 564, 576, 704, 768, 896, 1022, 1024};
    for (uint32_t B = 17; B <= 21; ++B)
        for (uint32_t nb : nbs) {
            if (check(B, nb, seen)) return 1;
            ++cases;
        }
    std::printf("visgeom ok: %u (B, buckets) cases, every id below 2^B has its own (bucket, entry)\n", cases);
    return 0;
}
