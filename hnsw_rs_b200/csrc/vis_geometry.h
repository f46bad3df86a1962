// Geometry of the bucketed visited set of search_kernel_fast (csrc/search_fast.cuh, VisB4): which bucket an id goes to and
// which 15-bit value stands for it there.  Shared by the launcher (host), the kernel (device) and the host test
// tests/cpp/visgeom_test.cpp, which checks exhaustively that (bucket, entry) names the id exactly -- the property that makes
// results.insert_visited (hnsw/src/template/results.rs:101-103) exact without storing the id itself.
//
//   h   = id * odd  mod 2^B                       a bijection on B-bit ids (ids < 2^B, B <= 21)
//   h32 = h << (32 - B)                           the same value top-aligned in 32 bits: id * mul, mul = odd << (32 - B)
//   home  = floor(h32 * nb / 2^32) = floor(h * nb / 2^B)               the bucket, 0 <= home < nb
//   entry = bits [s, B) of (h * nb) mod 2^B, shifted up by the displacement bits   (2^s < nb, s = 9 or 8)
// Two different ids have different h; if they share a bucket, their products h * nb lie in the same window of width 2^B and
// differ by a multiple of nb, i.e. by more than 2^s -- so bits [s, B) of (h * nb) mod 2^B differ.  The low `dbits` bits of
// an entry hold the displacement (how many buckets behind its home the entry sits); B - s + dbits <= 15.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HB_VG_HD __host__ __device__ __forceinline__
#else
#define HB_VG_HD inline
#endif

struct FastVisGeometry {
    uint32_t mul;   // odd << (32 - B)
    uint32_t rsh;   // (h32 * nb) >> rsh, displacement bits cleared = the entry with displacement 0
    uint32_t dmax;  // largest displacement an entry can hold: 2^dbits - 1
};

// bbits = B: ids < 2^B (10 <= B <= 21); nb: buckets (256 < nb <= 1024)
HB_VG_HD FastVisGeometry fast_vis_geometry(uint32_t bbits, uint32_t nb) {
    const uint32_t sbits = nb > 512 ? 9u : 8u;  // 2^s < nb keeps two ids of one bucket apart
    const uint32_t rembits = bbits > sbits ? bbits - sbits : 0u;
    const uint32_t dbits = rembits <= 12 ? 3u : 15u - rembits;  // >= 2 for B <= 21
    FastVisGeometry g;
    g.mul = 0x9E3779B1u << (32u - bbits);
    g.rsh = 32u - bbits + sbits - dbits;
    g.dmax = (1u << dbits) - 1u;
    return g;
}

// home bucket and entry value (displacement 0) of an id
HB_VG_HD void fast_vis_slot(uint32_t mul, uint32_t rsh, uint32_t dmax, uint32_t nb, uint32_t id, uint32_t& home, uint32_t& mine0) {
    const uint32_t h32 = id * mul;
#if defined(__CUDA_ARCH__)
    home = __umulhi(h32, nb);
#else
    home = (uint32_t)(((uint64_t)h32 * nb) >> 32);
#endif
    mine0 = ((h32 * nb) >> rsh) & (0x7FFFu & ~dmax);
}
