// K3, second generation: the kernel around search_query_fast (csrc/search_fast.cuh) and its launcher.
// HNSW::ann_by_vector (hnsw/src/template.rs:306-335) for a batch: one warp per query, persistent over the batch.
#include <stdio.h>
#include <stdlib.h>

#include "search_fast.cuh"

namespace hb {

// per warp: 32 candidate ids | spill list | worst key | scratch (accumulators / admitted keys + merge buffer / query) | visited buckets
__host__ __device__ inline size_t fast_warp_smem(uint32_t nb) { return (size_t)FAST_OFF_TABLE + (size_t)nb * 8; }

#ifndef HB_FAST_OPAQUE
#define HB_FAST_OPAQUE 1
#endif

template <class Q, int KPL, bool STATS, int MINB>
__global__ void __launch_bounds__(SEARCH_WPB * 32, MINB) search_kernel_fast(SearchParams p, uint32_t nb, uint32_t vmul,
                                                                           uint32_t vrsh, uint32_t* spill_ws,
                                                                           uint32_t spill_cap, uint32_t vdmax) {
    extern __shared__ __align__(16) unsigned char smem[];
#if HB_FAST_OPAQUE
    // lane and the warp's shared-memory offset are made opaque so that the compiler keeps them in registers instead of
    // recomputing them from %tid (S2R + 8 instructions, ~3 % of the executed instructions) all over the loop
    int lane = threadIdx.x & 31;
    uint32_t woff = (uint32_t)((threadIdx.x >> 5) * fast_warp_smem(nb));
    asm volatile("" : "+r"(lane), "+r"(woff));
    const int gl = lane & 3;
    unsigned char* wsm = smem + woff;
#else
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gl = lane & 3;
    unsigned char* wsm = smem + (size_t)wib * fast_warp_smem(nb);
#endif
    VisB4 vis;
    vis.sbase = (uint32_t)__cvta_generic_to_shared(wsm) + FAST_OFF_TABLE;
    vis.nb = nb;
    vis.mul = vmul;
    vis.rsh = vrsh;
    vis.dmax = vdmax;
    vis.spill = reinterpret_cast<uint32_t*>(wsm + FAST_OFF_SPILL);
    float* scratch = reinterpret_cast<float*>(wsm + FAST_OFF_SCRATCH);
    float* qd = scratch;
    // A following search (launched as programmatic dependent) reads nothing this grid writes: let its blocks
    // take over the SMs as soon as this grid's blocks retire, instead of waiting for the last long query.
    asm volatile("griddepcontrol.launch_dependents;");

    while (true) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(p.work_counter, 1u);
        qi = __shfl_sync(HB_FULL, qi, 0);
        if (qi >= p.nq) break;
        __syncwarp();
        // Point::new(vector): the query becomes a point exactly like a stored one (template.rs:313)
        // one read of the f32 query (it may live in pinned host memory: see hnswb200_search), then quantise in place
        {
            const float* src = (qi < p.split ? p.queries : p.queries_tail) + (size_t)qi * p.L.dim;
            for (uint32_t i = lane; i < p.L.dim; i += 32) qd[i] = src[i];
        }
        __syncwarp();
        // QuantVec::new on the query (quant.rs:41-66): dequantised values in place, the codes next to them
        float qmn, qdl;
        uint8_t* ctmp = reinterpret_cast<uint8_t*>(scratch) + 1024;  // natural-order codes (dim <= 128 bytes); qd ends below
        const bool ok = warp_quantise(qd, p.L.dim, lane, qd, ctmp, qmn, qdl);
        __syncwarp();
        uint32_t* oid = p.out_ids + (size_t)qi * p.topn;
        float* od = p.out_dists ? p.out_dists + (size_t)qi * p.topn : nullptr;
        if (!ok) {  // NaN in query: the reference panics; report through flags and an empty result
            for (uint32_t j = lane; j < p.topn; j += 32) put_result(p, oid, od, qi, j, EMPTY_ID, INFINITY);
            if (lane == 0) {
                if (p.out_counts) p.out_counts[qi] = 0;
                if (p.out_hops) p.out_hops[qi] = 0;
                if (p.out_evals) p.out_evals[qi] = 0;
                if (p.out_flags) p.out_flags[qi] = 1u;
                if (p.out_nbrs) p.out_nbrs[qi] = 0;
                if (p.nan_any) *reinterpret_cast<volatile uint32_t*>(p.nan_any) = 1u;
            }
            continue;
        }
        Q q;
        q.init(p.L, qd, gl, reinterpret_cast<u64*>(wsm + FAST_OFF_QTAB), lane);
        if (Q::kAux) q.init_filter(p.L, qd, ctmp, qmn, qdl, wsm + FAST_OFF_QCODE, gl, lane);
        SearchCounters cnt{0u, 0u, 0u, 0u};
        RegList<KPL> L;
        __syncwarp();  // every lane has copied its part of qd before the scratch area is reused
        SpillPool pool;
        pool.base = spill_ws;
        pool.npool = spill_cap >> 16;      // packed by the launcher: slices << 16 | log2(ids per slice)
        pool.cap = 1u << (spill_cap & 31u);
        search_query_fast<Q, KPL, STATS>(q, p.rec, p.L.stride, p.g, p.n_layers, p.ep, L, vis, wsm, pool, (int)p.ef, lane, cnt);
        // get_top_selected(n)   (results.rs:59-61): position lane*KPL + s
        uint32_t mine = 0;
        uint32_t* sid = reinterpret_cast<uint32_t*>(scratch);  // staged row for the peer stores
        float* sd = scratch + 32 * KPL;
#pragma unroll
        for (int s = 0; s < KPL; ++s) {
            const uint32_t j = (uint32_t)(lane * KPL + s);
            const u64 k = L.v[s];
            const bool real = (j < p.ef) && (k != RSENT);
            const uint32_t v = real ? rkey_id(k) + p.id_offset : EMPTY_ID;
            const float d = real ? __uint_as_float((uint32_t)(k >> 32)) : INFINITY;
            if (j < p.topn) {
                oid[j] = v;
                if (od) od[j] = d;
                mine += real ? 1u : 0u;
            }
            if (p.n_peers) { sid[j] = v; sd[j] = d; }
        }
        for (uint32_t j = 32 * KPL + lane; j < p.topn; j += 32) {
            oid[j] = EMPTY_ID;
            if (od) od[j] = INFINITY;
        }
        if (p.n_peers) {
            // fused all-gather: one result slot per lane, all peers in one sweep (row peer_row0 + qi of every buffer)
            __syncwarp();
            const uint32_t total = p.n_peers * p.topn;
            for (uint32_t t = lane; t < total; t += 32) {
                const uint32_t gq = t / p.topn, j = t - gq * p.topn;
                const size_t at = (p.peer_row0 + qi) * p.topn + j;
                p.peer_ids[gq][at] = j < 32u * KPL ? sid[j] : EMPTY_ID;
                if (p.peer_dists[gq]) p.peer_dists[gq][at] = j < 32u * KPL ? sd[j] : INFINITY;
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(HB_FULL, mine, o);
        if (lane == 0) {
            if (p.out_counts) p.out_counts[qi] = mine;
            if (p.out_hops) p.out_hops[qi] = cnt.hops;
            if (p.out_evals) p.out_evals[qi] = cnt.evals;
            if (p.out_flags) p.out_flags[qi] = cnt.overflow;  // bit1: counters may over-count; bit2: spill list used
            if (p.out_nbrs) p.out_nbrs[qi] = cnt.nbrs;
        }
    }
}

// resident blocks per SM a variant is compiled for, and the bucket count that lets that many fit:
// MINB * (4 * fast_warp_smem(nb) + 1024 reserved) <= 228 KB
static uint32_t buckets_for(int minb) {
    const size_t per_block = (size_t)233472 / minb - 1024;
    size_t nb = (per_block / SEARCH_WPB - FAST_OFF_TABLE) / 8;
    nb = nb / 4 * 4;
    if (nb > 1024) nb = 1024;
    return (uint32_t)nb;
}

static thread_local char g_variant[160] = "";
const char* last_search_variant() { return g_variant; }
void set_search_variant(const char* s) { snprintf(g_variant, sizeof g_variant, "%s", s); }

template <class Q, int KPL, bool STATS, int MINB>
static cudaError_t launch_fast_s(const SearchParams& p, int num_sms, cudaStream_t st, bool overlap_previous,
                                 uint32_t bbits, const char* qname, uint32_t* spill_ws, uint32_t spill_cap, uint32_t spill_warps) {
    uint32_t nb = buckets_for(MINB);
    if (const char* ev = getenv("HNSWB200_FAST_NB")) {  // test knob: a small table forces the slow path and the spill list
        const uint32_t v = (uint32_t)strtoul(ev, nullptr, 10);
        if (v > 256 && v <= nb && v % 2 == 0) nb = v;
    }
    const size_t smem = fast_warp_smem(nb) * SEARCH_WPB;
    static int occ_cache_d[64] = {};
    static size_t occ_smem_d[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int& occ_cache = occ_cache_d[dev & 63];
    size_t& occ_smem = occ_smem_d[dev & 63];
    cudaError_t e;
    auto kern = search_kernel_fast<Q, KPL, STATS, MINB>;
    if (occ_cache == 0 || occ_smem != smem) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, SEARCH_WPB * 32, smem);
        if (e != cudaSuccess) return e;
        occ_cache = occ < 1 ? 1 : occ;
        occ_smem = smem;
        if (getenv("HNSWB200_DEBUG_LAUNCH"))
            fprintf(stderr, "[hnswb200 search_fast] KPL=%d buckets=%u smem/block=%zu blocks/SM=%d (compiled for %d)\n", KPL,
                    nb, smem, occ_cache, MINB);
    }
    const uint64_t want = ((uint64_t)p.nq + SEARCH_WPB - 1) / SEARCH_WPB;
    int occ = occ_cache;
    if (const char* ev = getenv("HNSWB200_SEARCH_BLOCKS_PER_SM")) {  // experiment knob
        const int v = atoi(ev);
        if (v >= 1 && v < occ) occ = v;
    }
    const uint64_t cap = (uint64_t)num_sms * occ;
    int grid = (int)(want < cap ? want : cap);
    // spill pool geometry packed into one kernel argument: slices << 16 | log2(ids per slice)
    uint32_t lg = 5;
    while ((1u << (lg + 1)) <= spill_cap) ++lg;
    const uint32_t pool_arg = spill_ws && spill_warps && spill_cap >= 32 ? (spill_warps << 16) | lg : 0u;
    char name[160];
    snprintf(name, sizeof name, "hb::search_kernel_fast<FastQuery<%s>,KPL=%d,STATS=%d,blocks/SM=%d,VisB4 %u buckets>", qname, KPL,
             (int)STATS, occ_cache, nb);
    set_search_variant(name);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(SEARCH_WPB * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = overlap_previous ? 1 : 0;
    // entry = bits [s, B) of (h * nb) mod 2^B, then the displacement (csrc/vis_geometry.h; search_fast_supported admits B <= 21 only)
    const FastVisGeometry vg = fast_vis_geometry(bbits, nb);
    return cudaLaunchKernelEx(&cfg, kern, p, nb, vg.mul, vg.rsh, pool_arg ? spill_ws : (uint32_t*)nullptr, pool_arg, vg.dmax);
}

#ifndef HB_FAST_MINB2
#define HB_FAST_MINB2 9  // resident blocks per SM of the ef <= 64 variant: 56 registers, 460 visited buckets (8 blocks x 564 buckets: -3.9 %,
                         // profiles/r02_ab_variants.txt; the smaller table sends 5 % of the C2 queries through the exact spill set)
#endif
#ifndef HB_FAST_MINB4
#define HB_FAST_MINB4 5  // ef <= 128: the visited set of such a query holds 1,000-3,000 ids, so table size counts for more than a sixth block
                         // (C3, 1M x 128, ef = 100: 4.58 M q/s with 5 blocks x 1024 buckets, 4.15 M with 6 x 896, 3.06 M with 7 x 704)
#endif

template <class Q>
static cudaError_t launch_fast_q(const SearchParams& p, int num_sms, cudaStream_t st, bool overlap_previous, uint32_t bbits,
                                 const char* qname, uint32_t* sw, uint32_t sc, uint32_t sn) {
    const bool stats = p.out_hops || p.out_evals || p.out_flags || p.out_nbrs;
    if (p.ef <= 64) {
        if (stats) return launch_fast_s<Q, 2, true, HB_FAST_MINB2>(p, num_sms, st, overlap_previous, bbits, qname, sw, sc, sn);
        return launch_fast_s<Q, 2, false, HB_FAST_MINB2>(p, num_sms, st, overlap_previous, bbits, qname, sw, sc, sn);
    }
    if (stats) return launch_fast_s<Q, 4, true, HB_FAST_MINB4>(p, num_sms, st, overlap_previous, bbits, qname, sw, sc, sn);
    return launch_fast_s<Q, 4, false, HB_FAST_MINB4>(p, num_sms, st, overlap_previous, bbits, qname, sw, sc, sn);
}

// which searches run on this kernel: quantised records of one of the compile-time dimensions, ef <= 128, ids < 2^21
bool search_fast_supported(const RecLayout& L, uint32_t ef, uint64_t n_points) {
    // test / A-B knobs: HNSWB200_NO_FAST, and the knobs that select a variant of the round-1 kernels, run those
    if (getenv("HNSWB200_NO_FAST") || getenv("HNSWB200_VIS_POW2") || getenv("HNSWB200_VIS32") || getenv("HNSWB200_VIS_SLOTS"))
        return false;
    if (L.kind != HB_REC_QUANT || ef > 128 || n_points > (1ull << 21)) return false;
    return L.dim == 100 || L.dim == 128 || L.dim == 96 || L.dim == 50;
}

cudaError_t launch_search_fast(const SearchParams& p, uint64_t n_points, int num_sms, cudaStream_t st, bool overlap_previous,
                               uint32_t* sw, uint32_t sc, uint32_t sn) {
    uint32_t bbits = 10;
    while ((1ull << bbits) < n_points) ++bbits;
    switch (p.L.dim) {
        case 100: return launch_fast_q<FastQuery<12, 4>>(p, num_sms, st, overlap_previous, bbits, "12,4", sw, sc, sn);
        case 128: return launch_fast_q<FastQuery<16, 0>>(p, num_sms, st, overlap_previous, bbits, "16,0", sw, sc, sn);
        case 96: return launch_fast_q<FastQuery<12, 0>>(p, num_sms, st, overlap_previous, bbits, "12,0", sw, sc, sn);
        case 50: return launch_fast_q<FastQuery<6, 2>>(p, num_sms, st, overlap_previous, bbits, "6,2", sw, sc, sn);
    }
    return cudaErrorInvalidValue;
}

}  // namespace hb
