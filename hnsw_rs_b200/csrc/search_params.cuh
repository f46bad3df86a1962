// Parameters of the HNSW search kernels (csrc/kernels.cu: register-list and shared-memory-list variants;
// csrc/search_fast.cu: the bucketed-visited-set variant) and the result stores they share.
#pragma once
#include "kernels.h"
#include "search.cuh"

namespace hb {

struct SearchParams {
    const uint8_t* rec;
    RecLayout L;
    GraphView g;
    uint32_t n_layers, ep;
    const float* queries;
    const float* queries_tail;  // queries [split, nq) are read from here (same indexing); == queries when not split
    uint32_t split;
    uint32_t nq, topn, ef;
    uint32_t kpl;           // keys per lane of the result list (capacity 32*kpl >= ef)
    uint32_t tbits, bbits;  // visited table: 2^tbits entries; ids < 2^bbits
    uint32_t qd_cap;
    uint32_t* out_ids;
    float* out_dists;
    uint32_t* out_counts;
    uint32_t* out_hops;
    uint32_t* out_evals;
    uint32_t* out_flags;
    uint32_t* out_nbrs;
    uint32_t* work_counter;
    uint32_t* nan_any;  // may be null; set to 1 when a query holds a NaN (may live in pinned host memory)
    // fused all-gather of the id rows: row (peer_row0 + q) of every peer buffer also receives the ids of query q
    // (peer memory mapped into this device: stores travel over NVLink while the other queries keep computing)
    uint32_t* peer_ids[HB_MAX_PEERS];
    float* peer_dists[HB_MAX_PEERS];  // optional (base shards: the merge needs the distances too); all null or all set
    uint32_t n_peers;
    uint64_t peer_row0;
    uint32_t id_offset;  // added to every id written (global ids of a base shard); 0 for a whole index
};

// one result slot: id (+ id_offset unless padding) and distance, locally and to every peer buffer
__device__ __forceinline__ void put_result(const SearchParams& p, uint32_t* oid, float* od, uint32_t qi, uint32_t j,
                                           uint32_t id, float d) {
    const uint32_t v = id == EMPTY_ID ? id : id + p.id_offset;
    oid[j] = v;
    if (od) od[j] = d;
    for (uint32_t g = 0; g < p.n_peers; ++g) {
        const size_t at = (p.peer_row0 + qi) * p.topn + j;
        p.peer_ids[g][at] = v;
        if (p.peer_dists[g]) p.peer_dists[g][at] = d;
    }
}

constexpr int SEARCH_WPB = 4;

// csrc/search_fast.cu: the second-generation kernel and the name of the variant the last launch_search of this
// thread ran (hnswb200_last_search_variant)
bool search_fast_supported(const RecLayout& L, uint32_t ef, uint64_t n_points);
// spill_ws: optional global continuation of the visited set's spill list, spill_cap ids for each of spill_warps warps
cudaError_t launch_search_fast(const SearchParams& p, uint64_t n_points, int num_sms, cudaStream_t st, bool overlap_previous,
                               uint32_t* spill_ws, uint32_t spill_cap, uint32_t spill_warps);
const char* last_search_variant();
void set_search_variant(const char* s);

}  // namespace hb
