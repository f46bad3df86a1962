// Host-callable launchers of the sm_100a kernels (internal C++ interface; the
// public boundary is include/hnsw_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "layout.h"

namespace hb {

struct DevGraph {
    const uint32_t* adj0 = nullptr;
    uint32_t S0 = 0;
    const uint32_t* upper_off = nullptr;
    const uint32_t* upper_adj = nullptr;
    uint32_t SU = 0;
    uint32_t n_layers = 0;
};

constexpr uint32_t HB_MAX_PEERS = 8;  // GPUs of one box

struct SearchLaunch {
    const uint8_t* rec;
    RecLayout L;
    DevGraph g;
    uint32_t ep;
    uint64_t n_points;     // ids are < n_points (sizes the 16-bit visited table)
    const float* queries;  // device, nq * dim
    const float* queries_tail = nullptr;  // optional second source for queries [split, nq) (mapped host memory)
    uint32_t split = 0;
    uint32_t nq, topn, ef;
    uint32_t* out_ids;     // nq * topn (EMPTY padded)
    float* out_dists;      // nq * topn (+inf padded), may be null
    uint32_t* out_counts;  // nq, may be null
    uint32_t* out_hops;    // nq, may be null
    uint32_t* out_evals;   // nq, may be null
    uint32_t* out_flags;   // nq, may be null (bit0 NaN query, bit1 visited overflow)
    uint32_t* out_nbrs = nullptr;  // nq, may be null: neighbour ids read
    uint32_t* nan_any = nullptr;  // optional: set to 1 if any query holds a NaN
    uint32_t* peer_ids[HB_MAX_PEERS] = {};  // fused all-gather targets (device-visible peer buffers)
    float* peer_dists[HB_MAX_PEERS] = {};   // optional distance rows next to the id rows (base shards)
    uint32_t n_peers = 0;
    uint64_t peer_row0 = 0;
    uint32_t id_offset = 0;  // added to every returned id (global ids of a base shard)
    uint32_t* work_counter;  // device u32, zero on entry
    bool counter_is_fresh = false;  // true: the caller guarantees *work_counter == 0 (no memset is enqueued)
    bool overlap_previous = false;  // launch as programmatic dependent of the previous kernel in the stream
    // optional global continuation of the visited set's exact spill set (search_fast.cuh, SpillPool): spill_warps owner
    // words followed by spill_warps slices of spill_cap ids (a power of two)
    uint32_t* spill_ws = nullptr;
    uint32_t spill_cap = 0, spill_warps = 0;
};

// f32 rows -> lane-sliced records (+ optional flat codes/mins/deltas)
cudaError_t launch_quantise(const float* rows, uint64_t n, const RecLayout& L, uint8_t* rec,
                            uint8_t* codes, float* mins, float* deltas, uint32_t* nan_flag,
                            cudaStream_t st);
cudaError_t launch_pack(const uint8_t* codes, const float* mins, const float* deltas, uint64_t n,
                        const RecLayout& L, uint8_t* rec, cudaStream_t st);
// f32 records (FullVec): rows -> zero-padded records (flag raised on a non-finite value); records -> the values the
// distance arithmetic sees, natural order (works for both kinds: dequantised values of a QuantVec record)
cudaError_t launch_pack_f32(const float* rows, uint64_t n, const RecLayout& L, uint8_t* rec, uint32_t* bad_flag,
                            cudaStream_t st);
cudaError_t launch_record_values(const uint8_t* rec, uint64_t n, const RecLayout& L, float* rows, cudaStream_t st);
cudaError_t launch_unpack(const uint8_t* rec, uint64_t n, const RecLayout& L, uint8_t* codes,
                          float* mins, float* deltas, cudaStream_t st);
// rows / |row| (cosine mode); out may alias rows
cudaError_t launch_normalise(const float* rows, uint64_t n, uint32_t dim, float* out, cudaStream_t st);
// distance2point / dist2many: one f32 query (quantised on device) against ids[n]
cudaError_t launch_dist_query_many(const uint8_t* rec, const RecLayout& L, const float* query,
                                   const uint32_t* ids, uint64_t n, float* out, uint32_t* nan_flag,
                                   cudaStream_t st);
// Points::distance(a[i], b[i])
cudaError_t launch_dist_pairs(const uint8_t* rec, const RecLayout& L, const uint32_t* a,
                              const uint32_t* b, uint64_t n, float* out, cudaStream_t st);
// one stored point against a list of stored points, many such jobs:
// out[j] = distance(src[job], ids[j]) for j in [off[job], off[job+1])
cudaError_t launch_dist_one_to_many(const uint8_t* rec, const RecLayout& L, const uint32_t* src,
                                    const uint32_t* off, const uint32_t* ids, uint32_t njobs,
                                    float* out, cudaStream_t st);
// FullVec::distance: strictly sequential f32 sum, x[i*dim..], y[i*dim..]
cudaError_t launch_dist_full_pairs(const float* x, const float* y, uint64_t n, uint32_t dim,
                                   float* out, cudaStream_t st);
cudaError_t launch_search(const SearchLaunch& a, int num_sms, cudaStream_t st);
void choose_visited(uint32_t ef, uint32_t S0, uint64_t n_points, uint32_t* tbits, uint32_t* bbits, bool* use16);

// brute force: one pass over base records [b0, b1) for all queries
cudaError_t launch_bf_chunk(const uint8_t* base_rec, const RecLayout& L, uint64_t b0, uint64_t b1,
                            uint32_t id_offset, const uint8_t* qrec, uint32_t nq, const uint64_t* tau,
                            uint64_t* buf, uint32_t cap, uint32_t* cnt, uint32_t* overflow,
                            cudaStream_t st);
cudaError_t launch_bf_merge(uint64_t* topk, uint32_t k, uint64_t* tau, uint64_t* buf, uint32_t cap,
                            uint32_t* cnt, uint32_t nq, cudaStream_t st);
cudaError_t launch_keys_to_out(const uint64_t* topk, uint32_t k, uint32_t nq, uint32_t* ids,
                               float* dists, cudaStream_t st);
// brute force on the tensor cores (bf_tc.cu): tcgen05 kind::i8 contraction of the code bytes as a filter,
// exact re-rank of the survivors.  Only for records of exactly 128 bytes (dim 96 / 100).
bool bf_tc_supported(const RecLayout& L);
cudaError_t bf_tc_prepare(const uint8_t* base_rec, uint64_t n, const RecLayout& L, const uint8_t* qrec, uint32_t nq,
                          float4* bconst, uint8_t* amask, float4* qstat, int* qshift, cudaStream_t st);
cudaError_t bf_tc_first(const uint8_t* base_rec, const RecLayout& L, uint32_t first, uint32_t id_offset, const uint8_t* qrec,
                        uint32_t nq, unsigned long long* cand, uint32_t cap, uint32_t* cnt, cudaStream_t st);
cudaError_t bf_tc_chunk(const uint8_t* base_rec, uint64_t n_base, const RecLayout& L, uint64_t row0, uint64_t row_end,
                        uint32_t id_offset, const uint8_t* qrec, const uint8_t* amask, const float4* qstat,
                        const float4* bconst, float4* qconst, const int* qshift, uint32_t nq, const unsigned long long* tau,
                        unsigned long long* cand, uint32_t cap, uint32_t* cnt, uint32_t* overflow, int num_sms,
                        cudaStream_t st);
// merge G sorted lists of k (ids/dists [G][nq][k]) into one list of k per query
cudaError_t launch_topk_merge(const uint32_t* ids, const float* dists, uint32_t G, uint32_t nq,
                              uint32_t k, uint32_t* out_ids, float* out_dists, cudaStream_t st);

// peer exchange (kernels.cu): copy rows into peer buffers, raise / await per-rank flag words
cudaError_t launch_peer_put(const void* src, uint64_t bytes, uint32_t n_peers, void* const* dst, int num_sms, cudaStream_t st);
cudaError_t launch_peer_signal(uint32_t n_peers, uint32_t* const* flags, uint32_t slot, uint32_t epoch, cudaStream_t st);
cudaError_t launch_peer_wait(const uint32_t* flags, uint32_t n, uint32_t epoch, uint32_t* status, cudaStream_t st);
// name of the kernel variant the last launch_search of this thread ran (search_fast.cu)
const char* last_search_variant();

}  // namespace hb
