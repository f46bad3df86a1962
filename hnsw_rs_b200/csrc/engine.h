// Internal object model behind the opaque handles of include/hnsw_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/hnsw_b200.h"
#include "hostgraph.h"
#include "kernels.h"
#include "layout.h"

namespace hb {
void set_error(const std::string& s);
int cuda_fail(cudaError_t e, const char* what);
template <class T>
struct DevBuf {  // scoped device buffer
    T* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T)); }
};
template <typename T>
struct PinBuf {  // scoped page-locked host buffer
    T* p = nullptr;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    cudaError_t alloc(size_t n) { return cudaHostAlloc((void**)&p, (n ? n : 1) * sizeof(T), cudaHostAllocDefault); }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    T* data() { return p; }
};
}  // namespace hb

#define HB_CUDA(call)                                              \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return hb::cuda_fail(_e, #call);    \
    } while (0)

struct hnswb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    int vec_type = 0;  // HNSWB200_VEC_*: the vector type of the points created through this context (the reference's VecType)
    uint32_t* d_scratch = nullptr;  // [0] work counter (build), [1] nan flag, [2] overflow flag, ...
    // Work counters of the search launches: a ring of zeroed slots, one per launch, re-zeroed by ONE memset each
    // time the ring wraps.  No memset sits between two consecutive searches, so a search may be launched as the
    // programmatic dependent of the previous one and fill the SMs its tail leaves idle (hnswb200_search_dev).
    static constexpr uint32_t COUNTER_RING = 32;
    uint32_t* d_counters = nullptr;
    uint64_t search_seq = 0;
    // Programmatic dependent launch of a search is only safe behind another search of this library (a search reads
    // nothing a previous search writes).  last_was_search is true while the last operation this context enqueued was
    // a search kernel; every other entry point clears it (use()).  On an adopted stream (set_stream) the caller may
    // enqueue producers of the query buffer that the library cannot see, so overlap additionally needs the caller's
    // opt-in (hnswb200_ctx_set_overlap).
    mutable bool last_was_search = false;
    bool stream_adopted = false;
    bool overlap_opt_in = false;
    void* d_ws = nullptr;           // grow-only workspace for the host-buffer entry points
    size_t ws_bytes = 0;
    void* d_norm_ws = nullptr;      // grow-only scratch for the normalised queries of a cosine index
    size_t norm_ws_bytes = 0;
    int norm_ws_reserve(size_t bytes);
    void* d_bf_ws = nullptr;        // grow-only scratch of the brute-force entry points (cudaMalloc per call costs more than the kernels)
    size_t bf_ws_bytes = 0;
    int bf_ws_reserve(size_t bytes);
    // global continuation of the search kernel's visited spill set (csrc/search_fast.cuh, SpillPool): SPILL_SLICES owner
    // words (0 = free) followed by SPILL_SLICES hash sets of SPILL_CAP ids, every entry free (0xFF) while a slice is not lent
    uint32_t* d_spill_ws = nullptr;
    static constexpr uint32_t SPILL_CAP = 4096;  // power of two
    static constexpr uint32_t SPILL_SLICES = 1024;
    uint32_t spill_warps = 0;
    std::vector<uint32_t> h_flags;
    uint32_t* h_status = nullptr;   // pinned + mapped host word the search kernel raises on a NaN query
    uint32_t* d_status = nullptr;   // its device alias
    int ws_reserve(size_t bytes);
    int use() const;                // cudaSetDevice
};

struct hnswb200_points {
    hnswb200_ctx* ctx = nullptr;
    RecLayout L{};
    uint64_t n = 0, cap = 0;
    uint8_t* d_rec = nullptr;
    std::vector<uint8_t> levels;  // Point.level (points/src/point.rs:8)
    int metric = 0;               // HNSWB200_METRIC_*: cosine = rows and queries are L2-normalised on the device first
    int reserve(uint64_t want);   // grow device storage, keeps contents
};

struct hnswb200_graph {
    hnswb200_ctx* ctx = nullptr;
    hb::HostGraph h;
    // device mirror
    uint32_t* d_adj0 = nullptr;
    uint64_t rows0_cap = 0, chain0_cap = 0, chain0_used = 0;
    uint32_t* d_upper_off = nullptr;
    uint64_t upper_off_cap = 0;
    uint32_t* d_adju = nullptr;
    uint64_t rowsu_cap = 0, chainu_cap = 0, chainu_used = 0;
    std::unordered_map<uint32_t, std::vector<uint32_t>> chains0, chainsu;  // row -> chain rows
    bool device_valid = false;
    // staging for incremental row uploads (pinned host + device, grow-only)
    uint32_t *h_stage_rows = nullptr, *h_stage_data = nullptr, *d_stage_rows = nullptr, *d_stage_data = nullptr;
    size_t stage_cap = 0;
    uint32_t stage_S = 0;
    std::vector<uint8_t> mark0, marku;
    uint64_t n_full_uploads = 0, n_up_rows = 0, n_up_big = 0;
    double t_up[3] = {0, 0, 0};  // de-duplicate / staging allocation / row gather
    double t_stage = 0, t_xfer = 0;  // upload_rows: host gather into the staging buffer / copy + scatter + wait (build profile)
    hb::DevGraph view() const;
    int upload_full();                                        // (re)build the device mirror
    int upload_rows(std::vector<uint32_t>& dirty0, std::vector<uint32_t>& dirtyu);  // touched rows only
    int sync_new_nodes(uint64_t first_new);                   // after HostGraph::add_node calls
    void free_device();
};

struct hnswb200_index {
    hnswb200_ctx* ctx = nullptr;
    hnswb200_points* points = nullptr;
    hnswb200_graph* graph = nullptr;
    hnswb200_params params{};
};

namespace hb {
// builder.cu
int build_insert(hnswb200_ctx* ctx, hnswb200_index* ix, const std::vector<uint32_t>& new_ids, uint32_t batch);
void draw_levels(uint64_t m, uint64_t n, uint8_t* out);
// api.cu
int points_append_f32(hnswb200_ctx* c, hnswb200_points* p, const float* rows, uint64_t n, const uint8_t* levels);
cudaError_t launch_scatter_rows(uint32_t* dst, uint32_t S, const uint32_t* rows, const uint32_t* data,
                                uint32_t n, cudaStream_t st);
}  // namespace hb
