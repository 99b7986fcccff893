// host.h -- library-internal host state shared by the translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <vector>
#include <mutex>
#include <unordered_map>
#include "../../include/chroma_b200.h"
#include "engine.cuh"

namespace cb {

struct Context {
    int device = -1;
    int sm_count = 0;
    size_t l2_bytes = 0;
    size_t l2_persist_max = 0;            // cudaDevAttrMaxPersistingL2CacheSize
    size_t l2_window_max = 0;             // cudaDevAttrMaxAccessPolicyWindowSize
    const void* l2_window_base = nullptr; // what the kernel stream's access-policy window currently covers
    size_t l2_window_bytes = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;        // kernels
    cudaStream_t copy_stream = nullptr;   // host<->device copies, bank initialisation
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;      // cb_timer_*
    cudaEvent_t kev0 = nullptr, kev1 = nullptr;    // per-call kernel timing
    cudaEvent_t iev0 = nullptr, iev1 = nullptr;    // first-step traversal kernel timing
    std::vector<cudaEvent_t> class_ev;             // events around every launch of a propagate call (per-class kernel time)
    void* flush_buf = nullptr; size_t flush_bytes = 0;
    unsigned long long* d_counters = nullptr;      // 16 x u64 scratch (work counter, stats, flags)
    unsigned long long* h_counters = nullptr;      // pinned mirror
    uint32_t* d_block_counts = nullptr; size_t block_counts_cap = 0;   // compaction scratch
    uint32_t* d_queue[2] = {nullptr, nullptr}; int32_t* d_hit_tri = nullptr; float* d_hit_dist = nullptr;
    uint32_t* d_keys[2] = {nullptr, nullptr}; uint32_t* d_sorted = nullptr; void* d_sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    uint64_t scratch_cap = 0;                       // wavefront scratch (one rng-pool chunk)
    unsigned long long* d_step_counts = nullptr; size_t step_slots = 0;   // per-step alive counts + work counters
    // The scratch above, the kernel stream and its events are shared by every entry point that launches
    // kernels: those calls are serialised (CB_SERIALISE).  Host<->device copies (cb_memcpy_*) use per-thread
    // streams and stay concurrent, which is what the upload | propagate | read-back pipeline needs.
    std::recursive_mutex gpu_mu;
    bool blocking_sync = false;           // host waits yield the core instead of spinning (cb_set_blocking_sync)
};

Context& ctx();
int fail(int code, const char* fmt, ...);
// host waits until `s` has drained: spins (cudaStreamSynchronize) or, in blocking mode, sleeps on a
// cudaEventBlockingSync event so that many ranks / pipeline threads can share few cores
cudaError_t stream_wait(cudaStream_t s);
cudaError_t event_wait(cudaEvent_t e);
// p[0..count) = value on stream s (kernel launch only)
void fill32_launch(uint32_t* p, uint32_t value, uint64_t count, cudaStream_t s);
int cuda_fail(cudaError_t e, const char* what);

#define CB_CUDA(call)                                                         \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return cb::cuda_fail(_e, #call);               \
    } while (0)

// every entry point may be called from any host thread: bind it to the library's device once
inline void bind_thread()
{
    static thread_local bool bound = false;
    if (!bound) { cudaSetDevice(ctx().device); bound = true; }
}

#define CB_REQUIRE_INIT()                                                     \
    do {                                                                      \
        if (cb::ctx().device < 0) return cb::fail(CB_ERR_INVALID, "cb_init() has not been called"); \
        cb::bind_thread();                                                    \
    } while (0)

#define CB_SERIALISE() std::lock_guard<std::recursive_mutex> cb_serialise_guard(cb::ctx().gpu_mu)

struct Geometry {
    DevGeometry dev;                 // kernel-side view (pointers below)
    // reference-layout copies (API mirror + bank utilities)
    float* vertices = nullptr; uint32_t* triangles = nullptr; uint32_t* material_codes = nullptr;
    uint32_t* colors = nullptr; uint32_t* solid_id = nullptr; uint4* nodes = nullptr;
    // native
    uint4* native_nodes = nullptr; uint64_t nnative = 0;   // engine-built traversal tree
    float4* tri64 = nullptr; float* tables = nullptr; CbMaterial* materials = nullptr; CbSurface* surfaces = nullptr;
    WireFrame* wireframes = nullptr;
    uint64_t nvertices = 0, ntriangles = 0, nnodes = 0, table_floats = 0;
    // detector
    int32_t* solid_to_channel = nullptr; uint64_t nsolids = 0; int32_t nchannels = 0;
    float *time_cdf_x = nullptr, *time_cdf_y = nullptr, *charge_cdf_x = nullptr, *charge_cdf_y = nullptr;
    int32_t time_cdf_len = 0, charge_cdf_len = 0; float charge_unit = 0.f;
    uint64_t device_bytes = 0;
    uint32_t smem_table_bytes = 0;   // bytes of the table pool staged in shared memory
};

struct RngPool {
    uint32_t* states = nullptr;      // 6 words per state {d, v0..v4}
    RngPool* parent = nullptr;       // a view (cb_rng_view) borrows its parent's arrays
    uint64_t first_stream = 0;       // state i follows stream first_stream + i
    float* bm_extra = nullptr;       // Box-Muller cache (allocated lazily for run_daq_many)
    uint32_t* bm_flag = nullptr;
    uint64_t n = 0;
};

struct Daq {
    Geometry* geom = nullptr;
    int32_t ndaq = 1; uint64_t count = 0;
    float* earliest_time = nullptr; uint32_t* earliest_time_int = nullptr;
    uint32_t* channel_history = nullptr; uint32_t* channel_q_int = nullptr; float* channel_q = nullptr;
};

// handle registries (opaque 64-bit ids -> objects)
template <typename T>
struct Registry {
    std::mutex mu; std::unordered_map<uint64_t, T*> map; uint64_t next = 1;
    uint64_t add(T* p) { std::lock_guard<std::mutex> l(mu); uint64_t id = next++; map[id] = p; return id; }
    T* get(uint64_t id) { std::lock_guard<std::mutex> l(mu); auto it = map.find(id); return it == map.end() ? nullptr : it->second; }
    T* take(uint64_t id) { std::lock_guard<std::mutex> l(mu); auto it = map.find(id); if (it == map.end()) return nullptr; T* p = it->second; map.erase(it); return p; }
};
// mark the leading `bytes` of the traversal tree (breadth-first: its top levels) as persisting in L2
// for the kernels of the library stream; no-op when already set for this geometry
int l2_pin_tree_prefix(const Geometry* g);
Registry<Geometry>& geoms();
Registry<RngPool>& rngs();
Registry<Daq>& daqs();

} // namespace cb
