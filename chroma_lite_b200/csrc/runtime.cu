// runtime.cu -- context, memory, timers and the XORWOW state pool.
#include "host.h"
#include <string.h>
#include <stdlib.h>
#include <time.h>
#include <algorithm>
#ifndef CB_L2_WINDOW_MB_DEFAULT
#define CB_L2_WINDOW_MB_DEFAULT 16.0   /* measured r02: 16 MB -0.15 ms per 2.5 M-photon event, 48 MB the same, 96 MB +0.5 ms */
#endif

namespace cb {

static Context g_ctx;
static thread_local char g_err[1024] = "";
static char g_err_global[1024] = "";

Context& ctx() { return g_ctx; }

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    strncpy(g_err_global, g_err, sizeof(g_err_global) - 1);
    return code;
}
int cuda_fail(cudaError_t e, const char* what)
{
    int code = (e == cudaErrorMemoryAllocation) ? CB_ERR_NOMEM : CB_ERR_CUDA;
    return fail(code, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

// Host waits.  Spinning (the CUDA default) has the lowest latency but needs a core per waiting thread;
// a blocking wait yields the core but wakes up ~50 us late, which a DAQ acquisition with its ten short
// waits cannot afford (measured r02, two ranks on 8 cores: daq 0.19 -> 1.8 ms per event with plain
// blocking waits).  Blocking mode therefore polls for SPIN_US first -- short kernels and copies finish
// inside that -- and only then sleeps on a cudaEventBlockingSync event (the 5 ms propagate, the uploads).
static inline double now_us()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
constexpr double SPIN_US = 150.0;

cudaError_t stream_wait(cudaStream_t s)
{
    if (!g_ctx.blocking_sync) return cudaStreamSynchronize(s);
    const double t0 = now_us();
    for (;;) {
        const cudaError_t q = cudaStreamQuery(s);
        if (q != cudaErrorNotReady) return q;
        if (now_us() - t0 > SPIN_US) break;
    }
    static thread_local cudaEvent_t ev = nullptr;
    if (!ev) {
        cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) { ev = nullptr; return e; }
    }
    cudaError_t e = cudaEventRecord(ev, s);
    return e != cudaSuccess ? e : cudaEventSynchronize(ev);
}
// timing events are created spinning: in blocking mode poll, then sleep in short naps
cudaError_t event_wait(cudaEvent_t e)
{
    if (!g_ctx.blocking_sync) return cudaEventSynchronize(e);
    const double t0 = now_us();
    for (;;) {
        const cudaError_t q = cudaEventQuery(e);
        if (q != cudaErrorNotReady) return q;
        if (now_us() - t0 > SPIN_US) {
            struct timespec ts = {0, 30000};      // 30 us
            nanosleep(&ts, nullptr);
        }
    }
}

static Registry<Geometry> g_geoms;
static Registry<RngPool> g_rngs;
static Registry<Daq> g_daqs;
Registry<Geometry>& geoms() { return g_geoms; }
Registry<RngPool>& rngs() { return g_rngs; }
Registry<Daq>& daqs() { return g_daqs; }

// ---------------------------------------------------------------- L2 persistence
// The engine's tree is stored breadth-first, so its first bytes are its top levels: the part every
// ray walks.  An access-policy window on the kernel stream marks that prefix as persisting, so the
// per-photon streaming traffic (bank, queues, triangle records of one-off leaves) cannot evict it.
// CHROMA_B200_L2_WINDOW_MB: size of the prefix (0 = no window).
int l2_pin_tree_prefix(const Geometry* g)
{
    Context& c = ctx();
    const double want_mb = getenv("CHROMA_B200_L2_WINDOW_MB") ? atof(getenv("CHROMA_B200_L2_WINDOW_MB")) : CB_L2_WINDOW_MB_DEFAULT;
    const void* base = g->native_nodes ? (const void*)g->native_nodes : (const void*)g->nodes;
    const size_t have = (g->native_nodes ? g->nnative : g->nnodes) * sizeof(uint4);
    size_t bytes = std::min<size_t>((size_t)(want_mb * 1048576.0), have);
    bytes = std::min(bytes, c.l2_window_max);
    if (c.l2_persist_max == 0) bytes = 0;
    if (base == c.l2_window_base && bytes == c.l2_window_bytes) return CB_OK;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (bytes) {
        const size_t carve = std::min(bytes, c.l2_persist_max);
        CB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
        attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)bytes);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    CB_CUDA(cudaStreamSetAttribute(c.stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    if (!bytes && c.l2_window_bytes) { cudaCtxResetPersistingL2Cache(); cudaGetLastError(); }
    c.l2_window_base = base; c.l2_window_bytes = bytes;
    return CB_OK;
}

// ---------------------------------------------------------------- XORWOW skip matrices
// 160x160 GF(2) matrices in cuRAND's row layout (row i = image of state bit i,
// 5 words), generated here by repeated squaring of the one-step matrix instead
// of shipping curand_precalc.h: seq[k] = M^(2^67 * 2^k), off[k] = M^(2^k).
struct XwMat { uint32_t r[160][5]; };

static void xw_step(uint32_t v[5])
{
    uint32_t t = v[0] ^ (v[0] >> 2);
    v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}
static void xw_matvec(const XwMat& m, const uint32_t* v, uint32_t* out)
{
    uint32_t acc[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 160; i++)
        if (v[i >> 5] & (1u << (i & 31)))
            for (int j = 0; j < 5; j++) acc[j] ^= m.r[i][j];
    memcpy(out, acc, sizeof(acc));
}
static void xw_square(const XwMat& a, XwMat& out)
{
    XwMat tmp;
    for (int i = 0; i < 160; i++) xw_matvec(a, a.r[i], tmp.r[i]);
    out = tmp;
}

constexpr int XW_NMAT = 64;
static uint32_t* d_xw_seq = nullptr;   // [64][800]
static uint32_t* d_xw_off = nullptr;   // [64][800]

static int xw_upload_tables()
{
    if (d_xw_seq) return CB_OK;
    std::vector<XwMat> off(XW_NMAT), seq(XW_NMAT);
    for (int i = 0; i < 160; i++) {
        uint32_t v[5] = {0, 0, 0, 0, 0};
        v[i >> 5] = 1u << (i & 31);
        xw_step(v);
        memcpy(off[0].r[i], v, 20);
    }
    for (int k = 1; k < XW_NMAT; k++) xw_square(off[k - 1], off[k]);
    XwMat m = off[63];
    for (int k = 64; k <= 67; k++) xw_square(m, m);
    seq[0] = m;
    for (int k = 1; k < XW_NMAT; k++) xw_square(seq[k - 1], seq[k]);
    CB_CUDA(cudaMalloc(&d_xw_seq, sizeof(XwMat) * XW_NMAT));
    CB_CUDA(cudaMalloc(&d_xw_off, sizeof(XwMat) * XW_NMAT));
    CB_CUDA(cudaMemcpy(d_xw_seq, seq.data(), sizeof(XwMat) * XW_NMAT, cudaMemcpyHostToDevice));
    CB_CUDA(cudaMemcpy(d_xw_off, off.data(), sizeof(XwMat) * XW_NMAT, cudaMemcpyHostToDevice));
    return CB_OK;
}

__device__ __forceinline__ void xw_apply(const uint32_t* __restrict__ mat, uint32_t v[5])
{
    uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0;
#pragma unroll
    for (int w = 0; w < 5; w++) {
        uint32_t bits = v[w];
#pragma unroll 4
        for (int b = 0; b < 32; b++) {
            const uint32_t* row = mat + (w * 32 + b) * 5;
            uint32_t mask = 0u - ((bits >> b) & 1u);
            acc0 ^= __ldg(row + 0) & mask; acc1 ^= __ldg(row + 1) & mask; acc2 ^= __ldg(row + 2) & mask;
            acc3 ^= __ldg(row + 3) & mask; acc4 ^= __ldg(row + 4) & mask;
        }
    }
    v[0] = acc0; v[1] = acc1; v[2] = acc2; v[3] = acc3; v[4] = acc4;
}

// state[i] = curand_init(seed, first_stream + i, offset)  (curand_kernel.h:772-800)
__global__ void __launch_bounds__(256)
rng_init_kernel(uint32_t* __restrict__ states, uint64_t n, uint64_t seed, uint64_t first_stream,
                uint64_t offset, const uint32_t* __restrict__ seq, const uint32_t* __restrict__ off)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    uint32_t d = 6615241u + t1 + t0;
    uint32_t v[5];
    v[0] = 123456789u + t0;
    v[1] = 362436069u ^ t0;
    v[2] = 521288629u + t1;
    v[3] = 88675123u ^ t1;
    v[4] = 5783321u + t0;
    uint64_t stream = first_stream + i;
    for (int k = 0; stream; k++, stream >>= 1)
        if (stream & 1) xw_apply(seq + k * 800, v);
    uint64_t o = offset;
    for (int k = 0; o; k++, o >>= 1)
        if (o & 1) xw_apply(off + k * 800, v);
    d += 362437u * (uint32_t)offset;
    Rng r = {d, v[0], v[1], v[2], v[3], v[4]};
    rng_store(states, i, r);
}

__global__ void rng_fill_uniform_kernel(uint32_t* __restrict__ states, uint64_t n, float low, float high,
                                        float* __restrict__ out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Rng r = rng_load(states, i);
    out[i] = rng_range(r, low, high);
    rng_store(states, i, r);
}

__global__ void fill32_kernel(uint32_t* p, uint32_t v, uint64_t n)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

void fill32_launch(uint32_t* p, uint32_t value, uint64_t count, cudaStream_t s)
{
    if (!count) return;
    const int blocks = (int)std::min<uint64_t>((count + 255) / 256, (uint64_t)ctx().sm_count * 16);
    fill32_kernel<<<blocks, 256, 0, s>>>(p, value, count);
}

} // namespace cb

using namespace cb;

extern "C" {

int cb_abi_version(void) { return CB_ABI_VERSION; }
const char* cb_last_error(void) { return g_err[0] ? g_err : g_err_global; }

int cb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int cb_init(int device)
{
    Context& c = ctx();
    if (c.device == device && c.stream) return CB_OK;
    if (c.device >= 0 && c.device != device)
        return fail(CB_ERR_INVALID, "library already bound to device %d (one process per GPU)", c.device);
    int n = 0;
    CB_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(CB_ERR_INVALID, "device %d out of range (%d visible)", device, n);
    CB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(CB_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    c.sm_count = prop.multiProcessorCount;
    c.l2_bytes = (size_t)prop.l2CacheSize;
    c.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    CB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CB_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    CB_CUDA(cudaEventCreate(&c.ev0)); CB_CUDA(cudaEventCreate(&c.ev1));
    CB_CUDA(cudaEventCreate(&c.kev0)); CB_CUDA(cudaEventCreate(&c.kev1));
    CB_CUDA(cudaEventCreate(&c.iev0)); CB_CUDA(cudaEventCreate(&c.iev1));
    CB_CUDA(cudaMalloc(&c.d_counters, 16 * sizeof(unsigned long long)));
    CB_CUDA(cudaMemset(c.d_counters, 0, 16 * sizeof(unsigned long long)));
    CB_CUDA(cudaMallocHost(&c.h_counters, 32 * sizeof(unsigned long long)));   // [0..15] counters, [16] alive count read back, [17] alive count sent
    // L2 persistence: the carve-out is sized when a geometry's tree prefix is pinned (l2_pin_tree_prefix)
    c.l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    c.l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    if (const char* e = getenv("CHROMA_B200_SYNC")) c.blocking_sync = strcmp(e, "block") == 0;
    c.device = device;
    return CB_OK;
}

int cb_set_blocking_sync(int32_t on)
{
    ctx().blocking_sync = on != 0;
    return CB_OK;
}

int cb_sm_count(void) { return ctx().sm_count; }
int cb_device_pci_bus_id(char* out, int32_t len)
{
    CB_REQUIRE_INIT();
    if (!out || len < 16) return fail(CB_ERR_INVALID, "cb_device_pci_bus_id: buffer too small");
    CB_CUDA(cudaDeviceGetPCIBusId(out, len, ctx().device));
    return CB_OK;
}

int cb_synchronize(void)
{
    CB_REQUIRE_INIT();
    CB_CUDA(stream_wait(ctx().stream));
    CB_CUDA(cudaDeviceSynchronize());
    return CB_OK;
}

int cb_malloc(uint64_t bytes, void** dptr)
{
    CB_REQUIRE_INIT();
    if (!dptr) return fail(CB_ERR_INVALID, "cb_malloc: null out pointer");
    *dptr = nullptr;
    if (bytes == 0) bytes = 16;
    CB_CUDA(cudaMalloc(dptr, bytes));
    return CB_OK;
}
int cb_free(void* dptr)
{
    if (!dptr) return CB_OK;
    CB_REQUIRE_INIT();
    CB_CUDA(cudaFree(dptr));          // cudaFree itself waits for work that uses the block
    return CB_OK;
}
// host<->device copies run on per-host-thread streams, so that uploads issued from
// worker threads overlap each other and the kernels of a propagate call in flight
static cudaStream_t thread_copy_stream()
{
    static thread_local cudaStream_t s = nullptr;
    if (!s) {
        cudaSetDevice(ctx().device);
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); s = ctx().copy_stream; }
    }
    return s;
}
int cb_memcpy_h2d(void* d, const void* h, uint64_t bytes)
{
    CB_REQUIRE_INIT();
    if (bytes == 0) return CB_OK;
    cudaStream_t s = thread_copy_stream();
    CB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
    CB_CUDA(stream_wait(s));
    return CB_OK;
}
int cb_memcpy_d2h(void* h, const void* d, uint64_t bytes)
{
    CB_REQUIRE_INIT();
    if (bytes == 0) return CB_OK;
    cudaStream_t s = thread_copy_stream();
    CB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s));
    CB_CUDA(stream_wait(s));
    return CB_OK;
}
int cb_memcpy_d2d(void* dst, const void* src, uint64_t bytes)
{
    CB_REQUIRE_INIT();
    if (bytes == 0) return CB_OK;
    cudaStream_t s = thread_copy_stream();
    CB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
    CB_CUDA(stream_wait(s));
    return CB_OK;
}
int cb_memset32(void* dptr, uint32_t value, uint64_t count)
{
    CB_REQUIRE_INIT();
    if (count == 0) return CB_OK;
    int blocks = (int)std::min<uint64_t>((count + 255) / 256, (uint64_t)ctx().sm_count * 16);
    cudaStream_t s = thread_copy_stream();
    if (value == 0u || value == 0xFFFFFFFFu) {
        // byte patterns go through the copy engine and do not queue behind persistent kernels
        CB_CUDA(cudaMemsetAsync(dptr, value ? 0xFF : 0, count * 4, s));
    } else {
        fill32_kernel<<<blocks, 256, 0, s>>>((uint32_t*)dptr, value, count);
        CB_CUDA(cudaGetLastError());
    }
    CB_CUDA(stream_wait(s));
    return CB_OK;
}
// One event's photon bank from host arrays in ONE call: every copy and fill is issued on the calling
// thread's copy stream and waited for once.  (Array by array from Python, each call re-acquires the
// interpreter lock behind whatever the other pipeline threads are doing: measured 4.1 ms per 120 MB event
// inside the pipeline against 2.4 ms alone.)  Fills use the copy engine only -- a fill KERNEL would queue
// behind the persistent propagate kernels of the event in flight: byte patterns through cudaMemsetAsync,
// weights = 1.0f through a device-to-device copy from a buffer of ones.
static float* g_ones = nullptr;
static uint64_t g_ones_n = 0;
static std::mutex g_ones_mu;
static int ones_buffer(uint64_t n, const float** out)
{
    std::lock_guard<std::mutex> l(g_ones_mu);
    if (g_ones_n < n) {
        // the old buffer may still feed a copy in flight on another thread's stream: it is never freed, only outgrown
        float* p = nullptr;
        const uint64_t cap = std::max<uint64_t>(std::max<uint64_t>(n, 2 * g_ones_n), 1u << 20);    // geometric: outgrown buffers stay bounded
        CB_CUDA(cudaMalloc(&p, cap * 4));
        cudaStream_t s = ctx().copy_stream;
        fill32_launch((uint32_t*)p, 0x3F800000u, cap, s);
        CB_CUDA(cudaGetLastError());
        CB_CUDA(stream_wait(s));
        g_ones = p; g_ones_n = cap;
    }
    *out = g_ones;
    return CB_OK;
}

int cb_photon_bank_upload(const CbPhotonBank* dst, const CbPhotonBank* host, uint64_t n, uint32_t evidx_value)
{
    CB_REQUIRE_INIT();
    if (!dst || !host) return fail(CB_ERR_INVALID, "cb_photon_bank_upload: null bank");
    if (n == 0) return CB_OK;
    if (dst->n < n) return fail(CB_ERR_INVALID, "cb_photon_bank_upload: device bank too small");
    if (!dst->pos || !dst->dir || !dst->pol || !dst->wavelengths || !dst->t || !dst->last_hit_triangles || !dst->flags || !dst->weights)
        return fail(CB_ERR_INVALID, "cb_photon_bank_upload: device bank has null arrays");
    if (!host->pos || !host->dir || !host->pol || !host->wavelengths)
        return fail(CB_ERR_INVALID, "cb_photon_bank_upload: pos, dir, pol and wavelengths are required");
    cudaStream_t s = thread_copy_stream();
    CB_CUDA(cudaMemcpyAsync(dst->pos, host->pos, n * 12, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(dst->dir, host->dir, n * 12, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(dst->pol, host->pol, n * 12, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(dst->wavelengths, host->wavelengths, n * 4, cudaMemcpyHostToDevice, s));
    if (host->t) CB_CUDA(cudaMemcpyAsync(dst->t, host->t, n * 4, cudaMemcpyHostToDevice, s));
    else CB_CUDA(cudaMemsetAsync(dst->t, 0, n * 4, s));
    if (host->last_hit_triangles) CB_CUDA(cudaMemcpyAsync(dst->last_hit_triangles, host->last_hit_triangles, n * 4, cudaMemcpyHostToDevice, s));
    else CB_CUDA(cudaMemsetAsync(dst->last_hit_triangles, 0xFF, n * 4, s));                       // -1
    if (host->flags) CB_CUDA(cudaMemcpyAsync(dst->flags, host->flags, n * 4, cudaMemcpyHostToDevice, s));
    else CB_CUDA(cudaMemsetAsync(dst->flags, 0, n * 4, s));
    if (host->weights) CB_CUDA(cudaMemcpyAsync(dst->weights, host->weights, n * 4, cudaMemcpyHostToDevice, s));
    else {
        const float* ones = nullptr;
        int rc = ones_buffer(n, &ones);
        if (rc) return rc;
        CB_CUDA(cudaMemcpyAsync(dst->weights, ones, n * 4, cudaMemcpyDeviceToDevice, s));
    }
    if (dst->evidx) {
        if (host->evidx) CB_CUDA(cudaMemcpyAsync(dst->evidx, host->evidx, n * 4, cudaMemcpyHostToDevice, s));
        else if (evidx_value == 0u) CB_CUDA(cudaMemsetAsync(dst->evidx, 0, n * 4, s));
        else { fill32_launch(dst->evidx, evidx_value, n, s); CB_CUDA(cudaGetLastError()); }
    }
    CB_CUDA(stream_wait(s));
    return CB_OK;
}

int cb_host_alloc(uint64_t bytes, void** hptr)
{
    CB_REQUIRE_INIT();
    CB_CUDA(cudaMallocHost(hptr, bytes ? bytes : 16));
    return CB_OK;
}
int cb_host_alloc_flags(uint64_t bytes, int32_t write_combined, void** hptr)
{
    // write-combined page-locked memory: not snooped, so the DMA engine reads it faster on some hosts;
    // the CPU should only ever WRITE it (reads are uncached).  For upload-only staging buffers.
    CB_REQUIRE_INIT();
    CB_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 16, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    return CB_OK;
}
int cb_host_free(void* hptr)
{
    CB_REQUIRE_INIT();
    if (!hptr) return CB_OK;
    CB_CUDA(cudaFreeHost(hptr));
    return CB_OK;
}
int cb_mem_info(uint64_t* free_bytes, uint64_t* total_bytes)
{
    CB_REQUIRE_INIT();
    size_t f = 0, t = 0;
    CB_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return CB_OK;
}

// ---------------------------------------------------------------- completion events
// For callers that enqueue work with the *_async entry points: a marker on the library stream and a host
// wait for it.  The wait naps between polls (the waiter is off the critical path by construction: it
// consumes event k while the GPU works on event k+1), so it never occupies a core.
int cb_event_create(cb_event_t* out)
{
    CB_REQUIRE_INIT();
    if (!out) return fail(CB_ERR_INVALID, "cb_event_create: null out pointer");
    cudaEvent_t e = nullptr;
    CB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *out = (cb_event_t)(uintptr_t)e;
    return CB_OK;
}
int cb_event_record(cb_event_t ev)
{
    CB_REQUIRE_INIT();
    if (!ev) return fail(CB_ERR_INVALID, "cb_event_record: null event");
    CB_CUDA(cudaEventRecord((cudaEvent_t)(uintptr_t)ev, ctx().stream));
    return CB_OK;
}
int cb_event_wait(cb_event_t ev)
{
    CB_REQUIRE_INIT();
    if (!ev) return fail(CB_ERR_INVALID, "cb_event_wait: null event");
    const cudaEvent_t e = (cudaEvent_t)(uintptr_t)ev;
    const double t0 = now_us();
    for (;;) {
        const cudaError_t q = cudaEventQuery(e);
        if (q == cudaSuccess) return CB_OK;
        if (q != cudaErrorNotReady) return cuda_fail(q, "cudaEventQuery");
        if (now_us() - t0 > 20.0) {
            struct timespec ts = {0, 40000};      // 40 us
            nanosleep(&ts, nullptr);
        }
    }
}
int cb_event_destroy(cb_event_t ev)
{
    CB_REQUIRE_INIT();
    if (ev) CB_CUDA(cudaEventDestroy((cudaEvent_t)(uintptr_t)ev));
    return CB_OK;
}

int cb_timer_start(void)
{
    CB_REQUIRE_INIT();
    CB_CUDA(cudaEventRecord(ctx().ev0, ctx().stream));
    return CB_OK;
}
int cb_timer_stop(float* ms)
{
    CB_REQUIRE_INIT();
    CB_CUDA(cudaEventRecord(ctx().ev1, ctx().stream));
    CB_CUDA(event_wait(ctx().ev1));
    float t = 0.f;
    CB_CUDA(cudaEventElapsedTime(&t, ctx().ev0, ctx().ev1));
    if (ms) *ms = t;
    return CB_OK;
}
int cb_flush_l2(void)
{
    CB_REQUIRE_INIT();
    Context& c = ctx();
    if (!c.flush_buf) {
        c.flush_bytes = std::max<size_t>(c.l2_bytes * 2, (size_t)256 << 20);
        CB_CUDA(cudaMalloc(&c.flush_buf, c.flush_bytes));
    }
    static uint32_t tick = 0;
    fill32_kernel<<<c.sm_count * 8, 256, 0, c.stream>>>((uint32_t*)c.flush_buf, ++tick, c.flush_bytes / 4);
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}

// ---------------------------------------------------------------- RNG pool
int cb_rng_create_streams(uint64_t n, uint64_t seed, uint64_t first_stream, uint64_t offset, cb_rng_t* out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    if (!out) return fail(CB_ERR_INVALID, "cb_rng_create: null out pointer");
    int rc = xw_upload_tables();
    if (rc) return rc;
    RngPool* r = new RngPool();
    r->n = n;
    r->first_stream = first_stream;
    cudaError_t e = cudaMalloc(&r->states, std::max<uint64_t>(n, 1) * 24);
    if (e != cudaSuccess) { delete r; return cuda_fail(e, "cudaMalloc(rng states)"); }
    if (n) {
        unsigned blocks = (unsigned)((n + 255) / 256);
        rng_init_kernel<<<blocks, 256, 0, ctx().stream>>>(r->states, n, seed, first_stream, offset, d_xw_seq, d_xw_off);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = stream_wait(ctx().stream);
        if (e != cudaSuccess) { cudaFree(r->states); delete r; return cuda_fail(e, "rng_init_kernel"); }
    }
    *out = rngs().add(r);
    return CB_OK;
}
int cb_rng_create(uint64_t n, uint64_t seed, uint64_t offset, cb_rng_t* out)
{
    return cb_rng_create_streams(n, seed, 0, offset, out);
}
int cb_rng_view(cb_rng_t parent, uint64_t first, uint64_t count, cb_rng_t* out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    RngPool* p = rngs().get(parent);
    if (!p || !out) return fail(CB_ERR_INVALID, "cb_rng_view: bad argument");
    if (first + count > p->n) return fail(CB_ERR_INVALID, "cb_rng_view: window [%llu, %llu) exceeds the pool (%llu)",
                                          (unsigned long long)first, (unsigned long long)(first + count), (unsigned long long)p->n);
    // the Box-Muller cache of run_daq_many lives beside the states: allocate it now so that the window sees it
    if (!p->bm_flag) {
        CB_CUDA(cudaMalloc(&p->bm_flag, std::max<uint64_t>(p->n, 1) * 4));
        CB_CUDA(cudaMalloc(&p->bm_extra, std::max<uint64_t>(p->n, 1) * 4));
        CB_CUDA(cudaMemset(p->bm_flag, 0, std::max<uint64_t>(p->n, 1) * 4));
        CB_CUDA(cudaMemset(p->bm_extra, 0, std::max<uint64_t>(p->n, 1) * 4));
    }
    RngPool* v = new RngPool();
    v->parent = p->parent ? p->parent : p;
    v->states = p->states + 6 * first;
    v->bm_flag = p->bm_flag + first; v->bm_extra = p->bm_extra + first;
    v->n = count; v->first_stream = p->first_stream + first;
    *out = rngs().add(v);
    return CB_OK;
}
int cb_rng_destroy(cb_rng_t h)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    RngPool* r = rngs().take(h);
    if (!r) return fail(CB_ERR_INVALID, "cb_rng_destroy: bad handle");
    stream_wait(ctx().stream);
    if (!r->parent) { cudaFree(r->states); cudaFree(r->bm_extra); cudaFree(r->bm_flag); }
    delete r;
    return CB_OK;
}
int cb_rng_size(cb_rng_t h, uint64_t* n)
{
    CB_REQUIRE_INIT();
    RngPool* r = rngs().get(h);
    if (!r) return fail(CB_ERR_INVALID, "cb_rng_size: bad handle");
    *n = r->n;
    return CB_OK;
}
int cb_rng_download(cb_rng_t h, uint64_t first, uint64_t count, uint32_t* out6)
{
    CB_REQUIRE_INIT();
    RngPool* r = rngs().get(h);
    if (!r) return fail(CB_ERR_INVALID, "cb_rng_download: bad handle");
    if (first + count > r->n) return fail(CB_ERR_INVALID, "cb_rng_download: range exceeds pool");
    return cb_memcpy_d2h(out6, r->states + 6 * first, count * 24);
}
int cb_rng_fill_uniform(cb_rng_t h, uint64_t n, float low, float high, float* d_out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    RngPool* r = rngs().get(h);
    if (!r) return fail(CB_ERR_INVALID, "cb_rng_fill_uniform: bad handle");
    if (n > r->n) return fail(CB_ERR_INVALID, "cb_rng_fill_uniform: n exceeds pool");
    if (!n) return CB_OK;
    rng_fill_uniform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx().stream>>>(r->states, n, low, high, d_out);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(stream_wait(ctx().stream));
    return CB_OK;
}

} // extern "C"
