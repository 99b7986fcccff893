// comm.cu -- the one exchange step of the multi-GPU path: per-channel DAQ accumulators of
// sharded photon banks combined over NVLink.
//
// One process per GPU.  The reference accumulates hits of ONE device with atomics
// (atomicMin on the time bits, atomicAdd on the integer charge, atomicOr on the history word:
// chroma/cuda/daq.cu:73-75, 143-145); with photons partitioned over GPUs the same three
// operations become one reduction across ranks:
//     earliest_time_int  MIN    (non-negative floats order like their bit patterns)
//     channel_q_int      SUM    (uint32, wraps like atomicAdd)
//     channel_history    OR     NCCL has no OR: each of the 32 bits of the word travels as a
//                               counter of `bits` bits (4 up to 15 ranks), packed into uint32
//                               words of the SUM buffer; a bit is set in the result when its
//                               counter is non-zero
// daq_pack_kernel -> ncclGroupStart { AllReduce SUM, AllReduce MIN } ncclGroupEnd ->
// daq_unpack_finalize_kernel (also converts to the float arrays, the work of
// cb_daq_finalize), all on the library stream; nothing passes through the host.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: in a torch process that is the NCCL torch
// already loaded), so the library has no link-time dependency on it.
#include "host.h"
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <vector>

namespace cb {

struct Nccl {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    uint32_t* d_sum = nullptr; size_t sum_cap = 0;      // [q_int | packed history counters]
};
static Nccl g_nccl;

static int nccl_load()
{
    Nccl& n = g_nccl;
    if (n.so) return CB_OK;
    const char* names[] = {getenv("CHROMA_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* name : names) {
        if (!name || !*name) continue;
        n.so = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (n.so) break;
    }
    if (!n.so) return fail(CB_ERR_NCCL, "cannot load libnccl.so.2 (%s)", dlerror());
#define SYM(field, sym)                                                                          \
    *(void**)(&n.field) = dlsym(n.so, sym);                                                      \
    if (!n.field) { dlclose(n.so); n.so = nullptr; return fail(CB_ERR_NCCL, "libnccl lacks %s", sym); }
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return CB_OK;
}

#define CB_NCCL(call)                                                                            \
    do {                                                                                         \
        ncclResult_t _r = (call);                                                                \
        if (_r != ncclSuccess && _r != ncclInProgress)                                           \
            return cb::fail(CB_ERR_NCCL, "NCCL error %d (%s) in %s", (int)_r, g_nccl.GetErrorString(_r), #call); \
    } while (0)

// counter width for the OR: the smallest of 4 / 8 / 16 bits that holds nranks
__host__ __device__ inline int history_counter_bits(int nranks) { return nranks <= 15 ? 4 : (nranks <= 255 ? 8 : 16); }
__host__ __device__ inline int history_words(int bits) { return bits; }      // 32 flags x bits / 32

__global__ void __launch_bounds__(256)
daq_pack_kernel(uint64_t n, const uint32_t* __restrict__ q_int, const uint32_t* __restrict__ history, int bits,
                uint32_t* __restrict__ sum)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sum[i] = q_int[i];
    const uint32_t h = history[i];
    const int per_word = 32 / bits, words = history_words(bits);
    for (int w = 0; w < words; w++) {
        uint32_t packed = 0;
        for (int k = 0; k < per_word; k++) packed |= ((h >> (w * per_word + k)) & 1u) << (k * bits);
        sum[n * (1 + w) + i] = packed;
    }
}

__global__ void __launch_bounds__(256)
daq_unpack_finalize_kernel(uint64_t n, const uint32_t* __restrict__ sum, int bits, const uint32_t* __restrict__ time_int,
                           float charge_unit, uint32_t* __restrict__ q_int, uint32_t* __restrict__ history,
                           float* __restrict__ t_out, float* __restrict__ q_out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t q = sum[i];
    const int per_word = 32 / bits, words = history_words(bits);
    const uint32_t mask = (bits == 32) ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    uint32_t h = 0;
    for (int w = 0; w < words; w++) {
        const uint32_t packed = sum[n * (1 + w) + i];
        for (int k = 0; k < per_word; k++) h |= (((packed >> (k * bits)) & mask) != 0u ? 1u : 0u) << (w * per_word + k);
    }
    history[i] = h;
    q_int[i] = q;
    t_out[i] = __uint_as_float(time_int[i]);
    q_out[i] = q * charge_unit;
}

// the collective itself for accumulators that live on ONE device (ranks emulated as one kernel over
// all ranks' buffers): element-wise SUM of the packed buffers, MIN of the time words
__global__ void __launch_bounds__(256)
daq_fold_kernel(uint64_t words, uint64_t count, uint32_t* __restrict__ sum_into, const uint32_t* __restrict__ sum_from,
                uint32_t* __restrict__ time_into, const uint32_t* __restrict__ time_from)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words) sum_into[i] += sum_from[i];
    if (i < count) time_into[i] = min(time_into[i], time_from[i]);
}

} // namespace cb

using namespace cb;

extern "C" {

int cb_comm_unique_id(void* id_out)
{
    if (!id_out) return fail(CB_ERR_INVALID, "cb_comm_unique_id: null pointer");
    int rc = nccl_load();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == CB_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    CB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return CB_OK;
}

int cb_comm_init(int32_t nranks, int32_t rank, const void* id)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    if (nranks < 1 || rank < 0 || rank >= nranks || !id) return fail(CB_ERR_INVALID, "cb_comm_init: bad arguments");
    if (g_nccl.comm) return fail(CB_ERR_INVALID, "cb_comm_init: a communicator already exists (cb_comm_destroy first)");
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    CB_NCCL(g_nccl.CommInitRank(&g_nccl.comm, nranks, uid, rank));
    g_nccl.nranks = nranks; g_nccl.rank = rank;
    return CB_OK;
}

int cb_comm_destroy(void)
{
    CB_SERIALISE();
    if (g_nccl.comm) {
        stream_wait(ctx().stream);
        g_nccl.CommDestroy(g_nccl.comm);
        g_nccl.comm = nullptr;
    }
    cudaFree(g_nccl.d_sum); g_nccl.d_sum = nullptr; g_nccl.sum_cap = 0;
    g_nccl.nranks = 1; g_nccl.rank = 0;
    return CB_OK;
}

int cb_comm_size(int32_t* nranks, int32_t* rank)
{
    if (nranks) *nranks = g_nccl.comm ? g_nccl.nranks : 1;
    if (rank) *rank = g_nccl.comm ? g_nccl.rank : 0;
    return CB_OK;
}

int cb_daq_allreduce(cb_daq_t h)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    Daq* d = daqs().get(h);
    if (!d) return fail(CB_ERR_INVALID, "cb_daq_allreduce: bad handle");
    Context& c = ctx();
    Nccl& n = g_nccl;
    if (!n.comm || n.nranks == 1) return cb_daq_finalize(h);      // one rank: nothing to exchange
    const int bits = history_counter_bits(n.nranks);
    const size_t words = (size_t)d->count * (1 + history_words(bits));
    if (n.sum_cap < words) {
        cudaFree(n.d_sum); n.d_sum = nullptr; n.sum_cap = 0;
        CB_CUDA(cudaMalloc(&n.d_sum, words * sizeof(uint32_t)));
        n.sum_cap = words;
    }
    const unsigned blocks = (unsigned)((d->count + 255) / 256);
    daq_pack_kernel<<<blocks, 256, 0, c.stream>>>(d->count, d->channel_q_int, d->channel_history, bits, n.d_sum);
    CB_CUDA(cudaGetLastError());
    CB_NCCL(n.GroupStart());
    CB_NCCL(n.AllReduce(n.d_sum, n.d_sum, words, ncclUint32, ncclSum, n.comm, c.stream));
    CB_NCCL(n.AllReduce(d->earliest_time_int, d->earliest_time_int, d->count, ncclUint32, ncclMin, n.comm, c.stream));
    CB_NCCL(n.GroupEnd());
    daq_unpack_finalize_kernel<<<blocks, 256, 0, c.stream>>>(d->count, n.d_sum, bits, d->earliest_time_int,
                                                            d->geom->charge_unit, d->channel_q_int, d->channel_history,
                                                            d->earliest_time, d->channel_q);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(stream_wait(c.stream));
    return CB_OK;
}


// Test hook: what cb_daq_allreduce computes, for `n` accumulators on this device standing in for n
// ranks (same pack / SUM / MIN / unpack kernels, the exchange replaced by a local fold).  Result in
// daqs[0].  With fewer GPUs than ranks this is how the multi-rank arithmetic is checked (a spin on
// another rank's flag must not be emulated with several launches on one GPU).
int cb_daq_reduce_local(const cb_daq_t* handles, int32_t n)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    if (!handles || n < 1) return fail(CB_ERR_INVALID, "cb_daq_reduce_local: bad arguments");
    Context& c = ctx();
    std::vector<Daq*> ds;
    for (int i = 0; i < n; i++) {
        Daq* d = daqs().get(handles[i]);
        if (!d || (i && d->count != ds[0]->count)) return fail(CB_ERR_INVALID, "cb_daq_reduce_local: bad handle %d", i);
        ds.push_back(d);
    }
    const int bits = history_counter_bits(n);
    const uint64_t count = ds[0]->count;
    const size_t words = (size_t)count * (1 + history_words(bits));
    uint32_t *acc = nullptr, *tmp = nullptr;
    CB_CUDA(cudaMalloc(&acc, words * 4));
    cudaError_t e = cudaMalloc(&tmp, words * 4);
    if (e != cudaSuccess) { cudaFree(acc); return cuda_fail(e, "cudaMalloc"); }
    const unsigned blocks = (unsigned)((count + 255) / 256), wblocks = (unsigned)((words + 255) / 256);
    daq_pack_kernel<<<blocks, 256, 0, c.stream>>>(count, ds[0]->channel_q_int, ds[0]->channel_history, bits, acc);
    for (int i = 1; i < n; i++) {
        daq_pack_kernel<<<blocks, 256, 0, c.stream>>>(count, ds[i]->channel_q_int, ds[i]->channel_history, bits, tmp);
        daq_fold_kernel<<<wblocks, 256, 0, c.stream>>>(words, count, acc, tmp, ds[0]->earliest_time_int, ds[i]->earliest_time_int);
    }
    daq_unpack_finalize_kernel<<<blocks, 256, 0, c.stream>>>(count, acc, bits, ds[0]->earliest_time_int, ds[0]->geom->charge_unit,
                                                            ds[0]->channel_q_int, ds[0]->channel_history, ds[0]->earliest_time,
                                                            ds[0]->channel_q);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = stream_wait(c.stream);
    cudaFree(acc); cudaFree(tmp);
    if (e != cudaSuccess) return cuda_fail(e, "cb_daq_reduce_local");
    return CB_OK;
}

} // extern "C"
