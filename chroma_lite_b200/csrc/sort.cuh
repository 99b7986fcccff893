// sort.cuh -- hand-written device primitives for the builder and the optional ray sort:
// an exclusive scan and a stable LSD radix sort of (key, value) pairs.
//
// Radix sort, 8 bits per pass, three launches per pass:
//   digit_histogram_kernel   one CTA per tile of SORT_TILE keys: digit counts -> hist[digit][tile]
//   exclusive_scan           over hist (digit-major, so the scan yields, for every (digit, tile),
//                            the first output slot of that tile's keys with that digit)
//   digit_scatter_kernel     the same tiles again: keys are placed in index order (stable) --
//                            rank inside the warp from __match_any_sync, across the CTA's warps
//                            from per-warp digit counts in shared memory, across rounds from a
//                            running per-digit base
// Stability is what the BVH builder needs: triangles with equal Morton codes keep ascending
// triangle order (the reference's NumPy argsort is not stable).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;                       // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// ---- exclusive scan of uint32: tile sums -> scan of the sums -> rescan of each tile with its offset
static __global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_sums_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ sums)
{
    __shared__ uint32_t warp_sum[SCAN_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; w++) t += warp_sum[w];
        sums[blockIdx.x] = t;
    }
}

// one CTA: exclusive scan of `n` values in place, total -> data[n]
static __global__ void __launch_bounds__(1024)
scan_single_cta_kernel(uint32_t* data, uint32_t n)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < n) ? data[i] : 0;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t w = warp_sums[threadIdx.x];
            uint32_t s = w;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
                if (threadIdx.x >= o) s += y;
            }
            warp_sums[threadIdx.x] = s - w;
        }
        __syncthreads();
        const uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + x - v;
        if (i < n) data[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) data[n] = carry;
}

// each tile rescanned in index order: item i of the tile sits at k * SCAN_THREADS + thread
static __global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const uint32_t* __restrict__ in, uint64_t n, const uint32_t* __restrict__ tile_offsets,
                  uint32_t* __restrict__ out)
{
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    __shared__ uint32_t running;
    if (threadIdx.x == 0) running = tile_offsets[blockIdx.x];
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;
        const uint32_t v = (i < n) ? in[i] : 0;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (unsigned)o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        uint32_t before = running;
        for (unsigned w = 0; w < warp; w++) before += warp_sums[w];
        if (i < n) out[i] = before + x - v;
        __syncthreads();
        if (threadIdx.x == SCAN_THREADS - 1) running = before + x;
        __syncthreads();
    }
}

// out[i] = sum of in[0..i) ; out may alias in; `scratch` holds ceil(n / SCAN_TILE) + 1 words; the grand
// total is left in scratch[ntiles].  Returns the number of launches.
inline int exclusive_scan(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* scratch, cudaStream_t s)
{
    if (n == 0) return 0;
    const unsigned ntiles = (unsigned)((n + SCAN_TILE - 1) / SCAN_TILE);
    scan_tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, s>>>(in, n, scratch);
    scan_single_cta_kernel<<<1, 1024, 0, s>>>(scratch, ntiles);
    scan_apply_kernel<<<ntiles, SCAN_THREADS, 0, s>>>(in, n, scratch, out);
    return 3;
}
inline uint64_t exclusive_scan_scratch_words(uint64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

// ---- radix sort
constexpr int SORT_THREADS = 256;
constexpr int SORT_ROUNDS = 16;                      // rounds of SORT_THREADS consecutive keys per tile
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;

template <typename K>
__global__ void __launch_bounds__(SORT_THREADS)
digit_histogram_kernel(const K* __restrict__ keys, uint64_t n, int shift, uint32_t ntiles, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t count[256];
    count[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * SORT_TILE;
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint64_t i = base + (uint64_t)r * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&count[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x] = count[threadIdx.x];       // digit-major
}

template <typename K>
__global__ void __launch_bounds__(SORT_THREADS)
digit_scatter_kernel(const K* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t n, int shift,
                     uint32_t ntiles, const uint32_t* __restrict__ offsets, K* __restrict__ keys_out,
                     uint32_t* __restrict__ vals_out)
{
    constexpr int WARPS = SORT_THREADS / 32;
    __shared__ uint32_t base[256];                    // next output slot of every digit for this tile
    __shared__ uint32_t warp_count[WARPS][256];
    base[threadIdx.x] = offsets[(uint64_t)threadIdx.x * ntiles + blockIdx.x];
    for (int w = 0; w < WARPS; w++) warp_count[w][threadIdx.x] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t tile = (uint64_t)blockIdx.x * SORT_TILE;
    for (int r = 0; r < SORT_ROUNDS; r++) {
        const uint64_t i = tile + (uint64_t)r * SORT_THREADS + threadIdx.x;
        const bool live = i < n;
        K key = 0;
        uint32_t val = 0, digit = 256u + lane;        // dead lanes match nobody
        if (live) { key = keys_in[i]; val = vals_in[i]; digit = (uint32_t)(key >> shift) & 255u; }
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const unsigned rank = __popc(peers & ((1u << lane) - 1u));
        if (live && rank == 0) warp_count[warp][digit] = __popc(peers);
        __syncthreads();
        if (live) {
            uint32_t pos = base[digit] + rank;
            for (unsigned w = 0; w < warp; w++) pos += warp_count[w][digit];
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncthreads();
        uint32_t t = 0;
        for (int w = 0; w < WARPS; w++) { t += warp_count[w][threadIdx.x]; warp_count[w][threadIdx.x] = 0; }
        base[threadIdx.x] += t;
        __syncthreads();
    }
}

// scratch words the sort needs besides the ping-pong arrays
inline uint64_t radix_sort_scratch_words(uint64_t n)
{
    const uint64_t ntiles = (n + SORT_TILE - 1) / SORT_TILE;
    return 256 * ntiles + 1 + exclusive_scan_scratch_words(256 * ntiles + 1);
}

// Stable LSD sort of (keys, vals) by the low `bits` bits of the keys (rounded up to whole bytes).  Ping-pongs
// between (keys, vals) and (keys_alt, vals_alt); returns true when the result ended up in the *_alt arrays.
template <typename K>
inline bool radix_sort_pairs(K* keys, K* keys_alt, uint32_t* vals, uint32_t* vals_alt, uint64_t n, int bits,
                             uint32_t* scratch, cudaStream_t s, int* launches = nullptr)
{
    if (n == 0) return false;
    const uint32_t ntiles = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
    uint32_t* hist = scratch;
    uint32_t* scan_scratch = scratch + 256ull * ntiles + 1;
    bool in_alt = false;
    for (int shift = 0; shift < bits; shift += 8) {
        const K* kin = in_alt ? keys_alt : keys;
        const uint32_t* vin = in_alt ? vals_alt : vals;
        digit_histogram_kernel<K><<<ntiles, SORT_THREADS, 0, s>>>(kin, n, shift, ntiles, hist);
        exclusive_scan(hist, hist, 256ull * ntiles, scan_scratch, s);
        digit_scatter_kernel<K><<<ntiles, SORT_THREADS, 0, s>>>(kin, vin, n, shift, ntiles, hist, in_alt ? keys : keys_alt,
                                                                in_alt ? vals : vals_alt);
        if (launches) *launches += 5;
        in_alt = !in_alt;
    }
    return in_alt;
}

} // namespace cb
