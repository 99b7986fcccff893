// pdf.cu -- per-channel PDF / likelihood accumulators fed by the DAQ output
// (role of chroma/cuda/pdf.cu:9-368 behind chroma/gpu/pdf.py).
//
// All four entry points work on device arrays of the caller (one element per
// channel) and add ONE acquisition to them.  The three per-channel kernels are
// one thread per channel, coalesced.  cb_pdf_accumulate_eval replaces the
// reference's pair of launches with a device-wide work-queue array and a host
// synchronisation in between (gpu/pdf.py:297-330) by a single launch: the first
// blocks count Monte-Carlo hits per channel, the others own one hit channel per
// WARP, decide with ballots which copies still have to enter the nearest-
// neighbour list (the reference's running `bincount < min_bin_content` test is a
// prefix sum over the copies), and merge them into the channel's sorted list with
// a bitonic sort in shared memory instead of a one-thread insertion sort.
#include "host.h"

namespace cb {

constexpr int PDF_THREADS = 128;

__global__ void __launch_bounds__(PDF_THREADS)
pdf_bin_hits_kernel(int nchannels, const float* __restrict__ channel_q, const float* __restrict__ channel_time,
                    uint32_t* __restrict__ hitcount, int tbins, float tmin, float tmax, int qbins, float qmin,
                    float qmax, uint32_t* __restrict__ pdf)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nchannels) return;
    const unsigned int q = (unsigned int)channel_q[id];      // the DAQ charge as an integer (pdf.cu:19)
    const float t = channel_time[id];
    if (t < 1e8 && t >= tmin && t < tmax && q >= qmin && q < qmax) {
        hitcount[id] += 1;
        const int tbin = (t - tmin) / (tmax - tmin) * tbins;
        const int qbin = (q - qmin) / (qmax - qmin) * qbins;
        pdf[(size_t)id * (tbins * qbins) + tbin * qbins + qbin] += 1;   // (channel, t, q) row major
    }
}

__global__ void __launch_bounds__(PDF_THREADS)
pdf_moments_kernel(int time_only, int nchannels, const float* __restrict__ mc_time, const float* __restrict__ mc_charge,
                   float tmin, float tmax, float qmin, float qmax, uint32_t* __restrict__ mom0,
                   float* __restrict__ t_mom1, float* __restrict__ t_mom2, float* __restrict__ q_mom1,
                   float* __restrict__ q_mom2)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nchannels) return;
    const float t = mc_time[id];
    if (t < tmin || t > tmax) return;
    float q = 0.0f;
    if (!time_only) {
        q = mc_charge[id];
        if (q < qmin || q > qmax) return;
    }
    mom0[id] += 1;
    t_mom1[id] += t;
    t_mom2[id] += t * t;
    if (!time_only) {
        q_mom1[id] += q;
        q_mom2[id] += q * q;
    }
}

// Gaussian kernel of width 1/inv_bandwidth around an MC value, normalised inside [lo, hi]
// (pdf.cu:305-315 and :347-365; same expression order, the library is built with the same
// --use_fast_math)
__device__ __forceinline__ float window_norm(float lo, float hi, float mc, float inv_bandwidth)
{
    const float invroot2 = 0.70710678118654746f;
    const float rootPiBy2 = 1.2533141373155001f;
    float norm = hi - lo;
    if (inv_bandwidth > 0.0f) {
        const float loarg = (lo - mc) * inv_bandwidth * invroot2;
        const float hiarg = (hi - mc) * inv_bandwidth * invroot2;
        norm = (erff(hiarg) - erff(loarg)) * rootPiBy2;
    }
    return norm;
}

__global__ void __launch_bounds__(PDF_THREADS)
pdf_kernel_eval_kernel(int time_only, int nchannels, const uint32_t* __restrict__ event_hit,
                       const float* __restrict__ event_time, const float* __restrict__ event_charge,
                       const float* __restrict__ mc_time, const float* __restrict__ mc_charge, float tmin, float tmax,
                       float qmin, float qmax, const float* __restrict__ inv_time_bandwidths,
                       const float* __restrict__ inv_charge_bandwidths, uint32_t* __restrict__ hitcount,
                       float* __restrict__ time_pdf_values, float* __restrict__ charge_pdf_values)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nchannels) return;
    const float t = mc_time[id];
    if (t < tmin || t > tmax) return;
    float q = 0.0f;
    if (!time_only) {
        q = mc_charge[id];
        if (q < qmin || q > qmax) return;
    }
    hitcount[id] += 1;                       // this MC value is inside the PDF's range
    if (!event_hit[id]) return;              // nothing to evaluate for a channel the event did not hit
    {
        const float inv_bw = inv_time_bandwidths[id];
        const float arg = (t - event_time[id]) * inv_bw;
        if (time_only) {
            // 1-D: the kernel itself carries the 1/bandwidth (pdf.cu:303)
            const float term = expf(-0.5f * arg * arg) * inv_bw;
            time_pdf_values[id] += term / window_norm(tmin, tmax, t, inv_bw);
            return;
        }
        const float norm = window_norm(tmin, tmax, t, inv_bw);
        time_pdf_values[id] += expf(-0.5f * arg * arg) / norm;
    }
    {
        const float inv_bw = inv_charge_bandwidths[id];
        const float arg = (q - event_charge[id]) * inv_bw;
        const float norm = window_norm(qmin, qmax, q, inv_bw);
        charge_pdf_values[id] += expf(-0.5f * arg * arg) / norm;
    }
}

// One launch for GPUPDF.accumulate_pdf_eval.
//   blocks [0, count_blocks): one thread per channel, hitcount += MC copies inside [tmin, tmax]
//   the rest: one warp per hit channel (bincount + nearest-neighbour list)
__global__ void __launch_bounds__(PDF_THREADS)
pdf_accumulate_eval_kernel(int nchannels, int ndaq, int nhit, int count_blocks, const uint32_t* __restrict__ event_hit,
                           const float* __restrict__ event_time, const float* __restrict__ mc_time,
                           uint32_t* __restrict__ hitcount, uint32_t* __restrict__ bincount, float min_twidth,
                           float tmin, float tmax, int min_bin_content,
                           const uint32_t* __restrict__ map_hit_offset_to_channel_id, float* __restrict__ nearest_mc,
                           int table_cap)
{
    extern __shared__ float tables[];        // [warps per block][table_cap]
    if ((int)blockIdx.x < count_blocks) {
        const int ch = blockIdx.x * blockDim.x + threadIdx.x;
        if (ch >= nchannels) return;
        uint32_t n = 0;
        for (int i = 0; i < ndaq; i++) {
            const float t = mc_time[(size_t)nchannels * i + ch];
            n += (!(t >= 1e8f) && !(t < tmin) && !(t > tmax));      // the three tests of pdf.cu:62-70
        }
        hitcount[ch] += n;
        return;
    }
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hit_id = (blockIdx.x - count_blocks) * (PDF_THREADS / 32) + warp;
    if (hit_id >= nhit) return;              // (whole warp)
    const int ch = (int)map_hit_offset_to_channel_id[hit_id];
    if (!event_hit[ch]) return;
    float* table = tables + (size_t)warp * table_cap;
    float* list = nearest_mc + (size_t)min_bin_content * hit_id;
    const float INF = __int_as_float(0x7f800000);
    const float ev_t = event_time[ch];
    const float half_width = min_twidth / 2.0f;

    // the list as it stands: sorted, valid entries first (unused slots hold 1e9)
    int len = 0;
    for (int base = 0; base < min_bin_content; base += 32) {
        const int i = base + lane;
        const float d = (i < min_bin_content) ? list[i] : INF;
        const bool valid = (i < min_bin_content) && !(d > 1e8f);
        const unsigned m = __ballot_sync(FULL, valid);
        if (valid) table[i] = d;
        len += __popc(m);
        if (m != FULL) break;                // first unused slot seen (pdf.cu:127-128)
    }
    // the copies: |dt| inside the minimum bin counts; while the bin is short of
    // min_bin_content (counting this copy) the copy also goes to the list (pdf.cu:62-90)
    uint32_t bins = bincount[ch];
    int added = 0;
    for (int base = 0; base < ndaq; base += 32) {
        const int i = base + lane;
        float dist = INF;
        bool inside = false;
        if (i < ndaq) {
            const float t = mc_time[(size_t)nchannels * i + ch];
            inside = !(t >= 1e8f) && !(t < tmin) && !(t > tmax);
            dist = fabsf(t - ev_t);
        }
        const bool near = inside && dist < half_width;
        const unsigned near_m = __ballot_sync(FULL, near);
        const uint32_t bins_here = bins + __popc(near_m & (0xffffffffu >> (31 - lane)));   // inclusive prefix
        const bool queue = inside && bins_here < (uint32_t)min_bin_content;
        const unsigned queue_m = __ballot_sync(FULL, queue);
        if (queue) table[len + added + __popc(queue_m & ((1u << lane) - 1u))] = dist;
        added += __popc(queue_m);
        bins += __popc(near_m);
    }
    if (lane == 0) bincount[ch] = bins;
    if (added == 0) return;
    // ascending bitonic sort of the len + added entries, padded with +inf to a power of two
    const int n = len + added;
    int p = 1;
    while (p < n) p <<= 1;
    for (int i = n + lane; i < p; i += 32) table[i] = INF;
    __syncwarp();
    for (int k = 2; k <= p; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < p; i += 32) {
                const int partner = i ^ j;
                if (partner > i) {
                    const float a = table[i], b = table[partner];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { table[i] = b; table[partner] = a; }
                }
            }
            __syncwarp();
        }
    }
    const int keep = min(n, min_bin_content);
    for (int i = lane; i < keep; i += 32) list[i] = table[i];
}

static unsigned blocks_for(int n) { return (unsigned)((n + PDF_THREADS - 1) / PDF_THREADS); }

} // namespace cb

using namespace cb;

extern "C" {

int cb_pdf_bin_hits(int32_t nchannels, const float* q, const float* t, uint32_t* hitcount, int32_t tbins, float tmin,
                    float tmax, int32_t qbins, float qmin, float qmax, uint32_t* pdf)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    if (nchannels < 0 || tbins <= 0 || qbins <= 0 || !q || !t || !hitcount || !pdf)
        return fail(CB_ERR_INVALID, "cb_pdf_bin_hits: bad arguments");
    if (nchannels == 0) return CB_OK;
    Context& c = ctx();
    pdf_bin_hits_kernel<<<blocks_for(nchannels), PDF_THREADS, 0, c.stream>>>(nchannels, q, t, hitcount, tbins, tmin, tmax,
                                                                              qbins, qmin, qmax, pdf);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(stream_wait(c.stream));
    return CB_OK;
}

int cb_pdf_accumulate_moments(int32_t time_only, int32_t nchannels, const float* mc_time, const float* mc_charge,
                              float tmin, float tmax, float qmin, float qmax, uint32_t* mom0, float* t_mom1,
                              float* t_mom2, float* q_mom1, float* q_mom2)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    if (nchannels < 0 || !mc_time || !mom0 || !t_mom1 || !t_mom2 || (!time_only && (!mc_charge || !q_mom1 || !q_mom2)))
        return fail(CB_ERR_INVALID, "cb_pdf_accumulate_moments: bad arguments");
    if (nchannels == 0) return CB_OK;
    Context& c = ctx();
    pdf_moments_kernel<<<blocks_for(nchannels), PDF_THREADS, 0, c.stream>>>(time_only, nchannels, mc_time, mc_charge, tmin,
                                                                             tmax, qmin, qmax, mom0, t_mom1, t_mom2,
                                                                             q_mom1, q_mom2);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(stream_wait(c.stream));
    return CB_OK;
}

int cb_pdf_accumulate_kernel_eval(int32_t time_only, int32_t nchannels, const uint32_t* event_hit,
                                  const float* event_time, const float* event_charge, const float* mc_time,
                                  const float* mc_charge, float tmin, float tmax, float qmin, float qmax,
                                  const float* inv_time_bandwidths, const float* inv_charge_bandwidths,
                                  uint32_t* hitcount, float* time_pdf_values, float* charge_pdf_values)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    if (nchannels < 0 || !event_hit || !event_time || !mc_time || !inv_time_bandwidths || !hitcount || !time_pdf_values ||
        (!time_only && (!event_charge || !mc_charge || !inv_charge_bandwidths || !charge_pdf_values)))
        return fail(CB_ERR_INVALID, "cb_pdf_accumulate_kernel_eval: bad arguments");
    if (nchannels == 0) return CB_OK;
    Context& c = ctx();
    pdf_kernel_eval_kernel<<<blocks_for(nchannels), PDF_THREADS, 0, c.stream>>>(
        time_only, nchannels, event_hit, event_time, event_charge, mc_time, mc_charge, tmin, tmax, qmin, qmax,
        inv_time_bandwidths, inv_charge_bandwidths, hitcount, time_pdf_values, charge_pdf_values);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(stream_wait(c.stream));
    return CB_OK;
}

int cb_pdf_accumulate_eval(int32_t nchannels, int32_t ndaq, int32_t nhit, const uint32_t* event_hit,
                           const float* event_time, const float* mc_time, uint32_t* hitcount, uint32_t* bincount,
                           float min_twidth, float tmin, float tmax, int32_t min_bin_content,
                           const uint32_t* map_hit_offset_to_channel_id, float* nearest_mc)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    if (nchannels < 0 || ndaq <= 0 || nhit < 0 || min_bin_content <= 0 || !event_hit || !event_time || !mc_time ||
        !hitcount || !bincount || (nhit > 0 && (!map_hit_offset_to_channel_id || !nearest_mc)))
        return fail(CB_ERR_INVALID, "cb_pdf_accumulate_eval: bad arguments");
    if (nchannels == 0) return CB_OK;
    // shared-memory table per warp: the list plus one acquisition, rounded up to a power of two
    int cap = 32;
    while (cap < min_bin_content + ndaq) cap <<= 1;
    const size_t smem = (size_t)(PDF_THREADS / 32) * cap * sizeof(float);
    Context& c = ctx();
    if (smem > (size_t)c.max_smem_optin)
        return fail(CB_ERR_INVALID, "cb_pdf_accumulate_eval: min_bin_content + ndaq = %d does not fit in shared memory",
                    min_bin_content + ndaq);
    CB_CUDA(cudaFuncSetAttribute(pdf_accumulate_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int count_blocks = (int)blocks_for(nchannels);
    const int list_blocks = (nhit + PDF_THREADS / 32 - 1) / (PDF_THREADS / 32);
    pdf_accumulate_eval_kernel<<<(unsigned)(count_blocks + list_blocks), PDF_THREADS, smem, c.stream>>>(
        nchannels, ndaq, nhit, count_blocks, event_hit, event_time, mc_time, hitcount, bincount, min_twidth, tmin, tmax,
        min_bin_content, map_hit_offset_to_channel_id, nearest_mc, cap);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(stream_wait(c.stream));
    return CB_OK;
}

} // extern "C"
