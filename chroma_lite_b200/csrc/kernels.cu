// kernels.cu -- the hot path: ray intersection, photon propagation, photon-bank
// utilities and the DAQ, plus their C-ABI entry points.
#include "host.h"
#include "sort.cuh"
#include <algorithm>
#include <string.h>
#include <stdlib.h>

namespace cb {

#ifndef CB_PHYS_THREADS
#define CB_PHYS_THREADS 256
#endif
constexpr int PROP_THREADS = CB_PHYS_THREADS;   // physics kernel
constexpr int INT_THREADS = CB_INT_THREADS;   // traversal kernels

// ---------------------------------------------------------------- smem staging
// The wavelength tables (a few KB .. 48 KB) are staged once per CTA with the bulk
// async-copy engine (TMA 1-D: cp.async.bulk -> SASS UBLKCP) completing on an
// mbarrier, instead of 188-float tables being chased through three levels of
// global pointers per lookup as in the reference (photon.h:386-393).
__device__ __forceinline__ void stage_tables(float* smem_dst, const float* gsrc, uint32_t bytes,
                                             unsigned long long* mbar)
{
    if (bytes == 0) return;   // uniform across the CTA
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
        const char* src = reinterpret_cast<const char*>(gsrc);
        for (uint32_t off = 0; off < bytes; off += 16384) {
            uint32_t chunk = min(bytes - off, 16384u);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(dst + off), "l"(src + off), "r"(chunk), "r"(bar)
                : "memory");
        }
    }
    // every thread waits for phase 0 to complete
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar) : "memory");
    }
}

// ---------------------------------------------------------------- intersection
// A ray with a NaN / infinite component or a zero direction cannot be culled: every slab comparison is false,
// so the traversal would walk the WHOLE tree (139 ms for one such photon in the 37 M-triangle detector, seen as
// a 26x straggler event at 8 GPUs in round 2).  Such rays are not traced: they report "no hit", and the physics
// step flags the photon NAN_ABORT by the reference's own rule (propagate.cu:295-299) or NO_HIT.
__device__ __forceinline__ bool ray_is_traceable(const float3& o, const float3& d)
{
    const float s = o.x + o.y + o.z + d.x + d.y + d.z;                 // NaN or inf if any component is
    return isfinite(s) && isfinite(fabsf(o.x) + fabsf(o.y) + fabsf(o.z)) && (d.x != 0.0f || d.y != 0.0f || d.z != 0.0f);
}

struct RaySource {           // cb_intersect: free rays, direction normalised like distance_to_mesh
    const float* origins; const float* directions; const int32_t* last_hit;
    int32_t* tri_out; float* dist_out;
    __device__ __forceinline__ bool load(unsigned long long i, float3& o, float3& d, int& last) const
    {
        o = ld3(origins, i);
        d = ld3(directions, i);
        d = d / norm(d);
        last = last_hit ? last_hit[i] : -1;
        return ray_is_traceable(o, d);
    }
    __device__ __forceinline__ void store(unsigned long long i, int tri, float dist) const
    {
        tri_out[i] = tri;
        if (tri != -1) dist_out[i] = dist;          // untouched on a miss, like the reference
    }
};

// Persistent traversal with per-lane refill (PTrav).  Every iteration each
// lane takes ONE step: a lane with queued leaves tests a triangle, a lane without
// expands its node; both fetch from one 64-byte block, so the warp issues a single
// load sequence and waits for memory once per iteration.  Finished rays wait until
// `refill_min` lanes are free and are then finished (winner re-check + store) and
// replaced together with one atomic on the queue cursor.
//
// End game (ray splitting).  Once the queue is exhausted a launch used to end with a
// few lanes per warp walking the longest rays alone (up to ~460 iterations where the
// mean is 40; 0.6-0.7 ms, which was the whole duration of a step with few rays).  Now
// a lane that has run dry takes the top stack entry of a busy lane of its warp
// together with a copy of that ray's state, traverses that subtree as a HELPER and
// hands its best (distance, rank) back to the ray's owner lane, which stores the
// result once all its helpers are back.  Any visit order returns the same triangle
// (see PTrav), helpers start from the donor's culling limit, and the merge is the
// same lexicographic minimum, so results are unchanged bit for bit.
struct Tune { int refill_min; int split; };
#ifndef CB_TRI_NO_ALLOCATE
#define CB_TRI_NO_ALLOCATE 0
#endif
__device__ __forceinline__ uint4 ldg_no_allocate(const uint4* p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
#ifndef CB_EXP_PIPELINE
#define CB_EXP_PIPELINE 1   /* loads of children 4..7 issued while children 0..3 are being tested */
#endif
#ifndef CB_ROOMY
#define CB_ROOMY 1          /* cheap expansion when every expanding lane's stack has room for eight entries (PTrav::process4_roomy) */
#endif
#ifndef CB_INT_BLOCKS
#define CB_INT_BLOCKS 5     /* resident CTAs per SM of the traversal kernels: 5 x 128 threads at 96 registers */
#endif

template <bool COUNT, class Source>
__device__ __forceinline__ void persistent_intersect(const DevGeometry& g, const Source& src, unsigned long long n,
                                                      unsigned long long* cursor, uint32_t smem_base,
                                                      unsigned long long* counters, const Tune tune)
{
    static_assert(CB_PSTRIDE == INT_THREADS * 8u, "lane-interleaved stack stride");
    const unsigned FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t sbase = smem_base + threadIdx.x * 8u;            // this lane's stack, entry e at sbase + e*CB_PSTRIDE
    const uint32_t lbase = sbase + CB_PSTACK * CB_PSTRIDE;          // this lane's leaf queue
    PTrav tv;
    uint2 lstack[CB_PLSTACK];               // overflow of the shared-memory stack (local memory, rarely touched)
    tv.have = false; tv.sp = sbase; tv.lq = lbase; tv.lsp = 0;
    bool active = false, exhausted = false;
    int owner = (int)lane;                  // lane that owns the ray this lane works on (itself unless it is a helper)
    int pending = 0;                        // owner: helpers still out
    unsigned long long slot = 0;
    unsigned ray_iters = 0;                 // iterations spent on the current ray (statistics only)
    TraverseCounters cnt = {0, 0, 0};
    for (;;) {
        if (exhausted && tune.split) {
            // helpers that have finished their share report to the owner and become free
            unsigned hm = __ballot_sync(FULL, active && owner != (int)lane && !tv.have && tv.lq == lbase);
            while (hm) {
                const int h = __ffs(hm) - 1;
                hm &= hm - 1;
                const float ht = __shfl_sync(FULL, tv.best_t, h);
                const uint32_t hr = __shfl_sync(FULL, tv.best_rank, h);
                const int htri = __shfl_sync(FULL, tv.best_tri, h);
                const int hredo = __shfl_sync(FULL, (int)tv.redo, h);
                const int ho = __shfl_sync(FULL, owner, h);
                if ((int)lane == ho) {
                    if (htri != -1 && (ht < tv.best_t || (ht == tv.best_t && hr < tv.best_rank))) {
                        tv.best_t = ht; tv.best_rank = hr; tv.best_tri = htri;
                        tv.limit = ht + 2e-5f * ht;
                    }
                    tv.redo = tv.redo || (hredo != 0);
                    pending--;
                }
                if ((int)lane == h) { active = false; owner = (int)lane; }
            }
        }
        const bool done = active && !tv.have && tv.lq == lbase && owner == (int)lane && pending == 0;
        const unsigned dm = __ballot_sync(FULL, done);
        const unsigned fm = dm | __ballot_sync(FULL, !active);
        const int nfree = __popc(fm);
        if (nfree == 32 || nfree >= tune.refill_min || (exhausted && dm)) {
            if (done) {
                float dist;
                const int tri = tv.template finish<COUNT>(g, lbase, dist, (uint32_t*)(counters + 3), &cnt);
                if (COUNT) {                               // ray-length statistics (debug aid)
                    atomicMax(counters + 9, (unsigned long long)ray_iters);
                    if (ray_iters > 100) atomicAdd(counters + 10, 1ull);
                    if (ray_iters > 300) atomicAdd(counters + 11, 1ull);
                    if (ray_iters > 1000) atomicAdd(counters + 12, 1ull);
                    atomicAdd(counters + 13, (unsigned long long)ray_iters);
                    ray_iters = 0;
                }
                src.store(slot, tri, dist);
                active = false;
            }
            if (!exhausted) {
                unsigned long long base = 0;
                const int leader = __ffs(fm) - 1;
                if ((int)lane == leader) base = atomicAdd(cursor, (unsigned long long)nfree);
                base = __shfl_sync(FULL, base, leader);
                if (!active) {
                    const unsigned long long q = base + __popc(fm & lt_mask);
                    if (q < n) {
                        float3 o, d;
                        int last;
                        slot = q;
                        if (src.load(q, o, d, last)) {
                            active = tv.init(g, o, d, last, sbase, lbase);
                            if (!active) src.store(slot, -1, -1.0f);
                        } else {
                            src.store(slot, -1, -1.0f);      // skipped (terminal at step 0) or not traceable
                        }
                    }
                }
                exhausted = (base + nfree >= n);
            }
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }
        if (exhausted && tune.split) {
            // free lanes take the top stack entry of busy lanes (k-th free lane from the k-th busy one)
            const bool can_give = active && tv.have && tv.sp > sbase;
            const unsigned idle_m = __ballot_sync(FULL, !active);
            const unsigned donor_m = __ballot_sync(FULL, can_give);
            const int np = min(__popc(idle_m), __popc(donor_m));
            if (np > 0) {
                __syncwarp();
                const bool giving = can_give && __popc(donor_m & lt_mask) < np;
                const bool taking = !active && __popc(idle_m & lt_mask) < np;
                uint2 e = make_uint2(0u, 0u);
if (giving) { tv.sp -= CB_PSTRIDE; e = lds64(tv.sp); }
                const int src = taking ? (int)__fns(donor_m, 0, __popc(idle_m & lt_mask) + 1) : (int)lane;
                const uint32_t ex = __shfl_sync(FULL, e.x, src);
                const float et = __uint_as_float(__shfl_sync(FULL, e.y, src));
                PhasedRay r2;
                r2.sx = __shfl_sync(FULL, tv.r.sx, src); r2.sy = __shfl_sync(FULL, tv.r.sy, src); r2.sz = __shfl_sync(FULL, tv.r.sz, src);
                r2.nx = __shfl_sync(FULL, tv.r.nx, src); r2.ny = __shfl_sync(FULL, tv.r.ny, src); r2.nz = __shfl_sync(FULL, tv.r.nz, src);
                r2.fx = __shfl_sync(FULL, tv.r.fx, src); r2.fy = __shfl_sync(FULL, tv.r.fy, src); r2.fz = __shfl_sync(FULL, tv.r.fz, src);
                r2.selx = __shfl_sync(FULL, tv.r.selx, src); r2.sely = __shfl_sync(FULL, tv.r.sely, src);
                r2.selz = __shfl_sync(FULL, tv.r.selz, src);
                const float bt = __shfl_sync(FULL, tv.best_t, src);
                const uint32_t br = __shfl_sync(FULL, tv.best_rank, src);
                const float lim = __shfl_sync(FULL, tv.limit, src);
                const int lh = __shfl_sync(FULL, tv.last_hit, src);
                const int own = __shfl_sync(FULL, owner, src);
                const bool took = taking && !(et > lim);        // an entry behind the limit is simply dropped
                if (took) {
                    tv.r = r2;
                    tv.best_t = bt; tv.best_rank = br; tv.best_tri = -1; tv.limit = lim; tv.last_hit = lh;
                    tv.cur = ex; tv.cur_t = et; tv.have = true; tv.redo = false;
                    tv.sp = sbase; tv.lq = lbase; tv.lsp = 0;
                    // the ray itself (origin, direction) sits in three slots behind the donor's leaf queue
                    const uint32_t cold_src = lbase + ((uint32_t)src - lane) * 8u + CB_PLEAF * CB_PSTRIDE;
                    const uint32_t cold_dst = lbase + CB_PLEAF * CB_PSTRIDE;
#pragma unroll
                    for (int k = 0; k < CB_PCOLD; k++) {
                        const uint2 c = lds64(cold_src + k * CB_PSTRIDE);
                        sts64(cold_dst + k * CB_PSTRIDE, c.x, c.y);
                    }
                    active = true; owner = own;
                }
                // owners count the helpers that just left
                unsigned tm = __ballot_sync(FULL, took);
                while (tm) {
                    const int t = __ffs(tm) - 1;
                    tm &= tm - 1;
                    if ((int)lane == __shfl_sync(FULL, owner, t)) pending++;
                }
                __syncwarp();
            }
        }
        // One step per lane and iteration: a lane with queued leaves tests a triangle, a lane
        // without expands its node.  Both kinds fetch four 16-byte words from one 64-byte
        // block (a tri64 record / four child entries), so they share ONE load sequence and
        // the warp waits for memory once per iteration.
        if (COUNT) ray_iters += active;
        uint32_t tri = 0;
        const bool do_tri = active && tv.pop_leaf(lbase, tri);
        const bool do_exp = active && !do_tri && tv.lq == lbase && tv.have;
        const uint32_t first = tv.cur & 0x0FFFFFFFu, n = tv.cur >> 28;
#if CB_ROOMY
        // room for eight more entries on every expanding lane's shared-memory stack: the cheap form of the expansion
        const bool all_roomy = __all_sync(FULL, !do_exp || tv.sp + 8u * CB_PSTRIDE <= sbase + CB_PSTACK * CB_PSTRIDE);
#endif
        const uint4* blk = do_tri ? reinterpret_cast<const uint4*>(g.tri64) + 4ull * tri : g.nodes + first;
        const uint32_t last_k = do_tri ? 2u : n - 1u;
        uint4 q[4];
#if CB_TRI_NO_ALLOCATE
        // triangle records are read once per test and rarely again: they bypass L1 (no_allocate) so that it keeps
        // the node blocks, which neighbouring rays share.  Both load sequences are issued before either is
        // consumed, so the warp still waits for memory once per iteration.
        if (do_tri) {
#pragma unroll
            for (int k = 0; k < 3; k++) q[k] = ldg_no_allocate(blk + k);
        }
        if (do_exp) {
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = __ldg(blk + min((uint32_t)k, last_k));
        }
#else
        if (do_tri || do_exp) {
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = __ldg(blk + min((uint32_t)k, last_k));
        }
#endif
        if (do_tri) {
            if (COUNT) cnt.tris++;
            tv.test_triangle(lbase, tri, *reinterpret_cast<const float4*>(&q[0]), *reinterpret_cast<const float4*>(&q[1]),
                             *reinterpret_cast<const float4*>(&q[2]));
        } else if (do_exp) {
#if CB_ROOMY
            if (all_roomy) {
                PTrav::Picked pk = {__int_as_float(0x7f800000), 0u};
                const uint32_t sp0 = tv.sp;
#if CB_EXP_PIPELINE == 2
                const bool more = n > 4;
                tv.template roomy_child<COUNT, true>(q[0], 0, n, pk, &cnt, g.tri64);
                if (more) q[0] = __ldg(g.nodes + first + 4u);
                tv.template roomy_child<COUNT>(q[1], 1, n, pk, &cnt, g.tri64);
                if (more) q[1] = __ldg(g.nodes + first + min(5u, n - 1u));
                tv.template roomy_child<COUNT>(q[2], 2, n, pk, &cnt, g.tri64);
                if (more) q[2] = __ldg(g.nodes + first + min(6u, n - 1u));
                tv.template roomy_child<COUNT>(q[3], 3, n, pk, &cnt, g.tri64);
                if (more) {
                    q[3] = __ldg(g.nodes + first + min(7u, n - 1u));
                    tv.template process4_roomy<COUNT>(q, 4, n, pk, &cnt, g.tri64);
                }
#elif CB_EXP_PIPELINE
                // children 4..7 sit in the other two sectors of the node's 128-byte line: their loads are issued
                // as soon as a pair of registers is free, so that the second memory round trip of an expansion
                // (91 % of them have more than four children) overlaps the tests of children 0..3
                const bool more = n > 4;
                tv.template roomy_child<COUNT, true>(q[0], 0, n, pk, &cnt, g.tri64);
                tv.template roomy_child<COUNT>(q[1], 1, n, pk, &cnt, g.tri64);
                if (more) { q[0] = __ldg(g.nodes + first + 4u); q[1] = __ldg(g.nodes + first + min(5u, n - 1u)); }
                tv.template roomy_child<COUNT>(q[2], 2, n, pk, &cnt, g.tri64);
                tv.template roomy_child<COUNT>(q[3], 3, n, pk, &cnt, g.tri64);
                if (more) {
                    q[2] = __ldg(g.nodes + first + min(6u, n - 1u)); q[3] = __ldg(g.nodes + first + min(7u, n - 1u));
                    tv.template process4_roomy<COUNT>(q, 4, n, pk, &cnt, g.tri64);
                }
#else
                tv.template process4_roomy<COUNT>(q, 0, n, pk, &cnt, g.tri64);
                if (n > 4) {
#pragma unroll
                    for (int k = 0; k < 4; k++) q[k] = __ldg(g.nodes + first + min(4u + k, n - 1u));
                    tv.template process4_roomy<COUNT>(q, 4, n, pk, &cnt, g.tri64);
                }
#endif
                tv.expand_end_roomy(pk, sp0, sbase, lstack);
            } else
#endif
            {
                PTrav::Nearest nr = {0u, __int_as_float(0x7f800000)};
                tv.template process4<COUNT>(q, 0, n, nr, sbase, lstack, &cnt, g.nodes, g.tri64);
                if (n > 4) {
#pragma unroll
                    for (int k = 0; k < 4; k++) q[k] = __ldg(g.nodes + first + min(4u + k, n - 1u));
                    tv.template process4<COUNT>(q, 4, n, nr, sbase, lstack, &cnt, g.nodes, g.tri64);
                }
                tv.expand_end(nr, sbase, lstack);
            }
        }
        if (active && !do_exp) tv.after_leaves(sbase, lbase, lstack);
    }
    if (COUNT) {
        atomicAdd(counters + 1, (unsigned long long)cnt.nodes);
        atomicAdd(counters + 2, (unsigned long long)cnt.tris);
        atomicAdd(counters + 5, (unsigned long long)cnt.resolved);
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(INT_THREADS, CB_INT_BLOCKS)
intersect_kernel(const __grid_constant__ DevGeometry g, const __grid_constant__ RaySource src, uint64_t n,
                  unsigned long long* counters, Tune tune)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    persistent_intersect<COUNT>(g, src, n, counters, (uint32_t)__cvta_generic_to_shared(smem_raw), counters, tune);
}

// ---------------------------------------------------------------- propagation
// Wavefront scheduler.  One propagate call = a loop over physics steps; each step is
//   step_intersect_kernel : one ray per thread, traversal only (few registers, high
//                           occupancy), writes (triangle, distance) per photon;
//   step_physics_kernel   : bulk + surface physics for the same photons, then the
//                           survivors are appended to the next queue with one
//                           warp-aggregated atomic per warp (ballot + popc prefix).
// The queues hold chunk-local photon indices; RNG state k belongs to photon
// first+k (replay contract), so results do not depend on the
// order in which photons are scheduled.  When few photons are left the remaining
// steps run in ONE persistent launch (propagate_tail_kernel) instead of a launch
// pair per step.  This replaces the reference's host loop, which relaunches the
// monolithic kernel per step, reloads 108 B of state per photon per step and
// blocks on a 4-byte D2H copy between steps (gpu/photon.py:259-286).
struct PropParams {
    CbPhotonBank bank;
    uint32_t* rng;                  // state of photon (first + k) is rng[k]
    uint64_t first;                 // chunk base
    const uint32_t* queue_in;       // chunk-local indices, nullptr = identity
    uint32_t* queue_out;
    int32_t* hit_tri;               // result of the step's traversal, per QUEUE SLOT (dense for the physics kernel)
    float* hit_dist;
    const uint32_t* hit_slots;      // queue the physics kernel walks, when the traversal kernel walked another (sorted) one
    // The number of queued photons lives on the device, so that consecutive steps can be
    // launched without the host reading it back: step s reads n_in[0], appends its
    // survivors to queue_out and counts them in n_out[0]; `cursor` is this step's work
    // counter.  A wavefront kernel runs only while more than `tail_at` photons are queued;
    // the tail kernel launched behind it takes over (and ends the call) once there are
    // `tail_at` or fewer.
    const unsigned long long* n_in;
    unsigned long long* n_out;
    unsigned long long* cursor;
    unsigned long long tail_at;
    int32_t step;                   // steps already taken by every photon in queue_in
    int32_t max_steps, use_weights, scatter_first;
    unsigned long long* counters;   // [1] nodes, [2] tris, [3] overflow, [4] steps, [5] resolved, [6] wavefront rays,
                                    // [7] tail photons, [8] tail steps, [9..13] ray-length stats
};

// Queue entries: chunk-local photon index in the low 31 bits; bit 31 marks a photon that sits on
// a surface (last_hit_triangle >= 0), which the tail kernel schedules first.
constexpr uint32_t QUEUE_ON_SURFACE = 0x80000000u;
__device__ __forceinline__ uint32_t queue_index(const PropParams& P, unsigned long long q)
{
    return P.queue_in ? (P.queue_in[q] & ~QUEUE_ON_SURFACE) : (uint32_t)q;
}

__device__ __forceinline__ void load_photon(const CbPhotonBank& b, uint64_t id, uint32_t hist, bool normalise, Photon& p)
{
    p.pos = ld3(b.pos, id);
    p.dir = ld3(b.dir, id);
    p.pol = ld3(b.pol, id);
    if (normalise) {                 // once per propagate call, like the reference kernel's prologue
        p.dir = p.dir / norm(p.dir);
        p.pol = p.pol / norm(p.pol);
    }
    p.wavelength = b.wavelengths[id];
    p.time = b.t[id];
    p.last_hit_triangle = b.last_hit_triangles[id];
    p.history = hist;
    p.weight = b.weights[id];
}
__device__ __forceinline__ void store_photon(const CbPhotonBank& b, uint64_t id, const Photon& p)
{
    st3(b.pos, id, p.pos);
    st3(b.dir, id, p.dir);
    st3(b.pol, id, p.pol);
    b.wavelengths[id] = p.wavelength;
    b.t[id] = p.time;
    b.flags[id] = p.history;
    b.last_hit_triangles[id] = p.last_hit_triangle;
    b.weights[id] = p.weight;
}

// Coherence key of a ray: coarse origin cell (4 bits/axis of the world box) above
// the direction in an octahedral map (Morton-interleaved, 9 bits/axis).  Sorting
// the step's queue by this key puts rays that walk the same part of the tree into
// the same warp: fewer divergent iterations and the nodes they share stay in L1.
__device__ __forceinline__ uint32_t spread9(uint32_t x)
{
    x &= 0x000001ffu;                       // interleave zeros between the low 9 bits
    x = (x ^ (x << 8)) & 0x00ff00ffu;
    x = (x ^ (x << 4)) & 0x0f0f0f0fu;
    x = (x ^ (x << 2)) & 0x33333333u;
    x = (x ^ (x << 1)) & 0x55555555u;
    return x;
}
__global__ void __launch_bounds__(256)
ray_key_kernel(DevGeometry g, PropParams P, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *P.n_in) return;
    const uint32_t k = queue_index(P, i);
    const uint64_t id = P.first + k;
    const float3 pos = ld3(P.bank.pos, id);
    float3 d = ld3(P.bank.dir, id);
    const float extent = 65535.0f * g.world_scale;
    const float inv = 16.0f / extent;
    const uint32_t cx = (uint32_t)fminf(fmaxf((pos.x - g.world_origin.x) * inv, 0.0f), 15.0f);
    const uint32_t cy = (uint32_t)fminf(fmaxf((pos.y - g.world_origin.y) * inv, 0.0f), 15.0f);
    const uint32_t cz = (uint32_t)fminf(fmaxf((pos.z - g.world_origin.z) * inv, 0.0f), 15.0f);
    // octahedral map of the direction to [0,1]^2
    const float s = 1.0f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z) + 1e-30f);
    float u = d.x * s, v = d.y * s;
    if (d.z < 0.0f) {
        const float uu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f);
        const float vv = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f);
        u = uu; v = vv;
    }
    const uint32_t qu = (uint32_t)fminf(fmaxf((u * 0.5f + 0.5f) * 512.0f, 0.0f), 511.0f);
    const uint32_t qv = (uint32_t)fminf(fmaxf((v * 0.5f + 0.5f) * 512.0f, 0.0f), 511.0f);
    const uint32_t dirkey = spread9(qu) | (spread9(qv) << 1);            // 18 bits
    keys[i] = (((cx << 8) | (cy << 4) | cz) << 18) | dirkey;            // 30 bits
    vals[i] = i;                                                        // the queue slot: where the hit goes
}

struct PhotonRaySource {     // one propagation step: rays of the photons in the queue
    const PropParams& P;
    __device__ __forceinline__ bool load(unsigned long long q, float3& o, float3& d, int& last) const
    {
        const uint32_t k = queue_index(P, P.hit_slots ? (unsigned long long)P.hit_slots[q] : q);
        const uint64_t id = P.first + k;
        if (P.step == 0 && (P.bank.flags[id] & 0xFFFFu & CB_TERMINAL)) return false;   // never ran: untouched
        o = ld3(P.bank.pos, id);
        d = ld3(P.bank.dir, id);
        if (P.step == 0) d = d / norm(d);
        last = P.bank.last_hit_triangles[id];
        return ray_is_traceable(o, d);
    }
    __device__ __forceinline__ void store(unsigned long long q, int tri, float dist) const
    {
        // normally by queue slot; with a coherence-sorted queue (CHROMA_B200_SORT) the slot of the
        // physics kernel's queue comes from the sort's value array
        const unsigned long long slot = P.hit_slots ? (unsigned long long)P.hit_slots[q] : q;
        P.hit_tri[slot] = tri;
        P.hit_dist[slot] = dist;
    }
};

template <bool COUNT>
__global__ void __launch_bounds__(INT_THREADS, CB_INT_BLOCKS)
step_intersect_kernel(const __grid_constant__ DevGeometry g, const __grid_constant__ PropParams P, Tune tune)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const unsigned long long n = *P.n_in;
    if (n <= P.tail_at) return;              // the tail kernel behind this launch takes these
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.counters + 6, n);   // rays traced by the wavefront
    const PhotonRaySource src = {P};
    persistent_intersect<COUNT>(g, src, n, P.cursor, (uint32_t)__cvta_generic_to_shared(smem_raw), P.counters, tune);
}

// The thin-film optics are ONE shared, non-inlined pure function (physics.cuh thin_film): an inlined copy
// would be contracted differently in each kernel and results would then depend on the schedule
// (test_config3_scheduler_invariance_full_size).  It takes and returns values only, so no kernel has to
// keep its photon in local memory around the call.
#ifndef CB_PHYS_BLOCKS
#define CB_PHYS_BLOCKS 2
#endif
template <bool WIRES>
__global__ void __launch_bounds__(PROP_THREADS, CB_PHYS_BLOCKS)
step_physics_kernel(const __grid_constant__ DevGeometry g, const __grid_constant__ PropParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned long long mbar;
    const uint32_t n_in = (uint32_t)*P.n_in;
    if (n_in <= P.tail_at) return;           // (uniform) the tail kernel behind this launch takes these
    float* stab = reinterpret_cast<float*>(smem_raw);
    stage_tables(stab, g.tables, g.smem_bytes, &mbar);
    Tables T = {stab, g.tables, g.smem_floats};
    const unsigned lane = threadIdx.x & 31u;
    const bool last_step = (P.step + 1 >= P.max_steps);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.counters + 4, (unsigned long long)n_in);   // steps taken
    // grid-stride over whole warps so the ballot below always sees full warps.  Two rounds of
    // loads per photon instead of four dependent ones: (queue entry, hit triangle, hit distance) by
    // queue slot, then (flags, state, RNG, triangle record) together.
    const uint32_t n_round = (n_in + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        bool alive = false;
        uint32_t k = 0, mark = 0;
        if (i < n_in) {
            k = queue_index(P, i);
            const int hit_tri = P.hit_tri[i];
            const float hit_dist = P.hit_dist[i];
            const uint64_t id = P.first + k;
            if (hit_tri >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(g.tri64 + 4ull * (uint32_t)hit_tri));
            const uint32_t hist = P.bank.flags[id] & 0xFFFFu;
            Photon p;
            load_photon(P.bank, id, hist, P.step == 0, p);
            Rng rng = rng_load(P.rng, k);
            if (!(P.step == 0 && (hist & CB_TERMINAL))) {
                if (photon_is_nan(p)) {
                    p.history |= CB_NO_HIT | CB_NAN_ABORT;
                } else {
                    alive = physics_step<WIRES>(g, T, p, rng, hit_tri, hit_dist, P.use_weights != 0,
                                         P.step == 0 ? P.scatter_first : 0);
                }
                rng_store(P.rng, k, rng);
                store_photon(P.bank, id, p);
                alive = alive && !last_step;
                mark = (p.last_hit_triangle >= 0) ? QUEUE_ON_SURFACE : 0u;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, alive);
        if (m) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(P.n_out, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (alive) P.queue_out[base + __popc(m & ((1u << lane) - 1u))] = k | mark;
        }
    }
}

// Persistent tail, one photon per WARP: the warp traverses cooperatively
// (warp_traverse) and all lanes run the physics of their photon redundantly, so a
// step costs a few microseconds instead of the ~50 us of a single-thread step.
// Warps claim photons from the queue with one atomic each until it is empty.
#ifndef CB_TAIL_THREADS
#define CB_TAIL_THREADS 1024   /* one CTA per SM: the tables are staged once per SM and ~150 KB stay L1 */
#endif
constexpr int TAIL_THREADS = CB_TAIL_THREADS;
#ifndef CB_TAIL_BLOCKS
#define CB_TAIL_BLOCKS 1     /* 64 registers: 32 resident warps (photons) per SM */
#endif
template <bool COUNT, bool WIRES>
__global__ void __launch_bounds__(TAIL_THREADS, CB_TAIL_BLOCKS)
propagate_tail_kernel(const __grid_constant__ DevGeometry g, const __grid_constant__ PropParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned long long mbar;
    const unsigned long long n_in = *P.n_in;
    if (n_in > P.tail_at || n_in == 0) return;       // (uniform) still the wavefront kernels' turn, or nothing left
    float* stab = reinterpret_cast<float*>(smem_raw);
    const uint32_t tab_bytes = g.smem_bytes;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint2* wstack = reinterpret_cast<uint2*>(smem_raw + ((tab_bytes + 127u) & ~127u)) + warp * (CB_WSTACK + CB_WLEAF);
    uint2* wleaf = wstack + CB_WSTACK;
    stage_tables(stab, g.tables, tab_bytes, &mbar);
    Tables T = {stab, g.tables, g.smem_floats};
    TraverseCounters cnt = {0, 0, 0};
    unsigned long long nsteps_total = 0;

    // The queue is walked twice: photons sitting on a surface first, photons in the bulk second.
    // Scheduling only (every photon is taken exactly once, results do not depend on it): the rare
    // 50-step histories that end a propagate call are almost all photons rattling between
    // surfaces (90 % of those alive after 25 steps were on a surface when the tail began, against
    // 38 % of all its photons), and started early they finish under the cover of the bulk.
    for (;;) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(P.cursor, 1ull);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= 2 * n_in) break;
        const bool first_pass = q < n_in;
        if (!first_pass) q -= n_in;
        const uint32_t entry = P.queue_in ? P.queue_in[q] : ((uint32_t)q | QUEUE_ON_SURFACE);
        if (((entry & QUEUE_ON_SURFACE) != 0) != first_pass) continue;
        const uint32_t k = entry & ~QUEUE_ON_SURFACE;
        const uint64_t id = P.first + k;
        const uint32_t hist = P.bank.flags[id] & 0xFFFFu;
        if (P.step == 0 && (hist & CB_TERMINAL)) continue;
        Photon p;
        load_photon(P.bank, id, hist, P.step == 0, p);
        Rng rng = rng_load(P.rng, k);
        int steps = P.step;
        int sf = (P.step == 0) ? P.scatter_first : 0;
        bool alive = true;
        while (alive && steps < P.max_steps) {
            steps++;
            nsteps_total++;
            if (photon_is_nan(p)) {
                p.history |= CB_NO_HIT | CB_NAN_ABORT;
                alive = false;
            } else {
                float dist = -1.0f;
                const int tri = !ray_is_traceable(p.pos, p.dir) ? -1 :
                                warp_traverse<COUNT>(g, p.pos, p.dir, p.last_hit_triangle, dist, wstack, wleaf,
                                                     (uint32_t*)(P.counters + 3), &cnt);
                alive = physics_step<WIRES>(g, T, p, rng, tri, dist, P.use_weights != 0, sf);
                sf = 0;
            }
        }
        if (lane == 0) {
            rng_store(P.rng, k, rng);
            store_photon(P.bank, id, p);
#ifdef CB_TAIL_PROFILE
            // debug build only: steps taken here and finish time (x 64 ns, global timer) per photon
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            P.hit_tri[k] = steps - P.step;
            reinterpret_cast<uint32_t*>(P.hit_dist)[k] = (uint32_t)(now >> 6);
#endif
        }
    }
    if (lane == 0 && nsteps_total) { atomicAdd(P.counters + 4, nsteps_total); atomicAdd(P.counters + 8, nsteps_total); }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.counters + 7, n_in);
    if (COUNT) {
        atomicAdd(P.counters + 1, (unsigned long long)cnt.nodes);
        atomicAdd(P.counters + 2, (unsigned long long)cnt.tris);
        if (lane == 0) atomicAdd(P.counters + 5, (unsigned long long)cnt.resolved);
    }
}

// Persistent tail, second form: the warp still traverses ONE ray at a time cooperatively (warp_traverse), but
// it owns up to 32 photons, one per lane, and their physics runs lane-parallel.  In the one-photon-per-warp
// kernel above every lane executes the same ~1.1 k physics instructions of the same photon, a third of a
// step's instructions; here those are issued once for up to 32 photons.  A step of the warp is
//   refill     free lanes claim queue entries (one atomic per attempt for the whole warp)
//   traverse   for every lane that holds a photon: its ray is read from shared memory by all lanes and
//              traversed by the warp; the hit goes to the owning lane
//   physics    each lane advances its own photon; finished photons are stored and their lanes freed
// The photons live in shared memory between phases ([field][lane] per warp, 22 words per photon), so the
// traversal's registers are not shared with 20 words of photon state and the kernel keeps its 32 warps per SM.
constexpr int TAIL_SLOT_WORDS = 22;   // pos dir pol (9) wavelength time weight history last_hit (5) rng (6) index, steps|sf
template <bool COUNT, bool WIRES>
__global__ void __launch_bounds__(TAIL_THREADS, CB_TAIL_BLOCKS)
propagate_tail_lanes_kernel(const __grid_constant__ DevGeometry g, const __grid_constant__ PropParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned long long mbar;
    const unsigned long long n_in = *P.n_in;
    if (n_in > P.tail_at || n_in == 0) return;       // (uniform) still the wavefront kernels' turn, or nothing left
    constexpr int WARPS = TAIL_THREADS / 32;
    const unsigned FULL = 0xffffffffu;
    float* stab = reinterpret_cast<float*>(smem_raw);
    const uint32_t tab_bytes = g.smem_bytes;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned char* after_tables = smem_raw + ((tab_bytes + 127u) & ~127u);
    uint2* wstack = reinterpret_cast<uint2*>(after_tables) + warp * (CB_WSTACK + CB_WLEAF);
    uint2* wleaf = wstack + CB_WSTACK;
    uint32_t* slots = reinterpret_cast<uint32_t*>(after_tables + (size_t)WARPS * (CB_WSTACK + CB_WLEAF) * sizeof(uint2)) +
                      warp * (TAIL_SLOT_WORDS * 32);
    uint32_t* mine = slots + lane;                    // field f of this lane's photon: mine[f * 32]
    stage_tables(stab, g.tables, tab_bytes, &mbar);
    Tables T = {stab, g.tables, g.smem_floats};
    TraverseCounters cnt = {0, 0, 0};
    unsigned long long nsteps_total = 0;
    bool have = false, exhausted = false;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.counters + 7, n_in);
    // few photons: spread them over all warps of the grid (a warp traverses its photons' rays one after the
    // other, so 32 photons in one warp while other warps idle would serialise them)
    const unsigned long long warps_in_grid = (unsigned long long)gridDim.x * WARPS;
    const unsigned lanes_cap = (unsigned)min(32ull, (n_in + warps_in_grid - 1) / warps_in_grid);
    const unsigned may_hold = lanes_cap >= 32u ? FULL : ((1u << lanes_cap) - 1u);

    for (;;) {
        // ---- refill: the queue is walked twice, photons sitting on a surface first (see the kernel above)
        for (int attempt = 0; attempt < 4 && !exhausted; attempt++) {
            const unsigned need = __ballot_sync(FULL, !have) & may_hold;
            if (!need) break;
            const int want = __popc(need), leader = __ffs(need) - 1;
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(P.cursor, (unsigned long long)want);
            base = __shfl_sync(FULL, base, leader);
            exhausted = base + want >= 2 * n_in;
            unsigned long long q = base + __popc(need & lt_mask);
            if (!have && ((need >> lane) & 1u) && q < 2 * n_in) {
                const bool first_pass = q < n_in;
                if (!first_pass) q -= n_in;
                const uint32_t entry = P.queue_in ? P.queue_in[q] : ((uint32_t)q | QUEUE_ON_SURFACE);
                const uint32_t k = entry & ~QUEUE_ON_SURFACE;
                const uint64_t id = P.first + k;
                if (((entry & QUEUE_ON_SURFACE) != 0) == first_pass) {
                    const uint32_t hist = P.bank.flags[id] & 0xFFFFu;
                    if (!(P.step == 0 && (hist & CB_TERMINAL))) {
                        Photon p;
                        load_photon(P.bank, id, hist, P.step == 0, p);
                        const Rng rng = rng_load(P.rng, k);
                        mine[0 * 32] = __float_as_uint(p.pos.x); mine[1 * 32] = __float_as_uint(p.pos.y); mine[2 * 32] = __float_as_uint(p.pos.z);
                        mine[3 * 32] = __float_as_uint(p.dir.x); mine[4 * 32] = __float_as_uint(p.dir.y); mine[5 * 32] = __float_as_uint(p.dir.z);
                        mine[6 * 32] = __float_as_uint(p.pol.x); mine[7 * 32] = __float_as_uint(p.pol.y); mine[8 * 32] = __float_as_uint(p.pol.z);
                        mine[9 * 32] = __float_as_uint(p.wavelength); mine[10 * 32] = __float_as_uint(p.time);
                        mine[11 * 32] = __float_as_uint(p.weight); mine[12 * 32] = p.history; mine[13 * 32] = (uint32_t)p.last_hit_triangle;
                        mine[14 * 32] = rng.d; mine[15 * 32] = rng.v0; mine[16 * 32] = rng.v1; mine[17 * 32] = rng.v2;
                        mine[18 * 32] = rng.v3; mine[19 * 32] = rng.v4;
                        mine[20 * 32] = k;
                        const int sf = (P.step == 0) ? P.scatter_first : 0;
                        mine[21 * 32] = (uint32_t)P.step | ((uint32_t)(sf + 1) << 16);
                        have = true;
                    }
                }
            }
        }
        __syncwarp();
        if (!__any_sync(FULL, have)) {
            if (exhausted) break;
            continue;
        }
        // ---- traversal, one ray at a time, by the whole warp
        bool nan = false;
        if (have) {
            nan = isnan(__uint_as_float(mine[3 * 32]) * __uint_as_float(mine[4 * 32]) * __uint_as_float(mine[5 * 32]) *
                        __uint_as_float(mine[0 * 32]) * __uint_as_float(mine[1 * 32]) * __uint_as_float(mine[2 * 32]));
        }
        const bool traceable = have && !nan &&
            ray_is_traceable(f3(__uint_as_float(mine[0 * 32]), __uint_as_float(mine[1 * 32]), __uint_as_float(mine[2 * 32])),
                             f3(__uint_as_float(mine[3 * 32]), __uint_as_float(mine[4 * 32]), __uint_as_float(mine[5 * 32])));
        unsigned todo = __ballot_sync(FULL, traceable);
        int my_tri = -1;
        float my_dist = -1.0f;
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t* ray = slots + src;
            const float3 o = f3(__uint_as_float(ray[0 * 32]), __uint_as_float(ray[1 * 32]), __uint_as_float(ray[2 * 32]));
            const float3 d = f3(__uint_as_float(ray[3 * 32]), __uint_as_float(ray[4 * 32]), __uint_as_float(ray[5 * 32]));
            const int last = (int)ray[13 * 32];
            float dist;
            const int tri = warp_traverse<COUNT>(g, o, d, last, dist, wstack, wleaf, (uint32_t*)(P.counters + 3), &cnt);
            if ((int)lane == src) { my_tri = tri; my_dist = dist; }
        }
        // ---- physics, one photon per lane
        if (have) {
            Photon p;
            p.pos = f3(__uint_as_float(mine[0 * 32]), __uint_as_float(mine[1 * 32]), __uint_as_float(mine[2 * 32]));
            p.dir = f3(__uint_as_float(mine[3 * 32]), __uint_as_float(mine[4 * 32]), __uint_as_float(mine[5 * 32]));
            p.pol = f3(__uint_as_float(mine[6 * 32]), __uint_as_float(mine[7 * 32]), __uint_as_float(mine[8 * 32]));
            p.wavelength = __uint_as_float(mine[9 * 32]); p.time = __uint_as_float(mine[10 * 32]);
            p.weight = __uint_as_float(mine[11 * 32]); p.history = mine[12 * 32]; p.last_hit_triangle = (int)mine[13 * 32];
            Rng rng = {mine[14 * 32], mine[15 * 32], mine[16 * 32], mine[17 * 32], mine[18 * 32], mine[19 * 32]};
            const uint32_t k = mine[20 * 32];
            const uint32_t packed = mine[21 * 32];
            const int steps = (int)(packed & 0xFFFFu) + 1, sf = (int)(packed >> 16) - 1;
            nsteps_total++;
            bool alive;
            if (nan) {
                p.history |= CB_NO_HIT | CB_NAN_ABORT;
                alive = false;
            } else {
                alive = physics_step<WIRES>(g, T, p, rng, my_tri, my_dist, P.use_weights != 0, sf);
            }
            if (!alive || steps >= P.max_steps) {
                rng_store(P.rng, k, rng);
                store_photon(P.bank, P.first + k, p);
                have = false;
            } else {
                mine[0 * 32] = __float_as_uint(p.pos.x); mine[1 * 32] = __float_as_uint(p.pos.y); mine[2 * 32] = __float_as_uint(p.pos.z);
                mine[3 * 32] = __float_as_uint(p.dir.x); mine[4 * 32] = __float_as_uint(p.dir.y); mine[5 * 32] = __float_as_uint(p.dir.z);
                mine[6 * 32] = __float_as_uint(p.pol.x); mine[7 * 32] = __float_as_uint(p.pol.y); mine[8 * 32] = __float_as_uint(p.pol.z);
                mine[9 * 32] = __float_as_uint(p.wavelength); mine[10 * 32] = __float_as_uint(p.time);
                mine[11 * 32] = __float_as_uint(p.weight); mine[12 * 32] = p.history; mine[13 * 32] = (uint32_t)p.last_hit_triangle;
                mine[14 * 32] = rng.d; mine[15 * 32] = rng.v0; mine[16 * 32] = rng.v1; mine[17 * 32] = rng.v2;
                mine[18 * 32] = rng.v3; mine[19 * 32] = rng.v4;
                mine[21 * 32] = (uint32_t)steps | (1u << 16);          // scatter_first applies to the first step only
            }
        }
        __syncwarp();
    }
    for (int o = 16; o > 0; o >>= 1) nsteps_total += __shfl_down_sync(FULL, nsteps_total, o);
    if (lane == 0 && nsteps_total) { atomicAdd(P.counters + 4, nsteps_total); atomicAdd(P.counters + 8, nsteps_total); }
    if (COUNT) {
        atomicAdd(P.counters + 1, (unsigned long long)cnt.nodes);
        atomicAdd(P.counters + 2, (unsigned long long)cnt.tris);
        if (lane == 0) atomicAdd(P.counters + 5, (unsigned long long)cnt.resolved);
    }
}

// ---------------------------------------------------------------- bank utilities
struct BankPtrs { CbPhotonBank b; };

__device__ __forceinline__ void copy_photon(const CbPhotonBank& s, uint64_t i, const CbPhotonBank& d, uint64_t o)
{
    st3(d.pos, o, ld3(s.pos, i));
    st3(d.dir, o, ld3(s.dir, i));
    st3(d.pol, o, ld3(s.pol, i));
    d.wavelengths[o] = s.wavelengths[i];
    d.t[o] = s.t[i];
    d.flags[o] = s.flags[i];
    d.last_hit_triangles[o] = s.last_hit_triangles[i];
    d.weights[o] = s.weights[i];
    if (d.evidx && s.evidx) d.evidx[o] = s.evidx[i];
}

// replicate the first n photons `copies` more times at stride n (propagate.cu:29-68)
__global__ void duplicate_kernel(CbPhotonBank b, uint64_t n, int copies)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 1; c <= copies; c++) copy_photon(b, i, b, i + n * (uint64_t)c);
}

__device__ __forceinline__ int photon_channel(const CbPhotonBank& b, uint64_t id, uint32_t flag,
                                              const uint32_t* solid_map, const int32_t* solid_to_channel)
{
    int tri = b.last_hit_triangles[id];
    if ((b.flags[id] & flag) && tri > -1) return solid_to_channel[solid_map[tri]];
    return -1;
}

// Stable two-pass compaction (count per 1024-photon tile -> exclusive scan ->
// ordered scatter).  The reference appends in atomicAdd arrival order
// (propagate.cu:97-139, 201-251), which is run-to-run nondeterministic; keeping
// photon order makes hit lists reproducible at the same cost.
constexpr int TILE = 1024;
template <bool HITS>
__global__ void __launch_bounds__(256)
tile_count_kernel(CbPhotonBank b, uint64_t first, uint64_t n, uint32_t flag, const uint32_t* solid_map,
                  const int32_t* solid_to_channel, uint32_t* tile_counts)
{
    __shared__ uint32_t total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    uint32_t c = 0;
    uint64_t base = (uint64_t)blockIdx.x * TILE;
    for (int k = threadIdx.x; k < TILE; k += 256) {
        uint64_t i = base + k;
        if (i < n) {
            bool sel = HITS ? (photon_channel(b, first + i, flag, solid_map, solid_to_channel) >= 0)
                            : ((b.flags[first + i] & flag) != 0);
            c += sel;
        }
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&total, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

// single-CTA exclusive scan of the tile counts; total -> counts[ntiles]
__global__ void __launch_bounds__(1024)
tile_scan_kernel(uint32_t* counts, uint32_t ntiles)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = (i < ntiles) ? counts[i] : 0;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sums[threadIdx.x], s = w;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
                if (threadIdx.x >= o) s += y;
            }
            warp_sums[threadIdx.x] = s - w;
        }
        __syncthreads();
        uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + x - v;
        if (i < ntiles) counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[ntiles] = carry;
}

// PACKED: the destination is ONE block holding the ten hit arrays back to back, each as long as the
// number of hits (words per hit 3,3,3,1,1,1,1,1,1 + channel; the layout GPUPhotons.get_flat_hits reads back in
// one copy).  The total is only known on the device (tile_offsets[ntiles]), so the array bases are derived
// here: the host enqueues count, scan and scatter without reading anything back in between.
__device__ __forceinline__ CbPhotonBank packed_hit_bank(uint32_t* block, uint64_t total, int32_t** channels)
{
    const uint64_t m = total ? total : 1;
    CbPhotonBank d;
    float* f = reinterpret_cast<float*>(block);
    d.pos = f; d.dir = f + 3 * m; d.pol = f + 6 * m; d.wavelengths = f + 9 * m; d.t = f + 10 * m;
    d.last_hit_triangles = reinterpret_cast<int32_t*>(block + 11 * m);
    d.flags = block + 12 * m; d.weights = f + 13 * m; d.evidx = block + 14 * m;
    d.n = total;
    *channels = reinterpret_cast<int32_t*>(block + 15 * m);
    return d;
}

template <bool HITS, bool PACKED = false>
__global__ void __launch_bounds__(256)
tile_scatter_kernel(CbPhotonBank src, uint64_t first, uint64_t n, uint32_t flag, const uint32_t* solid_map,
                    const int32_t* solid_to_channel, const uint32_t* tile_offsets, CbPhotonBank dst,
                    int32_t* channels_out)
{
    __shared__ uint32_t warp_base[8];
    __shared__ uint32_t running;
    if (PACKED) dst = packed_hit_bank(reinterpret_cast<uint32_t*>(dst.pos), tile_offsets[gridDim.x], &channels_out);
    if (threadIdx.x == 0) running = tile_offsets[blockIdx.x];
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t base = (uint64_t)blockIdx.x * TILE;
    for (int k0 = 0; k0 < TILE; k0 += 256) {
        uint64_t i = base + k0 + threadIdx.x;
        int ch = -1;
        bool sel = false;
        if (i < n) {
            if (HITS) { ch = photon_channel(src, first + i, flag, solid_map, solid_to_channel); sel = ch >= 0; }
            else sel = (src.flags[first + i] & flag) != 0;
        }
        unsigned m = __ballot_sync(0xffffffffu, sel);
        if (lane == 0) warp_base[warp] = __popc(m);
        __syncthreads();
        uint32_t off = running;
        for (unsigned w = 0; w < warp; w++) off += warp_base[w];
        if (sel) {
            uint64_t o = off + __popc(m & ((1u << lane) - 1u));
            copy_photon(src, first + i, dst, o);
            if (HITS) channels_out[o] = ch;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < 8; w++) t += warp_base[w];
            running += t;
        }
        __syncthreads();
    }
}

// gather by queue (propagate.cu:141-169)
__global__ void gather_queue_kernel(CbPhotonBank src, const uint32_t* queue, uint64_t n, CbPhotonBank dst)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    copy_photon(src, queue[i], dst, i);
}

// ---------------------------------------------------------------- DAQ
// one photon per thread: weight test, smeared time, charge -> per-channel
// atomicMin / atomicAdd / atomicOr (behaviour of daq.cu:35-86; time ordering on
// raw float bits, valid for t >= 0)
__global__ void __launch_bounds__(256)
daq_kernel(uint32_t* __restrict__ rng_states, uint32_t detection_state, uint64_t first_photon, uint64_t nphotons,
           CbPhotonBank b, const uint32_t* __restrict__ solid_map, const int32_t* __restrict__ solid_to_channel,
           const float* __restrict__ time_cdf_x, const float* __restrict__ time_cdf_y, int time_cdf_len,
           const float* __restrict__ charge_cdf_x, const float* __restrict__ charge_cdf_y, int charge_cdf_len,
           float charge_unit, uint32_t* earliest_time_int, uint32_t* channel_q_int, uint32_t* channel_histories,
           float global_weight)
{
    uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nphotons) return;
    uint64_t photon_id = id + first_photon;
    int triangle_id = b.last_hit_triangles[photon_id];
    if (triangle_id > -1) {
        int solid_id = solid_map[triangle_id];
        uint32_t history = b.flags[photon_id];
        int channel_index = solid_to_channel[solid_id];
        if (channel_index >= 0 && (history & detection_state)) {
            Rng rng = rng_load(rng_states, id);
            float weight = b.weights[photon_id] * global_weight;
            if (rng_uniform(rng) < weight) {
                float time = b.t[photon_id] + interp_xy(rng_uniform(rng), time_cdf_len, time_cdf_y, time_cdf_x);
                uint32_t time_int = __float_as_uint(time);
                float charge = interp_xy(rng_uniform(rng), charge_cdf_len, charge_cdf_y, charge_cdf_x);
                uint32_t charge_int = roundf(charge / charge_unit);
                atomicMin(earliest_time_int + channel_index, time_int);
                atomicAdd(channel_q_int + channel_index, charge_int);
                atomicOr(channel_histories + channel_index, history);
            }
            rng_store(rng_states, id, rng);
        }
    }
}

// cuRAND's Box-Muller normal on the XORWOW stream (curand_normal.h: two 32-bit
// draws -> two normals, second one cached in the state)
__device__ __forceinline__ float rng_normal(Rng& s, uint32_t& flag, float& extra)
{
    if (flag) { flag = 0; return extra; }
    uint32_t x = rng_next(s), y = rng_next(s);
    float u = x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
    float v = y * (2.3283064e-10f * 6.2831855f) + (2.3283064e-10f * 6.2831855f / 2.0f);
    float r = sqrtf(-2.0f * logf(u));
    float sv, cv;
    __sincosf(v, &sv, &cv);
    extra = r * cv;
    flag = 1;
    return r * sv;
}

// one CTA per photon, threads stride over the ndaq virtual DAQ copies (daq.cu:88-150)
__global__ void daq_many_kernel(uint32_t* __restrict__ rng_states, float* bm_extra, uint32_t* bm_flag,
                                uint32_t detection_state, uint64_t first_photon, CbPhotonBank b,
                                const uint32_t* __restrict__ solid_map, const int32_t* __restrict__ solid_to_channel,
                                const float* __restrict__ time_cdf_x, const float* __restrict__ time_cdf_y, int time_cdf_len,
                                const float* __restrict__ charge_cdf_x, const float* __restrict__ charge_cdf_y, int charge_cdf_len,
                                float charge_unit, uint32_t* earliest_time_int, uint32_t* channel_q_int,
                                uint32_t* channel_histories, int ndaq, int channel_stride, float global_weight)
{
    uint64_t photon_id = first_photon + blockIdx.x;
    int triangle_id = b.last_hit_triangles[photon_id];
    if (triangle_id == -1) return;
    int channel_index = (triangle_id > -1) ? solid_to_channel[solid_map[triangle_id]] : -1;
    uint32_t history = b.flags[photon_id];
    if (channel_index < 0 || !(history & detection_state)) return;
    float photon_time = b.t[photon_id];
    float weight = b.weights[photon_id] * global_weight;
    uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    Rng rng = rng_load(rng_states, id);
    uint32_t flag = bm_flag[id];
    float extra = bm_extra[id];
    for (int i = threadIdx.x; i < ndaq; i += blockDim.x) {
        int channel_offset = channel_index + i * channel_stride;
        if (rng_uniform(rng) < weight) {
            float time = photon_time + rng_normal(rng, flag, extra) +
                         interp_xy(rng_uniform(rng), time_cdf_len, time_cdf_y, time_cdf_x);
            uint32_t time_int = __float_as_uint(time);
            float charge = interp_xy(rng_uniform(rng), charge_cdf_len, charge_cdf_y, charge_cdf_x);
            uint32_t charge_int = roundf(charge / charge_unit);
            atomicMin(earliest_time_int + channel_offset, time_int);
            atomicAdd(channel_q_int + channel_offset, charge_int);
            atomicOr(channel_histories + channel_offset, history);
        }
    }
    rng_store(rng_states, id, rng);
    bm_flag[id] = flag;
    bm_extra[id] = extra;
}

// fused finaliser: time bits -> float, integer charge -> float (daq.cu:152-173)
__global__ void daq_finalize_kernel(uint64_t n, const uint32_t* time_int, const uint32_t* q_int, float charge_unit,
                                    float* t_out, float* q_out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    t_out[i] = __uint_as_float(time_int[i]);
    q_out[i] = q_int[i] * charge_unit;
}

// dst (+)= src, the three atomics of run_daq applied array to array
__global__ void daq_fold_into_kernel(uint64_t n, uint32_t* __restrict__ t_dst, const uint32_t* __restrict__ t_src,
                                     uint32_t* __restrict__ q_dst, const uint32_t* __restrict__ q_src,
                                     uint32_t* __restrict__ h_dst, const uint32_t* __restrict__ h_src)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    t_dst[i] = min(t_dst[i], t_src[i]);
    q_dst[i] += q_src[i];
    h_dst[i] |= h_src[i];
}

static int check_bank(const CbPhotonBank* b, const char* who)
{
    if (!b) return fail(CB_ERR_INVALID, "%s: null photon bank", who);
    if (b->n && (!b->pos || !b->dir || !b->pol || !b->wavelengths || !b->t || !b->last_hit_triangles || !b->flags || !b->weights))
        return fail(CB_ERR_INVALID, "%s: photon bank has null arrays", who);
    return CB_OK;
}

static size_t stack_smem_bytes() { return (size_t)(CB_PSTACK + CB_PLEAF + CB_PCOLD) * INT_THREADS * sizeof(uint2); }
static Tune tune_from_env()
{
    Tune t = {10, 1};   // refill once 10 lanes of a warp are free (swept 6..20: 8-10 best, 12 is 4 % slower)
    if (const char* e = getenv("CHROMA_B200_REFILL_MIN")) t.refill_min = atoi(e);
    if (const char* e = getenv("CHROMA_B200_SPLIT")) t.split = atoi(e);
    return t;
}

// cudaFuncSetAttribute + occupancy query of a kernel, done once per (kernel, CTA size, dynamic shared memory):
// together they cost ~50 us per propagate call, 1 % of a 2.5 M-photon event
static int kernel_setup(const void* fn, int threads, size_t smem, int* per_sm)
{
    struct Key { const void* fn; int threads; size_t smem; int per_sm; };
    struct Limit { const void* fn; size_t smem; };
    static std::vector<Key> cache;
    static std::vector<Limit> limits;          // dynamic shared memory each kernel has been allowed so far (only ever raised)
    static std::mutex mu;
    std::lock_guard<std::mutex> l(mu);
    Limit* lim = nullptr;
    for (Limit& x : limits) if (x.fn == fn) lim = &x;
    if (!lim) { limits.push_back({fn, 0}); lim = &limits.back(); }
    if (smem > lim->smem || lim->smem == 0) {
        CB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 16)));
        lim->smem = std::max<size_t>(smem, 16);
    }
    for (const Key& k : cache)
        if (k.fn == fn && k.threads == threads && k.smem == smem) { *per_sm = k.per_sm; return CB_OK; }
    int n = 0;
    CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, threads, smem));
    cache.push_back({fn, threads, smem, n});
    *per_sm = n;
    return CB_OK;
}

static int ensure_tile_scratch(uint64_t ntiles)
{
    Context& c = ctx();
    if (c.block_counts_cap < ntiles + 1) {
        cudaFree(c.d_block_counts);
        c.d_block_counts = nullptr;
        c.block_counts_cap = 0;
        CB_CUDA(cudaMalloc(&c.d_block_counts, (ntiles + 1) * sizeof(uint32_t)));
        c.block_counts_cap = ntiles + 1;
    }
    return CB_OK;
}

template <bool HITS>
static int compact(const CbPhotonBank* src, uint64_t first, uint64_t n, uint32_t flag, Geometry* g,
                   const CbPhotonBank* dst, int32_t* d_channels, uint32_t* count_out)
{
    Context& c = ctx();
    if (count_out) *count_out = 0;
    if (n == 0) return CB_OK;
    if (first + n > src->n) return fail(CB_ERR_INVALID, "compaction range exceeds the photon bank");
    const uint64_t ntiles = (n + TILE - 1) / TILE;
    int rc = ensure_tile_scratch(ntiles);
    if (rc) return rc;
    const uint32_t* solid_map = g ? g->solid_id : nullptr;
    const int32_t* s2c = g ? g->solid_to_channel : nullptr;
    tile_count_kernel<HITS><<<(unsigned)ntiles, 256, 0, c.stream>>>(*src, first, n, flag, solid_map, s2c, c.d_block_counts);
    tile_scan_kernel<<<1, 1024, 0, c.stream>>>(c.d_block_counts, (uint32_t)ntiles);
    CB_CUDA(cudaGetLastError());
    uint32_t total = 0;
    CB_CUDA(cudaMemcpyAsync(&total, c.d_block_counts + ntiles, 4, cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(stream_wait(c.stream));
    if (count_out) *count_out = total;
    if (dst && total) {
        if (dst->n < total) return fail(CB_ERR_INVALID, "destination bank too small (%llu < %u)", (unsigned long long)dst->n, total);
        tile_scatter_kernel<HITS><<<(unsigned)ntiles, 256, 0, c.stream>>>(*src, first, n, flag, solid_map, s2c,
                                                                         c.d_block_counts, *dst, d_channels);
        CB_CUDA(cudaGetLastError());
        CB_CUDA(stream_wait(c.stream));
    }
    return CB_OK;
}

} // namespace cb

using namespace cb;

extern "C" {

int cb_intersect(cb_geom_t gh, const float* d_origins, const float* d_directions, const int32_t* d_last_hit,
                 uint64_t n, int32_t* d_triangle_out, float* d_distance_out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    Geometry* g = geoms().get(gh);
    if (!g) return fail(CB_ERR_INVALID, "cb_intersect: bad geometry handle");
    if (n == 0) return CB_OK;
    if (!d_origins || !d_directions || !d_triangle_out || !d_distance_out)
        return fail(CB_ERR_INVALID, "cb_intersect: null array");
    Context& c = ctx();
    if (int lrc = l2_pin_tree_prefix(g)) return lrc;
    CB_CUDA(cudaMemsetAsync(c.d_counters, 0, 16 * sizeof(unsigned long long), c.stream));
    const size_t smem = stack_smem_bytes();
    int per_sm = 0;
    if (int src = kernel_setup((const void*)intersect_kernel<false>, INT_THREADS, smem, &per_sm)) return src;
    if (per_sm < 1) per_sm = 1;
    unsigned blocks = (unsigned)std::min<uint64_t>((n + INT_THREADS - 1) / INT_THREADS, (uint64_t)c.sm_count * per_sm);
    RaySource src = {d_origins, d_directions, d_last_hit, d_triangle_out, d_distance_out};
    intersect_kernel<false><<<blocks, INT_THREADS, smem, c.stream>>>(g->dev, src, n, c.d_counters, tune_from_env());
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cudaMemcpyAsync(c.h_counters, c.d_counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
    CB_CUDA(stream_wait(c.stream));
    if (c.h_counters[3]) return fail(CB_ERR_CUDA, "cb_intersect: traversal stack overflow");
    return CB_OK;
}

// Scheduler of one propagate call (replaces the host loop of gpu/photon.py:259-286).
// Step s = [step_intersect, step_physics] while many photons are alive, then ONE
// warp-cooperative persistent launch for everything that is left.  The alive count
// stays on the device (PropParams::n_in/n_out): up to STEP_BATCH steps are enqueued
// back to back, each followed by a guarded tail launch, and the host reads the count
// once per batch.  Kernels whose guard fails return at once (a few microseconds).
constexpr int STEP_BATCH = 4;

int cb_propagate(const CbPhotonBank* bank, cb_geom_t gh, cb_rng_t rh, int32_t nthreads_per_block,
                 int32_t max_blocks, int32_t max_steps, int32_t use_weights, int32_t scatter_first,
                 CbPropagateStats* stats)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    int rc = check_bank(bank, "cb_propagate");
    if (rc) return rc;
    Geometry* g = geoms().get(gh);
    RngPool* r = rngs().get(rh);
    if (!g) return fail(CB_ERR_INVALID, "cb_propagate: bad geometry handle");
    if (!r) return fail(CB_ERR_INVALID, "cb_propagate: bad rng handle");
    if (nthreads_per_block <= 0 || max_blocks <= 0) return fail(CB_ERR_INVALID, "cb_propagate: bad launch parameters");
    const uint64_t pool = std::min<uint64_t>((uint64_t)nthreads_per_block * (uint64_t)max_blocks, r->n);
    if (pool == 0) return fail(CB_ERR_INVALID, "cb_propagate: empty rng pool");
    if (pool >= (1ull << 31)) return fail(CB_ERR_INVALID, "cb_propagate: rng pool too large");
    if (stats) { memset(stats, 0, sizeof(*stats)); }
    if (bank->n == 0 || max_steps <= 0) return CB_OK;
    Context& c = ctx();
    if (int lrc = l2_pin_tree_prefix(g)) return lrc;
    const bool count = getenv("CHROMA_B200_STATS") != nullptr;
    const Tune tune = tune_from_env();
    const bool trace = getenv("CHROMA_B200_TRACE") != nullptr;   // per-step timing to stderr (debug aid)
    const bool timeline = getenv("CHROMA_B200_TIMELINE") != nullptr;   // event after every launch, printed at the end
    std::vector<cudaEvent_t> tl_ev; std::vector<const char*> tl_name;
    auto mark = [&](const char* name) {
        if (!timeline) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c.stream); tl_ev.push_back(e); tl_name.push_back(name);
    };
    cudaEvent_t tev[3] = {nullptr, nullptr, nullptr};
    if (trace) for (auto& e : tev) cudaEventCreate(&e);
    // hand the rest to the persistent kernel once the survivors fit on the chip ~1.3 times over
    const uint64_t tail_threshold = getenv("CHROMA_B200_TAIL") ? (uint64_t)atoll(getenv("CHROMA_B200_TAIL")) : (uint64_t)(1.3 * 2048 * ctx().sm_count);
    const uint64_t sort_threshold = getenv("CHROMA_B200_SORT") ? (uint64_t)atoll(getenv("CHROMA_B200_SORT")) : 0;
    // the host follows the alive count step by step only when it needs it (tracing, ray sorting)
    const int batch = (trace || sort_threshold) ? 1 : STEP_BATCH;

    // scratch: two queues + hit arrays, sized for one chunk; per-step counters
    const uint64_t cap = std::min<uint64_t>(pool, bank->n);
    if (c.scratch_cap < cap) {
        cudaFree(c.d_queue[0]); cudaFree(c.d_queue[1]); cudaFree(c.d_hit_tri); cudaFree(c.d_hit_dist);
        c.d_queue[0] = c.d_queue[1] = nullptr; c.d_hit_tri = nullptr; c.d_hit_dist = nullptr; c.scratch_cap = 0;
        CB_CUDA(cudaMalloc(&c.d_queue[0], cap * 4)); CB_CUDA(cudaMalloc(&c.d_queue[1], cap * 4));
        CB_CUDA(cudaMalloc(&c.d_hit_tri, cap * 4)); CB_CUDA(cudaMalloc(&c.d_hit_dist, cap * 4));
        cudaFree(c.d_keys[0]); cudaFree(c.d_keys[1]); cudaFree(c.d_sorted); cudaFree(c.d_sort_tmp);
        c.d_keys[0] = c.d_keys[1] = c.d_sorted = nullptr; c.d_sort_tmp = nullptr;
        CB_CUDA(cudaMalloc(&c.d_keys[0], cap * 4)); CB_CUDA(cudaMalloc(&c.d_keys[1], cap * 4));
        CB_CUDA(cudaMalloc(&c.d_sorted, cap * 4));
        c.sort_tmp_bytes = radix_sort_scratch_words(cap) * 4;
        CB_CUDA(cudaMalloc(&c.d_sort_tmp, c.sort_tmp_bytes));
        c.scratch_cap = cap;
    }
    const size_t nslots = (size_t)max_steps + 2;          // alive count before step s, s = 0..max_steps
    if (c.step_slots < nslots) {
        cudaFree(c.d_step_counts); c.d_step_counts = nullptr; c.step_slots = 0;
        CB_CUDA(cudaMalloc(&c.d_step_counts, 2 * nslots * sizeof(unsigned long long)));
        c.step_slots = nslots;
    }
    unsigned long long* d_alive = c.d_step_counts;                 // [s]: photons queued for step s
    unsigned long long* d_cursor = c.d_step_counts + c.step_slots; // [s]: work counter of step s

    auto k_int = count ? step_intersect_kernel<true> : step_intersect_kernel<false>;
    const bool wires = g->dev.nwireplanes > 0;
    // CHROMA_B200_TAIL_MODE=warp: one photon per warp, physics executed by all lanes (round 1);
    // default: 32 photons per warp, physics lane-parallel (propagate_tail_lanes_kernel)
    const bool tail_lanes = !(getenv("CHROMA_B200_TAIL_MODE") && strcmp(getenv("CHROMA_B200_TAIL_MODE"), "warp") == 0);
    auto k_tail = tail_lanes
        ? (wires ? (count ? propagate_tail_lanes_kernel<true, true> : propagate_tail_lanes_kernel<false, true>)
                 : (count ? propagate_tail_lanes_kernel<true, false> : propagate_tail_lanes_kernel<false, false>))
        : (wires ? (count ? propagate_tail_kernel<true, true> : propagate_tail_kernel<false, true>)
                 : (count ? propagate_tail_kernel<true, false> : propagate_tail_kernel<false, false>));
    auto k_phys = wires ? step_physics_kernel<true> : step_physics_kernel<false>;
    const size_t smem_int = stack_smem_bytes();
    const size_t smem_tab = (g->smem_table_bytes + 127u) & ~127u;
    const size_t smem_tail = smem_tab + (size_t)(TAIL_THREADS / 32) * (CB_WSTACK + CB_WLEAF) * sizeof(uint2) +
                             (tail_lanes ? (size_t)(TAIL_THREADS / 32) * TAIL_SLOT_WORDS * 32 * sizeof(uint32_t) : 0);
    int phys_per_sm = 0, tail_per_sm = 0, int_per_sm = 0;
    if ((rc = kernel_setup((const void*)k_int, INT_THREADS, smem_int, &int_per_sm))) return rc;
    if ((rc = kernel_setup((const void*)k_phys, PROP_THREADS, smem_tab, &phys_per_sm))) return rc;
    if ((rc = kernel_setup((const void*)k_tail, TAIL_THREADS, smem_tail, &tail_per_sm))) return rc;
    if (int_per_sm < 1 || phys_per_sm < 1 || tail_per_sm < 1) return fail(CB_ERR_CUDA, "cb_propagate: kernels do not fit on an SM");

    unsigned long long tot[16] = {0};
    uint32_t launches = 0;
    // per-class kernel time: an event before and after every launch (no host synchronisation);
    // class of interval k in cls[k]: 0 traversal, 1 physics, 2 tail
    std::vector<int> cls;
    size_t nev = 0;
    auto stamp = [&]() -> cudaEvent_t {
        if (nev == c.class_ev.size()) { cudaEvent_t e; cudaEventCreate(&e); c.class_ev.push_back(e); }
        cudaEvent_t e = c.class_ev[nev++];
        cudaEventRecord(e, c.stream);
        return e;
    };
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
    uint64_t int0_rays = 0;
    CB_CUDA(cudaEventRecord(c.kev0, c.stream));
    mark("kev0");
    // photons beyond the pool reuse states chunk by chunk, in order, exactly as
    // the reference's chunk_iterator does for one step (gpu/photon.py:266-268)
    for (uint64_t first = 0; first < bank->n; first += pool) {
        const uint64_t cnt = std::min<uint64_t>(pool, bank->n - first);
        PropParams P;
        P.bank = *bank; P.rng = r->states; P.first = first;
        P.hit_tri = c.d_hit_tri; P.hit_dist = c.d_hit_dist; P.hit_slots = nullptr;
        P.max_steps = max_steps; P.use_weights = use_weights; P.scatter_first = scatter_first;
        P.counters = c.d_counters;
        CB_CUDA(cudaMemsetAsync(c.d_counters, 0, 16 * sizeof(unsigned long long), c.stream));
        CB_CUDA(cudaMemsetAsync(c.d_step_counts, 0, 2 * c.step_slots * sizeof(unsigned long long), c.stream));
        c.h_counters[17] = cnt;
        CB_CUDA(cudaMemcpyAsync(d_alive, c.h_counters + 17, sizeof(unsigned long long), cudaMemcpyHostToDevice, c.stream));
        uint64_t n_alive = cnt;            // host's view: exact after every read-back, an upper bound in between
        bool exact = true;
        bool counters_read = false;        // the call's counters came back with the last alive count
        int step = 0;
        while (step < max_steps && n_alive > 0) {
            const int nb = std::min(batch, max_steps - step);
            for (int b = 0; b < nb && n_alive > 0; b++, step++) {
                P.step = step;
                P.queue_in = (step == 0) ? nullptr : c.d_queue[(step - 1) & 1];
                P.queue_out = c.d_queue[step & 1];
                P.n_in = d_alive + step; P.n_out = d_alive + step + 1; P.cursor = d_cursor + step;
                const bool last = (step + 1 == max_steps);
                P.tail_at = last ? ~0ull : tail_threshold;
                const bool wave = !last && !(exact && n_alive <= tail_threshold);
                const bool tail = last || !exact || n_alive <= tail_threshold;
                if (wave) {
                    const uint64_t blocks = (n_alive + PROP_THREADS - 1) / PROP_THREADS;
                    if (sort_threshold && exact && n_alive >= sort_threshold) {
                        // regroup the queue so that warps hold rays of similar origin and direction
                        // (the intersect kernel alone walks the sorted order; the physics keeps the queue's)
                        uint32_t* unsorted_vals = c.d_queue[step & 1];     // free until the physics kernel writes it
                        ray_key_kernel<<<(unsigned)((n_alive + 255) / 256), 256, 0, c.stream>>>(g->dev, P, c.d_keys[0], unsorted_vals);
                        // 30-bit keys: four 8-bit passes, the result is back in (d_keys[0], unsorted_vals)
                        int sort_launches = 0;
                        const bool in_alt = radix_sort_pairs<uint32_t>(c.d_keys[0], c.d_keys[1], unsorted_vals, c.d_sorted, n_alive, 32,
                                                                       (uint32_t*)c.d_sort_tmp, c.stream, &sort_launches);
                        if (!in_alt)
                            CB_CUDA(cudaMemcpyAsync(c.d_sorted, unsorted_vals, n_alive * 4, cudaMemcpyDeviceToDevice, c.stream));
                        CB_CUDA(cudaGetLastError());
                        launches += 1 + sort_launches;
                    }
                    const unsigned iblocks = (unsigned)std::min<uint64_t>((n_alive + INT_THREADS - 1) / INT_THREADS,
                                                                          (uint64_t)c.sm_count * int_per_sm);
                    const bool time_it = (step == 0 && first == 0);
                    if (time_it) CB_CUDA(cudaEventRecord(c.iev0, c.stream));
                    if (trace) cudaEventRecord(tev[0], c.stream);
                    PropParams PI = P;
                    if (sort_threshold && exact && n_alive >= sort_threshold) PI.hit_slots = c.d_sorted;
                    mark("pre-int");
                    cudaEvent_t e0 = stamp();
                    k_int<<<iblocks, INT_THREADS, smem_int, c.stream>>>(g->dev, PI, tune);
                    cudaEvent_t e1 = stamp();
                    spans.push_back({e0, e1}); cls.push_back(0);
                    mark("int");
                    if (trace) cudaEventRecord(tev[1], c.stream);
                    if (time_it) { CB_CUDA(cudaEventRecord(c.iev1, c.stream)); int0_rays = n_alive; }
                    const unsigned pblocks = (unsigned)std::min<uint64_t>(blocks, (uint64_t)c.sm_count * phys_per_sm);
                    k_phys<<<pblocks, PROP_THREADS, smem_tab, c.stream>>>(g->dev, P);
                    cudaEvent_t e2 = stamp();
                    spans.push_back({e1, e2}); cls.push_back(1);
                    mark("phys");
                    if (trace) cudaEventRecord(tev[2], c.stream);
                    CB_CUDA(cudaGetLastError());
                    launches += 2;
                }
                if (tail) {
                    // everything that is left, in one persistent launch; one photon per warp
                    const uint64_t per_block = TAIL_THREADS / 32;      // warps first; the lanes kernel then fills lanes
                    const uint64_t most = (last || exact) ? n_alive : std::min<uint64_t>(n_alive, tail_threshold);
                    const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((most + per_block - 1) / per_block,
                                                                                             (uint64_t)c.sm_count * tail_per_sm));
                    if (trace) cudaEventRecord(tev[0], c.stream);
                    cudaEvent_t t0 = stamp();
                    k_tail<<<blocks, TAIL_THREADS, smem_tail, c.stream>>>(g->dev, P);
                    cudaEvent_t t1 = stamp();
                    spans.push_back({t0, t1}); cls.push_back(2);
                    mark("tail");
                    CB_CUDA(cudaGetLastError());
                    if (trace) {
                        cudaEventRecord(tev[1], c.stream); cudaEventSynchronize(tev[1]);
                        float ms = 0; cudaEventElapsedTime(&ms, tev[0], tev[1]);
                        fprintf(stderr, "[cb trace] step %d tail: %llu photons %.3f ms\n", step, (unsigned long long)n_alive, ms);
                    }
                    launches++;
#ifdef CB_TAIL_PROFILE
                    if (!wave && P.queue_in) {
                        // debug build only: when did the tail's photons finish, and after how many steps?
                        stream_wait(c.stream);
                        std::vector<uint32_t> q(n_alive), st(cap), tm(cap);
                        cudaMemcpy(q.data(), P.queue_in, n_alive * 4, cudaMemcpyDeviceToHost);
                        cudaMemcpy(st.data(), c.d_hit_tri, cap * 4, cudaMemcpyDeviceToHost);
                        cudaMemcpy(tm.data(), c.d_hit_dist, cap * 4, cudaMemcpyDeviceToHost);
                        std::vector<std::pair<uint32_t, uint32_t>> v;       // (finish time, steps)
                        uint32_t t0 = ~0u;
                        for (uint64_t i = 0; i < n_alive; i++) { uint32_t k = q[i] & 0x7fffffffu; t0 = std::min(t0, tm[k]); }
                        for (uint64_t i = 0; i < n_alive; i++) { uint32_t k = q[i] & 0x7fffffffu; v.push_back({tm[k] - t0, st[k]}); }
                        std::sort(v.begin(), v.end());
                        fprintf(stderr, "[tail profile] %llu photons; finish-time percentiles (us):", (unsigned long long)n_alive);
                        for (double f : {0.5, 0.9, 0.99, 0.999, 0.9999, 1.0})
                            fprintf(stderr, " p%g=%.0f", f * 100, v[std::min<size_t>(v.size() - 1, (size_t)(f * v.size()))].first * 0.064);
                        fprintf(stderr, "\n[tail profile] last 12 finishers (us, steps):");
                        for (size_t i = v.size() >= 12 ? v.size() - 12 : 0; i < v.size(); i++) fprintf(stderr, " (%.0f,%u)", v[i].first * 0.064, v[i].second);
                        unsigned long long hist[8] = {0}; unsigned long long stepsum = 0;
                        for (auto& x : v) { stepsum += x.second; int b = x.second <= 1 ? 0 : x.second <= 2 ? 1 : x.second <= 4 ? 2 : x.second <= 8 ? 3 : x.second <= 16 ? 4 : x.second <= 32 ? 5 : x.second <= 64 ? 6 : 7; hist[b]++; }
                        fprintf(stderr, "\n[tail profile] steps histogram (<=1,2,4,8,16,32,64,more):");
                        for (int b = 0; b < 8; b++) fprintf(stderr, " %llu", hist[b]);
                        fprintf(stderr, "  total steps %llu\n", stepsum);
                        // photons alive (still being worked on or not yet started) at a few times
                        fprintf(stderr, "[tail profile] unfinished at (us):");
                        for (double us : {250.0, 500.0, 750.0, 1000.0, 1250.0, 1500.0, 1750.0, 2000.0, 2250.0}) {
                            size_t lo = std::lower_bound(v.begin(), v.end(), std::make_pair((uint32_t)(us / 0.064), 0u)) - v.begin();
                            fprintf(stderr, " %g:%zu", us, v.size() - lo);
                        }
                        fprintf(stderr, "\n");
                    }
#endif
                    if (!wave) { n_alive = 0; step++; break; }       // the tail ran for certain: done
                }
                exact = false;
            }
            CB_CUDA(cudaEventRecord(c.kev1, c.stream));      // kernel time ends behind the last launch (re-recorded per batch)
            if (n_alive == 0) break;
            // one read-back per batch: photons queued for the next step (0 once a tail launch has run), and the
            // chunk's counters, which are final if that count is 0 -- one host round trip instead of two
            CB_CUDA(cudaMemcpyAsync(c.h_counters + 16, d_alive + step, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
            CB_CUDA(cudaMemcpyAsync(c.h_counters, c.d_counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
            CB_CUDA(stream_wait(c.stream));
            counters_read = true;
            if (trace) {
                float a = 0, b2 = 0;
                cudaEventElapsedTime(&a, tev[0], tev[1]); cudaEventElapsedTime(&b2, tev[1], tev[2]);
                fprintf(stderr, "[cb trace] step %d: %llu rays intersect %.3f ms physics %.3f ms -> %llu alive\n",
                        step - 1, (unsigned long long)n_alive, a, b2, c.h_counters[16]);
            }
            n_alive = c.h_counters[16];
            exact = true;
            if (n_alive) counters_read = false;
        }
        if (!counters_read) {
            CB_CUDA(cudaMemcpyAsync(c.h_counters, c.d_counters, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
            CB_CUDA(stream_wait(c.stream));
        }
        for (int i = 1; i < 9; i++) tot[i] += c.h_counters[i];
        if (trace && count)
            fprintf(stderr, "[cb trace] ray iterations: max %llu, >100: %llu, >300: %llu, >1000: %llu, total %llu\n",
                    c.h_counters[9], c.h_counters[10], c.h_counters[11], c.h_counters[12], c.h_counters[13]);
    }
    mark("end");
    CB_CUDA(event_wait(c.kev1));
    if (timeline) {
        for (size_t i = 1; i < tl_ev.size(); i++) {
            float ms = 0; cudaEventElapsedTime(&ms, tl_ev[i - 1], tl_ev[i]);
            fprintf(stderr, "[cb timeline] %-8s +%.3f ms\n", tl_name[i], ms);
        }
        for (auto& e : tl_ev) cudaEventDestroy(e);
    }
    if (trace) for (auto& e : tev) cudaEventDestroy(e);
    if (stats) {
        stats->photons = bank->n; stats->steps = tot[4]; stats->nodes_visited = tot[1]; stats->tris_tested = tot[2];
        stats->rays_resolved = tot[5];
        stats->launches = launches;
        cudaEventElapsedTime(&stats->kernel_ms, c.kev0, c.kev1);
        if (int0_rays) { cudaEventElapsedTime(&stats->intersect0_ms, c.iev0, c.iev1); stats->intersect0_rays = int0_rays; }
        float by_class[3] = {0.f, 0.f, 0.f};
        for (size_t k = 0; k < spans.size(); k++) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, spans[k].first, spans[k].second) == cudaSuccess) by_class[cls[k]] += ms;
        }
        stats->intersect_ms = by_class[0]; stats->physics_ms = by_class[1]; stats->tail_ms = by_class[2];
        stats->intersect_rays = tot[6]; stats->tail_photons = tot[7]; stats->tail_steps = tot[8];
        stats->physics_steps = tot[4] - tot[8];
    }
    if (tot[3]) return fail(CB_ERR_CUDA, "cb_propagate: traversal stack overflow");
    return CB_OK;
}

int cb_photon_duplicate(const CbPhotonBank* bank, uint64_t nphotons, int32_t ncopies)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    int rc = check_bank(bank, "cb_photon_duplicate");
    if (rc) return rc;
    if (ncopies <= 1 || nphotons == 0) return CB_OK;
    if (nphotons * (uint64_t)ncopies > bank->n) return fail(CB_ERR_INVALID, "cb_photon_duplicate: bank too small");
    duplicate_kernel<<<(unsigned)((nphotons + 255) / 256), 256, 0, ctx().stream>>>(*bank, nphotons, ncopies - 1);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(stream_wait(ctx().stream));
    return CB_OK;
}

int cb_count_photons(const CbPhotonBank* bank, uint64_t first, uint64_t n, uint32_t flag, uint32_t* count_out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    int rc = check_bank(bank, "cb_count_photons");
    if (rc) return rc;
    return compact<false>(bank, first, n, flag, nullptr, nullptr, nullptr, count_out);
}
int cb_copy_photons(const CbPhotonBank* src, uint64_t first, uint64_t n, uint32_t flag, const CbPhotonBank* dst,
                    uint32_t* count_out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    int rc = check_bank(src, "cb_copy_photons");
    if (rc) return rc;
    if ((rc = check_bank(dst, "cb_copy_photons(dst)"))) return rc;
    return compact<false>(src, first, n, flag, nullptr, dst, nullptr, count_out);
}
int cb_count_photon_hits(const CbPhotonBank* bank, uint64_t first, uint64_t n, uint32_t flag, cb_geom_t gh,
                         uint32_t* count_out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    int rc = check_bank(bank, "cb_count_photon_hits");
    if (rc) return rc;
    Geometry* g = geoms().get(gh);
    if (!g || !g->solid_to_channel) return fail(CB_ERR_INVALID, "cb_count_photon_hits: geometry has no detector attached");
    return compact<true>(bank, first, n, flag, g, nullptr, nullptr, count_out);
}
int cb_copy_photon_hits(const CbPhotonBank* src, uint64_t first, uint64_t n, uint32_t flag, cb_geom_t gh,
                        const CbPhotonBank* dst, int32_t* d_channels_out, uint32_t* count_out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    int rc = check_bank(src, "cb_copy_photon_hits");
    if (rc) return rc;
    if ((rc = check_bank(dst, "cb_copy_photon_hits(dst)"))) return rc;
    Geometry* g = geoms().get(gh);
    if (!g || !g->solid_to_channel) return fail(CB_ERR_INVALID, "cb_copy_photon_hits: geometry has no detector attached");
    if (!d_channels_out) return fail(CB_ERR_INVALID, "cb_copy_photon_hits: null channel array");
    return compact<true>(src, first, n, flag, g, dst, d_channels_out, count_out);
}
int cb_copy_photon_hits_async(const CbPhotonBank* src, uint64_t first, uint64_t n, uint32_t flag, cb_geom_t gh,
                              uint32_t* d_block, uint32_t* d_count_out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    int rc = check_bank(src, "cb_copy_photon_hits_async");
    if (rc) return rc;
    Geometry* g = geoms().get(gh);
    if (!g || !g->solid_to_channel) return fail(CB_ERR_INVALID, "cb_copy_photon_hits_async: geometry has no detector attached");
    if (!d_block || !d_count_out) return fail(CB_ERR_INVALID, "cb_copy_photon_hits_async: null output");
    if (first + n > src->n) return fail(CB_ERR_INVALID, "cb_copy_photon_hits_async: range exceeds the photon bank");
    Context& c = ctx();
    if (n == 0) {
        CB_CUDA(cudaMemsetAsync(d_count_out, 0, 4, c.stream));
        return CB_OK;
    }
    const uint64_t ntiles = (n + TILE - 1) / TILE;
    if ((rc = ensure_tile_scratch(ntiles))) return rc;
    CbPhotonBank dst;
    memset(&dst, 0, sizeof(dst));
    dst.pos = reinterpret_cast<float*>(d_block);           // the kernel derives the ten array bases from the total
    tile_count_kernel<true><<<(unsigned)ntiles, 256, 0, c.stream>>>(*src, first, n, flag, g->solid_id, g->solid_to_channel,
                                                                    c.d_block_counts);
    tile_scan_kernel<<<1, 1024, 0, c.stream>>>(c.d_block_counts, (uint32_t)ntiles);
    tile_scatter_kernel<true, true><<<(unsigned)ntiles, 256, 0, c.stream>>>(*src, first, n, flag, g->solid_id, g->solid_to_channel,
                                                                            c.d_block_counts, dst, nullptr);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(cudaMemcpyAsync(d_count_out, c.d_block_counts + ntiles, 4, cudaMemcpyDeviceToDevice, c.stream));
    return CB_OK;
}
int cb_copy_photon_queue(const CbPhotonBank* src, const uint32_t* d_queue, uint64_t n, const CbPhotonBank* dst)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    int rc = check_bank(src, "cb_copy_photon_queue");
    if (rc) return rc;
    if ((rc = check_bank(dst, "cb_copy_photon_queue(dst)"))) return rc;
    if (n == 0) return CB_OK;
    if (dst->n < n) return fail(CB_ERR_INVALID, "cb_copy_photon_queue: destination too small");
    gather_queue_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx().stream>>>(*src, d_queue, n, *dst);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(stream_wait(ctx().stream));
    return CB_OK;
}

// ---------------------------------------------------------------- DAQ
int cb_daq_create(cb_geom_t gh, int32_t ndaq, cb_daq_t* out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    Geometry* g = geoms().get(gh);
    if (!g || !out) return fail(CB_ERR_INVALID, "cb_daq_create: bad argument");
    if (g->nchannels <= 0 || !g->solid_to_channel)
        return fail(CB_ERR_INVALID, "Geometry has no detectors, DAQ can't be initialized.");
    if (ndaq < 1) return fail(CB_ERR_INVALID, "cb_daq_create: ndaq must be >= 1");
    Daq* d = new Daq();
    d->geom = g; d->ndaq = ndaq; d->count = (uint64_t)g->nchannels * ndaq;
    cudaError_t e;
    if ((e = cudaMalloc(&d->earliest_time, d->count * 4)) || (e = cudaMalloc(&d->earliest_time_int, d->count * 4)) ||
        (e = cudaMalloc(&d->channel_history, d->count * 4)) || (e = cudaMalloc(&d->channel_q_int, d->count * 4)) ||
        (e = cudaMalloc(&d->channel_q, d->count * 4))) {
        cudaFree(d->earliest_time); cudaFree(d->earliest_time_int); cudaFree(d->channel_history);
        cudaFree(d->channel_q_int); cudaFree(d->channel_q);
        delete d;
        return cuda_fail(e, "cudaMalloc(daq)");
    }
    if ((e = cudaMemset(d->channel_history, 0, d->count * 4)) || (e = cudaMemset(d->channel_q_int, 0, d->count * 4)) ||
        (e = cudaMemset(d->channel_q, 0, d->count * 4))) {
        cudaFree(d->earliest_time); cudaFree(d->earliest_time_int); cudaFree(d->channel_history);
        cudaFree(d->channel_q_int); cudaFree(d->channel_q);
        delete d;
        return cuda_fail(e, "cudaMemset(daq)");
    }
    *out = daqs().add(d);
    return CB_OK;
}
int cb_daq_destroy(cb_daq_t h)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    Daq* d = daqs().take(h);
    if (!d) return fail(CB_ERR_INVALID, "cb_daq_destroy: bad handle");
    stream_wait(ctx().stream);
    cudaFree(d->earliest_time); cudaFree(d->earliest_time_int); cudaFree(d->channel_history);
    cudaFree(d->channel_q_int); cudaFree(d->channel_q);
    delete d;
    return CB_OK;
}
// launch-only pieces of an acquisition (library stream, no host wait); the entry points below add the waits
static int daq_begin_launch(Daq* d)
{
    Context& c = ctx();
    const float maxtime = 1e9f;   // gpu/daq.py:56
    uint32_t bits;
    memcpy(&bits, &maxtime, 4);
    CB_CUDA(cudaMemsetAsync(d->channel_q_int, 0, d->count * 4, c.stream));
    CB_CUDA(cudaMemsetAsync(d->channel_q, 0, d->count * 4, c.stream));
    CB_CUDA(cudaMemsetAsync(d->channel_history, 0, d->count * 4, c.stream));
    fill32_launch(d->earliest_time_int, bits, d->count, c.stream);
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}
static int daq_acquire_launch(Daq* d, const CbPhotonBank* bank, RngPool* r, int32_t nthreads_per_block, int32_t max_blocks,
                              uint64_t start_photon, uint64_t nphotons, float weight)
{
    int rc = check_bank(bank, "cb_daq_acquire");
    if (rc) return rc;
    if (start_photon + nphotons > bank->n) return fail(CB_ERR_INVALID, "cb_daq_acquire: photon range exceeds bank");
    if (nthreads_per_block <= 0 || max_blocks <= 0) return fail(CB_ERR_INVALID, "cb_daq_acquire: bad launch parameters");
    Geometry* g = d->geom;
    Context& c = ctx();
    if (d->ndaq == 1) {
        // chunks reuse rng states [0, chunk) in order, as chunk_iterator does (gpu/daq.py:68-78)
        uint64_t chunk = std::min<uint64_t>((uint64_t)nthreads_per_block * (uint64_t)max_blocks, r->n);
        if (chunk == 0) return fail(CB_ERR_INVALID, "cb_daq_acquire: empty rng pool");
        for (uint64_t first = 0; first < nphotons; first += chunk) {
            uint64_t cnt = std::min<uint64_t>(chunk, nphotons - first);
            daq_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, c.stream>>>(
                r->states, CB_SURFACE_DETECT, start_photon + first, cnt, *bank, g->solid_id, g->solid_to_channel,
                g->time_cdf_x, g->time_cdf_y, g->time_cdf_len, g->charge_cdf_x, g->charge_cdf_y, g->charge_cdf_len,
                g->charge_unit, d->earliest_time_int, d->channel_q_int, d->channel_history, weight);
            CB_CUDA(cudaGetLastError());
        }
    } else {
        // one CTA per photon (chunk_iterator(nphotons, 1, max_blocks), gpu/daq.py:80-91)
        if (!r->bm_flag) {
            CB_CUDA(cudaMalloc(&r->bm_flag, std::max<uint64_t>(r->n, 1) * 4));
            CB_CUDA(cudaMalloc(&r->bm_extra, std::max<uint64_t>(r->n, 1) * 4));
            CB_CUDA(cudaMemsetAsync(r->bm_flag, 0, r->n * 4, c.stream));
            CB_CUDA(cudaMemsetAsync(r->bm_extra, 0, r->n * 4, c.stream));
        }
        uint64_t max_ctas = std::min<uint64_t>((uint64_t)max_blocks, r->n / (uint64_t)nthreads_per_block);
        if (max_ctas == 0) return fail(CB_ERR_INVALID, "cb_daq_acquire: rng pool smaller than one block");
        for (uint64_t first = 0; first < nphotons; first += max_ctas) {
            uint64_t cnt = std::min<uint64_t>(max_ctas, nphotons - first);
            daq_many_kernel<<<(unsigned)cnt, nthreads_per_block, 0, c.stream>>>(
                r->states, r->bm_extra, r->bm_flag, CB_SURFACE_DETECT, start_photon + first, *bank, g->solid_id,
                g->solid_to_channel, g->time_cdf_x, g->time_cdf_y, g->time_cdf_len, g->charge_cdf_x, g->charge_cdf_y,
                g->charge_cdf_len, g->charge_unit, d->earliest_time_int, d->channel_q_int, d->channel_history,
                d->ndaq, g->nchannels, weight);
            CB_CUDA(cudaGetLastError());
        }
    }
    return CB_OK;
}
static int daq_finalize_launch(Daq* d)
{
    Context& c = ctx();
    daq_finalize_kernel<<<(unsigned)((d->count + 255) / 256), 256, 0, c.stream>>>(
        d->count, d->earliest_time_int, d->channel_q_int, d->geom->charge_unit, d->earliest_time, d->channel_q);
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}
int cb_daq_begin_acquire(cb_daq_t h)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    Daq* d = daqs().get(h);
    if (!d) return fail(CB_ERR_INVALID, "cb_daq_begin_acquire: bad handle");
    int rc = daq_begin_launch(d);
    if (rc) return rc;
    CB_CUDA(stream_wait(ctx().stream));
    return CB_OK;
}
int cb_daq_acquire(cb_daq_t h, const CbPhotonBank* bank, cb_rng_t rh, int32_t nthreads_per_block,
                   int32_t max_blocks, uint64_t start_photon, uint64_t nphotons, float weight)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    Daq* d = daqs().get(h);
    RngPool* r = rngs().get(rh);
    if (!d) return fail(CB_ERR_INVALID, "cb_daq_acquire: bad daq handle");
    if (!r) return fail(CB_ERR_INVALID, "cb_daq_acquire: bad rng handle");
    int rc = daq_acquire_launch(d, bank, r, nthreads_per_block, max_blocks, start_photon, nphotons, weight);
    if (rc) return rc;
    CB_CUDA(stream_wait(ctx().stream));
    return CB_OK;
}
int cb_daq_acquire_async(cb_daq_t h, const CbPhotonBank* bank, cb_rng_t rh, int32_t nthreads_per_block,
                         int32_t max_blocks, uint64_t start_photon, uint64_t nphotons, float weight,
                         int32_t begin, int32_t finalize)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    Daq* d = daqs().get(h);
    RngPool* r = rngs().get(rh);
    if (!d) return fail(CB_ERR_INVALID, "cb_daq_acquire_async: bad daq handle");
    if (!r) return fail(CB_ERR_INVALID, "cb_daq_acquire_async: bad rng handle");
    int rc = CB_OK;
    if (begin && (rc = daq_begin_launch(d))) return rc;
    if ((rc = daq_acquire_launch(d, bank, r, nthreads_per_block, max_blocks, start_photon, nphotons, weight))) return rc;
    if (finalize && (rc = daq_finalize_launch(d))) return rc;
    return CB_OK;
}
int cb_daq_finalize(cb_daq_t h)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    Daq* d = daqs().get(h);
    if (!d) return fail(CB_ERR_INVALID, "cb_daq_finalize: bad handle");
    int rc = daq_finalize_launch(d);
    if (rc) return rc;
    CB_CUDA(stream_wait(ctx().stream));
    return CB_OK;
}
int cb_daq_end_acquire(cb_daq_t h) { return cb_daq_finalize(h); }
int cb_daq_fold(cb_daq_t dst, cb_daq_t src)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    Daq* a = daqs().get(dst);
    Daq* b = daqs().get(src);
    if (!a || !b || a->count != b->count) return fail(CB_ERR_INVALID, "cb_daq_fold: bad or mismatched handles");
    Context& c = ctx();
    daq_fold_into_kernel<<<(unsigned)((a->count + 255) / 256), 256, 0, c.stream>>>(
        a->count, a->earliest_time_int, b->earliest_time_int, a->channel_q_int, b->channel_q_int, a->channel_history,
        b->channel_history);
    CB_CUDA(cudaGetLastError());
    CB_CUDA(stream_wait(c.stream));
    return CB_OK;
}
int cb_daq_fold_async(cb_daq_t dst, cb_daq_t src)
{
    // Enqueue only, and without the library lock: the launch touches nothing but the two accumulators, and a
    // caller that consumes event k while the pipeline propagates event k+1 must not wait for that call to return.
    CB_REQUIRE_INIT();
    Daq* a = daqs().get(dst);
    Daq* b = daqs().get(src);
    if (!a || !b || a->count != b->count) return fail(CB_ERR_INVALID, "cb_daq_fold_async: bad or mismatched handles");
    daq_fold_into_kernel<<<(unsigned)((a->count + 255) / 256), 256, 0, ctx().stream>>>(
        a->count, a->earliest_time_int, b->earliest_time_int, a->channel_q_int, b->channel_q_int, a->channel_history,
        b->channel_history);
    CB_CUDA(cudaGetLastError());
    return CB_OK;
}
int cb_daq_pointers(cb_daq_t h, void** t, void** q, void** flags, void** time_int, void** q_int, uint64_t* count)
{
    Daq* d = daqs().get(h);
    if (!d) return fail(CB_ERR_INVALID, "cb_daq_pointers: bad handle");
    if (t) *t = d->earliest_time;
    if (q) *q = d->channel_q;
    if (flags) *flags = d->channel_history;
    if (time_int) *time_int = d->earliest_time_int;
    if (q_int) *q_int = d->channel_q_int;
    if (count) *count = d->count;
    return CB_OK;
}

} // extern "C"
