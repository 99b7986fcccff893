// engine.cuh -- device-side building blocks of the B200 photon transport engine:
// vector helpers, XORWOW, the BVH traversal and the optical physics.
//
// Reference behaviour followed (NOT its code structure): chroma/cuda/mesh.h,
// intersect.h, geometry.h, photon.h, random.h, interpolate.h, rotate.h.  The
// arithmetic expression shapes deliberately match the reference's so that, built
// with the same nvcc flags (--use_fast_math, default -fmad), decisions replay.
#pragma once
#include <cuda_runtime.h>
#include <cuComplex.h>
#include <stdint.h>
#include <float.h>
#include "../../include/chroma_b200.h"

namespace cb {

// ------------------------------------------------------------------ vectors
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator-(const float3& a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator+(const float3& a, const float3& b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(const float3& a, const float3& b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(const float3& a, float c) { return f3(a.x * c, a.y * c, a.z * c); }
__device__ __forceinline__ float3 operator*(float c, const float3& a) { return f3(c * a.x, c * a.y, c * a.z); }
__device__ __forceinline__ float3 operator/(const float3& a, float c) { return f3(a.x / c, a.y / c, a.z / c); }
__device__ __forceinline__ float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross(const float3& a, const float3& b)
{
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float norm(const float3& a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ float3 normalize(const float3& a) { return a / norm(a); }

#define CB_SPEED_OF_LIGHT 299.792458f
#define CB_PI 3.141592653589793f
#define CB_WEIGHT_LOWER_THRESHOLD 0.0001f
#define CB_TERMINAL (CB_NO_HIT | CB_BULK_ABSORB | CB_SURFACE_DETECT | CB_SURFACE_ABSORB | CB_NAN_ABORT)

// rotate `a` through phi about axis n (behaviour of chroma/cuda/rotate.h:22-28)
__device__ __forceinline__ float3 rotate(const float3& a, float phi, const float3& n)
{
    float cos_phi = cosf(phi);
    float sin_phi = sinf(phi);
    return a * cos_phi + n * dot(a, n) * (1.0f - cos_phi) + cross(a, n) * sin_phi;
}

// ------------------------------------------------------------------ XORWOW
// Compact 24-byte state {d, v[5]}: bit-compatible stream with cuRAND's
// curandStateXORWOW (curand_kernel.h:863-874); Box-Muller cache kept separately
// by the one caller that needs normals (run_daq_many).
struct Rng {
    uint32_t d, v0, v1, v2, v3, v4;
};

__device__ __forceinline__ uint32_t rng_next(Rng& s)
{
    uint32_t t = s.v0 ^ (s.v0 >> 2);
    s.v0 = s.v1; s.v1 = s.v2; s.v2 = s.v3; s.v3 = s.v4;
    s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
    s.d += 362437u;
    return s.v4 + s.d;
}
// curand_uniform: (0, 1]  (curand_uniform.h:69-72)
__device__ __forceinline__ float rng_uniform(Rng& s)
{
    return rng_next(s) * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}
__device__ __forceinline__ float rng_range(Rng& s, float low, float high)
{
    return low + rng_uniform(s) * (high - low);
}
__device__ __forceinline__ Rng rng_load(const uint32_t* __restrict__ states, uint64_t i)
{
    const uint2* p = reinterpret_cast<const uint2*>(states + 6 * i);
    uint2 a = p[0], b = p[1], c = p[2];
    Rng r = {a.x, a.y, b.x, b.y, c.x, c.y};
    return r;
}
__device__ __forceinline__ void rng_store(uint32_t* __restrict__ states, uint64_t i, const Rng& r)
{
    uint2* p = reinterpret_cast<uint2*>(states + 6 * i);
    p[0] = make_uint2(r.d, r.v0); p[1] = make_uint2(r.v1, r.v2); p[2] = make_uint2(r.v3, r.v4);
}
// isotropic direction: theta first, then u (random.h:15-23)
__device__ __forceinline__ float3 rng_sphere(Rng& s)
{
    float theta = rng_range(s, 0.0f, 2 * CB_PI);
    float u = rng_range(s, -1.0f, 1.0f);
    float c = sqrtf(1.0f - u * u);
    return f3(c * cosf(theta), c * sinf(theta), u);
}

// ------------------------------------------------------------------ geometry view
struct DevGeometry {
    const uint4*  nodes;      // the engine's traversal tree (reference packing, children contiguous)
    const uint4*  ref_nodes;  // the reference tree itself: visit-order fallback only
    uint32_t ref_root_w;
    const float4* tri64;      // 4 x float4 per triangle: v0.xyz v1.x | v1.yz v2.xy | v2.z rank code - | reference leaf box xyz -
    const float*  tables;     // global table pool
    const CbMaterial* materials;
    const CbSurface*  surfaces;
    float3 world_origin;
    float  world_scale;
    int32_t wavelength_n; float wavelength_start, wavelength_step;
    int32_t time_n;       float time_start, time_step;
    uint32_t root_w;          // root node's child word
    uint32_t root_x, root_y, root_z;  // root node's packed box
    uint32_t ref_root_x, ref_root_y, ref_root_z;   // the reference tree's root box
    uint32_t smem_floats;     // leading floats of the pool staged into shared memory
    uint32_t nmaterials, nsurfaces;
    const CbWirePlane* wireplanes;   // analytic wire planes (cold path; usually none)
    int32_t nwireplanes;
};

// shared-memory view handed to the physics (tables staged by the kernel prologue)
struct Tables {
    const float* smem;        // may be nullptr when nothing is staged
    const float* gmem;
    uint32_t smem_floats;
    __device__ __forceinline__ const float* at(int32_t off) const
    {
        return ((uint32_t)off < smem_floats) ? (smem + off) : (gmem + off);
    }
};

struct TraverseCounters { uint32_t nodes, tris, resolved; };

// ------------------------------------------------------------------ triangle test
// Moller-Trumbore with the reference's tolerances and mixed precision
// (behaviour of chroma/cuda/intersect.h:26-101): reject
// |a| < FLT_EPSILON, reciprocal in double, u/v bounds +-1e-6 compared in
// double, accept 1e-6 < t < inf.
// The arithmetic is pinned instruction by instruction to what nvcc emits for the
// reference (SASS of intersect_triangle in oracle/_ref: cross = fma(a.y,b.z,-(a.z*b.y)),
// dot = fma(a.z,b.z,fma(a.x,b.x,a.y*b.y)), and u+v contracted to fma(f,dot(s,h),v)),
// because rays aimed at shared edges/vertices sit right on the +-1e-6 tolerances.
__device__ __forceinline__ float3 cross_pinned(const float3& a, const float3& b)
{
    return f3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
              __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float dot_pinned(const float3& a, const float3& b)
{
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
}
__device__ __forceinline__ bool hit_triangle(const float3& origin, const float3& direction,
                                             const float3& v0, const float3& v1, const float3& v2,
                                             float& distance)
{
    const float3 edge1 = f3(__fadd_rn(v1.x, -v0.x), __fadd_rn(v1.y, -v0.y), __fadd_rn(v1.z, -v0.z));
    const float3 edge2 = f3(__fadd_rn(v2.x, -v0.x), __fadd_rn(v2.y, -v0.y), __fadd_rn(v2.z, -v0.z));
    const float3 h = cross_pinned(direction, edge2);
    const float a = dot_pinned(edge1, h);
    if (a > -FLT_EPSILON && a < FLT_EPSILON) return false;
    const float f = 1.0 / a;
    const float3 s = f3(__fadd_rn(origin.x, -v0.x), __fadd_rn(origin.y, -v0.y), __fadd_rn(origin.z, -v0.z));
    const float sh = dot_pinned(s, h);
    const float u = __fmul_rn(f, sh);
    if (u < -1e-6 || u > 1.0 + 1e-6) return false;
    const float3 q = cross_pinned(s, edge1);
    const float v = __fmul_rn(f, dot_pinned(direction, q));
    const float upv = __fmaf_rn(f, sh, v);
    if (v < -1e-6 || upv > 1.0 + 1e-6) return false;
    const float t = __fmul_rn(f, dot_pinned(edge2, q));
    if (t > 1e-6 && t < __int_as_float(0x7f800000)) {
        distance = t;
        return true;
    }
    return false;
}

// exact uint16 -> float without the (quarter-rate) I2F unit
__device__ __forceinline__ float u16f_lo(uint32_t w) { return __uint_as_float((w & 0xFFFFu) | 0x4B000000u) - 8388608.0f; }
__device__ __forceinline__ float u16f_hi(uint32_t w) { return __uint_as_float((w >> 16) | 0x4B000000u) - 8388608.0f; }

struct RaySetup {
    float3 inv, noid;          // 1/d and -o/d  (mesh.h:57-58)
    bool fx, fy, fz;           // isfinite(inv.*)  (intersect.h:120-146)
};

// slab test on a packed node, same dequantisation and arithmetic as
// geometry.h:31-47 + intersect.h:112-157.  Returns tmin through `tnear`.
__device__ __forceinline__ bool hit_box(const DevGeometry& g, const RaySetup& r, uint32_t px,
                                        uint32_t py, uint32_t pz, float& tnear)
{
    const float INF = __int_as_float(0x7f800000);
    float tmin = 0.0f, tmax = INF;
    if (r.fx) {
        float lo = __fmaf_rn(u16f_lo(px), g.world_scale, g.world_origin.x);
        float hi = __fmaf_rn(u16f_hi(px), g.world_scale, g.world_origin.x);
        float t0 = __fmaf_rn(lo, r.inv.x, r.noid.x);
        float t1 = __fmaf_rn(hi, r.inv.x, r.noid.x);
        tmin = fmaxf(tmin, fminf(t0, t1));
        tmax = fminf(tmax, fmaxf(t0, t1));
    }
    if (r.fy) {
        float lo = __fmaf_rn(u16f_lo(py), g.world_scale, g.world_origin.y);
        float hi = __fmaf_rn(u16f_hi(py), g.world_scale, g.world_origin.y);
        float t0 = __fmaf_rn(lo, r.inv.y, r.noid.y);
        float t1 = __fmaf_rn(hi, r.inv.y, r.noid.y);
        tmin = fmaxf(tmin, fminf(t0, t1));
        tmax = fminf(tmax, fmaxf(t0, t1));
    }
    if (r.fz) {
        float lo = __fmaf_rn(u16f_lo(pz), g.world_scale, g.world_origin.z);
        float hi = __fmaf_rn(u16f_hi(pz), g.world_scale, g.world_origin.z);
        float t0 = __fmaf_rn(lo, r.inv.z, r.noid.z);
        float t1 = __fmaf_rn(hi, r.inv.z, r.noid.z);
        tmin = fmaxf(tmin, fminf(t0, t1));
        tmax = fminf(tmax, fmaxf(t0, t1));
    }
    tnear = tmin;
    return !(tmin > tmax);
}

// slab test in the reference's arithmetic, recomputing the ray setup (rare path)
static __device__ __noinline__ bool hit_box_exact(const DevGeometry& g, const float3 o, const float3 d,
                                                  uint32_t px, uint32_t py, uint32_t pz, float& tnear)
{
    RaySetup rr;
    rr.noid = f3(-o.x / d.x, -o.y / d.y, -o.z / d.z);
    rr.inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    rr.fx = isfinite(rr.inv.x); rr.fy = isfinite(rr.inv.y); rr.fz = isfinite(rr.inv.z);
    return hit_box(g, rr, px, py, pz, tnear);
}

// ------------------------------------------------------------------ traversal
constexpr int CB_RSTACK = 64;   // local stack of the reference-order fallback

// Exact emulation of the reference's visit order (behaviour of mesh.h:45-126:
// children ascending, leaves tested on the spot, internal children pushed and
// popped LIFO, prune against the current minimum).  Only used for the rare rays
// the traversal flags as order-sensitive (see PTrav::finish).
template <bool COUNT>
static __device__ __noinline__ int traverse_reference_order(const DevGeometry& g, const float3 origin,
                                                            const float3 direction, int last_hit,
                                                            float& min_distance, uint32_t* overflow_flag,
                                                            TraverseCounters* cnt)
{
    RaySetup r;
    r.noid = f3(-origin.x / direction.x, -origin.y / direction.y, -origin.z / direction.z);
    r.inv = f3(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
    r.fx = isfinite(r.inv.x); r.fy = isfinite(r.inv.y); r.fz = isfinite(r.inv.z);
    int triangle_index = -1;
    min_distance = -1.0f;
    uint32_t stack[CB_RSTACK];
    {
        const uint4 root = __ldg(&g.ref_nodes[0]);
        float tn;
        if (!hit_box(g, r, root.x, root.y, root.z, tn)) return -1;
    }
    int sp = 0;
    stack[sp++] = g.ref_root_w;
    while (sp > 0) {
        const uint32_t cur = stack[--sp];
        const uint32_t first = cur & 0x0FFFFFFFu, n = cur >> 28;
        for (uint32_t i = first; i < first + n; i++) {
            const uint4 nd = __ldg(&g.ref_nodes[i]);
            if (COUNT) cnt->nodes++;
            float tmin;
            if (hit_box(g, r, nd.x, nd.y, nd.z, tmin) && (min_distance < 0.0f || !(tmin > min_distance))) {
                const uint32_t w = nd.w;
                if ((w >> 28) == 0) {
                    if ((int)w != last_hit) {
                        if (COUNT) cnt->tris++;
                        const float4* tp = g.tri64 + 4ull * w;
                        float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                        float t;
                        if (hit_triangle(origin, direction, f3(a.x, a.y, a.z), f3(a.w, b.x, b.y), f3(b.z, b.w, c.x), t)) {
                            if (triangle_index == -1 || t < min_distance) {
                                triangle_index = (int)w;
                                min_distance = t;
                            }
                        }
                    }
                } else if (sp < CB_RSTACK) {
                    stack[sp++] = w;
                } else {
                    atomicOr(overflow_flag, 1u);
                }
            }
        }
    }
    return triangle_index;
}

// ------------------------------------------------------------------ phased traversal
// One ray per lane over the engine's tree (bvh_native.cu: same 16-byte entry format
// as the reference, <= 8 children per node stored contiguously).
//
// Exactness (what makes ANY visit order return the reference's triangle, SURVEY A-1):
//   * the result is the lexicographic minimum of (distance, reference test rank) over
//     all triangle hits, the triangle test being pinned to the reference's arithmetic;
//   * boxes are culled against limit = best_t * (1 + 2e-5) with slabs widened by the
//     rounding bound of the plane arithmetic (phased_ray_axis), so no box that the
//     reference's arithmetic would enter with a candidate inside is ever skipped;
//   * the reference returns the same minimum unless it never tests the winner, which
//     needs the winner's hit to lie in FRONT of its own reference leaf box in float
//     arithmetic; finish() re-evaluates exactly that in the reference's arithmetic
//     and such rays (~5e-7) are redone in the reference's own visit order on the
//     reference tree (traverse_reference_order).
//
// Scheduling.  A first version (while-while, leaves on the stack, one branch per
// child hit) ran its box tests with ~9 of 32 lanes and the push logic with ~3.  Here
//   * leaves never go on the stack: an expansion drops its leaf hits into a small
//     per-lane queue; a lane with queued leaves tests one triangle per iteration, a
//     lane without expands its node, and both share one load sequence (if-if
//     traversal with a unified 64-byte fetch), so lanes do not wait for each other's
//     phase;
//   * the child loop is straight-line: hit tests, nearest-child selection and the
//     pushes are predicated, no branch per child;
//   * stack and leaf queue live in shared memory, addressed by running 32-bit shared
//     addresses (st.shared/ld.shared), lane-interleaved, with a local-memory overflow
//     area behind the stack (touched by ~2 % of the rays of the 29k-PMT detector);
//   * a plane test is PRMT + FFMA: the byte-permute builds the float 2^23+q straight
//     from the packed uint16, the affine map folds 2^23 into its offset.
#ifndef CB_PSTACK_N
#define CB_PSTACK_N 16
#endif
constexpr int CB_PSTACK = CB_PSTACK_N;    // internal entries per lane
constexpr int CB_PLEAF = 8;      // leaf queue per lane: one expansion's worth
constexpr int CB_PCOLD = 3;      // slots per lane for rarely used ray state (origin, direction)
constexpr int CB_PLSTACK = 48;   // overflow entries per lane in local memory (rarely touched)
#ifndef CB_INT_THREADS
#define CB_INT_THREADS 128       /* threads per CTA of the traversal kernels */
#endif
constexpr uint32_t CB_PSTRIDE = CB_INT_THREADS * 8u;   // bytes between consecutive entries of a lane (lanes x uint2)

__device__ __forceinline__ void sts64(uint32_t addr, uint32_t x, uint32_t y)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Per-ray constants of the plane test t = (2^23 + q) * s + off, q = packed uint16 plane.
// `off` folds the 2^23 and is widened by a bound on the difference between this
// arithmetic and the reference's (intersect.h:112-157 on geometry.h:31-47 boxes), so
// a box the reference would hit is never missed:
//   ours:      inv 1 ulp (rcp.approx), s = scale*inv 1.5 ulp, a = (origin-o)*inv 2 ulp,
//              off = a - 2^23 s half an ulp of (|a|+|b|)         -> 3.0e-7|a| + 2.4e-7|b|
//   reference: lo = fma(q,scale,origin), t = fma(lo, 1/d, -o/d)   -> < 2.4e-7 (|lo|+|o|)/|d|
struct PhasedRay {
    float sx, sy, sz;
    float nx, ny, nz, fx, fy, fz;     // near / far offsets (-inf / +inf when the axis is ignored)
    uint32_t selx, sely, selz;        // PRMT selector of the near plane; far = near ^ 0x22
};
__device__ __forceinline__ void phased_ray_axis(float o, float d, float worigin, float wscale, float& s, float& n,
                                                float& f, uint32_t& sel)
{
    const float INF = __int_as_float(0x7f800000);
    const float inv = 1.0f / d;
    if (isfinite(inv)) {
        s = wscale * inv;
        const float a = (worigin - o) * inv, b = 8388608.0f * s;
        const float off = a - b;
        const float c = (fabsf(worigin) + fabsf(o) + 65536.0f * wscale) * fabsf(inv);
        const float e = 6e-7f * c + 3e-7f * fabsf(b);
        n = off - e; f = off + e;
        sel = (inv >= 0.0f) ? 0x7610u : 0x7632u;   // low half first when the ray runs towards +axis
    } else {                                       // the reference skips such an axis (intersect.h:120)
        s = 0.0f; n = -INF; f = INF; sel = 0x7610u;
    }
}
__device__ __forceinline__ bool hit_box_phased(const PhasedRay& r, uint32_t px, uint32_t py, uint32_t pz, float& tnear)
{
    const uint32_t K = 0x4B000000u;                // bytes {q_lo, q_hi, 0x00, 0x4B} = 2^23 + q
    const float tnx = __fmaf_rn(__uint_as_float(prmt(px, K, r.selx)), r.sx, r.nx);
    const float tfx = __fmaf_rn(__uint_as_float(prmt(px, K, r.selx ^ 0x22u)), r.sx, r.fx);
    const float tny = __fmaf_rn(__uint_as_float(prmt(py, K, r.sely)), r.sy, r.ny);
    const float tfy = __fmaf_rn(__uint_as_float(prmt(py, K, r.sely ^ 0x22u)), r.sy, r.fy);
    const float tnz = __fmaf_rn(__uint_as_float(prmt(pz, K, r.selz)), r.sz, r.nz);
    const float tfz = __fmaf_rn(__uint_as_float(prmt(pz, K, r.selz ^ 0x22u)), r.sz, r.fz);
    const float tmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
    const float tmax = fminf(fminf(tfx, tfy), tfz);
    tnear = tmin;
    return !(tmin > tmax);
}

struct PTrav {
    // (the ray's origin and direction are only needed by the triangle test and finish():
    //  they live in three extra shared-memory slots behind the leaf queue, not in registers)
    PhasedRay r;
    float best_t, limit, cur_t;
    uint32_t best_rank, cur;
    int best_tri, last_hit;
    uint32_t sp, lq;           // shared addresses one past the top of this lane's stack / leaf queue
    int lsp;                   // entries in the local-memory overflow area (on top of the shared ones)
    bool have, redo;

    __device__ __forceinline__ void pop_next(uint32_t sbase, const uint2* lstack)
    {
        have = false;
        while (lsp > 0 || sp > sbase) {
            uint2 e;
            if (lsp > 0) e = lstack[--lsp];
            else { sp -= CB_PSTRIDE; e = lds64(sp); }
            if (!(__uint_as_float(e.y) > limit)) { cur = e.x; cur_t = __uint_as_float(e.y); have = true; break; }
        }
    }

    // returns false when the ray misses the world box (result: no hit)
    __device__ __forceinline__ bool init(const DevGeometry& g, const float3& o, const float3& d, int last,
                                         uint32_t sbase, uint32_t lbase)
    {
        const float INF = __int_as_float(0x7f800000);
        last_hit = last;
        const uint32_t cold = lbase + CB_PLEAF * CB_PSTRIDE;
        sts64(cold, __float_as_uint(o.x), __float_as_uint(o.y));
        sts64(cold + CB_PSTRIDE, __float_as_uint(o.z), __float_as_uint(d.x));
        sts64(cold + 2 * CB_PSTRIDE, __float_as_uint(d.y), __float_as_uint(d.z));
        phased_ray_axis(o.x, d.x, g.world_origin.x, g.world_scale, r.sx, r.nx, r.fx, r.selx);
        phased_ray_axis(o.y, d.y, g.world_origin.y, g.world_scale, r.sy, r.ny, r.fy, r.sely);
        phased_ray_axis(o.z, d.z, g.world_origin.z, g.world_scale, r.sz, r.nz, r.fz, r.selz);
        best_tri = -1; best_rank = 0xFFFFFFFFu; best_t = INF; limit = INF;
        sp = sbase; lq = lbase; lsp = 0; cur = g.root_w; cur_t = 0.0f; redo = false;
        float tn;                                  // the world-box test stays in the reference's arithmetic (mesh.h:60)
        have = hit_box_exact(g, o, d, g.ref_root_x, g.ref_root_y, g.ref_root_z, tn) && (g.root_w >> 28) != 0;
        return have;
    }

    // ---- expansion of the current internal entry, in bursts of four children
    struct Nearest { uint32_t w; float t; };     // nearest internal hit so far (w == 0: none; an internal word is never 0)

    // test children i..i+3 of the current entry (already fetched into nd; slots >= n hold a copy of child n-1)
    template <bool COUNT>
    __device__ __forceinline__ void process4(const uint4 (&nd)[4], uint32_t i, uint32_t n, Nearest& nr,
                                             uint32_t sbase, uint2* lstack, TraverseCounters* cnt)
    {
        const uint32_t stop = sbase + CB_PSTACK * CB_PSTRIDE;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float tmin;
            const bool ok = hit_box_phased(r, nd[k].x, nd[k].y, nd[k].z, tmin) && !(tmin > limit) && (i + k < n);
            if (COUNT) cnt->nodes += (i + k < n);
            const uint32_t w = nd[k].w;
            const bool is_leaf = ok && w < 0x10000000u;
            const bool is_int = ok && w >= 0x10000000u;
            if (is_leaf) sts64(lq, w, __float_as_uint(tmin));
            lq += is_leaf ? CB_PSTRIDE : 0u;
            // keep the nearest internal hit in registers, the others go to the stack
            const bool better = is_int && tmin < nr.t;
            const uint32_t pw = better ? nr.w : w;
            const float pt = better ? nr.t : tmin;
            nr.w = better ? w : nr.w;
            nr.t = better ? tmin : nr.t;
            const bool want_push = is_int && pw != 0;
            const bool do_push = want_push && sp < stop;
            if (do_push) sts64(sp, pw, __float_as_uint(pt));
            sp += do_push ? CB_PSTRIDE : 0u;
            if (want_push && !do_push) {               // rare: deeper than the shared-memory stack
                if (lsp < CB_PLSTACK) lstack[lsp++] = make_uint2(pw, __float_as_uint(pt));
                else redo = true;                      // redone in reference order
            }
        }
    }
    __device__ __forceinline__ void expand_end(const Nearest& nr, uint32_t sbase, const uint2* lstack)
    {
        if (nr.w) { cur = nr.w; cur_t = nr.t; }
        else pop_next(sbase, lstack);
    }

    // ---- leaves
    // next queued leaf worth testing (entries behind `limit` and the excluded triangle are dropped)
    __device__ __forceinline__ bool pop_leaf(uint32_t lbase, uint32_t& tri)
    {
        while (lq > lbase) {
            lq -= CB_PSTRIDE;
            const uint2 e = lds64(lq);
            if (!(__uint_as_float(e.y) > limit) && (int)e.x != last_hit) { tri = e.x; return true; }
        }
        return false;
    }
    // a, b, c: the first three float4 of the triangle's tri64 record
    __device__ __forceinline__ void load_ray(uint32_t lbase, float3& origin, float3& direction) const
    {
        const uint32_t cold = lbase + CB_PLEAF * CB_PSTRIDE;
        const uint2 a = lds64(cold), b = lds64(cold + CB_PSTRIDE), c = lds64(cold + 2 * CB_PSTRIDE);
        origin = f3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(b.x));
        direction = f3(__uint_as_float(b.y), __uint_as_float(c.x), __uint_as_float(c.y));
    }
    __device__ __forceinline__ void test_triangle(uint32_t lbase, uint32_t tri, const float4& a, const float4& b, const float4& c)
    {
        float3 origin, direction;
        load_ray(lbase, origin, direction);
        float t;
        if (hit_triangle(origin, direction, f3(a.x, a.y, a.z), f3(a.w, b.x, b.y), f3(b.z, b.w, c.x), t)) {
            const uint32_t rank = __float_as_uint(c.y);
            if (t < best_t || (t == best_t && rank < best_rank)) {
                best_t = t; best_tri = (int)tri; best_rank = rank;
                limit = best_t + 2e-5f * best_t;     // slack for the triangle test's own rounding
            }
        }
    }
    // after the leaf queue has drained: the pending node may have fallen behind the new limit
    __device__ __forceinline__ void after_leaves(uint32_t sbase, uint32_t lbase, const uint2* lstack)
    {
        if (lq == lbase && have && cur_t > limit) pop_next(sbase, lstack);
    }

    // call once !have and the leaf queue is empty; returns the triangle (or -1) and its distance
    template <bool COUNT>
    __device__ __forceinline__ int finish(const DevGeometry& g, uint32_t lbase, float& dist, uint32_t* overflow_flag,
                                          TraverseCounters* cnt)
    {
        float3 origin, direction;
        load_ray(lbase, origin, direction);
        if (best_tri != -1 && !redo) {
            const float4 lb = __ldg(g.tri64 + 4ull * (uint32_t)best_tri + 3);
            float box_t;
            const bool in_box = hit_box_exact(g, origin, direction, __float_as_uint(lb.x), __float_as_uint(lb.y),
                                              __float_as_uint(lb.z), box_t);
            redo = !in_box || best_t < box_t;
        }
        if (redo) {
            TraverseCounters local = {0, 0, 0};
            float rd;
            const int tri = traverse_reference_order<COUNT>(g, origin, direction, last_hit, rd, overflow_flag, &local);
            if (COUNT) { cnt->nodes += local.nodes; cnt->tris += local.tris; cnt->resolved++; }
            dist = rd;
            return tri;
        }
        dist = (best_tri == -1) ? -1.0f : best_t;
        return best_tri;
    }
};

// ------------------------------------------------------------------ warp-cooperative traversal
// One ray per WARP: the 32 lanes pop up to four entries from a shared-memory stack
// and test their (<= 8) children in parallel, one child per lane; surviving leaves
// go to a shared leaf queue and are tested 32 triangles at a time, the nearest hit
// found by a shuffle reduction on (distance, rank).  A single ray's latency drops
// from ~150 dependent box tests to ~10 warp iterations, which is what bounds the
// late steps of a propagate call (few photons left, some with 100-step histories
// or rays that graze dozens of PMTs).  Same exactness rule as Trav.
constexpr int CB_WSTACK = 192;   // stack entries per warp
constexpr int CB_WLEAF = 64;     // leaf-queue entries per warp
#ifndef CB_WTRI_MIN
#define CB_WTRI_MIN 24            /* queued leaves that trigger a triangle batch */
#endif

template <bool COUNT>
__device__ __forceinline__ int warp_traverse(const DevGeometry& g, const float3& origin, const float3& direction,
                                             int last_hit, float& dist, uint2* wstack, uint2* wleaf,
                                             uint32_t* overflow_flag, TraverseCounters* cnt)
{
    const float INF = __int_as_float(0x7f800000);
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    PhasedRay r;
    phased_ray_axis(origin.x, direction.x, g.world_origin.x, g.world_scale, r.sx, r.nx, r.fx, r.selx);
    phased_ray_axis(origin.y, direction.y, g.world_origin.y, g.world_scale, r.sy, r.ny, r.fy, r.sely);
    phased_ray_axis(origin.z, direction.z, g.world_origin.z, g.world_scale, r.sz, r.nz, r.fz, r.selz);
    RaySetup rr;
    rr.noid = f3(-origin.x / direction.x, -origin.y / direction.y, -origin.z / direction.z);
    rr.inv = f3(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
    rr.fx = isfinite(rr.inv.x); rr.fy = isfinite(rr.inv.y); rr.fz = isfinite(rr.inv.z);
    float tn;
    if (!hit_box(g, rr, g.ref_root_x, g.ref_root_y, g.ref_root_z, tn) || (g.root_w >> 28) == 0) { dist = -1.0f; return -1; }
    float best_t = INF, limit = INF;
    int best_tri = -1;
    uint32_t best_rank = 0xFFFFFFFFu;
    bool redo = false;
    int sp = 0, nleaf = 0;                 // warp-uniform
    if (lane == 0) wstack[0] = make_uint2(g.root_w, 0u);
    sp = 1;
    __syncwarp();

    while (sp > 0 || nleaf > 0) {
        if (nleaf >= CB_WTRI_MIN || sp == 0) {
            // ---- triangle phase: up to 32 queued leaves at once
            const int m = min(nleaf, 32);
            nleaf -= m;
            float t = INF;
            uint32_t rank = 0xFFFFFFFFu;
            int tri = -1;
            if ((int)lane < m) {
                const uint2 e = wleaf[nleaf + lane];
                if (!(__uint_as_float(e.y) > limit) && (int)e.x != last_hit) {
                    if (COUNT) cnt->tris++;
                    const float4* tp = g.tri64 + 4ull * e.x;
                    const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                    float tt;
                    if (hit_triangle(origin, direction, f3(a.x, a.y, a.z), f3(a.w, b.x, b.y), f3(b.z, b.w, c.x), tt)) {
                        t = tt; rank = __float_as_uint(c.y); tri = (int)e.x;
                    }
                }
            }
            // nearest hit of the batch, ties by reference rank: two REDUX.MIN (a hit distance is
            // positive, so its bit pattern orders like the float) + one ballot instead of a
            // five-level shuffle tree on (t, rank, tri)
            const uint32_t tmin_bits = __reduce_min_sync(0xffffffffu, __float_as_uint(t));
            if (tmin_bits != 0x7f800000u) {
                const bool cand = __float_as_uint(t) == tmin_bits;
                const uint32_t rmin = __reduce_min_sync(0xffffffffu, cand ? rank : 0xFFFFFFFFu);
                const int src = __ffs(__ballot_sync(0xffffffffu, cand && rank == rmin)) - 1;
                const int wtri = __shfl_sync(0xffffffffu, tri, src);
                const float wt = __uint_as_float(tmin_bits);
                if (wt < best_t || (wt == best_t && rmin < best_rank)) {
                    best_t = wt; best_tri = wtri; best_rank = rmin;
                    limit = best_t + 2e-5f * best_t;
                }
            }
            __syncwarp();
            continue;
        }
        // ---- expand phase: pop up to four entries, one child per lane
        const int take = min(sp, 4);
        const int which = lane >> 3, c = lane & 7;
        uint2 e = make_uint2(0u, 0u);
        if (which < take) e = wstack[sp - 1 - which];
        sp -= take;
        __syncwarp();
        const uint32_t n = e.x >> 28, first = e.x & 0x0FFFFFFFu;
        bool ok = false;
        float tmin = 0.0f;
        uint32_t w = 0;
        if (which < take && (uint32_t)c < n && !(__uint_as_float(e.y) > limit)) {
            const uint4 nd = __ldg(&g.nodes[first + c]);
            if (COUNT) cnt->nodes++;
            ok = hit_box_phased(r, nd.x, nd.y, nd.z, tmin) && !(tmin > limit);
            w = nd.w;
        }
        const bool is_leaf = ok && (w >> 28) == 0;
        const bool is_int = ok && (w >> 28) != 0;
        const unsigned lm = __ballot_sync(0xffffffffu, is_leaf);
        const unsigned im = __ballot_sync(0xffffffffu, is_int);
        if (is_leaf) wleaf[nleaf + __popc(lm & lt_mask)] = make_uint2(w, __float_as_uint(tmin));
        nleaf += __popc(lm);
        const int ni = __popc(im);
        if (sp + ni > CB_WSTACK) { redo = true; break; }
        // push the internal hits far -> near, so that the nearest ends up on top: selection sort
        // with one REDUX.MIN + ballot per hit (tmin >= 0, so its bit pattern orders like the float)
        int pos = 0;
        uint32_t key = is_int ? __float_as_uint(tmin) : 0xFFFFFFFFu;
        for (int r = ni - 1; r >= 0; r--) {
            const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
            const int leader = __ffs(__ballot_sync(0xffffffffu, key == kmin)) - 1;
            if ((int)lane == leader) { pos = r; key = 0xFFFFFFFFu; }
        }
        if (is_int) wstack[sp + pos] = make_uint2(w, __float_as_uint(tmin));
        sp += ni;
        __syncwarp();
    }
    if (best_tri != -1 && !redo) {
        const float4 lb = __ldg(g.tri64 + 4ull * (uint32_t)best_tri + 3);
        float box_t;
        const bool in_box = hit_box(g, rr, __float_as_uint(lb.x), __float_as_uint(lb.y), __float_as_uint(lb.z), box_t);
        redo = !in_box || best_t < box_t;
    }
    if (redo) {
        if (COUNT) cnt->resolved++;
        return traverse_reference_order<COUNT>(g, origin, direction, last_hit, dist, overflow_flag, cnt);
    }
    dist = (best_tri == -1) ? -1.0f : best_t;
    return best_tri;
}

// ------------------------------------------------------------------ physics
struct Photon {
    float3 pos, dir, pol;
    float wavelength, time, weight;
    uint32_t history;        // 16 significant bits (photon.h:29)
    int last_hit_triangle;
};

struct StepState {
    float3 normal;
    float n1, n2, absorption_length, scattering_length;
    const CbMaterial* material1;
    int surface_index;
    float distance;
};

enum { CMD_BREAK = 0, CMD_CONTINUE = 1, CMD_PASS = 2 };

// clamp-then-linear table lookup on the uniform wavelength grid (geometry.h:61-74)
__device__ __forceinline__ float interp_property(const DevGeometry& g, float x, const float* fp)
{
    if (x < g.wavelength_start) return fp[0];
    if (x > (g.wavelength_start + (g.wavelength_n - 1) * g.wavelength_step)) return fp[g.wavelength_n - 1];
    int jl = (x - g.wavelength_start) / g.wavelength_step;
    return fp[jl] + (x - (g.wavelength_start + jl * g.wavelength_step)) * (fp[jl + 1] - fp[jl]) / g.wavelength_step;
}

// inverse-CDF sampling on a uniform x grid (random.h:33-55)
__device__ __forceinline__ float sample_cdf_uniform(Rng& rng, int ncdf, float x0, float delta, const float* cdf_y)
{
    float u = rng_uniform(rng);
    int lower = 0, upper = ncdf - 1;
    while (lower < upper - 1) {
        int half = (lower + upper) / 2;
        if (u < cdf_y[half]) upper = half; else lower = half;
    }
    float delta_cdf_y = cdf_y[upper] - cdf_y[lower];
    return x0 + delta * lower + delta * (u - cdf_y[lower]) / delta_cdf_y;
}

// piecewise-linear interpolation by bisection (interpolate.h:33-58)
__device__ __forceinline__ float interp_xy(float x, int n, const float* xp, const float* fp)
{
    int lower = 0, upper = n - 1;
    if (x <= xp[lower]) return fp[lower];
    if (x >= xp[upper]) return fp[upper];
    while (lower < upper - 1) {
        int half = (lower + upper) / 2;
        if (x < xp[half]) upper = half; else lower = half;
    }
    float df = fp[upper] - fp[lower];
    float dx = xp[upper] - xp[lower];
    return fp[lower] + df * (x - xp[lower]) / dx;
}

// fractional index of x in xp (interpolate.h:5-29)
__device__ __forceinline__ float interp_idx(float x, int n, const float* xp)
{
    int lower = 0, upper = n - 1;
    if (x <= xp[lower]) return lower;
    if (x >= xp[upper]) return upper;
    while (lower < upper - 1) {
        int half = (lower + upper) / 2;
        if (x < xp[half]) upper = half; else lower = half;
    }
    float dx = xp[upper] - xp[lower];
    return lower + 1.0 * (x - xp[lower]) / dx;
}

__device__ __forceinline__ int sext8(int c) { return (c & 0x80) ? (0xFFFFFF00 | c) : c; }
__device__ __forceinline__ float get_theta(const float3& a, const float3& b)
{
    return acosf(fmaxf(-1.0f, fminf(1.0f, dot(a, b))));
}

// Classify the boundary found by the traversal (mesh branch of fill_state,
// photon.h:355-394).  `tri` >= 0.
__device__ __forceinline__ void classify_hit(const DevGeometry& g, const Tables& T, Photon& p,
                                             StepState& s, int tri)
{
    p.last_hit_triangle = tri;
    const float4* tp = g.tri64 + 4ull * (uint32_t)tri;
    float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
    float3 v0 = f3(a.x, a.y, a.z), v1 = f3(a.w, b.x, b.y), v2 = f3(b.z, b.w, c.x);
    uint32_t material_code = __float_as_uint(c.z);
    int inner = sext8(0xFF & (material_code >> 24));
    int outer = sext8(0xFF & (material_code >> 16));
    s.surface_index = sext8(0xFF & (material_code >> 8));

    float3 v01 = v1 - v0;
    float3 v12 = v2 - v1;
    s.normal = normalize(cross(v01, v12));

    const CbMaterial *m1, *m2;
    if (dot(s.normal, -p.dir) > 0.0f) {
        m1 = &g.materials[outer]; m2 = &g.materials[inner];
    } else {
        m1 = &g.materials[inner]; m2 = &g.materials[outer];
        s.normal = -s.normal;
    }
    s.n1 = interp_property(g, p.wavelength, T.at(m1->refractive_index));
    s.n2 = interp_property(g, p.wavelength, T.at(m2->refractive_index));
    s.absorption_length = interp_property(g, p.wavelength, T.at(m1->absorption_length));
    s.scattering_length = interp_property(g, p.wavelength, T.at(m1->scattering_length));
    s.material1 = m1;
}

// new direction at polar angle theta / azimuth phi about `axis` (photon.h:399-424)
__device__ __forceinline__ float3 pick_new_direction(float3 axis, float theta, float phi)
{
    float cos_theta, sin_theta;
    sincosf(theta, &sin_theta, &cos_theta);
    float cos_phi, sin_phi;
    sincosf(phi, &sin_phi, &cos_phi);
    float sin_axis_theta = sqrt(1.0f - axis.z * axis.z);
    float cos_axis_phi, sin_axis_phi;
    if (isnan(sin_axis_theta) || sin_axis_theta < 0.00001f) {
        cos_axis_phi = 1.0f;
        sin_axis_phi = 0.0f;
    } else {
        cos_axis_phi = axis.x / sin_axis_theta;
        sin_axis_phi = axis.y / sin_axis_theta;
    }
    float dirx = cos_theta * axis.x + sin_theta * (axis.z * cos_phi * cos_axis_phi - sin_phi * sin_axis_phi);
    float diry = cos_theta * axis.y + sin_theta * (cos_phi * axis.z * sin_axis_phi + sin_phi * cos_axis_phi);
    float dirz = cos_theta * axis.z - sin_theta * cos_phi * sin_axis_theta;
    return f3(dirx, diry, dirz);
}

// Rayleigh scattering about the polarisation axis (photon.h:426-453)
__device__ __forceinline__ void rayleigh_scatter(Photon& p, Rng& rng)
{
    float cos_theta = 2.0f * cosf((acosf(1.0f - 2.0f * rng_uniform(rng)) - 2 * CB_PI) / 3.0f);
    if (cos_theta > 1.0f) cos_theta = 1.0f;
    else if (cos_theta < -1.0f) cos_theta = -1.0f;
    float theta = acosf(cos_theta);
    float phi = rng_range(rng, 0.0f, 2.0f * CB_PI);
    p.dir = pick_new_direction(p.pol, theta, phi);
    if (1.0f - fabsf(cos_theta) < 1e-6f) p.pol = pick_new_direction(p.pol, CB_PI / 2.0f, phi);
    else p.pol = p.pol - cos_theta * p.dir;
    p.dir = p.dir / norm(p.dir);
    p.pol = p.pol / norm(p.pol);
}

// bulk step: absorption / re-emission / Rayleigh / reach boundary (photon.h:455-570)
// to_boundary and at_boundary are inlined into the kernels (physics kernel 0.40 -> 0.35 ms: the photon
// stays in registers across them).  The thin-film surface model stays ONE shared call: inlined copies differ
// in the last bit between kernels (FMA contraction across the call boundary), which would make results depend
// on the schedule (physics_step<WIRES, INLINE_SURFACES> exists for experiments only).
#ifndef CB_PHYS_CALL
#define CB_PHYS_CALL __forceinline__
#endif
#ifndef CB_PHYS_CALL2
#define CB_PHYS_CALL2 __forceinline__
#endif
static __device__ CB_PHYS_CALL int to_boundary(const DevGeometry& g, const Tables& T, Photon& p, StepState& s,
                                        Rng& rng, bool use_weights, int scatter_first)
{
    float absorption_distance = -s.absorption_length * logf(rng_uniform(rng));
    float scattering_distance = -s.scattering_length * logf(rng_uniform(rng));

    if (use_weights && p.weight > CB_WEIGHT_LOWER_THRESHOLD) absorption_distance = 1e30;
    else use_weights = false;

    if (scatter_first == 1) {
        float scatter_prob = 1.0f - expf(-s.distance / s.scattering_length);
        if (scatter_prob > CB_WEIGHT_LOWER_THRESHOLD) {
            int i = 0;
            while (i < 1000 && scattering_distance > s.distance) {
                scattering_distance = -s.scattering_length * logf(rng_uniform(rng));
                i++;
            }
            p.weight *= scatter_prob;
        }
    } else if (scatter_first == -1) {
        float no_scatter_prob = expf(-s.distance / s.scattering_length);
        if (no_scatter_prob > CB_WEIGHT_LOWER_THRESHOLD) {
            int i = 0;
            while (i < 1000 && scattering_distance <= s.distance) {
                scattering_distance = -s.scattering_length * logf(rng_uniform(rng));
                i++;
            }
            p.weight *= no_scatter_prob;
        }
    }

    if (absorption_distance <= scattering_distance) {
        if (absorption_distance <= s.distance) {
            p.time += absorption_distance / (CB_SPEED_OF_LIGHT / s.n1);
            p.pos = p.pos + absorption_distance * p.dir;
            const CbMaterial* m = s.material1;
            if (m->num_comp == 0) {
                p.last_hit_triangle = -1;
                p.history |= CB_BULK_ABSORB;
                return CMD_BREAK;
            }
            float uniform_sample_comp = rng_uniform(rng);
            float prob = 0.0f;
            int comp;
            for (comp = 0;; comp++) {
                float comp_abs = interp_property(g, p.wavelength, T.at(m->comp_absorption_length + comp * g.wavelength_n));
                prob += s.absorption_length / comp_abs;
                if (uniform_sample_comp < prob || comp + 1 == m->num_comp) break;
            }
            float uniform_sample_reemit = rng_uniform(rng);
            float comp_reemit_prob = interp_property(g, p.wavelength, T.at(m->comp_reemission_prob + comp * g.wavelength_n));
            if (uniform_sample_reemit < comp_reemit_prob) {
                p.wavelength = sample_cdf_uniform(rng, g.wavelength_n, g.wavelength_start, g.wavelength_step,
                                                  T.at(m->comp_reemission_wvl_cdf + comp * g.wavelength_n));
                p.time += sample_cdf_uniform(rng, g.time_n, g.time_start, g.time_step,
                                             T.at(m->comp_reemission_time_cdf + comp * g.time_n));
                p.dir = rng_sphere(rng);
                p.pol = cross(rng_sphere(rng), p.dir);
                p.pol = p.pol / norm(p.pol);
                p.last_hit_triangle = -1;
                p.history |= CB_BULK_REEMIT;
                return CMD_CONTINUE;
            }
            p.last_hit_triangle = -1;
            p.history |= CB_BULK_ABSORB;
            return CMD_BREAK;
        }
    } else {
        if (scattering_distance <= s.distance) {
            if (use_weights) p.weight *= expf(-scattering_distance / s.absorption_length);
            p.time += scattering_distance / (CB_SPEED_OF_LIGHT / s.n1);
            p.pos = p.pos + scattering_distance * p.dir;
            rayleigh_scatter(p, rng);
            p.history |= CB_RAYLEIGH_SCATTER;
            p.last_hit_triangle = -1;
            return CMD_CONTINUE;
        }
    }
    if (use_weights) p.weight *= expf(-s.distance / s.absorption_length);
    p.pos = p.pos + s.distance * p.dir;
    p.time += s.distance / (CB_SPEED_OF_LIGHT / s.n1);
    return CMD_PASS;
}

// Fresnel reflection / refraction (photon.h:572-632)
static __device__ CB_PHYS_CALL2 void at_boundary(Photon& p, StepState& s, Rng& rng)
{
    float incident_angle = get_theta(s.normal, -p.dir);
    float refracted_angle = asinf(sinf(incident_angle) * s.n1 / s.n2);

    float3 incident_plane_normal = cross(p.dir, s.normal);
    float incident_plane_normal_length = norm(incident_plane_normal);
    if (incident_plane_normal_length < 1e-6f) incident_plane_normal = p.pol;
    else incident_plane_normal = incident_plane_normal / incident_plane_normal_length;

    float normal_coefficient = dot(p.pol, incident_plane_normal);
    float normal_probability = normal_coefficient * normal_coefficient;

    float reflection_coefficient;
    if (rng_uniform(rng) < normal_probability) {
        reflection_coefficient = -sinf(incident_angle - refracted_angle) / sinf(incident_angle + refracted_angle);
        if ((rng_uniform(rng) < reflection_coefficient * reflection_coefficient) || isnan(refracted_angle)) {
            p.dir = rotate(s.normal, incident_angle, incident_plane_normal);
            p.history |= CB_REFLECT_SPECULAR;
        } else {
            p.dir = rotate(s.normal, CB_PI - refracted_angle, incident_plane_normal);
        }
        p.pol = incident_plane_normal;
    } else {
        reflection_coefficient = tanf(incident_angle - refracted_angle) / tanf(incident_angle + refracted_angle);
        if ((rng_uniform(rng) < reflection_coefficient * reflection_coefficient) || isnan(refracted_angle)) {
            p.dir = rotate(s.normal, incident_angle, incident_plane_normal);
            p.history |= CB_REFLECT_SPECULAR;
        } else {
            p.dir = rotate(s.normal, CB_PI - refracted_angle, incident_plane_normal);
        }
        p.pol = cross(incident_plane_normal, p.dir);
        p.pol = p.pol / norm(p.pol);
    }
}

// mirror reflection (photon.h:634-646)
__device__ __forceinline__ int specular_reflect(Photon& p, const StepState& s)
{
    float incident_angle = get_theta(s.normal, -p.dir);
    float3 incident_plane_normal = cross(p.dir, s.normal);
    incident_plane_normal = incident_plane_normal / norm(incident_plane_normal);
    p.dir = rotate(s.normal, incident_angle, incident_plane_normal);
    p.history |= CB_REFLECT_SPECULAR;
    return CMD_CONTINUE;
}

// Lambertian reflection by rejection (photon.h:648-667)
__device__ __forceinline__ int diffuse_reflect(Photon& p, const StepState& s, Rng& rng)
{
    float ndotv;
    do {
        p.dir = rng_sphere(rng);
        ndotv = dot(p.dir, s.normal);
        if (ndotv < 0.0f) {
            p.dir = -p.dir;
            ndotv = -ndotv;
        }
    } while (!(rng_uniform(rng) < ndotv));
    p.pol = cross(rng_sphere(rng), p.dir);
    p.pol = p.pol / norm(p.pol);
    p.history |= CB_REFLECT_DIFFUSE;
    return CMD_CONTINUE;
}

// complex helpers the reference adds to cuComplex.h (cx.h:29-35)
__device__ __forceinline__ float cx_arg(cuFloatComplex x) { return atan2f(x.y, x.x); }
__device__ __forceinline__ cuFloatComplex cx_sqrt(cuFloatComplex x)
{
    float r = sqrtf(cuCabsf(x));
    float t = cx_arg(x) / 2.0f;
    return make_cuFloatComplex(r * cosf(t), r * sinf(t));
}

// thin-film surface (n1 | eta+ik, thickness | n3), photon.h:669-827
__device__ __forceinline__ int surface_complex_body(const DevGeometry& g, const Tables& T, Photon& p, StepState& s,
                                                    Rng& rng, const CbSurface* surface, bool use_weights)
{
    float detect = interp_property(g, p.wavelength, T.at(surface->detect));
    float reflect_diffuse = interp_property(g, p.wavelength, T.at(surface->reflect_diffuse));
    float n2_eta = interp_property(g, p.wavelength, T.at(surface->eta));
    float n2_k = interp_property(g, p.wavelength, T.at(surface->k));

    cuFloatComplex n1 = make_cuFloatComplex(s.n1, 0.0f);
    cuFloatComplex n2 = make_cuFloatComplex(n2_eta, n2_k);
    cuFloatComplex n3 = make_cuFloatComplex(s.n2, 0.0f);

    float cos_t1 = dot(p.dir, s.normal);
    if (cos_t1 < 0.0f) cos_t1 = -cos_t1;
    float theta = acosf(cos_t1);

    cuFloatComplex cos1 = make_cuFloatComplex(cosf(theta), 0.0f);
    cuFloatComplex sin1 = make_cuFloatComplex(sinf(theta), 0.0f);

    float e = 2.0f * CB_PI * surface->thickness / p.wavelength;
    cuFloatComplex ratio13sin = cuCmulf(cuCmulf(cuCdivf(n1, n3), cuCdivf(n1, n3)), cuCmulf(sin1, sin1));
    cuFloatComplex cos3 = cx_sqrt(cuCsubf(make_cuFloatComplex(1.0f, 0.0f), ratio13sin));
    cuFloatComplex ratio12sin = cuCmulf(cuCmulf(cuCdivf(n1, n2), cuCdivf(n1, n2)), cuCmulf(sin1, sin1));
    cuFloatComplex cos2 = cx_sqrt(cuCsubf(make_cuFloatComplex(1.0f, 0.0f), ratio12sin));
    float u = cuCrealf(cuCmulf(n2, cos2));
    float v = cuCimagf(cuCmulf(n2, cos2));

    // s polarisation
    cuFloatComplex s_n1c1 = cuCmulf(n1, cos1);
    cuFloatComplex s_n2c2 = cuCmulf(n2, cos2);
    cuFloatComplex s_n3c3 = cuCmulf(n3, cos3);
    cuFloatComplex s_r12 = cuCdivf(cuCsubf(s_n1c1, s_n2c2), cuCaddf(s_n1c1, s_n2c2));
    cuFloatComplex s_r23 = cuCdivf(cuCsubf(s_n2c2, s_n3c3), cuCaddf(s_n2c2, s_n3c3));
    cuFloatComplex s_t12 = cuCdivf(cuCmulf(make_cuFloatComplex(2.0f, 0.0f), s_n1c1), cuCaddf(s_n1c1, s_n2c2));
    cuFloatComplex s_t23 = cuCdivf(cuCmulf(make_cuFloatComplex(2.0f, 0.0f), s_n2c2), cuCaddf(s_n2c2, s_n3c3));
    cuFloatComplex s_g = cuCdivf(s_n3c3, s_n1c1);

    float s_abs_r12 = cuCabsf(s_r12);
    float s_abs_r23 = cuCabsf(s_r23);
    float s_abs_t12 = cuCabsf(s_t12);
    float s_abs_t23 = cuCabsf(s_t23);
    float s_arg_r12 = cx_arg(s_r12);
    float s_arg_r23 = cx_arg(s_r23);
    float s_exp1 = exp(2.0f * v * e);
    float s_exp2 = 1.0f / s_exp1;
    float s_denom = s_exp1 + s_abs_r12 * s_abs_r12 * s_abs_r23 * s_abs_r23 * s_exp2 +
                    2.0f * s_abs_r12 * s_abs_r23 * cosf(s_arg_r23 + s_arg_r12 + 2.0f * u * e);
    float s_r = s_abs_r12 * s_abs_r12 * s_exp1 + s_abs_r23 * s_abs_r23 * s_exp2 +
                2.0f * s_abs_r12 * s_abs_r23 * cosf(s_arg_r23 - s_arg_r12 + 2.0f * u * e);
    s_r /= s_denom;
    float s_t = cuCrealf(s_g) * s_abs_t12 * s_abs_t12 * s_abs_t23 * s_abs_t23;
    s_t /= s_denom;

    // p polarisation
    cuFloatComplex p_n2c1 = cuCmulf(n2, cos1);
    cuFloatComplex p_n3c2 = cuCmulf(n3, cos2);
    cuFloatComplex p_n2c3 = cuCmulf(n2, cos3);
    cuFloatComplex p_n1c2 = cuCmulf(n1, cos2);
    cuFloatComplex p_r12 = cuCdivf(cuCsubf(p_n2c1, p_n1c2), cuCaddf(p_n2c1, p_n1c2));
    cuFloatComplex p_r23 = cuCdivf(cuCsubf(p_n3c2, p_n2c3), cuCaddf(p_n3c2, p_n2c3));
    cuFloatComplex p_t12 = cuCdivf(cuCmulf(cuCmulf(make_cuFloatComplex(2.0f, 0.0f), n1), cos1), cuCaddf(p_n2c1, p_n1c2));
    cuFloatComplex p_t23 = cuCdivf(cuCmulf(cuCmulf(make_cuFloatComplex(2.0f, 0.0f), n2), cos2), cuCaddf(p_n3c2, p_n2c3));
    cuFloatComplex p_g = cuCdivf(cuCmulf(n3, cos3), cuCmulf(n1, cos1));

    float p_abs_r12 = cuCabsf(p_r12);
    float p_abs_r23 = cuCabsf(p_r23);
    float p_abs_t12 = cuCabsf(p_t12);
    float p_abs_t23 = cuCabsf(p_t23);
    float p_arg_r12 = cx_arg(p_r12);
    float p_arg_r23 = cx_arg(p_r23);
    float p_exp1 = exp(2.0f * v * e);
    float p_exp2 = 1.0f / p_exp1;
    float p_denom = p_exp1 + p_abs_r12 * p_abs_r12 * p_abs_r23 * p_abs_r23 * p_exp2 +
                    2.0f * p_abs_r12 * p_abs_r23 * cosf(p_arg_r23 + p_arg_r12 + 2.0f * u * e);
    float p_r = p_abs_r12 * p_abs_r12 * p_exp1 + p_abs_r23 * p_abs_r23 * p_exp2 +
                2.0f * p_abs_r12 * p_abs_r23 * cosf(p_arg_r23 - p_arg_r12 + 2.0f * u * e);
    p_r /= p_denom;
    float p_t = cuCrealf(p_g) * p_abs_t12 * p_abs_t12 * p_abs_t23 * p_abs_t23;
    p_t /= p_denom;

    // s-polarisation fraction, as in at_boundary
    float incident_angle = get_theta(s.normal, -p.dir);
    float refracted_angle = asinf(sinf(incident_angle) * s.n1 / s.n2);
    float3 incident_plane_normal = cross(p.dir, s.normal);
    float incident_plane_normal_length = norm(incident_plane_normal);
    if (incident_plane_normal_length < 1e-6f) incident_plane_normal = p.pol;
    else incident_plane_normal = incident_plane_normal / incident_plane_normal_length;
    float normal_coefficient = dot(p.pol, incident_plane_normal);
    float normal_probability = normal_coefficient * normal_coefficient;

    float transmit = normal_probability * s_t + (1.0f - normal_probability) * p_t;
    if (!surface->transmissive) transmit = 0.0f;
    float reflect = normal_probability * s_r + (1.0f - normal_probability) * p_r;
    float absorb = 1.0f - transmit - reflect;

    if (use_weights && p.weight > CB_WEIGHT_LOWER_THRESHOLD && absorb < (1.0f - CB_WEIGHT_LOWER_THRESHOLD)) {
        float survive = 1.0f - absorb;
        absorb = 0.0f;
        p.weight *= survive;
        detect /= survive;
        reflect /= survive;
        transmit /= survive;
    }
    if (use_weights && detect > 0.0f) {
        p.history |= CB_SURFACE_DETECT;
        p.weight *= detect;
        return CMD_BREAK;
    }

    float uniform_sample = rng_uniform(rng);
    if (uniform_sample < absorb) {
        float uniform_sample_detect = rng_uniform(rng);
        if (uniform_sample_detect < detect) p.history |= CB_SURFACE_DETECT;
        else p.history |= CB_SURFACE_ABSORB;
        return CMD_BREAK;
    } else if (uniform_sample < absorb + reflect || !surface->transmissive) {
        float uniform_sample_reflect = rng_uniform(rng);
        if (uniform_sample_reflect < reflect_diffuse) return diffuse_reflect(p, s, rng);
        return specular_reflect(p, s);
    } else {
        p.dir = rotate(s.normal, CB_PI - refracted_angle, incident_plane_normal);
        p.pol = cross(incident_plane_normal, p.dir);
        p.pol = p.pol / norm(p.pol);
        p.history |= CB_SURFACE_TRANSMIT;
        return CMD_CONTINUE;
    }
}

static __device__ __noinline__ int surface_complex_call(const DevGeometry& g, const Tables& T, Photon& p, StepState& s,
                                                        Rng& rng, const CbSurface* surface, bool use_weights)
{
    return surface_complex_body(g, T, p, s, rng, surface, use_weights);
}
template <bool INLINE>
__device__ __forceinline__ int surface_complex(const DevGeometry& g, const Tables& T, Photon& p, StepState& s,
                                               Rng& rng, const CbSurface* surface, bool use_weights)
{
    if (INLINE) return surface_complex_body(g, T, p, s, rng, surface, use_weights);
    return surface_complex_call(g, T, p, s, rng, surface, use_weights);
}

// wavelength-shifting surface (photon.h:829-874)
__device__ __forceinline__ int surface_wls(const DevGeometry& g, const Tables& T, Photon& p, StepState& s,
                                           Rng& rng, const CbSurface* surface, bool use_weights)
{
    float absorb = interp_property(g, p.wavelength, T.at(surface->absorb));
    float reflect_specular = interp_property(g, p.wavelength, T.at(surface->reflect_specular));
    float reflect_diffuse = interp_property(g, p.wavelength, T.at(surface->reflect_diffuse));
    float reemit = interp_property(g, p.wavelength, T.at(surface->reemit));

    float uniform_sample = rng_uniform(rng);
    if (use_weights && p.weight > CB_WEIGHT_LOWER_THRESHOLD && absorb < (1.0f - CB_WEIGHT_LOWER_THRESHOLD)) {
        float survive = 1.0f - absorb;
        absorb = 0.0f;
        p.weight *= survive;
        reflect_diffuse /= survive;
        reflect_specular /= survive;
    }
    if (uniform_sample < absorb) {
        float uniform_sample_reemit = rng_uniform(rng);
        if (uniform_sample_reemit < reemit) {
            p.history |= CB_SURFACE_REEMIT;
            p.wavelength = sample_cdf_uniform(rng, g.wavelength_n, g.wavelength_start, g.wavelength_step,
                                              T.at(surface->reemission_cdf));
            p.dir = rng_sphere(rng);
            p.pol = cross(rng_sphere(rng), p.dir);
            p.pol = p.pol / norm(p.pol);
            return CMD_CONTINUE;
        }
        p.history |= CB_SURFACE_ABSORB;
        return CMD_BREAK;
    } else if (uniform_sample < absorb + reflect_specular + reflect_diffuse) {
        float uniform_sample_reflect = rng_uniform(rng) * (reflect_specular + reflect_diffuse);
        if (uniform_sample_reflect < reflect_specular) return specular_reflect(p, s);
        return diffuse_reflect(p, s, rng);
    }
    p.history |= CB_SURFACE_TRANSMIT;
    return CMD_PASS;
}

// dichroic filter: angle-of-incidence blended reflect/transmit tables (photon.h:877-907)
__device__ __forceinline__ int surface_dichroic(const DevGeometry& g, const Tables& T, Photon& p, StepState& s,
                                                Rng& rng, const CbSurface* surface)
{
    float incident_angle = get_theta(s.normal, -p.dir);
    float idx = interp_idx(incident_angle, surface->dichroic_nangles, T.at(surface->dichroic_angles));
    unsigned int iidx = (int)idx;
    const int W = g.wavelength_n;
    float reflect_prob_low = interp_property(g, p.wavelength, T.at(surface->dichroic_reflect + iidx * W));
    float reflect_prob_high = interp_property(g, p.wavelength, T.at(surface->dichroic_reflect + (iidx + 1) * W));
    float transmit_prob_low = interp_property(g, p.wavelength, T.at(surface->dichroic_transmit + iidx * W));
    float transmit_prob_high = interp_property(g, p.wavelength, T.at(surface->dichroic_transmit + (iidx + 1) * W));
    float reflect_prob = reflect_prob_low + (reflect_prob_high - reflect_prob_low) * (idx - iidx);
    float transmit_prob = transmit_prob_low + (transmit_prob_high - transmit_prob_low) * (idx - iidx);

    float uniform_sample = rng_uniform(rng);
    if (uniform_sample < reflect_prob) return specular_reflect(p, s);
    if (uniform_sample < transmit_prob + reflect_prob) {
        p.history |= CB_SURFACE_TRANSMIT;
        return CMD_PASS;
    }
    p.history |= CB_SURFACE_ABSORB;
    return CMD_BREAK;
}

// angle-tabulated surface (photon.h:909-951)
__device__ __forceinline__ int surface_angular(const DevGeometry& g, const Tables& T, Photon& p, StepState& s,
                                               Rng& rng, const CbSurface* surface, bool use_weights)
{
    float incident_angle = get_theta(s.normal, -p.dir);
    float idx = interp_idx(incident_angle, surface->angular_nangles, T.at(surface->angular_angles));
    unsigned int iidx = (int)idx;
    float t = idx - iidx;
    const float* tr = T.at(surface->angular_transmit);
    const float* rs = T.at(surface->angular_reflect_specular);
    const float* rd = T.at(surface->angular_reflect_diffuse);
    float transmit_prob = tr[iidx] + t * (tr[iidx + 1] - tr[iidx]);
    float reflect_spec_prob = rs[iidx] + t * (rs[iidx + 1] - rs[iidx]);
    float reflect_diff_prob = rd[iidx] + t * (rd[iidx + 1] - rd[iidx]);
    float absorb_prob = 1.0f - transmit_prob - reflect_spec_prob - reflect_diff_prob;

    if (use_weights && p.weight > CB_WEIGHT_LOWER_THRESHOLD && absorb_prob < (1.0f - CB_WEIGHT_LOWER_THRESHOLD)) {
        float survive = 1.0f - absorb_prob;
        absorb_prob = 0.0f;
        p.weight *= survive;
        transmit_prob /= survive;
        reflect_spec_prob /= survive;
        reflect_diff_prob /= survive;
    }
    float uniform_sample = rng_uniform(rng);
    if (uniform_sample < absorb_prob) {
        p.history |= CB_SURFACE_ABSORB;
        return CMD_BREAK;
    }
    if (uniform_sample < absorb_prob + transmit_prob) {
        p.history |= CB_SURFACE_TRANSMIT;
        return CMD_PASS;
    }
    if (uniform_sample < absorb_prob + transmit_prob + reflect_spec_prob) return specular_reflect(p, s);
    return diffuse_reflect(p, s, rng);
}

// surface dispatch + default model (photon.h:953-1037; the reference's
// effective default is CHROMA_FORCE_SCATTER_AT_PASS == 0)
template <bool INLINE_SURFACES>
__device__ __forceinline__ int at_surface(const DevGeometry& g, const Tables& T, Photon& p, StepState& s,
                                          Rng& rng, bool use_weights)
{
    const CbSurface* surface = &g.surfaces[s.surface_index];
    const int model = surface->model;
    if (model == CB_SURFACE_COMPLEX) return surface_complex<INLINE_SURFACES>(g, T, p, s, rng, surface, use_weights);
    if (model == CB_SURFACE_WLS) return surface_wls(g, T, p, s, rng, surface, use_weights);
    if (model == CB_SURFACE_DICHROIC) return surface_dichroic(g, T, p, s, rng, surface);
    if (model == CB_SURFACE_ANGULAR) return surface_angular(g, T, p, s, rng, surface, use_weights);

    float detect = interp_property(g, p.wavelength, T.at(surface->detect));
    float absorb = interp_property(g, p.wavelength, T.at(surface->absorb));
    float reflect_diffuse = interp_property(g, p.wavelength, T.at(surface->reflect_diffuse));
    float reflect_specular = interp_property(g, p.wavelength, T.at(surface->reflect_specular));

    float uniform_sample = rng_uniform(rng);
    if (use_weights && p.weight > CB_WEIGHT_LOWER_THRESHOLD && absorb < (1.0f - CB_WEIGHT_LOWER_THRESHOLD)) {
        float survive = 1.0f - absorb;
        absorb = 0.0f;
        p.weight *= survive;
        detect /= survive;
        reflect_diffuse /= survive;
        reflect_specular /= survive;
    }
    if (use_weights && detect > 0.0f) {
        p.history |= CB_SURFACE_DETECT;
        p.weight *= detect;
        return CMD_BREAK;
    }
    if (uniform_sample < absorb) {
        p.history |= CB_SURFACE_ABSORB;
        return CMD_BREAK;
    } else if (uniform_sample < absorb + detect) {
        p.history |= CB_SURFACE_DETECT;
        return CMD_BREAK;
    } else if (uniform_sample < absorb + detect + reflect_diffuse)
        return diffuse_reflect(p, s, rng);
    else if (uniform_sample < absorb + detect + reflect_diffuse + reflect_specular)
        return specular_reflect(p, s);
    return CMD_PASS;
}

// Nearest analytic wire-plane boundary along the photon's ray (behaviour of the
// analytic branch of fill_state, photon.h:108-270): per plane an orthonormal (u, v, n)
// frame in double precision, the ray clipped to the plane's u extent and to the slab
// |n| <= radius, the range of wire indices k the clipped segment can reach, and for
// each of them the roots of the ray-cylinder quadratic in the (v, n) plane.  The root
// taken depends on whether the origin is outside (entry root), inside (exit root) or
// numerically on the cylinder (a 1e-4 mm step).  Candidates never prune against each
// other except through `distance` (strictly nearer wins, compared in float).
struct WireHit {
    float distance;            // 1e30f: none
    int surface, material_inner, material_outer;
    float3 normal;             // outward cylinder normal at the hit (unoriented)
    float dot_raw;             // normal . (-direction)
};
static __device__ __noinline__ void wire_planes_nearest(const DevGeometry& g, const float3 pos, const float3 dir,
                                                        float best_distance, WireHit& hit)
{
    hit.distance = 1e30f;
    hit.surface = -1; hit.material_inner = -1; hit.material_outer = -1;
    hit.normal = f3(0.0f, 0.0f, 0.0f);
    hit.dot_raw = 0.0f;
    for (int ip = 0; ip < g.nwireplanes; ip++) {
        const CbWirePlane& wp = g.wireplanes[ip];
        // orthonormal frame: u normalised, v made orthogonal to u and normalised, n = u x v
        const double ux = (double)wp.u[0], uy = (double)wp.u[1], uz = (double)wp.u[2];
        const double vx0 = (double)wp.v[0], vy0 = (double)wp.v[1], vz0 = (double)wp.v[2];
        const double un = 1.0 / sqrt(ux * ux + uy * uy + uz * uz);
        const double ux1 = ux * un, uy1 = uy * un, uz1 = uz * un;
        const double vdotu = vx0 * ux1 + vy0 * uy1 + vz0 * uz1;
        const double vx1 = vx0 - vdotu * ux1;
        const double vy1 = vy0 - vdotu * uy1;
        const double vz1 = vz0 - vdotu * uz1;
        const double vn = 1.0 / sqrt(vx1 * vx1 + vy1 * vy1 + vz1 * vz1);
        const double vx = vx1 * vn, vy = vy1 * vn, vz = vz1 * vn;
        const double nx = uy1 * vz - uz1 * vy;
        const double ny = uz1 * vx - ux1 * vz;
        const double nz = ux1 * vy - uy1 * vx;

        const float3 w = pos - f3(wp.origin[0], wp.origin[1], wp.origin[2]);
        const double du = (double)dir.x * ux1 + (double)dir.y * uy1 + (double)dir.z * uz1;
        const double dv = (double)dir.x * vx + (double)dir.y * vy + (double)dir.z * vz;
        const double dn = (double)dir.x * nx + (double)dir.y * ny + (double)dir.z * nz;
        const double wu = (double)w.x * ux1 + (double)w.y * uy1 + (double)w.z * uz1;
        const double wv0 = (double)w.x * vx + (double)w.y * vy + (double)w.z * vz - (double)wp.v0;
        const double wn0 = (double)w.x * nx + (double)w.y * ny + (double)w.z * nz;

        // the ray's parameter window inside the plane's u extent
        double t_in = -1.0e300, t_out = 1.0e300;
        if (fabs(du) < 1e-15) {
            if (wu < (double)wp.umin || wu > (double)wp.umax) continue;
        } else {
            double t1 = ((double)wp.umin - wu) / du;
            double t2 = ((double)wp.umax - wu) / du;
            if (t1 > t2) { const double tmp = t1; t1 = t2; t2 = tmp; }
            if (t1 > t_in) t_in = t1;
            if (t2 < t_out) t_out = t2;
            if (t_in > t_out) continue;
        }

        const double pitch = (double)wp.pitch;
        const double inv_pitch = (pitch != 0.0) ? (1.0 / pitch) : 0.0;
        const double wire_radius = (double)wp.radius;
        const double wire_thickness = 2.0 * wire_radius;
        const double pad_v = 0.5 * wire_thickness + 1e-6;
        const double pad_n = 0.5 * wire_thickness + 1e-6;

        const int kmin = (int)ceil(((double)wp.vmin - (double)wp.v0) / pitch);
        const int kmax = (int)floor(((double)wp.vmax - (double)wp.v0) / pitch);
        const double A = dv * dv + dn * dn;
        int k_start = kmin, k_stop = kmax;

        if (kmin <= kmax) {
            // wires the segment can reach: clip to the slab around the plane, then to v
            const double t_eps = 1.0e-4;
            double t_lo = fmax(t_in, t_eps);
            double t_hi = t_out;
            const double best_cap = (double)best_distance;
            if (best_cap < t_hi) t_hi = best_cap;
            if (fabs(dn) > 1e-12) {
                double tn1 = (-pad_n - wn0) / dn;
                double tn2 = (pad_n - wn0) / dn;
                if (tn1 > tn2) { const double tmp = tn1; tn1 = tn2; tn2 = tmp; }
                t_lo = fmax(t_lo, tn1);
                t_hi = fmin(t_hi, tn2);
            } else if (fabs(wn0) > pad_n) {
                continue;
            }
            if (t_hi < t_lo) continue;
            if (fabs(dn) <= 1e-12 && fabs(dv) > 1e-12) {
                const double t_span = (pitch + wire_thickness) / fabs(dv);
                t_hi = fmin(t_hi, t_lo + t_span);
            }
            const double v_entry = wv0 + dv * t_lo;
            const double v_exit = wv0 + dv * t_hi;
            double v_lo = fmin(v_entry, v_exit) - pad_v;
            double v_hi = fmax(v_entry, v_exit) + pad_v;
            if (wv0 - pad_v < v_lo) v_lo = wv0 - pad_v;
            if (wv0 + pad_v > v_hi) v_hi = wv0 + pad_v;
            long long k_lo = (long long)floor(v_lo * inv_pitch);
            long long k_hi = (long long)ceil(v_hi * inv_pitch);
            if (k_lo < kmin) k_lo = kmin;
            if (k_hi > kmax) k_hi = kmax;
            if (k_lo > k_hi) continue;
            k_start = (int)k_lo;
            k_stop = (int)k_hi;
        }

        for (int k = k_start; k <= k_stop; k++) {
            const double wv = wv0 - (double)k * pitch;
            const double B = wv * dv + wn0 * dn;
            const double Cq = wv * wv + wn0 * wn0 - wire_radius * wire_radius;
            const double disc = B * B - A * Cq;
            if (disc < 0.0) continue;
            const double sqrt_disc = sqrt(disc);
            const double t_small = (-B - sqrt_disc) / A;
            const double t_large = (-B + sqrt_disc) / A;
            const double t_min = 1.0e-4;
            const double r2_wire = wire_radius * wire_radius;
            const double r2_0 = wv * wv + wn0 * wn0;
            const double eps0 = fmax(1e-18, 1e-12 * r2_wire);
            double t;
            if (r2_0 > r2_wire + eps0) {            // origin outside: forward entry root
                if (t_small <= t_min) continue;
                t = t_small;
            } else if (r2_0 < r2_wire - eps0) {     // origin inside: forward exit root
                if (t_large <= t_min) continue;
                t = t_large;
            } else {                                // numerically on the surface
                t = t_min;
            }
            const double uc = wu + du * t;
            if (uc < wp.umin || uc > wp.umax) continue;
            if ((float)t >= hit.distance) continue;
            if (t < t_in || t > t_out) continue;
            const double vn_hit = wv + dv * t;
            const double nn_hit = wn0 + dn * t;
            const double len = sqrt(vn_hit * vn_hit + nn_hit * nn_hit);
            if (len <= 0.0) continue;
            const float3 n_local = f3((float)((vn_hit / len) * vx + (nn_hit / len) * nx),
                                      (float)((vn_hit / len) * vy + (nn_hit / len) * ny),
                                      (float)((vn_hit / len) * vz + (nn_hit / len) * nz));
            hit.distance = (float)t;
            hit.surface = wp.surface_index;
            hit.material_inner = wp.material_inner_index;
            hit.material_outer = wp.material_outer_index;
            hit.normal = n_local;
            hit.dot_raw = dot(n_local, -dir);
        }
    }
}

// The analytic candidate competes with the mesh hit (photon.h:272-330): it wins when it
// is nearer (in double, by more than 1e-12) and its plane has a surface; the photon's
// last_hit_triangle becomes -2 and the materials follow the side the photon comes from.
static __device__ __noinline__ bool wire_plane_boundary(const DevGeometry& g, const Tables& T, Photon& p, StepState& s,
                                                        int tri, float distance)
{
    WireHit wh;
    const float best = (tri == -1) ? 1e30f : distance;
    wire_planes_nearest(g, p.pos, p.dir, best, wh);
    if (!(wh.surface >= 0 && (double)wh.distance + 1e-12 < (double)best)) return false;
    s.distance = wh.distance;
    s.surface_index = wh.surface;
    p.last_hit_triangle = -2;
    const CbMaterial *m1, *m2;
    if (wh.dot_raw > 0.0f) {                 // outside -> inside: the outward normal already faces the photon
        m1 = &g.materials[wh.material_outer]; m2 = &g.materials[wh.material_inner];
        s.normal = wh.normal;
    } else {
        m1 = &g.materials[wh.material_inner]; m2 = &g.materials[wh.material_outer];
        s.normal = -wh.normal;
    }
    s.n1 = interp_property(g, p.wavelength, T.at(m1->refractive_index));
    s.n2 = interp_property(g, p.wavelength, T.at(m2->refractive_index));
    s.absorption_length = interp_property(g, p.wavelength, T.at(m1->absorption_length));
    s.scattering_length = interp_property(g, p.wavelength, T.at(m1->scattering_length));
    s.material1 = m1;
    return true;
}

// everything after the intersection for one step (propagate.cu:312-336).
// Returns true when the photon continues to another step.
// WIRES: geometry with analytic wire planes (a separate instantiation, so that the cold
// path costs the usual kernels neither registers nor instructions).
template <bool WIRES, bool INLINE_SURFACES = false>
__device__ __forceinline__ bool physics_step(const DevGeometry& g, const Tables& T, Photon& p, Rng& rng,
                                             int tri, float distance, bool use_weights, int scatter_first)
{
    StepState s;
    const bool analytic = WIRES && g.nwireplanes > 0 && wire_plane_boundary(g, T, p, s, tri, distance);
    if (!analytic) {
        if (tri == -1) {
            p.last_hit_triangle = -1;
            p.history |= CB_NO_HIT;
            return false;
        }
        s.distance = distance;
        classify_hit(g, T, p, s, tri);
    }
    int command = to_boundary(g, T, p, s, rng, use_weights, scatter_first);
    if (command == CMD_BREAK) return false;
    if (command == CMD_CONTINUE) return true;
    if (s.surface_index != -1) {
        command = at_surface<INLINE_SURFACES>(g, T, p, s, rng, use_weights);
        if (command == CMD_BREAK) return false;
        if (command == CMD_CONTINUE) return true;
    }
    at_boundary(p, s, rng);
    return true;
}

__device__ __forceinline__ bool photon_is_nan(const Photon& p)
{
    return isnan(p.dir.x * p.dir.y * p.dir.z * p.pos.x * p.pos.y * p.pos.z);
}

// ------------------------------------------------------------------ bank I/O
__device__ __forceinline__ float3 ld3(const float* __restrict__ a, uint64_t i)
{
    return f3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
}
__device__ __forceinline__ void st3(float* __restrict__ a, uint64_t i, const float3& v)
{
    a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z;
}

} // namespace cb
