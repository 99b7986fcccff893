// engine.cuh -- device-side building blocks of the B200 photon transport engine:
// vector helpers, XORWOW, the BVH traversal and the optical physics.
//
// Reference behaviour followed (NOT its code structure): chroma/cuda/mesh.h,
// intersect.h, geometry.h, photon.h, random.h, interpolate.h, rotate.h.  The
// arithmetic expression shapes deliberately match the reference's so that, built
// with the same nvcc flags (--use_fast_math, default -fmad), decisions replay.
#pragma once
#include <cuda_runtime.h>
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
    uint32_t smem_floats;     // leading floats of the pool served from shared memory (ends on a table boundary)
    uint32_t smem_bytes;      // bytes the kernels copy for that (one float beyond, rounded to 16 B)
    uint32_t nmaterials, nsurfaces;
    const struct WireFrame* wireframes;   // analytic wire planes in their own frames (cold path; usually none)
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
// The reference's excursions into double are reproduced in float, bit for bit, at a third of the
// instructions (no F2F / MUFU.RCP64H / 5 DFMA / 6 DSETP per test):
//   * f = (float)(1.0 / (double)a) is the correctly rounded float reciprocal: rounding a quotient to 53
//     bits and then to 24 is innocuous when 53 >= 2*24 + 2 (Figueroa), so __frcp_rn(a) is the same float
//     (checked on 10^8 random floats, and by the full-size parity tests against the reference kernel);
//   * (double)u < -1e-6 holds for exactly the floats below the float nearest to -1e-6 (which lies above
//     it), and likewise t > 1e-6 and u > 1.0 + 1e-6 for the floats above the nearest floats (which lie
//     below): the comparisons against the float-rounded constants select the same floats.
__device__ __forceinline__ float3 cross_pinned(const float3& a, const float3& b)
{
    return f3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
              __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float dot_pinned(const float3& a, const float3& b)
{
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
}
#ifndef CB_TRI_DOUBLE
#define CB_TRI_DOUBLE 0      /* 1: the reference's literal mixed-precision expressions (A/B and bisecting aid) */
#endif
__device__ __forceinline__ bool hit_triangle(const float3& origin, const float3& direction,
                                             const float3& v0, const float3& v1, const float3& v2,
                                             float& distance)
{
    const float3 edge1 = f3(__fadd_rn(v1.x, -v0.x), __fadd_rn(v1.y, -v0.y), __fadd_rn(v1.z, -v0.z));
    const float3 edge2 = f3(__fadd_rn(v2.x, -v0.x), __fadd_rn(v2.y, -v0.y), __fadd_rn(v2.z, -v0.z));
    const float3 h = cross_pinned(direction, edge2);
    const float a = dot_pinned(edge1, h);
    if (a > -FLT_EPSILON && a < FLT_EPSILON) return false;
#if CB_TRI_DOUBLE
    const float f = 1.0 / a;
    const double LO = -1e-6, HI = 1.0 + 1e-6, TMIN = 1e-6;
#else
    const float f = __frcp_rn(a);
    const float LO = -1e-6f, HI = (float)(1.0 + 1e-6), TMIN = 1e-6f;
#endif
    const float3 s = f3(__fadd_rn(origin.x, -v0.x), __fadd_rn(origin.y, -v0.y), __fadd_rn(origin.z, -v0.z));
    const float sh = dot_pinned(s, h);
    const float u = __fmul_rn(f, sh);
    if (u < LO || u > HI) return false;
    const float3 q = cross_pinned(s, edge1);
    const float v = __fmul_rn(f, dot_pinned(direction, q));
    const float upv = __fmaf_rn(f, sh, v);
    if (v < LO || upv > HI) return false;
    const float t = __fmul_rn(f, dot_pinned(edge2, q));
    if (t > TMIN && t < __int_as_float(0x7f800000)) {
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
struct WorldGrid { float3 origin; float scale; };
__device__ __forceinline__ bool hit_box(const WorldGrid& g, const RaySetup& r, uint32_t px,
                                        uint32_t py, uint32_t pz, float& tnear)
{
    const float INF = __int_as_float(0x7f800000);
    float tmin = 0.0f, tmax = INF;
    if (r.fx) {
        float lo = __fmaf_rn(u16f_lo(px), g.scale, g.origin.x);
        float hi = __fmaf_rn(u16f_hi(px), g.scale, g.origin.x);
        float t0 = __fmaf_rn(lo, r.inv.x, r.noid.x);
        float t1 = __fmaf_rn(hi, r.inv.x, r.noid.x);
        tmin = fmaxf(tmin, fminf(t0, t1));
        tmax = fminf(tmax, fmaxf(t0, t1));
    }
    if (r.fy) {
        float lo = __fmaf_rn(u16f_lo(py), g.scale, g.origin.y);
        float hi = __fmaf_rn(u16f_hi(py), g.scale, g.origin.y);
        float t0 = __fmaf_rn(lo, r.inv.y, r.noid.y);
        float t1 = __fmaf_rn(hi, r.inv.y, r.noid.y);
        tmin = fmaxf(tmin, fminf(t0, t1));
        tmax = fminf(tmax, fmaxf(t0, t1));
    }
    if (r.fz) {
        float lo = __fmaf_rn(u16f_lo(pz), g.scale, g.origin.z);
        float hi = __fmaf_rn(u16f_hi(pz), g.scale, g.origin.z);
        float t0 = __fmaf_rn(lo, r.inv.z, r.noid.z);
        float t1 = __fmaf_rn(hi, r.inv.z, r.noid.z);
        tmin = fmaxf(tmin, fminf(t0, t1));
        tmax = fminf(tmax, fmaxf(t0, t1));
    }
    tnear = tmin;
    return !(tmin > tmax);
}

__device__ __forceinline__ RaySetup ray_setup(const float3& o, const float3& d)
{
    RaySetup r;
    r.noid = f3(-o.x / d.x, -o.y / d.y, -o.z / d.z);
    r.inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    r.fx = isfinite(r.inv.x); r.fy = isfinite(r.inv.y); r.fz = isfinite(r.inv.z);
    return r;
}
__device__ __forceinline__ WorldGrid world_grid(const DevGeometry& g)
{
    WorldGrid w = {g.world_origin, g.world_scale};
    return w;
}

// slab test in the reference's arithmetic, recomputing the ray setup (rare path).  Everything
// by value: a caller never has to park its ray in local memory for this call.  Returns the
// entry distance (>= 0), or -1 when the ray misses the box.
static __device__ __noinline__ float box_entry_exact(const float3 worigin, const float wscale, const float3 o, const float3 d,
                                                     uint32_t px, uint32_t py, uint32_t pz)
{
    const WorldGrid w = {worigin, wscale};
    float tnear;
    return hit_box(w, ray_setup(o, d), px, py, pz, tnear) ? tnear : -1.0f;
}

// ------------------------------------------------------------------ traversal
constexpr int CB_RSTACK = 64;   // local stack of the reference-order fallback

// Exact emulation of the reference's visit order (behaviour of mesh.h:45-126:
// children ascending, leaves tested on the spot, internal children pushed and
// popped LIFO, prune against the current minimum).  Only used for the rare rays
// the traversal flags as order-sensitive (see PTrav::finish).
struct RayHit { int tri; float dist; };      // tri == -1: no hit (dist -1)

template <bool COUNT>
static __device__ __noinline__ RayHit traverse_reference_order(const DevGeometry& g, const float3 origin,
                                                               const float3 direction, int last_hit,
                                                               uint32_t* overflow_flag, TraverseCounters* cnt)
{
    const RaySetup r = ray_setup(origin, direction);
    const WorldGrid w = world_grid(g);
    RayHit best = {-1, -1.0f};
    uint32_t stack[CB_RSTACK];
    {
        const uint4 root = __ldg(&g.ref_nodes[0]);
        float tn;
        if (!hit_box(w, r, root.x, root.y, root.z, tn)) return best;
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
            if (!hit_box(w, r, nd.x, nd.y, nd.z, tmin) || (best.dist >= 0.0f && tmin > best.dist)) continue;
            const uint32_t child = nd.w;
            if ((child >> 28) != 0) {
                if (sp < CB_RSTACK) stack[sp++] = child;
                else atomicOr(overflow_flag, 1u);
                continue;
            }
            if ((int)child == last_hit) continue;
            if (COUNT) cnt->tris++;
            const float4* tp = g.tri64 + 4ull * child;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            float t;
            if (hit_triangle(origin, direction, f3(a.x, a.y, a.z), f3(a.w, b.x, b.y), f3(b.z, b.w, c.x), t) &&
                (best.tri == -1 || t < best.dist)) {
                best.tri = (int)child;
                best.dist = t;
            }
        }
    }
    return best;
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
#define CB_PSTACK_N 24   /* the single-level tree holds up to 31 entries; 0.013 % of its expansions exceed 24 */
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
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
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
    } else {
        // The ray runs parallel to this axis' planes (d == 0 exactly: 3e-8 of the isotropic directions drawn by
        // rng_sphere are (0, 0, +-1)).  The reference skips such an axis (intersect.h:120) and therefore enters every
        // box that lies in the ray's way on the OTHER axes, wherever it is on this one: for a ray along z that is a
        // slab through the whole detector -- 164 ms for one photon in the 37 M-triangle detector, the straggler event
        // of the 8-GPU runs of round 2.  A triangle can only be hit if its extent on this axis contains the ray, so
        // the axis still culls here: with the planes scaled by BIG, "the ray is inside the slab (+- 8 grid quanta,
        // far more than the triangle test's rounding)" becomes tnear = -huge, tfar = +huge, "outside" becomes
        // tnear = +huge or tfar = -huge.  Same hits (the triangle test decides), found in microseconds.
        const float BIG = 1e30f;
        const float xq = (o - worigin) / wscale;   // where the ray is, in grid quanta
        s = BIG; sel = 0x7610u;
        n = -(8388608.0f + xq + 8.0f) * BIG;
        f = -(8388608.0f + xq - 8.0f) * BIG;
        (void)INF;
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

#ifndef CB_LANE_PREFETCH
#define CB_LANE_PREFETCH 0
#endif
#ifndef CB_LEAF_PREFETCH
#define CB_LEAF_PREFETCH 0   /* cheap expansion: prefetch the records of queued triangles (1: to L2, 2: to L1) */
#endif
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
        tn = box_entry_exact(g.world_origin, g.world_scale, o, d, g.ref_root_x, g.ref_root_y, g.ref_root_z);
        have = tn >= 0.0f && (g.root_w >> 28) != 0;
        return have;
    }

    // ---- expansion of the current internal entry, in bursts of four children
    struct Nearest { uint32_t w; float t; };     // nearest internal hit so far (w == 0: none; an internal word is never 0)

    // test children i..i+3 of the current entry (already fetched into nd; slots >= n hold a copy of child n-1)
    template <bool COUNT>
    __device__ __forceinline__ void process4(const uint4 (&nd)[4], uint32_t i, uint32_t n, Nearest& nr,
                                             uint32_t sbase, uint2* lstack, TraverseCounters* cnt,
                                             const uint4* nodes_base = nullptr, const float4* tri64_base = nullptr)
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
            if (CB_LANE_PREFETCH && is_leaf) {
                const char* rec = reinterpret_cast<const char*>(tri64_base + 4ull * w);
                prefetch_l2(rec); prefetch_l2(rec + 32);
            }
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
            if (CB_LANE_PREFETCH && do_push) prefetch_l2(nodes_base + (pw & 0x0FFFFFFFu));
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

    // ---- the same expansion when the shared-memory stack has room for all eight children (the caller
    // checks that for the whole warp): EVERY internal hit is stored on the stack, only the slot and the
    // distance of the nearest one are tracked, and at the end the nearest is taken out by moving the top
    // entry into its slot.  13 instead of 24 bookkeeping instructions per child (r02 ncu: the child loop was
    // 44 instructions per child, 20 of them the plane test); the other entries end up in a slightly
    // different order, which changes nothing but the order of visits.
    struct Picked { float t; uint32_t at; };
    // EXISTS: the caller knows that child idx exists (children 0 and, behind the n > 4 test, 4)
    template <bool COUNT, bool EXISTS = false>
    __device__ __forceinline__ void roomy_child(const uint4& nd, uint32_t idx, uint32_t n, Picked& pk, TraverseCounters* cnt,
                                                const float4* tri_base = nullptr)
    {
        float tmin;
        const bool ok = hit_box_phased(r, nd.x, nd.y, nd.z, tmin) && !(tmin > limit) && (EXISTS || idx < n);
        if (COUNT) cnt->nodes += (idx < n);
        const uint32_t w = nd.w;
        const bool is_leaf = ok && w < 0x10000000u;
        const bool is_int = ok && w >= 0x10000000u;
        if (is_leaf) {
            sts64(lq, w, __float_as_uint(tmin)); lq += CB_PSTRIDE;
#if CB_LEAF_PREFETCH == 1
            { const char* rec = reinterpret_cast<const char*>(tri_base + 4ull * w); prefetch_l2(rec); prefetch_l2(rec + 32); }
#elif CB_LEAF_PREFETCH == 2
            { const char* rec = reinterpret_cast<const char*>(tri_base + 4ull * w);
              asm volatile("prefetch.global.L1 [%0];" ::"l"(rec)); asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + 32)); }
#endif
        }
        if (is_int) {
            sts64(sp, w, __float_as_uint(tmin));
            if (tmin < pk.t) { pk.t = tmin; pk.at = sp; }
            sp += CB_PSTRIDE;
        }
    }
    template <bool COUNT>
    __device__ __forceinline__ void process4_roomy(const uint4 (&nd)[4], uint32_t i, uint32_t n, Picked& pk, TraverseCounters* cnt,
                                                   const float4* tri_base = nullptr)
    {
        roomy_child<COUNT, true>(nd[0], i, n, pk, cnt, tri_base);          // i is 0, or 4 behind the caller's n > 4 test
#pragma unroll
        for (int k = 1; k < 4; k++) roomy_child<COUNT>(nd[k], i + k, n, pk, cnt, tri_base);
    }
    __device__ __forceinline__ void expand_end_roomy(const Picked& pk, uint32_t sp0, uint32_t sbase, const uint2* lstack)
    {
        if (sp != sp0) {
            sp -= CB_PSTRIDE;
            const uint2 top = lds64(sp);
            const uint2 e = lds64(pk.at);
            cur = e.x; cur_t = pk.t;
            if (pk.at != sp) sts64(pk.at, top.x, top.y);
        } else {
            pop_next(sbase, lstack);
        }
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
            const float box_t = box_entry_exact(g.world_origin, g.world_scale, origin, direction, __float_as_uint(lb.x),
                                                __float_as_uint(lb.y), __float_as_uint(lb.z));
            redo = box_t < 0.0f || best_t < box_t;
        }
        if (redo) {
            TraverseCounters local = {0, 0, 0};
            const RayHit h = traverse_reference_order<COUNT>(g, origin, direction, last_hit, overflow_flag, COUNT ? &local : nullptr);
            if (COUNT) { cnt->nodes += local.nodes; cnt->tris += local.tris; cnt->resolved++; }
            dist = h.dist;
            return h.tri;
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

// prefetch hints (no register, no dependency): the warp-cooperative traversal is bound by the latency of
// its dependent fetches, 30 % of which hit L2 in the tail (ncu r02); entries and triangle records are
// requested as soon as they are known to be wanted
#ifndef CB_TAIL_PREFETCH
#define CB_TAIL_PREFETCH 0   /* measured r02: 1.616 ms with, 1.594 ms without (29k-PMT tail) */
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
    // world-box test and the winner's leaf-box check use the reference's arithmetic through a by-value
    // call, so that its ray setup (9 registers) is not kept alive across the traversal loop
    if (box_entry_exact(g.world_origin, g.world_scale, origin, direction, g.ref_root_x, g.ref_root_y, g.ref_root_z) < 0.0f ||
        (g.root_w >> 28) == 0) { dist = -1.0f; return -1; }
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
        if (CB_TAIL_PREFETCH) {
            if (is_leaf) {                               // the 48 bytes the triangle test reads span two sectors
                const char* rec = reinterpret_cast<const char*>(g.tri64 + 4ull * w);
                prefetch_l2(rec); prefetch_l2(rec + 32);
            } else if (is_int) {                         // the children of a node that is about to be pushed
                const uint4* kids = g.nodes + (w & 0x0FFFFFFFu);
                prefetch_l2(kids);
                if ((w >> 28) > 2) prefetch_l2(kids + 2);
                if ((w >> 28) > 4) prefetch_l2(kids + 4);
                if ((w >> 28) > 6) prefetch_l2(kids + 6);
            }
        }
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
        const float box_t = box_entry_exact(g.world_origin, g.world_scale, origin, direction, __float_as_uint(lb.x),
                                            __float_as_uint(lb.y), __float_as_uint(lb.z));
        redo = box_t < 0.0f || best_t < box_t;
    }
    if (redo) {
        if (COUNT) cnt->resolved++;
        const RayHit h = traverse_reference_order<COUNT>(g, origin, direction, last_hit, overflow_flag, COUNT ? cnt : nullptr);
        dist = h.dist;
        return h.tri;
    }
    dist = (best_tri == -1) ? -1.0f : best_t;
    return best_tri;
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

#include "physics.cuh"
