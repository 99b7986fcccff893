// physics.cuh -- optical physics of one photon step, the engine's own organisation.
//
// Behaviour follows chroma/cuda/photon.h (cited per function); the structure does not:
//
//   Incidence        how the photon meets the boundary (angles, s axis, s-polarised share);
//                    computed once and shared by the dielectric boundary and the thin film
//   bulk_step        free paths -> absorbed / re-emitted / Rayleigh scattered / at the boundary
//   fresnel_step     dielectric boundary
//   Fate             what a surface model decides; every model only PICKS a fate, one shared
//                    epilogue (apply_fate) changes the photon
//   thin_film        a pure function (by value in, reflectance / transmittance out): the one
//                    non-inlined piece, shared by all kernels so its arithmetic is identical
//                    in each of them, and free of by-reference arguments so that no caller has
//                    to keep its photon in local memory
//   WireFrame        per-plane orthonormal frame and wire range, prepared once on the host
//                    (cb_geometry_create) instead of per photon and step in double precision
//
// What is kept on purpose: the ORDER of the random draws and the shape of the float
// expressions that decide between outcomes (--use_fast_math, default -fmad), because
// seeded runs must replay the reference's histories.
#pragma once

namespace cb {

struct Photon {
    float3 pos, dir, pol;
    float wavelength, time, weight;
    uint32_t history;        // 16 significant bits (photon.h:29)
    int last_hit_triangle;
};

// the boundary ahead of the photon and the medium it travels in
struct StepState {
    float3 normal;           // faces the incoming photon
    float n1, n2, absorption_length, scattering_length;
    const CbMaterial* material1;
    int surface_index;
    float distance;
};

enum Step { STEP_ENDS = 0, STEP_CONTINUES = 1, STEP_AT_BOUNDARY = 2 };

// ------------------------------------------------------------------ table lookups
// clamp-then-linear lookup on the uniform wavelength grid (geometry.h:61-74)
__device__ __forceinline__ float interp_property(const DevGeometry& g, float x, const float* fp)
{
    if (x < g.wavelength_start) return fp[0];
    if (x > (g.wavelength_start + (g.wavelength_n - 1) * g.wavelength_step)) return fp[g.wavelength_n - 1];
    int jl = (x - g.wavelength_start) / g.wavelength_step;
    return fp[jl] + (x - (g.wavelength_start + jl * g.wavelength_step)) * (fp[jl + 1] - fp[jl]) / g.wavelength_step;
}

// bracket [lo, lo+1] of x in an ascending array by bisection (shared by the three lookups below)
__device__ __forceinline__ int bisect(float x, int n, const float* xs)
{
    int lo = 0, hi = n - 1;
    while (lo < hi - 1) {
        const int mid = (lo + hi) / 2;
        if (x < xs[mid]) hi = mid; else lo = mid;
    }
    return lo;
}

// inverse-CDF sampling on a uniform x grid (random.h:33-55)
__device__ __forceinline__ float sample_cdf_uniform(Rng& rng, int ncdf, float x0, float delta, const float* cdf_y)
{
    const float u = rng_uniform(rng);
    const int lo = bisect(u, ncdf, cdf_y);
    const float rise = cdf_y[lo + 1] - cdf_y[lo];
    return x0 + delta * lo + delta * (u - cdf_y[lo]) / rise;
}

// piecewise-linear y(x) through (xp, fp), flat outside (interpolate.h:33-58)
__device__ __forceinline__ float interp_xy(float x, int n, const float* xp, const float* fp)
{
    if (x <= xp[0]) return fp[0];
    if (x >= xp[n - 1]) return fp[n - 1];
    const int lo = bisect(x, n, xp);
    const float df = fp[lo + 1] - fp[lo];
    const float dx = xp[lo + 1] - xp[lo];
    return fp[lo] + df * (x - xp[lo]) / dx;
}

// fractional index of x in xp (interpolate.h:5-29)
__device__ __forceinline__ float interp_idx(float x, int n, const float* xp)
{
    if (x <= xp[0]) return 0;
    if (x >= xp[n - 1]) return n - 1;
    const int lo = bisect(x, n, xp);
    const float dx = xp[lo + 1] - xp[lo];
    return lo + 1.0 * (x - xp[lo]) / dx;
}

__device__ __forceinline__ int sext8(int c) { return (c & 0x80) ? (0xFFFFFF00 | c) : c; }
__device__ __forceinline__ float get_theta(const float3& a, const float3& b)
{
    return acosf(fmaxf(-1.0f, fminf(1.0f, dot(a, b))));
}
__device__ __forceinline__ float3 unit(const float3& a) { return a / norm(a); }

// the four bulk properties the step needs, looked up for the medium the photon is in
__device__ __forceinline__ void enter_media(const DevGeometry& g, const Tables& T, float wavelength,
                                            const CbMaterial* from, const CbMaterial* into, StepState& s)
{
    s.n1 = interp_property(g, wavelength, T.at(from->refractive_index));
    s.n2 = interp_property(g, wavelength, T.at(into->refractive_index));
    s.absorption_length = interp_property(g, wavelength, T.at(from->absorption_length));
    s.scattering_length = interp_property(g, wavelength, T.at(from->scattering_length));
    s.material1 = from;
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
    const bool from_outside = dot(s.normal, -p.dir) > 0.0f;
    if (!from_outside) s.normal = -s.normal;
    enter_media(g, T, p.wavelength, &g.materials[from_outside ? outer : inner], &g.materials[from_outside ? inner : outer], s);
}

// ------------------------------------------------------------------ directions
// unit vector at polar angle theta / azimuth phi about `axis` (photon.h:399-424)
__device__ __forceinline__ float3 direction_about(float3 axis, float theta, float phi)
{
    float ct, st, cp, sp;
    sincosf(theta, &st, &ct);
    sincosf(phi, &sp, &cp);
    // azimuth of the axis itself; an axis along z has none
    const float axis_st = sqrt(1.0f - axis.z * axis.z);
    const bool polar = isnan(axis_st) || axis_st < 0.00001f;
    const float axis_cp = polar ? 1.0f : axis.x / axis_st;
    const float axis_sp = polar ? 0.0f : axis.y / axis_st;
    return f3(ct * axis.x + st * (axis.z * cp * axis_cp - sp * axis_sp),
              ct * axis.y + st * (cp * axis.z * axis_sp + sp * axis_cp),
              ct * axis.z - st * cp * axis_st);
}

// isotropic re-emission: fresh direction, polarisation perpendicular to it
// (the tail of bulk re-emission photon.h:531-541 and of WLS re-emission :850-858)
__device__ __forceinline__ void emit_isotropic(Photon& p, Rng& rng)
{
    p.dir = rng_sphere(rng);
    p.pol = unit(cross(rng_sphere(rng), p.dir));
}

// Rayleigh scattering about the polarisation axis (photon.h:426-453)
__device__ __forceinline__ void rayleigh_scatter(Photon& p, Rng& rng)
{
    const float c = 2.0f * cosf((acosf(1.0f - 2.0f * rng_uniform(rng)) - 2 * CB_PI) / 3.0f);
    const float cos_theta = fminf(1.0f, fmaxf(-1.0f, c));
    const float phi = rng_range(rng, 0.0f, 2.0f * CB_PI);
    const float3 old_pol = p.pol;
    p.dir = direction_about(old_pol, acosf(cos_theta), phi);
    // the new polarisation lies in the plane of the old one and the new direction; straight
    // forward / backward that plane is undefined and any perpendicular at azimuth phi does
    p.pol = (1.0f - fabsf(cos_theta) < 1e-6f) ? direction_about(old_pol, CB_PI / 2.0f, phi) : old_pol - cos_theta * p.dir;
    p.dir = unit(p.dir);
    p.pol = unit(p.pol);
}

// ------------------------------------------------------------------ bulk
__device__ __forceinline__ float free_path(float mean, Rng& rng) { return -mean * logf(rng_uniform(rng)); }
__device__ __forceinline__ void advance(Photon& p, float d, float n)
{
    p.time += d / (CB_SPEED_OF_LIGHT / n);
    p.pos = p.pos + d * p.dir;
}

// forced (+1) / forbidden (-1) first scatter: redraw the scattering length until it falls on
// the wanted side of the boundary (at most 1000 times) and carry the probability of that in
// the photon's weight (photon.h:467-495)
__device__ __forceinline__ float bias_first_scatter(Photon& p, const StepState& s, Rng& rng, int mode, float scatter_at)
{
    const float miss = expf(-s.distance / s.scattering_length);          // no scatter before the boundary
    const float prob = (mode == 1) ? 1.0f - miss : miss;
    if (!(prob > CB_WEIGHT_LOWER_THRESHOLD)) return scatter_at;
    for (int tries = 0; tries < 1000 && ((scatter_at > s.distance) == (mode == 1)); tries++)
        scatter_at = free_path(s.scattering_length, rng);
    p.weight *= prob;
    return scatter_at;
}

// absorbed in a medium with re-emitting components: pick the component by its share of the
// absorption, then re-emit with its probability (photon.h:506-549).  Returns false = absorbed.
__device__ __forceinline__ bool bulk_reemit(const DevGeometry& g, const Tables& T, Photon& p, const StepState& s, Rng& rng)
{
    const CbMaterial* m = s.material1;
    const int W = g.wavelength_n;
    const float pick = rng_uniform(rng);
    float share = 0.0f;
    int comp = 0;
    for (;; comp++) {
        share += s.absorption_length / interp_property(g, p.wavelength, T.at(m->comp_absorption_length + comp * W));
        if (pick < share || comp + 1 == m->num_comp) break;
    }
    const float u = rng_uniform(rng);
    if (!(u < interp_property(g, p.wavelength, T.at(m->comp_reemission_prob + comp * W)))) return false;
    p.wavelength = sample_cdf_uniform(rng, W, g.wavelength_start, g.wavelength_step, T.at(m->comp_reemission_wvl_cdf + comp * W));
    p.time += sample_cdf_uniform(rng, g.time_n, g.time_start, g.time_step, T.at(m->comp_reemission_time_cdf + comp * g.time_n));
    emit_isotropic(p, rng);
    return true;
}

// bulk step: absorption / re-emission / Rayleigh / reach the boundary (photon.h:455-570)
__device__ __forceinline__ int bulk_step(const DevGeometry& g, const Tables& T, Photon& p, StepState& s,
                                         Rng& rng, bool use_weights, int scatter_first)
{
    float absorb_at = free_path(s.absorption_length, rng);
    float scatter_at = free_path(s.scattering_length, rng);
    const bool weighted = use_weights && p.weight > CB_WEIGHT_LOWER_THRESHOLD;   // absorption becomes a weight
    if (weighted) absorb_at = 1e30;
    if (scatter_first != 0) scatter_at = bias_first_scatter(p, s, rng, scatter_first, scatter_at);

    const bool absorbs_first = absorb_at <= scatter_at;
    const float stop = absorbs_first ? absorb_at : scatter_at;
    if (!(stop <= s.distance)) {                        // nothing happens before the boundary
        if (weighted) p.weight *= expf(-s.distance / s.absorption_length);
        p.pos = p.pos + s.distance * p.dir;
        p.time += s.distance / (CB_SPEED_OF_LIGHT / s.n1);
        return STEP_AT_BOUNDARY;
    }
    if (!absorbs_first && weighted) p.weight *= expf(-scatter_at / s.absorption_length);
    advance(p, stop, s.n1);
    p.last_hit_triangle = -1;
    if (!absorbs_first) {
        rayleigh_scatter(p, rng);
        p.history |= CB_RAYLEIGH_SCATTER;
        return STEP_CONTINUES;
    }
    if (s.material1->num_comp != 0 && bulk_reemit(g, T, p, s, rng)) {
        p.history |= CB_BULK_REEMIT;
        return STEP_CONTINUES;
    }
    p.history |= CB_BULK_ABSORB;
    return STEP_ENDS;
}

// ------------------------------------------------------------------ meeting a surface
struct Incidence {
    float theta_i, theta_t;  // angle of incidence; angle of refraction (NaN beyond the critical angle)
    float3 s_axis;           // unit normal of the plane of incidence = direction of s polarisation
    float s_share;           // probability of finding the photon s-polarised
};
// (prologue of propagate_at_boundary photon.h:575-590, repeated in propagate_complex :757-770)
__device__ __forceinline__ Incidence meet_boundary(const Photon& p, const StepState& s)
{
    Incidence in;
    in.theta_i = get_theta(s.normal, -p.dir);
    in.theta_t = asinf(sinf(in.theta_i) * s.n1 / s.n2);
    const float3 axis = cross(p.dir, s.normal);
    const float len = norm(axis);
    in.s_axis = (len < 1e-6f) ? p.pol : axis / len;      // normal incidence: any axis, take the polarisation's
    const float along = dot(p.pol, in.s_axis);
    in.s_share = along * along;
    return in;
}
// refracted ray: direction turned to theta_t behind the surface, polarisation kept in its plane
__device__ __forceinline__ void refract(Photon& p, const StepState& s, const Incidence& in, bool s_polarised)
{
    p.dir = rotate(s.normal, CB_PI - in.theta_t, in.s_axis);
    p.pol = s_polarised ? in.s_axis : unit(cross(in.s_axis, p.dir));
}

// Fresnel reflection / refraction at a dielectric boundary (photon.h:572-632)
__device__ __forceinline__ void fresnel_step(Photon& p, const StepState& s, Rng& rng)
{
    const Incidence in = meet_boundary(p, s);
    const bool s_polarised = rng_uniform(rng) < in.s_share;
    const float down = in.theta_i - in.theta_t, up = in.theta_i + in.theta_t;
    const float amplitude = s_polarised ? -sinf(down) / sinf(up) : tanf(down) / tanf(up);
    const bool reflected = (rng_uniform(rng) < amplitude * amplitude) || isnan(in.theta_t);
    if (reflected) {
        p.dir = rotate(s.normal, in.theta_i, in.s_axis);
        p.history |= CB_REFLECT_SPECULAR;
        p.pol = s_polarised ? in.s_axis : unit(cross(in.s_axis, p.dir));
    } else {
        refract(p, s, in, s_polarised);
    }
}

// ------------------------------------------------------------------ surface fates
// What a surface model decides.  FATE_THROUGH marks SURFACE_TRANSMIT and lets the dielectric
// boundary refract the photon (WLS, dichroic, angular); FATE_REFRACT is the thin film's own
// refraction; FATE_BOUNDARY is "nothing happened here" (the default model's fall-through).
enum Fate { FATE_ABSORB, FATE_DETECT, FATE_DIFFUSE, FATE_MIRROR, FATE_REEMIT, FATE_THROUGH, FATE_REFRACT, FATE_BOUNDARY,
            FATE_DETECT_WEIGHTED };

// mirror reflection (photon.h:634-646)
__device__ __forceinline__ void reflect_mirror(Photon& p, const StepState& s)
{
    const float theta = get_theta(s.normal, -p.dir);
    const float3 axis = unit(cross(p.dir, s.normal));
    p.dir = rotate(s.normal, theta, axis);
    p.history |= CB_REFLECT_SPECULAR;
}
// Lambertian reflection by rejection (photon.h:648-667)
__device__ __forceinline__ void reflect_lambertian(Photon& p, const StepState& s, Rng& rng)
{
    float cosine;
    do {
        p.dir = rng_sphere(rng);
        cosine = dot(p.dir, s.normal);
        if (cosine < 0.0f) { p.dir = -p.dir; cosine = -cosine; }       // into the hemisphere of the normal
    } while (!(rng_uniform(rng) < cosine));
    p.pol = unit(cross(rng_sphere(rng), p.dir));
    p.history |= CB_REFLECT_DIFFUSE;
}

__device__ __forceinline__ int apply_fate(const DevGeometry& g, const Tables& T, Photon& p, const StepState& s, Rng& rng,
                                          const CbSurface* surface, int fate, const Incidence* in, float detect_weight)
{
    switch (fate) {
    case FATE_ABSORB: p.history |= CB_SURFACE_ABSORB; return STEP_ENDS;
    case FATE_DETECT: p.history |= CB_SURFACE_DETECT; return STEP_ENDS;
    case FATE_DETECT_WEIGHTED: p.history |= CB_SURFACE_DETECT; p.weight *= detect_weight; return STEP_ENDS;
    case FATE_DIFFUSE: reflect_lambertian(p, s, rng); return STEP_CONTINUES;
    case FATE_MIRROR: reflect_mirror(p, s); return STEP_CONTINUES;
    case FATE_REEMIT:
        p.history |= CB_SURFACE_REEMIT;
        p.wavelength = sample_cdf_uniform(rng, g.wavelength_n, g.wavelength_start, g.wavelength_step, T.at(surface->reemission_cdf));
        emit_isotropic(p, rng);
        return STEP_CONTINUES;
    case FATE_THROUGH: p.history |= CB_SURFACE_TRANSMIT; return STEP_AT_BOUNDARY;
    case FATE_REFRACT:
        refract(p, s, *in, false);
        p.history |= CB_SURFACE_TRANSMIT;
        return STEP_CONTINUES;
    default: return STEP_AT_BOUNDARY;
    }
}

// survival biasing shared by the weighted modes: absorption is taken out of the lottery, the
// photon's weight carries the survival probability, the other shares are renormalised
__device__ __forceinline__ bool bias_survival(Photon& p, bool use_weights, float& absorb, float& a, float& b, float& c)
{
    if (!(use_weights && p.weight > CB_WEIGHT_LOWER_THRESHOLD && absorb < (1.0f - CB_WEIGHT_LOWER_THRESHOLD))) return false;
    const float survive = 1.0f - absorb;
    absorb = 0.0f;
    p.weight *= survive;
    a /= survive; b /= survive; c /= survive;
    return true;
}

// ---- thin film ----------------------------------------------------------------------
// Reflectance R and transmittance T of a film (index eta + i k, thickness d) between media n1
// and n3 at vacuum wavelength lambda, for s and p polarisation (the optics of propagate_complex,
// photon.h:669-755, there written with moduli and arguments).  With x_j = n_j cos_j and
// E = exp(2 i beta), beta = 2 pi d / lambda * n2 cos2:
//     r = (r12 + r23 E) / (1 + r12 r23 E),   R = |r|^2
//     T = Re(n3 cos3 / n1 cos1) |t12|^2 |t23|^2 |E| / |1 + r12 r23 E|^2
struct Cx { float re, im; };
__device__ __forceinline__ Cx cx(float re, float im = 0.0f) { Cx z = {re, im}; return z; }
__device__ __forceinline__ Cx operator+(Cx a, Cx b) { return cx(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ Cx operator-(Cx a, Cx b) { return cx(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ Cx operator*(Cx a, Cx b) { return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ float norm2(Cx a) { return a.re * a.re + a.im * a.im; }
__device__ __forceinline__ Cx operator/(Cx a, Cx b)
{
    const float inv = 1.0f / norm2(b);
    return cx((a.re * b.re + a.im * b.im) * inv, (a.im * b.re - a.re * b.im) * inv);
}
__device__ __forceinline__ Cx cx_sqrt(Cx a)                 // principal root
{
    const float r = sqrtf(sqrtf(norm2(a)));
    const float half = atan2f(a.im, a.re) / 2.0f;
    return cx(r * cosf(half), r * sinf(half));
}

struct FilmRT { float r_s, t_s, r_p, t_p; };

template <bool S_POL>
__device__ __forceinline__ void film_polarisation(Cx n1, Cx n2, Cx n3, Cx c1, Cx c2, Cx c3, Cx E, float absE, float flux,
                                                  float& R, float& Tr)
{
    // interface 1|2 and 2|3: s couples n_j cos_j on both sides, p swaps the cosines
    const Cx a12 = S_POL ? n1 * c1 : n2 * c1, b12 = S_POL ? n2 * c2 : n1 * c2;
    const Cx a23 = S_POL ? n2 * c2 : n3 * c2, b23 = S_POL ? n3 * c3 : n2 * c3;
    const Cx r12 = (a12 - b12) / (a12 + b12), r23 = (a23 - b23) / (a23 + b23);
    const Cx t12 = (cx(2.0f) * n1 * c1) / (a12 + b12), t23 = (cx(2.0f) * n2 * c2) / (a23 + b23);
    const float inv = 1.0f / norm2(cx(1.0f) + r12 * r23 * E);
    R = norm2(r12 + r23 * E) * inv;
    Tr = flux * norm2(t12) * norm2(t23) * absE * inv;
}

static __device__ __noinline__ FilmRT thin_film(float n1f, float eta, float k, float n3f, float cos_i, float sin_i,
                                                float phase)      // phase = 2 pi d / lambda
{
    const Cx n1 = cx(n1f), n2 = cx(eta, k), n3 = cx(n3f), c1 = cx(cos_i);
    const Cx s2 = cx(sin_i * sin_i);
    const Cx q13 = n1 / n3, q12 = n1 / n2;
    const Cx c3 = cx_sqrt(cx(1.0f) - q13 * q13 * s2);       // Snell in each medium
    const Cx c2 = cx_sqrt(cx(1.0f) - q12 * q12 * s2);
    const Cx x2 = n2 * c2;
    const float absE = expf(-2.0f * x2.im * phase);
    float se, ce;
    sincosf(2.0f * x2.re * phase, &se, &ce);
    const Cx E = cx(absE * ce, absE * se);
    const float flux = ((n3 * c3) / (n1 * c1)).re;
    FilmRT out;
    film_polarisation<true>(n1, n2, n3, c1, c2, c3, E, absE, flux, out.r_s, out.t_s);
    film_polarisation<false>(n1, n2, n3, c1, c2, c3, E, absE, flux, out.r_p, out.t_p);
    return out;
}

// thin-film surface (photocathode models), decisions of photon.h:757-827
__device__ __forceinline__ int film_fate(const DevGeometry& g, const Tables& T, Photon& p, const StepState& s, Rng& rng,
                                         const CbSurface* surface, bool use_weights, const Incidence& in, float& detect)
{
    detect = interp_property(g, p.wavelength, T.at(surface->detect));
    const float diffuse_share = interp_property(g, p.wavelength, T.at(surface->reflect_diffuse));
    const float eta = interp_property(g, p.wavelength, T.at(surface->eta));
    const float k = interp_property(g, p.wavelength, T.at(surface->k));
    const float theta = acosf(fabsf(dot(p.dir, s.normal)));
    const FilmRT f = thin_film(s.n1, eta, k, s.n2, cosf(theta), sinf(theta), 2.0f * CB_PI * surface->thickness / p.wavelength);

    float transmit = surface->transmissive ? in.s_share * f.t_s + (1.0f - in.s_share) * f.t_p : 0.0f;
    float reflect = in.s_share * f.r_s + (1.0f - in.s_share) * f.r_p;
    float absorb = 1.0f - transmit - reflect;
    bias_survival(p, use_weights, absorb, detect, reflect, transmit);
    if (use_weights && detect > 0.0f) return FATE_DETECT_WEIGHTED;

    const float u = rng_uniform(rng);
    if (u < absorb) return (rng_uniform(rng) < detect) ? FATE_DETECT : FATE_ABSORB;      // detection is conditional on absorption
    if (u < absorb + reflect || !surface->transmissive) return (rng_uniform(rng) < diffuse_share) ? FATE_DIFFUSE : FATE_MIRROR;
    return FATE_REFRACT;
}

// wavelength-shifting surface (photon.h:829-874)
__device__ __forceinline__ int wls_fate(const DevGeometry& g, const Tables& T, Photon& p, Rng& rng,
                                        const CbSurface* surface, bool use_weights)
{
    float absorb = interp_property(g, p.wavelength, T.at(surface->absorb));
    float mirror = interp_property(g, p.wavelength, T.at(surface->reflect_specular));
    float diffuse = interp_property(g, p.wavelength, T.at(surface->reflect_diffuse));
    const float reemit = interp_property(g, p.wavelength, T.at(surface->reemit));
    const float u = rng_uniform(rng);
    float unused = 0.0f;
    bias_survival(p, use_weights, absorb, diffuse, mirror, unused);
    if (u < absorb) return (rng_uniform(rng) < reemit) ? FATE_REEMIT : FATE_ABSORB;
    if (u < absorb + mirror + diffuse) return (rng_uniform(rng) * (mirror + diffuse) < mirror) ? FATE_MIRROR : FATE_DIFFUSE;
    return FATE_THROUGH;
}

// angle of incidence -> (bracket, weight) in a surface's angle table
struct AngleBin { unsigned lo; float w; };
__device__ __forceinline__ AngleBin angle_bin(const Photon& p, const StepState& s, int n, const float* angles)
{
    const float idx = interp_idx(get_theta(s.normal, -p.dir), n, angles);
    AngleBin b;
    b.lo = (int)idx;
    b.w = idx - b.lo;
    return b;
}

// dichroic filter: reflect / transmit tables per angle of incidence, blended between the two
// nearest angles (photon.h:877-907)
__device__ __forceinline__ int dichroic_fate(const DevGeometry& g, const Tables& T, const Photon& p, const StepState& s, Rng& rng,
                                             const CbSurface* surface)
{
    const AngleBin b = angle_bin(p, s, surface->dichroic_nangles, T.at(surface->dichroic_angles));
    const int W = g.wavelength_n;
    const float r_lo = interp_property(g, p.wavelength, T.at(surface->dichroic_reflect + b.lo * W));
    const float r_hi = interp_property(g, p.wavelength, T.at(surface->dichroic_reflect + (b.lo + 1) * W));
    const float t_lo = interp_property(g, p.wavelength, T.at(surface->dichroic_transmit + b.lo * W));
    const float t_hi = interp_property(g, p.wavelength, T.at(surface->dichroic_transmit + (b.lo + 1) * W));
    const float reflect = r_lo + (r_hi - r_lo) * b.w;
    const float transmit = t_lo + (t_hi - t_lo) * b.w;
    const float u = rng_uniform(rng);
    if (u < reflect) return FATE_MIRROR;
    if (u < transmit + reflect) return FATE_THROUGH;
    return FATE_ABSORB;
}

// angle-tabulated surface (photon.h:909-951)
__device__ __forceinline__ int angular_fate(const Tables& T, Photon& p, const StepState& s, Rng& rng,
                                            const CbSurface* surface, bool use_weights)
{
    const AngleBin b = angle_bin(p, s, surface->angular_nangles, T.at(surface->angular_angles));
    const float* tr = T.at(surface->angular_transmit);
    const float* rs = T.at(surface->angular_reflect_specular);
    const float* rd = T.at(surface->angular_reflect_diffuse);
    float transmit = tr[b.lo] + b.w * (tr[b.lo + 1] - tr[b.lo]);
    float mirror = rs[b.lo] + b.w * (rs[b.lo + 1] - rs[b.lo]);
    float diffuse = rd[b.lo] + b.w * (rd[b.lo + 1] - rd[b.lo]);
    float absorb = 1.0f - transmit - mirror - diffuse;
    bias_survival(p, use_weights, absorb, transmit, mirror, diffuse);
    const float u = rng_uniform(rng);
    if (u < absorb) return FATE_ABSORB;
    if (u < absorb + transmit) return FATE_THROUGH;
    if (u < absorb + transmit + mirror) return FATE_MIRROR;
    return FATE_DIFFUSE;
}

// default model: four tabulated shares, one draw (photon.h:953-1037; the reference's
// effective default is CHROMA_FORCE_SCATTER_AT_PASS == 0)
__device__ __forceinline__ int default_fate(const DevGeometry& g, const Tables& T, Photon& p, Rng& rng,
                                            const CbSurface* surface, bool use_weights, float& detect)
{
    detect = interp_property(g, p.wavelength, T.at(surface->detect));
    float absorb = interp_property(g, p.wavelength, T.at(surface->absorb));
    float diffuse = interp_property(g, p.wavelength, T.at(surface->reflect_diffuse));
    float mirror = interp_property(g, p.wavelength, T.at(surface->reflect_specular));
    const float u = rng_uniform(rng);
    bias_survival(p, use_weights, absorb, detect, diffuse, mirror);
    if (use_weights && detect > 0.0f) return FATE_DETECT_WEIGHTED;
    if (u < absorb) return FATE_ABSORB;
    if (u < absorb + detect) return FATE_DETECT;
    if (u < absorb + detect + diffuse) return FATE_DIFFUSE;
    if (u < absorb + detect + diffuse + mirror) return FATE_MIRROR;
    return FATE_BOUNDARY;
}

__device__ __forceinline__ int surface_step(const DevGeometry& g, const Tables& T, Photon& p, const StepState& s,
                                            Rng& rng, bool use_weights)
{
    const CbSurface* surface = &g.surfaces[s.surface_index];
    float detect = 0.0f;
    int fate;
    switch (surface->model) {
    case CB_SURFACE_COMPLEX: {
        const Incidence in = meet_boundary(p, s);
        fate = film_fate(g, T, p, s, rng, surface, use_weights, in, detect);
        return apply_fate(g, T, p, s, rng, surface, fate, &in, detect);
    }
    case CB_SURFACE_WLS: fate = wls_fate(g, T, p, rng, surface, use_weights); break;
    case CB_SURFACE_DICHROIC: fate = dichroic_fate(g, T, p, s, rng, surface); break;
    case CB_SURFACE_ANGULAR: fate = angular_fate(T, p, s, rng, surface, use_weights); break;
    default: fate = default_fate(g, T, p, rng, surface, use_weights, detect); break;
    }
    return apply_fate(g, T, p, s, rng, surface, fate, nullptr, detect);
}

// ------------------------------------------------------------------ analytic wire planes
// A wire plane is a periodic row of parallel cylinders (struct WirePlane,
// chroma/cuda/geometry_types.h:42-58).  Per plane the host prepares (cb_geometry_create,
// double precision, once): the orthonormal frame (U along the wires, V across them in the
// plane, N = U x V), the wire index range and the squared radius.  In that frame the problem
// is two-dimensional: the ray's (v, n) trace against circles of radius r at (k pitch, 0).
// Candidates compete through the distance alone (strictly nearer in float wins); which wires
// to look at follows the reference's windowing (photon.h:108-270) so that the same candidates
// are considered.
struct WireFrame {
    double U[3], V[3], N[3];
    double v0, pitch, inv_pitch, radius, radius2, umin, umax;
    float origin[3];
    int kmin, kmax;
    int surface, material_inner, material_outer;
};

struct Window {
    double lo, hi;
    __device__ __forceinline__ bool empty() const { return lo > hi; }
    __device__ __forceinline__ void clip(double a, double b) { lo = fmax(lo, fmin(a, b)); hi = fmin(hi, fmax(a, b)); }
};

struct WireHit {
    float distance;            // 1e30f: none
    int surface, material_inner, material_outer;
    float3 normal;             // outward cylinder normal at the hit (unoriented)
    float facing;              // normal . (-direction): > 0 when the photon arrives from outside the wire
};

__device__ __forceinline__ double dot3(const double a[3], const float3& b)
{
    return (double)b.x * a[0] + (double)b.y * a[1] + (double)b.z * a[2];
}

static __device__ __noinline__ WireHit nearest_wire(const WireFrame* __restrict__ frames, int nframes, const float3 pos,
                                                    const float3 dir, float mesh_distance)
{
    WireHit hit;
    hit.distance = 1e30f;
    hit.surface = -1; hit.material_inner = -1; hit.material_outer = -1;
    hit.normal = f3(0.0f, 0.0f, 0.0f);
    hit.facing = 0.0f;
    const double T_MIN = 1.0e-4;                 // never re-hit the wire the photon sits on
    for (int ip = 0; ip < nframes; ip++) {
        const WireFrame& w = frames[ip];
        // ray in the plane's frame: origin a, direction b
        const float3 rel = pos - f3(w.origin[0], w.origin[1], w.origin[2]);
        const double bu = dot3(w.U, dir), bv = dot3(w.V, dir), bn = dot3(w.N, dir);
        const double au = dot3(w.U, rel), av = dot3(w.V, rel) - w.v0, an = dot3(w.N, rel);

        // parameter window in which the ray is within the wires' length
        Window along = {-1.0e300, 1.0e300};
        if (fabs(bu) < 1e-15) {
            if (au < w.umin || au > w.umax) continue;
        } else {
            along.clip((w.umin - au) / bu, (w.umax - au) / bu);
            if (along.empty()) continue;
        }

        int k_first = w.kmin, k_last = w.kmax;
        if (w.kmin <= w.kmax) {
            // wires the ray can touch: the stretch of the ray inside the slab |n| <= r (+1e-6) around the plane
            const double pad = w.radius + 1e-6;
            Window reach = {fmax(along.lo, T_MIN), fmin(along.hi, (double)mesh_distance)};
            if (fabs(bn) > 1e-12) reach.clip((-pad - an) / bn, (pad - an) / bn);
            else if (fabs(an) > pad) continue;
            if (reach.hi < reach.lo) continue;
            if (fabs(bn) <= 1e-12 && fabs(bv) > 1e-12)            // in-plane ray: one pitch further at most
                reach.hi = fmin(reach.hi, reach.lo + (w.pitch + 2.0 * w.radius) / fabs(bv));
            const double v_a = av + bv * reach.lo, v_b = av + bv * reach.hi;
            const double v_lo = fmin(fmin(v_a, v_b), av) - pad, v_hi = fmax(fmax(v_a, v_b), av) + pad;
            const long long k_lo = max((long long)floor(v_lo * w.inv_pitch), (long long)w.kmin);
            const long long k_hi = min((long long)ceil(v_hi * w.inv_pitch), (long long)w.kmax);
            if (k_lo > k_hi) continue;
            k_first = (int)k_lo; k_last = (int)k_hi;
        }

        const double A = bv * bv + bn * bn;
        const double on_surface = fmax(1e-18, 1e-12 * w.radius2);
        for (int k = k_first; k <= k_last; k++) {
            const double cv = av - (double)k * w.pitch;            // ray origin relative to wire k
            const double B = cv * bv + an * bn;
            const double r0 = cv * cv + an * an;
            const double disc = B * B - A * (r0 - w.radius2);
            if (disc < 0.0) continue;
            const double root = sqrt(disc);
            double t;
            if (r0 > w.radius2 + on_surface) {            // from outside: where the ray enters
                t = (-B - root) / A;
                if (t <= T_MIN) continue;
            } else if (r0 < w.radius2 - on_surface) {     // from inside: where it leaves
                t = (-B + root) / A;
                if (t <= T_MIN) continue;
            } else {
                t = T_MIN;                                // sitting on the surface: a small step forward
            }
            const double u_hit = au + bu * t;
            if (u_hit < w.umin || u_hit > w.umax) continue;
            if ((float)t >= hit.distance) continue;
            if (t < along.lo || t > along.hi) continue;
            const double hv = cv + bv * t, hn = an + bn * t;      // hit point relative to the wire's axis
            const double len = sqrt(hv * hv + hn * hn);
            if (len <= 0.0) continue;
            const double ev = hv / len, en = hn / len;
            hit.normal = f3((float)(ev * w.V[0] + en * w.N[0]), (float)(ev * w.V[1] + en * w.N[1]),
                            (float)(ev * w.V[2] + en * w.N[2]));
            hit.distance = (float)t;
            hit.surface = w.surface;
            hit.material_inner = w.material_inner;
            hit.material_outer = w.material_outer;
            hit.facing = dot(hit.normal, -dir);
        }
    }
    return hit;
}

// The analytic candidate competes with the mesh hit (photon.h:272-330): it wins when it is
// nearer (in double, by more than 1e-12) and its plane has a surface; last_hit_triangle
// becomes -2 and the media follow the side the photon comes from.
__device__ __forceinline__ bool wire_boundary(const DevGeometry& g, const Tables& T, Photon& p, StepState& s,
                                              int tri, float distance)
{
    const float mesh = (tri == -1) ? 1e30f : distance;
    const WireHit wh = nearest_wire(g.wireframes, g.nwireplanes, p.pos, p.dir, mesh);
    if (!(wh.surface >= 0 && (double)wh.distance + 1e-12 < (double)mesh)) return false;
    s.distance = wh.distance;
    s.surface_index = wh.surface;
    p.last_hit_triangle = -2;
    const bool from_outside = wh.facing > 0.0f;
    s.normal = from_outside ? wh.normal : -wh.normal;
    enter_media(g, T, p.wavelength, &g.materials[from_outside ? wh.material_outer : wh.material_inner],
                &g.materials[from_outside ? wh.material_inner : wh.material_outer], s);
    return true;
}

// ------------------------------------------------------------------ one step
// everything after the intersection (propagate.cu:312-336).  Returns true when the photon
// continues to another step.  WIRES: geometry with analytic wire planes (a separate
// instantiation, so that the cold path costs the usual kernels neither registers nor instructions).
template <bool WIRES>
__device__ __forceinline__ bool physics_step(const DevGeometry& g, const Tables& T, Photon& p, Rng& rng,
                                             int tri, float distance, bool use_weights, int scatter_first)
{
    StepState s;
    const bool on_wire = WIRES && g.nwireplanes > 0 && wire_boundary(g, T, p, s, tri, distance);
    if (!on_wire) {
        if (tri == -1) {
            p.last_hit_triangle = -1;
            p.history |= CB_NO_HIT;
            return false;
        }
        s.distance = distance;
        classify_hit(g, T, p, s, tri);
    }
    int step = bulk_step(g, T, p, s, rng, use_weights, scatter_first);
    if (step == STEP_AT_BOUNDARY && s.surface_index != -1) step = surface_step(g, T, p, s, rng, use_weights);
    if (step == STEP_AT_BOUNDARY) {
        fresnel_step(p, s, rng);
        return true;
    }
    return step == STEP_CONTINUES;
}

__device__ __forceinline__ bool photon_is_nan(const Photon& p)
{
    return isnan(p.dir.x * p.dir.y * p.dir.z * p.pos.x * p.pos.y * p.pos.z);
}

} // namespace cb
