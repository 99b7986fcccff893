// scratch tool: like ordered_emul.c, plus a split of the triangle tests by solid class
// (solid 0 = liner vs PMTs) and of the winning triangle.
#include "../oracle/chroma_oracle.c"
static float g_widen = 0.0f;   /* world units added to every box side (the engine's plane test is widened by ~1.9 mm) */
ORC_EXPORT void emu2_set_widen(float w) { g_widen = w; }
static float *g_dist_out = NULL;   /* optional: nearest-hit distance per ray */
ORC_EXPORT void emu2_set_dist_out(float *p) { g_dist_out = p; }
static int g_phased = 0;      /* 1: the engine's plane test (phased_ray_axis / hit_box_phased of engine.cuh) on the packed boxes */
ORC_EXPORT void emu2_set_phased(int m) { g_phased = m; }
typedef struct { float s[3], n[3], f[3]; int pos[3]; } phased_ray;
static void phased_axis(float o, float d, float worigin, float wscale, float *s, float *n, float *f, int *pos)
{
    const float inv = 1.0f / d;
    if (isfinite(inv)) {
        *s = wscale * inv;
        const float a = (worigin - o) * inv, b = 8388608.0f * *s;
        const float off = a - b;
        const float c = (fabsf(worigin) + fabsf(o) + 65536.0f * wscale) * fabsf(inv);
        const float e = 6e-7f * c + 3e-7f * fabsf(b);
        *n = off - e; *f = off + e; *pos = inv >= 0.0f;
    } else { *s = 0.0f; *n = -INFINITY; *f = INFINITY; *pos = 1; }
}
static int phased_box(const phased_ray *r, const uint32_t *w, float *tnear)
{
    float tn[3], tf[3];
    for (int a = 0; a < 3; a++) {
        const float lo = 8388608.0f + (float)(w[a] & 0xFFFFu), hi = 8388608.0f + (float)(w[a] >> 16);
        tn[a] = fmaf(r->pos[a] ? lo : hi, r->s[a], r->n[a]);
        tf[a] = fmaf(r->pos[a] ? hi : lo, r->s[a], r->f[a]);
    }
    const float tmin = fmaxf(fmaxf(tn[0], tn[1]), fmaxf(tn[2], 0.0f));
    const float tmax = fminf(fminf(tf[0], tf[1]), tf[2]);
    *tnear = tmin;
    return !(tmin > tmax);
}
static uint64_t g_occ_hist[513]; static int g_occ_max = 0;   /* stack entries held after a node expansion (nearest child taken directly) */
ORC_EXPORT int emu2_get_occupancy(uint64_t *hist) { memcpy(hist, g_occ_hist, sizeof(g_occ_hist)); return g_occ_max; }
ORC_EXPORT void emu2_reset_occupancy(void) { memset(g_occ_hist, 0, sizeof(g_occ_hist)); g_occ_max = 0; }
static int g_leaf_mode = 0;   /* 0: test leaves as found; 1: per node, nearest box first, re-checked against the best hit; 2: LIFO queue re-checked (the kernel) */
ORC_EXPORT void emu2_set_leaf_mode(int m) { g_leaf_mode = m; }
static int64_t g_dbg = -1; static float g_rec[4096]; static int g_nrec = 0;
ORC_EXPORT void emu2_set_debug(int64_t ray) { g_dbg = ray; g_nrec = 0; }
ORC_EXPORT int emu2_get_debug(float *out) { memcpy(out, g_rec, sizeof(float) * 4 * g_nrec); return g_nrec; }
ORC_EXPORT void emu2_intersect(const CbGeometryDesc *g, const uint32_t *solid_of, const float *origins, const float *dirs, uint64_t n,
                              int32_t *tri_out, uint64_t *counters, uint16_t *per_ray)
{
    uint64_t rounds=0, entries=0, tris0=0, tris1=0, halves=0, win0=0, win1=0, inner_hits=0;
    for (uint64_t i = 0; i < n; i++) {
        f3 o = mk(origins[3*i], origins[3*i+1], origins[3*i+2]);
        f3 d = mk(dirs[3*i], dirs[3*i+1], dirs[3*i+2]);
        d = divs(d, norm(d));
        f3 noid = mk(-o.x/d.x, -o.y/d.y, -o.z/d.z), inv = mk(1.0f/d.x, 1.0f/d.y, 1.0f/d.z);
        phased_ray pr_;
        phased_axis(o.x, d.x, g->world_origin[0], g->world_scale, &pr_.s[0], &pr_.n[0], &pr_.f[0], &pr_.pos[0]);
        phased_axis(o.y, d.y, g->world_origin[1], g->world_scale, &pr_.s[1], &pr_.n[1], &pr_.f[1], &pr_.pos[1]);
        phased_axis(o.z, d.z, g->world_origin[2], g->world_scale, &pr_.s[2], &pr_.n[2], &pr_.f[2], &pr_.pos[2]);
        Node root = get_node(g, 0);
        float best = INFINITY; int best_tri = -1; float tb; uint16_t pr0=0, pr1=0, prr=0;
        if (!intersect_box(noid, inv, root.lower, root.upper, &tb)) { tri_out[i] = -1; continue; }
        uint32_t sw[512]; float st[512]; int sp = 0;
        uint32_t cur = g->nodes[3];
        for (;;) {
            uint32_t first = cur & 0x0FFFFFFF, k = cur >> 28;
            rounds++; prr++; halves += (k + 3) / 4;
            uint32_t hw[16]; float ht[16]; int nh = 0;
            uint32_t lq_tri[16]; float lq_t[16]; int nl = 0;
            for (uint32_t j = first; j < first + k; j++) {
                Node nd = get_node(g, j); entries++;
                nd.lower.x -= g_widen; nd.lower.y -= g_widen; nd.lower.z -= g_widen; nd.upper.x += g_widen; nd.upper.y += g_widen; nd.upper.z += g_widen;
                float tmin;
                const float limit_ = g_phased ? best + 2e-5f * best : best;
                if ((g_phased ? phased_box(&pr_, g->nodes + 4ull * j, &tmin) : intersect_box(noid, inv, nd.lower, nd.upper, &tmin)) && !(tmin > limit_)) {
                    if (nd.nchild == 0) {
                        lq_tri[nl] = nd.child; lq_t[nl] = tmin; nl++;
                    } else { hw[nh] = g->nodes[4ull*j+3]; ht[nh] = tmin; nh++; inner_hits++; }
                }
            }
            if (g_leaf_mode == 1) for (int a = 0; a < nl; a++) for (int b = a+1; b < nl; b++) if (lq_t[b] < lq_t[a]) { float tt=lq_t[a]; lq_t[a]=lq_t[b]; lq_t[b]=tt; uint32_t ww=lq_tri[a]; lq_tri[a]=lq_tri[b]; lq_tri[b]=ww; }
            for (int q = 0; q < nl; q++) {
                int a = (g_leaf_mode == 2) ? nl - 1 - q : q;
                if (g_leaf_mode != 0 && lq_t[a] > (g_phased ? best + 2e-5f * best : best)) continue;
                uint32_t tri_id = lq_tri[a];
                if (solid_of[tri_id] == 0) { tris0++; pr0++; } else { tris1++; pr1++; }
                const uint32_t *t = g->triangles + 3ull*tri_id; float dist;
                int hit = intersect_triangle(o, d, vtx(g,t[0]), vtx(g,t[1]), vtx(g,t[2]), &dist);
                if (hit && dist < best) { best = dist; best_tri = tri_id; }
            }
            for (int a = 0; a < nh; a++) for (int b = a+1; b < nh; b++) if (ht[b] > ht[a]) { float tt=ht[a]; ht[a]=ht[b]; ht[b]=tt; uint32_t ww=hw[a]; hw[a]=hw[b]; hw[b]=ww; }
            for (int a = 0; a < nh; a++) { sw[sp]=hw[a]; st[sp]=ht[a]; sp++; }
            { int occ = sp > 0 ? sp - 1 : 0; if (occ > 512) occ = 512; g_occ_hist[occ]++; if (occ > g_occ_max) g_occ_max = occ; }
            int found = 0;
            while (sp > 0) { sp--; if (!(st[sp] > (g_phased ? best + 2e-5f * best : best))) { cur = sw[sp]; found = 1; break; } }
            if (!found) break;
        }
        if (g_dist_out) g_dist_out[i] = best_tri >= 0 ? best : -1.0f;
        tri_out[i] = best_tri; per_ray[3*i]=pr0; per_ray[3*i+1]=pr1; per_ray[3*i+2]=prr;
        if (best_tri >= 0) { if (solid_of[best_tri] == 0) win0++; else win1++; }
    }
    counters[0]=rounds; counters[1]=entries; counters[2]=tris0; counters[3]=tris1; counters[4]=halves; counters[5]=win0; counters[6]=win1; counters[7]=inner_hits;
}
