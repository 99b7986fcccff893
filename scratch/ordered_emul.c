// scratch tool (not product, not oracle): CPU emulation of the engine's ordered traversal to
// count dependent rounds / entries / triangle tests for a given tree.
#include "../oracle/chroma_oracle.c"

typedef struct { uint64_t rounds, entries, tris, rays; } emu_counters;

ORC_EXPORT void emu_intersect(const CbGeometryDesc *g, const float *origins, const float *dirs, uint64_t n,
                              int32_t *tri_out, float *dist_out, uint64_t *counters, int sort_all)
{
    emu_counters c = {0,0,0,0};
    for (uint64_t i = 0; i < n; i++) {
        f3 o = mk(origins[3*i], origins[3*i+1], origins[3*i+2]);
        f3 d = mk(dirs[3*i], dirs[3*i+1], dirs[3*i+2]);
        d = divs(d, norm(d));
        f3 noid = mk(-o.x/d.x, -o.y/d.y, -o.z/d.z), inv = mk(1.0f/d.x, 1.0f/d.y, 1.0f/d.z);
        Node root = get_node(g, 0);
        float best = INFINITY; int best_tri = -1; float tb;
        c.rays++;
        if (!intersect_box(noid, inv, root.lower, root.upper, &tb)) { tri_out[i] = -1; continue; }
        uint32_t sw[256]; float st[256]; int sp = 0;
        uint32_t cur = g->nodes[3];
        for (;;) {
            uint32_t first = cur & 0x0FFFFFFF, k = cur >> 28;
            c.rounds++;
            uint32_t hw[16]; float ht[16]; int nh = 0;
            for (uint32_t j = first; j < first + k; j++) {
                Node nd = get_node(g, j); c.entries++;
                float tmin;
                if (intersect_box(noid, inv, nd.lower, nd.upper, &tmin) && !(tmin > best)) {
                    if (nd.nchild == 0) {
                        c.tris++;
                        const uint32_t *t = g->triangles + 3ull*nd.child; float dist;
                        if (intersect_triangle(o, d, vtx(g,t[0]), vtx(g,t[1]), vtx(g,t[2]), &dist) && dist < best) { best = dist; best_tri = nd.child; }
                    } else { hw[nh] = g->nodes[4ull*j+3]; ht[nh] = tmin; nh++; }
                }
            }
            if (sort_all) { // push far -> near
                for (int a = 0; a < nh; a++) for (int b = a+1; b < nh; b++) if (ht[b] > ht[a]) { float tt=ht[a]; ht[a]=ht[b]; ht[b]=tt; uint32_t ww=hw[a]; hw[a]=hw[b]; hw[b]=ww; }
                for (int a = 0; a < nh; a++) { sw[sp]=hw[a]; st[sp]=ht[a]; sp++; }
            } else {      // engine v1: keep nearest on top only
                for (int a = 0; a < nh; a++) {
                    if (sp > 0 && ht[a] > st[sp-1]) { sw[sp]=sw[sp-1]; st[sp]=st[sp-1]; sw[sp-1]=hw[a]; st[sp-1]=ht[a]; sp++; }
                    else { sw[sp]=hw[a]; st[sp]=ht[a]; sp++; }
                }
            }
            int found = 0;
            while (sp > 0) { sp--; if (!(st[sp] > best)) { cur = sw[sp]; found = 1; break; } }
            if (!found) break;
        }
        tri_out[i] = best_tri; if (best_tri >= 0) dist_out[i] = best;
    }
    counters[0]=c.rounds; counters[1]=c.entries; counters[2]=c.tris; counters[3]=c.rays;
}
