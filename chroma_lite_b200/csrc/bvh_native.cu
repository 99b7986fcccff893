// bvh_native.cu -- the engine's own traversal tree.
//
// The reference tree (Morton grid, mean fan-out ~4, chroma/bvh/grid.py) is an
// INPUT: it defines the leaf boxes and the tie-break rank.  For traversal the
// engine builds its own hierarchy over the same leaves:
//   * top-down binned SAH, collapsed on the fly to <= 8 children per node
//     (always split the sub-range with the largest surface area next),
//   * two-level when the mesh comes as many solids (29k PMTs): a SAH tree over
//     solids, then one independent SAH subtree per solid, built in parallel on
//     the host cores,
//   * same 16-byte entry format as the reference (6 x uint16 box on the world
//     grid + nchild<<28|child), children of a node contiguous and 64/128-byte
//     aligned so a node's children arrive in one or two 128-bit-vector bursts,
//   * leaf entries are the reference's leaf nodes verbatim, which is what makes
//     the order-independence argument of traverse() exact,
//   * breadth-first storage: the top of the tree is one contiguous prefix (L2
//     persistence window).
#include "host.h"
#include <algorithm>
#include <atomic>
#include <thread>
#include <string.h>
#include <math.h>

namespace cb {

struct Entry { uint32_t x, y, z, w; };   // reference node packing

static inline void box_of(const Entry& e, int lo[3], int hi[3])
{
    lo[0] = e.x & 0xFFFF; hi[0] = e.x >> 16;
    lo[1] = e.y & 0xFFFF; hi[1] = e.y >> 16;
    lo[2] = e.z & 0xFFFF; hi[2] = e.z >> 16;
}
struct IBox {
    int lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; a++) { lo[a] = 0x7fffffff; hi[a] = -1; } }
    void grow(const int l[3], const int h[3]) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]); } }
    void grow(const IBox& b) { grow(b.lo, b.hi); }
    double area() const
    {
        if (hi[0] < lo[0]) return 0.0;
        double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
};

struct Range { uint64_t b, e; IBox box; };

constexpr int NBINS = 16;
constexpr int WIDE = 8;

// run fn(t, lo, hi) over [b,e) cut into nthreads contiguous chunks
template <class F>
static void parallel_chunks(uint64_t b, uint64_t e, unsigned nthreads, F fn)
{
    if (nthreads <= 1 || e - b < (1u << 16)) { fn(0u, b, e); return; }
    std::vector<std::thread> pool;
    const uint64_t n = e - b;
    for (unsigned t = 1; t < nthreads; t++) pool.emplace_back(fn, t, b + n * t / nthreads, b + n * (t + 1) / nthreads);
    fn(0u, b, b + n / nthreads);
    for (auto& th : pool) th.join();
}

// Binned SAH split of prims[b,e) (partitioned in place).  Returns the split point.
// nthreads > 1: the passes over a large range run on that many host threads (the top of a
// single-level build over tens of millions of leaves).
static uint64_t sah_split(Entry* prims, uint64_t b, uint64_t e, IBox& left, IBox& right, unsigned nthreads = 1)
{
    const uint64_t n = e - b;
    if (n < (1u << 18)) nthreads = 1;
    const unsigned T = std::max(1u, nthreads);
    int clo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, chi[3] = {-1, -1, -1};
    {
        std::vector<int> plo(3 * T, 0x7fffffff), phi(3 * T, -1);
        parallel_chunks(b, e, T, [&](unsigned t, uint64_t lo_i, uint64_t hi_i) {
            int l3[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, h3[3] = {-1, -1, -1};
            for (uint64_t i = lo_i; i < hi_i; i++) {
                int lo[3], hi[3];
                box_of(prims[i], lo, hi);
                for (int a = 0; a < 3; a++) { int c = lo[a] + hi[a]; l3[a] = std::min(l3[a], c); h3[a] = std::max(h3[a], c); }
            }
            for (int a = 0; a < 3; a++) { plo[3 * t + a] = l3[a]; phi[3 * t + a] = h3[a]; }
        });
        for (unsigned t = 0; t < T; t++)
            for (int a = 0; a < 3; a++) { clo[a] = std::min(clo[a], plo[3 * t + a]); chi[a] = std::max(chi[a], phi[3 * t + a]); }
    }
    // one pass bins all three axes
    struct Bins { IBox bb[3][NBINS]; uint64_t cnt[3][NBINS]; };
    double scale3[3];
    for (int a = 0; a < 3; a++) scale3[a] = (double)NBINS / ((double)(chi[a] - clo[a]) + 1.0);
    std::vector<Bins> bins(T);
    for (auto& B : bins)
        for (int a = 0; a < 3; a++) for (int k = 0; k < NBINS; k++) { B.bb[a][k].reset(); B.cnt[a][k] = 0; }
    parallel_chunks(b, e, T, [&](unsigned t, uint64_t lo_i, uint64_t hi_i) {
        Bins& B = bins[t];
        for (uint64_t i = lo_i; i < hi_i; i++) {
            int lo[3], hi[3];
            box_of(prims[i], lo, hi);
            for (int a = 0; a < 3; a++) {
                int k = (int)((double)(lo[a] + hi[a] - clo[a]) * scale3[a]);
                k = std::min(std::max(k, 0), NBINS - 1);
                B.bb[a][k].grow(lo, hi);
                B.cnt[a][k]++;
            }
        }
    });
    for (unsigned t = 1; t < T; t++)
        for (int a = 0; a < 3; a++) for (int k = 0; k < NBINS; k++) { bins[0].bb[a][k].grow(bins[t].bb[a][k]); bins[0].cnt[a][k] += bins[t].cnt[a][k]; }
    double best = 1e300;
    int best_axis = -1, best_bin = -1;
    for (int a = 0; a < 3; a++) {
        if (chi[a] == clo[a]) continue;
        const IBox* bb = bins[0].bb[a];
        const uint64_t* cnt = bins[0].cnt[a];
        double ra[NBINS];
        uint64_t rc[NBINS];
        IBox acc; acc.reset();
        uint64_t c = 0;
        for (int k = NBINS - 1; k > 0; k--) { acc.grow(bb[k]); c += cnt[k]; ra[k] = acc.area(); rc[k] = c; }
        acc.reset(); c = 0;
        for (int k = 0; k < NBINS - 1; k++) {
            acc.grow(bb[k]); c += cnt[k];
            if (c == 0 || rc[k + 1] == 0) continue;
            double cost = acc.area() * (double)c + ra[k + 1] * (double)rc[k + 1];
            if (cost < best) { best = cost; best_axis = a; best_bin = k; }
        }
    }
    uint64_t mid;
    if (best_axis < 0) {
        mid = b + n / 2;                              // all centroids coincide: split by index
    } else {
        const int a = best_axis;
        const double scale = scale3[a];
        Entry* m = std::partition(prims + b, prims + e, [&](const Entry& p) {
            int lo[3], hi[3];
            box_of(p, lo, hi);
            int k = (int)((double)(lo[a] + hi[a] - clo[a]) * scale);
            k = std::min(std::max(k, 0), NBINS - 1);
            return k <= best_bin;
        });
        mid = (uint64_t)(m - prims);
        if (mid == b || mid == e) mid = b + n / 2;
    }
    auto bound = [&](uint64_t lo_i, uint64_t hi_i, IBox& out) {
        std::vector<IBox> part(T);
        for (auto& bx : part) bx.reset();
        parallel_chunks(lo_i, hi_i, T, [&](unsigned t, uint64_t x, uint64_t y) {
            IBox bx; bx.reset();
            for (uint64_t i = x; i < y; i++) { int lo[3], hi[3]; box_of(prims[i], lo, hi); bx.grow(lo, hi); }
            part[t] = bx;
        });
        out.reset();
        for (unsigned t = 0; t < T; t++) if (part[t].hi[0] >= part[t].lo[0]) out.grow(part[t]);   // skip idle chunks
    };
    bound(b, mid, left);
    bound(mid, e, right);
    return mid;
}

struct Arena { uint64_t cur = 0, end = 0; };   // per-thread bump allocator over the node pool
constexpr uint64_t ARENA_CHUNK = 1 << 16;

struct Builder {
    Entry* prims;                     // leaf entries (reordered in place)
    std::vector<Entry>* out;          // wide nodes (pre-sized)
    std::atomic<uint64_t> next{8};    // entry 0 is the root
    std::atomic<int> overflow{0};

    uint64_t alloc(Arena& A, int k)
    {
        // 8-entry (128 B line) alignment for wide blocks, 2-entry (32 B sector) for small ones
        const uint64_t align = (k > 4) ? 8 : 2;
        uint64_t start = (A.cur + align - 1) / align * align;
        if (start + k > A.end) {
            A.cur = next.fetch_add(ARENA_CHUNK);
            A.end = A.cur + ARENA_CHUNK;
            start = A.cur;                         // chunk starts are multiples of 8
        }
        if (start + k > out->size()) { overflow = 1; return 0; }
        A.cur = start + k;
        return start;
    }

    static Entry pack(const IBox& b, uint32_t w)
    {
        Entry e;
        e.x = (uint32_t)b.lo[0] | ((uint32_t)b.hi[0] << 16);
        e.y = (uint32_t)b.lo[1] | ((uint32_t)b.hi[1] << 16);
        e.z = (uint32_t)b.lo[2] | ((uint32_t)b.hi[2] << 16);
        e.w = w;
        return e;
    }

    // split [b,e) into <= WIDE sub-ranges, always refining the largest-area one
    int widen(uint64_t b, uint64_t e, const IBox& box, Range* r, unsigned nthreads = 1)
    {
        int k = 1;
        r[0].b = b; r[0].e = e; r[0].box = box;
        while (k < WIDE) {
            int pick = -1;
            double pa = -1.0;
            for (int i = 0; i < k; i++)
                if (r[i].e - r[i].b > 1) {
                    double a = r[i].box.area() + 1e-9 * (double)(r[i].e - r[i].b);
                    if (a > pa) { pa = a; pick = i; }
                }
            if (pick < 0) break;
            IBox L, R;
            uint64_t mid = sah_split(prims, r[pick].b, r[pick].e, L, R, nthreads);
            Range right; right.b = mid; right.e = r[pick].e; right.box = R;
            r[pick].e = mid; r[pick].box = L;
            for (int i = k; i > pick + 1; i--) r[i] = r[i - 1];     // keep spatial order
            r[pick + 1] = right;
            k++;
        }
        return k;
    }

    void build(Arena& A, uint64_t b, uint64_t e, const IBox& box, uint64_t slot)
    {
        if (e - b == 1) { (*out)[slot] = prims[b]; return; }
        Range r[WIDE + 1];
        int k = widen(b, e, box, r);
        uint64_t blk = alloc(A, k);
        if (overflow) return;
        (*out)[slot] = pack(box, ((uint32_t)k << 28) | (uint32_t)blk);
        for (int i = 0; i < k; i++) build(A, r[i].b, r[i].e, r[i].box, blk + i);
    }
};

// ---- optional leaf splitting ("early split clipping") -------------------------------------
// A long triangle that lies obliquely to the axes has a box many times its own size (a PMT
// lathed with 6 steps: every ray that enters a PMT's box tests ~27 triangles, a ray that
// grazes one without hitting anything ~66).  With SplitInput::max_pieces > 1 such a leaf is
// replaced by up to max_pieces leaves of the SAME triangle, each bounding the part of the
// triangle inside one cell of a recursive midpoint subdivision of the leaf box (the triangle
// is clipped in double precision on the world grid, boxes are rounded outward with the
// reference's one-quantum padding, bvh.cu:148-203, and clamped to the reference leaf box).
// What this keeps and what it does not.  Geometrically nothing is lost: every point of the triangle lies
// in one of the pieces' boxes, a triangle tested twice returns the same (distance, rank), and finish()
// keeps checking the reference leaf box, which comes from the tri64 record, not from the tree.  But the
// engine's contract is the REFERENCE's answer, float32 quirks included, and there splitting is not exact:
// for a ray that grazes a sliver the reference's float32 Moeller-Trumbore can report a hit several mm away
// from the triangle (u, v and t lose their digits when the ray is nearly parallel to the plane); the
// reference finds that hit because the ray is inside the sliver's big leaf box, a split tree never tests
// the triangle there (or prunes the piece behind a nearer true hit).  Emulation (scratch/emu_phased_check.py):
// 0 of 2 M random rays, 1 of 40,000 rays aimed at triangle corners / edge midpoints.  Hence off by default.
struct SplitInput {
    const float* vertices;        // world-space, 3 floats per vertex
    const uint32_t* triangles;    // 3 indices per triangle
    float origin[3], scale;       // world grid (WorldCoords)
    int max_pieces;               // <= 1: no splitting
    int min_extent;               // boxes whose longest side is shorter (grid quanta) are kept whole
    double min_ratio;             // split only while box area > min_ratio * (2 x triangle-part area + rim)
};

constexpr int POLY_MAX = 16;   // a triangle clipped to a box has <= 9 corners; points on a plane can repeat
struct Poly { int n; double p[POLY_MAX][3]; };

// false when the result does not fit (the caller then keeps the piece whole)
static bool clip_poly(const Poly& in, int axis, double plane, bool keep_below, Poly& out)
{
    out.n = 0;
    for (int i = 0; i < in.n; i++) {
        const double* a = in.p[i];
        const double* b = in.p[(i + 1) % in.n];
        const bool ina = keep_below ? a[axis] <= plane : a[axis] >= plane;
        const bool inb = keep_below ? b[axis] <= plane : b[axis] >= plane;
        if (ina) {
            if (out.n == POLY_MAX) return false;
            memcpy(out.p[out.n++], a, sizeof(double) * 3);
        }
        if (ina != inb) {
            if (out.n == POLY_MAX) return false;
            const double t = (plane - a[axis]) / (b[axis] - a[axis]);
            double* q = out.p[out.n++];
            for (int k = 0; k < 3; k++) q[k] = a[k] + t * (b[k] - a[k]);
            q[axis] = plane;
        }
    }
    return true;
}

static double poly_area(const Poly& P)
{
    double ax = 0, ay = 0, az = 0;
    for (int i = 1; i + 1 < P.n; i++) {
        const double u[3] = {P.p[i][0] - P.p[0][0], P.p[i][1] - P.p[0][1], P.p[i][2] - P.p[0][2]};
        const double v[3] = {P.p[i + 1][0] - P.p[0][0], P.p[i + 1][1] - P.p[0][1], P.p[i + 1][2] - P.p[0][2]};
        ax += u[1] * v[2] - u[2] * v[1]; ay += u[2] * v[0] - u[0] * v[2]; az += u[0] * v[1] - u[1] * v[0];
    }
    return 0.5 * sqrt(ax * ax + ay * ay + az * az);
}

// padded grid box of a polygon, clamped to `within`
static IBox poly_box(const Poly& P, const IBox& within)
{
    IBox b;
    for (int a = 0; a < 3; a++) {
        double lo = P.p[0][a], hi = P.p[0][a];
        for (int i = 1; i < P.n; i++) { lo = std::min(lo, P.p[i][a]); hi = std::max(hi, P.p[i][a]); }
        b.lo[a] = std::max(within.lo[a], (int)floor(lo) - 1);
        b.hi[a] = std::min(within.hi[a], (int)floor(hi) + 1);
        if (b.hi[a] < b.lo[a]) b.hi[a] = b.lo[a];
    }
    return b;
}

// Surface measure of a leaf box as the traversal sees it: the engine's plane test widens every slab by
// ~2.5 grid quanta per side (phased_ray_axis), so a box that is a few quanta thin does not get thinner.
constexpr double SLAB_SLACK = 5.0;
static double seen_area(const IBox& b)
{
    const double dx = b.hi[0] - b.lo[0] + SLAB_SLACK, dy = b.hi[1] - b.lo[1] + SLAB_SLACK, dz = b.hi[2] - b.lo[2] + SLAB_SLACK;
    return dx * dy + dy * dz + dz * dx;
}

struct Piece { Poly poly; IBox box; double area; bool done; };

// pieces of one leaf (appended to out); returns the number of pieces
static int split_leaf(const Entry& leaf, const SplitInput& S, std::vector<Entry>& out)
{
    IBox L;
    box_of(leaf, L.lo, L.hi);
    const uint32_t tri = leaf.w & 0x0FFFFFFFu;
    Piece pc[32];
    int np = 1;
    const int maxp = std::min(S.max_pieces, 32);
    for (int v = 0; v < 3; v++) {
        const float* x = S.vertices + 3ull * S.triangles[3ull * tri + v];
        for (int a = 0; a < 3; a++) pc[0].poly.p[v][a] = ((double)x[a] - (double)S.origin[a]) / (double)S.scale;
    }
    for (int v = 0; v < 3; v++)
        for (int a = 0; a < 3; a++) {
            const double q = pc[0].poly.p[v][a];
            // a corner off the grid (or not a number) cannot be clipped on it: keep the leaf whole
            if (!(q >= -2.0 && q <= 65538.0)) { out.push_back(leaf); return 1; }
        }
    pc[0].poly.n = 3;
    pc[0].box = L;
    pc[0].area = poly_area(pc[0].poly);
    pc[0].done = false;
    while (np < maxp) {
        // refine the piece with the largest box that is still worth splitting
        int pick = -1;
        double pa = 0.0;
        for (int i = 0; i < np; i++) {
            if (pc[i].done) continue;
            const IBox& b = pc[i].box;
            const int ex = std::max(b.hi[0] - b.lo[0], std::max(b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]));
            // a flat axis-aligned part has a box of 2 x its area + a rim from the padding
            const double rim = (2.0 + SLAB_SLACK) * 2.0 * ((b.hi[0] - b.lo[0]) + (b.hi[1] - b.lo[1]) + (b.hi[2] - b.lo[2]));
            const double tight = 2.0 * pc[i].area + rim;
            if (ex < std::max(S.min_extent, 4) || 2.0 * seen_area(b) <= S.min_ratio * tight) { pc[i].done = true; continue; }
            if (seen_area(b) > pa) { pa = seen_area(b); pick = i; }
        }
        if (pick < 0) break;
        Piece& P = pc[pick];
        int axis = 0;
        for (int a = 1; a < 3; a++) if (P.box.hi[a] - P.box.lo[a] > P.box.hi[axis] - P.box.lo[axis]) axis = a;
        const double plane = (double)((P.box.lo[axis] + P.box.hi[axis]) / 2);
        Piece A, B;
        if (!clip_poly(P.poly, axis, plane, true, A.poly) || !clip_poly(P.poly, axis, plane, false, B.poly) ||
            A.poly.n < 3 || B.poly.n < 3) { P.done = true; continue; }
        A.box = poly_box(A.poly, P.box); B.box = poly_box(B.poly, P.box);
        A.area = poly_area(A.poly); B.area = poly_area(B.poly);
        A.done = B.done = false;
        if (seen_area(A.box) + seen_area(B.box) > 0.8 * seen_area(P.box)) { P.done = true; continue; }   // no gain
        pc[pick] = A;
        pc[np++] = B;
    }
    for (int i = 0; i < np; i++) out.push_back(Builder::pack(pc[i].box, leaf.w));
    return np;
}

// Replace `leaves` by their pieces (parallel over host threads, order preserved).
static void presplit_leaves(std::vector<Entry>& leaves, const SplitInput& S)
{
    const uint64_t n = leaves.size();
    unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (n < 4096) nthreads = 1;
    std::vector<std::vector<Entry>> part(nthreads);
    auto worker = [&](unsigned t) {
        const uint64_t b = n * t / nthreads, e = n * (t + 1) / nthreads;
        part[t].reserve((e - b) * 2);
        for (uint64_t i = b; i < e; i++) split_leaf(leaves[i], S, part[t]);
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; t++) pool.emplace_back(worker, t);
    worker(0);
    for (auto& t : pool) t.join();
    uint64_t total = 0;
    for (auto& p : part) total += p.size();
    std::vector<Entry> all;
    all.reserve(total);
    for (auto& p : part) { all.insert(all.end(), p.begin(), p.end()); std::vector<Entry>().swap(p); }
    leaves.swap(all);
}

// Build the native tree.  leaves: one reference leaf entry per triangle that the
// reference tree can reach; solid_of[tri] (may be null) groups triangles.
// Returns entries in `nodes` (root at 0).
int build_native_tree(std::vector<Entry>& leaves, const uint32_t* solid_of, std::vector<Entry>& nodes,
                      const SplitInput* split)
{
    if (split && split->max_pieces > 1 && split->vertices && split->triangles && !leaves.empty())
        presplit_leaves(leaves, *split);
    const uint64_t n = leaves.size();
    if (n + n / 8 >= (1ull << 28))
        return fail(CB_ERR_INVALID, "native BVH: too many leaf entries for the 28-bit child field (lower the split count)");
    nodes.clear();
    if (n == 0) { nodes.push_back(Entry{0, 0, 0, 0}); return CB_OK; }
    // group by solid (stable counting sort keeps triangle order inside a solid)
    std::vector<uint64_t> solid_begin;
    if (solid_of) {
        uint32_t nsolids = 0;
        for (uint64_t i = 0; i < n; i++) nsolids = std::max(nsolids, solid_of[leaves[i].w & 0x0FFFFFFFu] + 1);
        std::vector<uint64_t> count(nsolids + 1, 0);
        for (uint64_t i = 0; i < n; i++) count[solid_of[leaves[i].w & 0x0FFFFFFFu] + 1]++;
        for (uint32_t s = 0; s < nsolids; s++) count[s + 1] += count[s];
        std::vector<Entry> sorted(n);
        std::vector<uint64_t> cursor(count.begin(), count.end() - 1);
        for (uint64_t i = 0; i < n; i++) sorted[cursor[solid_of[leaves[i].w & 0x0FFFFFFFu]]++] = leaves[i];
        leaves.swap(sorted);
        for (uint32_t s = 0; s < nsolids; s++)
            if (count[s + 1] > count[s]) solid_begin.push_back(count[s]);
        solid_begin.push_back(n);
    } else {
        solid_begin = {0, n};
    }
    const uint64_t nsol = solid_begin.size() - 1;
    // (a tree needs ~1.3 entries per leaf; the slack is for the per-thread arenas; child indices are 28 bits)
    nodes.assign(std::min<uint64_t>(2 * n + n / 4 + 64 * ARENA_CHUNK, (1ull << 28) - 1), Entry{0, 0, 0, 0});
    Builder B;
    B.prims = leaves.data();
    B.out = &nodes;

    struct Task { uint64_t b, e; IBox box; uint64_t slot; };
    std::vector<Task> tasks;
    if (nsol == 1) {
        IBox box; box.reset();
        for (uint64_t i = 0; i < n; i++) { int lo[3], hi[3]; box_of(leaves[i], lo, hi); box.grow(lo, hi); }
        tasks.push_back(Task{0, n, box, 0});
    } else {
        // top level: SAH over solids, each solid represented by its bounding box
        std::vector<Entry> tl(nsol);
        std::vector<IBox> sbox(nsol);
        for (uint64_t s = 0; s < nsol; s++) {
            IBox box; box.reset();
            for (uint64_t i = solid_begin[s]; i < solid_begin[s + 1]; i++) { int lo[3], hi[3]; box_of(leaves[i], lo, hi); box.grow(lo, hi); }
            sbox[s] = box;
            tl[s] = Builder::pack(box, (uint32_t)s);
        }
        Builder T;
        T.prims = tl.data();
        std::vector<Entry> top(nsol * 2 + 2 * ARENA_CHUNK, Entry{0, 0, 0, 0});
        T.out = &top;
        Arena TA;
        IBox all; all.reset();
        for (uint64_t s = 0; s < nsol; s++) all.grow(sbox[s]);
        T.build(TA, 0, nsol, all, 0);
        if (T.overflow) return fail(CB_ERR_NOMEM, "native BVH: top-level node pool overflow");
        // copy the top tree into the output; its leaves (w = solid id, nchild = 0) become tasks
        const uint64_t ntop = std::min<uint64_t>(T.next.load(), top.size());
        for (uint64_t i = 0; i < ntop; i++) nodes[i] = top[i];
        B.next = (ntop + 7) / 8 * 8;
        // walk the top tree to find referenced leaf slots
        std::vector<uint64_t> stack = {0};
        if ((top[0].w >> 28) == 0) {
            tasks.push_back(Task{solid_begin[top[0].w], solid_begin[top[0].w + 1], sbox[top[0].w], 0});
        } else {
            while (!stack.empty()) {
                uint64_t i = stack.back();
                stack.pop_back();
                uint32_t first = top[i].w & 0x0FFFFFFFu, k = top[i].w >> 28;
                for (uint32_t c = first; c < first + k; c++) {
                    if ((top[c].w >> 28) == 0) {
                        uint32_t s = top[c].w;
                        tasks.push_back(Task{solid_begin[s], solid_begin[s + 1], sbox[s], c});
                    } else {
                        stack.push_back(c);
                    }
                }
            }
        }
    }
    unsigned nthreads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    // Ranges too large for one thread (a single-level build starts with ONE range over all leaves)
    // are expanded here, one node at a time with the passes of sah_split spread over the threads;
    // their children join the task list.
    {
        constexpr uint64_t BIG = 1u << 19;
        Arena A0;
        std::vector<Task> pending;
        pending.swap(tasks);
        while (!pending.empty()) {
            Task t = pending.back();
            pending.pop_back();
            if (t.e - t.b <= BIG || nthreads == 1) { tasks.push_back(t); continue; }
            Range r[WIDE + 1];
            const int k = B.widen(t.b, t.e, t.box, r, nthreads);
            const uint64_t blk = B.alloc(A0, k);
            if (B.overflow) break;
            nodes[t.slot] = Builder::pack(t.box, ((uint32_t)k << 28) | (uint32_t)blk);
            for (int i = 0; i < k; i++) pending.push_back(Task{r[i].b, r[i].e, r[i].box, blk + i});
        }
    }
    // biggest tasks first, then a simple work-stealing loop over host threads
    std::sort(tasks.begin(), tasks.end(), [](const Task& a, const Task& b) { return (a.e - a.b) > (b.e - b.b); });
    std::atomic<size_t> cursor{0};
    auto worker = [&]() {
        Arena A;
        for (;;) {
            size_t i = cursor.fetch_add(1);
            if (i >= tasks.size()) break;
            B.build(A, tasks[i].b, tasks[i].e, tasks[i].box, tasks[i].slot);
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; t++) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    if (B.overflow) return fail(CB_ERR_NOMEM, "native BVH: node pool overflow");
    nodes.resize(std::min<uint64_t>(B.next.load() + 16, nodes.size()));

    // breadth-first relayout: children blocks in level order, same alignment rule
    std::vector<Entry> bfs(nodes.size() + 8 * 1024, Entry{0, 0, 0, 0});
    std::vector<std::pair<uint64_t, uint64_t>> level = {{0, 0}}, nextlevel;   // (old slot, new slot)
    bfs[0] = nodes[0];
    uint64_t cur = 1;
    while (!level.empty()) {
        nextlevel.clear();
        for (auto& pr : level) {
            const Entry& en = nodes[pr.first];
            uint32_t k = en.w >> 28, first = en.w & 0x0FFFFFFFu;
            if (k == 0) continue;
            uint64_t align = (k > 4) ? 8 : 2;
            uint64_t start = (cur + align - 1) / align * align;
            if (start + k + 8 > bfs.size()) bfs.resize(bfs.size() * 2, Entry{0, 0, 0, 0});
            for (uint32_t c = 0; c < k; c++) {
                bfs[start + c] = nodes[first + c];
                nextlevel.push_back({first + c, start + c});
            }
            bfs[pr.second].w = (k << 28) | (uint32_t)start;
            cur = start + k;
        }
        level.swap(nextlevel);
    }
    if (cur + 16 >= (1ull << 28)) return fail(CB_ERR_INVALID, "native BVH: tree exceeds the 28-bit child field (lower the split count)");
    bfs.resize(cur + 16);
    nodes.swap(bfs);
    return CB_OK;
}

} // namespace cb
