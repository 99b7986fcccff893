// hostmesh.cu -- host-side mesh preparation (no GPU needed).
//
// cb_unique_vertices: the vertex de-duplication of Geometry.flatten / Mesh
// (chroma/geometry.py:59-69: np.unique over the rows of the vertex array, triangles
// remapped through the inverse).  NumPy sorts 18.6 M structured rows of the 29k-PMT
// detector on one core in 15-30 s, which is most of the reference's CPU-side geometry
// build; here the rows are sorted as 16-byte records {x, y, z, index} by all host
// threads (chunk sort + parallel pairwise merges), runs of equal rows are numbered
// with a prefix sum, and the result is identical to np.unique(axis=0,
// return_inverse=True): unique rows in lexicographic (x, then y, then z) float order.
#include "host.h"
#include <algorithm>
#include <thread>
#include <string.h>

namespace cb {

struct Row { float x, y, z; uint32_t i; };

static inline bool row_less(const Row& a, const Row& b)
{
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    if (a.z != b.z) return a.z < b.z;
    return a.i < b.i;                       // deterministic; NumPy leaves the order inside a run open
}
static inline bool row_same(const Row& a, const Row& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

template <class F>
static void parallel_for(unsigned nthreads, F f)
{
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; t++) pool.emplace_back(f, t);
    f(0u);
    for (auto& th : pool) th.join();
}

} // namespace cb

using namespace cb;

extern "C" int cb_unique_vertices(const float* vertices, uint64_t n, float* unique_out, uint32_t* inverse_out,
                                  uint64_t* nunique_out)
{
    if (!vertices || !unique_out || !inverse_out || !nunique_out) return fail(CB_ERR_INVALID, "cb_unique_vertices: null argument");
    if (n >= (1ull << 32)) return fail(CB_ERR_INVALID, "cb_unique_vertices: too many vertices");
    *nunique_out = 0;
    if (n == 0) return CB_OK;
    unsigned T = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    while (T > 1 && n / T < 65536) T /= 2;
    unsigned P = 1;
    while (P * 2 <= T) P *= 2;              // power-of-two number of chunks: clean merge tree
    std::vector<Row> a(n), b(n);
    std::vector<uint64_t> cut(P + 1);
    for (unsigned c = 0; c <= P; c++) cut[c] = n * c / P;
    parallel_for(P, [&](unsigned c) {
        for (uint64_t i = cut[c]; i < cut[c + 1]; i++) a[i] = Row{vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2], (uint32_t)i};
        std::sort(a.begin() + cut[c], a.begin() + cut[c + 1], row_less);
    });
    Row* src = a.data();
    Row* dst = b.data();
    for (unsigned width = 1; width < P; width *= 2) {
        const unsigned pairs = P / (2 * width);
        // every merge is split in two at the median of its left run so that `2 * pairs` threads work
        parallel_for(2 * pairs, [&](unsigned w) {
            const unsigned pr = w / 2, half = w % 2;
            const uint64_t lo = cut[pr * 2 * width], mid = cut[pr * 2 * width + width], hi = cut[(pr + 1) * 2 * width];
            const uint64_t lm = lo + (mid - lo) / 2;
            const Row* rsplit = std::lower_bound(src + mid, src + hi, src[lm], row_less);
            const uint64_t rm = (uint64_t)(rsplit - src);
            if (half == 0) std::merge(src + lo, src + lm, src + mid, src + rm, dst + lo, row_less);
            else std::merge(src + lm, src + mid, src + rm, src + hi, dst + lo + (lm - lo) + (rm - mid), row_less);
        });
        std::swap(src, dst);
    }
    // number the runs of equal rows: per-chunk counts, exclusive scan, then write
    std::vector<uint64_t> starts(P + 1, 0);
    parallel_for(P, [&](unsigned c) {
        uint64_t k = 0;
        for (uint64_t i = cut[c]; i < cut[c + 1]; i++) k += (i == 0) || !row_same(src[i - 1], src[i]);
        starts[c + 1] = k;
    });
    for (unsigned c = 0; c < P; c++) starts[c + 1] += starts[c];
    parallel_for(P, [&](unsigned c) {
        uint64_t u = starts[c];             // runs started before this chunk
        for (uint64_t i = cut[c]; i < cut[c + 1]; i++) {
            if ((i == 0) || !row_same(src[i - 1], src[i])) {
                unique_out[3 * u] = src[i].x; unique_out[3 * u + 1] = src[i].y; unique_out[3 * u + 2] = src[i].z;
                u++;
            }
            inverse_out[src[i].i] = (uint32_t)(u - 1);
        }
    });
    *nunique_out = starts[P];
    return CB_OK;
}
