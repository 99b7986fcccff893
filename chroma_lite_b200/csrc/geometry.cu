// geometry.cu -- geometry / detector upload into the engine's native layout and
// the recursive-grid BVH builder in the reference node format.
#include "host.h"
#include "sort.cuh"
#include <algorithm>
#include <string.h>
#include <math.h>
#include <thread>
#include <chrono>
#include <string>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>

namespace cb {

struct Entry { uint32_t x, y, z, w; };
struct SplitInput {               // bvh_native.cu: optional splitting of loosely bounded leaves
    const float* vertices;
    const uint32_t* triangles;
    float origin[3], scale;
    int max_pieces, min_extent;
    double min_ratio;
};
int build_native_tree(std::vector<Entry>& leaves, const uint32_t* solid_of, std::vector<Entry>& nodes,
                      const SplitInput* split);

// CHROMA_B200_LEAF_SPLIT=<max pieces per triangle>[,<min extent in grid quanta>[,<min box/part area ratio>]]
// turns leaf splitting on (default: off, one leaf per triangle).
static bool split_from_env(SplitInput& s)
{
    const char* e = getenv("CHROMA_B200_LEAF_SPLIT");
    s.max_pieces = 0; s.min_extent = 8; s.min_ratio = 2.0;
    if (!e || !*e) return false;
    double r = s.min_ratio;
    int n = sscanf(e, "%d,%d,%lf", &s.max_pieces, &s.min_extent, &r);
    if (n >= 3) s.min_ratio = r;
    return n >= 1 && s.max_pieces > 1;
}

// Pack triangles for the traversal: 64 B = 4 x float4 per triangle holding the
// three world-space vertices, the reference test rank (tie-break, SURVEY A-1),
// the material code and the triangle's reference leaf box, so a leaf test is
// three 128-bit loads from two 32-byte sectors instead of an index fetch + three
// scattered 12-byte gathers.
__global__ void __launch_bounds__(256)
pack_triangles_kernel(const float* __restrict__ vertices, const uint32_t* __restrict__ triangles,
                      const uint32_t* __restrict__ material_codes, const uint32_t* __restrict__ rank,
                      const uint32_t* __restrict__ leafbox, uint64_t ntriangles, float4* __restrict__ tri64)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntriangles) return;
    uint32_t i0 = triangles[3 * t], i1 = triangles[3 * t + 1], i2 = triangles[3 * t + 2];
    float3 v0 = ld3(vertices, i0), v1 = ld3(vertices, i1), v2 = ld3(vertices, i2);
    tri64[4 * t + 0] = make_float4(v0.x, v0.y, v0.z, v1.x);
    tri64[4 * t + 1] = make_float4(v1.y, v1.z, v2.x, v2.y);
    tri64[4 * t + 2] = make_float4(v2.z, __uint_as_float(rank[t]), __uint_as_float(material_codes[t]), 0.0f);
    tri64[4 * t + 3] = make_float4(__uint_as_float(leafbox[3 * t]), __uint_as_float(leafbox[3 * t + 1]),
                                   __uint_as_float(leafbox[3 * t + 2]), 0.0f);
}

// Position at which the reference traversal (mesh.h:75-117) would test each
// triangle if nothing were pruned: leaf children of a group in ascending index,
// then internal children in LIFO order.  Ray independent.
static void reference_test_rank(const uint32_t* nodes, uint64_t nnodes, uint64_t ntriangles,
                                std::vector<uint32_t>& rank)
{
    rank.assign(ntriangles, 0xFFFFFFFFu);
    std::vector<uint32_t> stack;
    stack.reserve(4096);
    uint32_t r = 0;
    stack.push_back(nodes[3]);
    while (!stack.empty()) {
        uint32_t w = stack.back();
        stack.pop_back();
        uint32_t first = w & 0x0FFFFFFFu, n = w >> 28;
        for (uint32_t i = first; i < first + n && i < nnodes; i++) {
            uint32_t cw = nodes[4ull * i + 3];
            if ((cw >> 28) == 0) {
                uint32_t tri = cw & 0x0FFFFFFFu;
                if (tri < ntriangles && rank[tri] == 0xFFFFFFFFu) rank[tri] = r++;
            } else {
                stack.push_back(cw);
            }
        }
    }
}

// one leaf entry per triangle the reference traversal can reach, verbatim
static void collect_reference_leaves(const uint32_t* nodes, uint64_t nnodes, uint64_t ntriangles,
                                     const std::vector<uint32_t>& rank, std::vector<Entry>& leaves)
{
    leaves.clear();
    leaves.reserve(ntriangles);
    // a root that is itself a leaf is never tested by the reference (mesh.h:55-68)
    if ((nodes[3] >> 28) == 0) return;
    std::vector<uint8_t> seen(ntriangles, 0);
    for (uint64_t i = 0; i < nnodes; i++) {
        const uint32_t w = nodes[4 * i + 3];
        if ((w >> 28) == 0 && w < ntriangles && rank[w] != 0xFFFFFFFFu && !seen[w]) {
            seen[w] = 1;
            leaves.push_back(Entry{nodes[4 * i], nodes[4 * i + 1], nodes[4 * i + 2], w});
        }
    }
}

// ---------------------------------------------------------------- prepared-tree cache
// What cb_geometry_create derives from the reference-format tree on the host -- the reference test rank
// and leaf box of every triangle and the engine's own traversal tree -- costs 5-8 s for the 29k-PMT
// detector, and with one process per GPU every rank of a node would derive the same thing at the same
// time with all host threads.  With CHROMA_B200_TREE_CACHE=<dir> the result is kept in a file keyed by a
// hash of the inputs; the first process to ask builds and writes it (create-exclusive lock file), the
// others wait for the file and read it (role of chroma/cache.py:209-236 for BVHs, one level further).
struct PreparedTree {
    std::vector<uint32_t> rank;       // [ntriangles]
    std::vector<uint32_t> leafbox;    // [3 * ntriangles]
    std::vector<Entry> native;        // engine tree (empty: traverse the reference tree)
};

static uint64_t mix64(uint64_t h, uint64_t v)
{
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xFF51AFD7ED558CCDull;
    return h ^ (h >> 32);
}
static uint64_t hash_words(uint64_t h, const void* p, uint64_t bytes)
{
    const uint64_t* w = static_cast<const uint64_t*>(p);
    const uint64_t n = bytes / 8;
    uint64_t a = h, b = ~h, c = h * 3, d2 = h ^ 0x5555555555555555ull;      // four lanes: the loop is memory-bound
    uint64_t i = 0;
    for (; i + 4 <= n; i += 4) { a = mix64(a, w[i]); b = mix64(b, w[i + 1]); c = mix64(c, w[i + 2]); d2 = mix64(d2, w[i + 3]); }
    for (; i < n; i++) a = mix64(a, w[i]);
    const unsigned char* tail = static_cast<const unsigned char*>(p) + n * 8;
    uint64_t t = 0;
    for (uint64_t k = 0; k < bytes % 8; k++) t = (t << 8) | tail[k];
    return mix64(mix64(mix64(a, b), mix64(c, d2)), t ^ bytes);
}

static bool read_all(FILE* f, void* p, size_t bytes) { return bytes == 0 || fread(p, 1, bytes, f) == bytes; }
static bool write_all(FILE* f, const void* p, size_t bytes) { return bytes == 0 || fwrite(p, 1, bytes, f) == bytes; }

static bool prepared_tree_load(const std::string& path, uint64_t key, uint64_t ntriangles, PreparedTree& out)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    uint64_t head[4] = {0, 0, 0, 0};       // magic, key, ntriangles, native entries
    bool ok = read_all(f, head, sizeof(head)) && head[0] == 0x43423230305452ull && head[1] == key && head[2] == ntriangles;
    if (ok) {
        out.rank.resize(ntriangles); out.leafbox.resize(3 * ntriangles); out.native.resize(head[3]);
        ok = read_all(f, out.rank.data(), ntriangles * 4) && read_all(f, out.leafbox.data(), ntriangles * 12) &&
             read_all(f, out.native.data(), head[3] * sizeof(Entry));
    }
    fclose(f);
    return ok;
}

static void prepared_tree_store(const std::string& path, uint64_t key, uint64_t ntriangles, const PreparedTree& t)
{
    const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const uint64_t head[4] = {0x43423230305452ull, key, ntriangles, (uint64_t)t.native.size()};
    const bool ok = write_all(f, head, sizeof(head)) && write_all(f, t.rank.data(), ntriangles * 4) &&
                    write_all(f, t.leafbox.data(), ntriangles * 12) && write_all(f, t.native.data(), t.native.size() * sizeof(Entry));
    fclose(f);
    if (ok) rename(tmp.c_str(), path.c_str()); else remove(tmp.c_str());
}

template <typename T>
static int upload(T** dst, const T* src, uint64_t count, uint64_t& total)
{
    *dst = nullptr;
    uint64_t bytes = std::max<uint64_t>(count, 1) * sizeof(T);
    CB_CUDA(cudaMalloc((void**)dst, bytes));
    if (count && src) CB_CUDA(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
    total += bytes;
    return CB_OK;
}

// Per-plane constants of the analytic wire planes, in double precision (physics.cuh WireFrame):
// orthonormal frame (U normalised, V made orthogonal to U and normalised, N = U x V), the
// wire index range inside [vmin, vmax] and the squared radius.  The reference recomputes
// all of this per photon and step (chroma/cuda/photon.h:113-160).
static WireFrame wire_frame(const CbWirePlane& wp)
{
    WireFrame f;
    memset(&f, 0, sizeof(f));
    double u[3] = {wp.u[0], wp.u[1], wp.u[2]}, v[3] = {wp.v[0], wp.v[1], wp.v[2]};
    const double ul = 1.0 / sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    for (int a = 0; a < 3; a++) f.U[a] = u[a] * ul;
    const double along = v[0] * f.U[0] + v[1] * f.U[1] + v[2] * f.U[2];
    for (int a = 0; a < 3; a++) v[a] -= along * f.U[a];
    const double vl = 1.0 / sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int a = 0; a < 3; a++) f.V[a] = v[a] * vl;
    f.N[0] = f.U[1] * f.V[2] - f.U[2] * f.V[1];
    f.N[1] = f.U[2] * f.V[0] - f.U[0] * f.V[2];
    f.N[2] = f.U[0] * f.V[1] - f.U[1] * f.V[0];
    f.v0 = wp.v0; f.pitch = wp.pitch; f.inv_pitch = (f.pitch != 0.0) ? 1.0 / f.pitch : 0.0;
    f.radius = wp.radius; f.radius2 = f.radius * f.radius;
    f.umin = wp.umin; f.umax = wp.umax;
    for (int a = 0; a < 3; a++) f.origin[a] = wp.origin[a];
    f.kmin = (int)ceil(((double)wp.vmin - (double)wp.v0) / f.pitch);
    f.kmax = (int)floor(((double)wp.vmax - (double)wp.v0) / f.pitch);
    f.surface = wp.surface_index;
    f.material_inner = wp.material_inner_index; f.material_outer = wp.material_outer_index;
    return f;
}

static void free_geometry(Geometry* g)
{
    cudaFree(g->vertices); cudaFree(g->triangles); cudaFree(g->material_codes); cudaFree(g->colors);
    cudaFree(g->solid_id); cudaFree(g->nodes); cudaFree(g->native_nodes); cudaFree(g->tri64); cudaFree(g->tables);
    cudaFree(g->materials); cudaFree(g->surfaces); cudaFree(g->wireframes); cudaFree(g->solid_to_channel);
    cudaFree(g->time_cdf_x); cudaFree(g->time_cdf_y); cudaFree(g->charge_cdf_x); cudaFree(g->charge_cdf_y);
    delete g;
}

// rank + leaf box per triangle + the engine's traversal tree, computed from the descriptor
static int prepare_tree_compute(const CbGeometryDesc* d, bool use_reference_tree, PreparedTree& out)
{
    reference_test_rank(d->nodes, d->nnodes, d->ntriangles, out.rank);
    // reference leaf box of every triangle (first leaf entry that names it)
    out.leafbox.assign(3 * std::max<uint64_t>(d->ntriangles, 1), 0u);
    {
        std::vector<uint8_t> seen(d->ntriangles, 0);
        for (uint64_t i = 0; i < d->nnodes; i++) {
            const uint32_t w = d->nodes[4 * i + 3];
            if ((w >> 28) == 0 && w < d->ntriangles && !seen[w]) {
                seen[w] = 1;
                for (int a = 0; a < 3; a++) out.leafbox[3ull * w + a] = d->nodes[4 * i + a];
            }
        }
    }
    out.native.clear();
    if (use_reference_tree || d->ntriangles == 0) return CB_OK;
    std::vector<Entry> leaves;
    collect_reference_leaves(d->nodes, d->nnodes, d->ntriangles, out.rank, leaves);
    SplitInput split;
    const bool do_split = split_from_env(split);
    split.vertices = d->vertices; split.triangles = d->triangles; split.scale = d->world_scale;
    for (int a = 0; a < 3; a++) split.origin[a] = d->world_origin[a];
    // One SAH hierarchy over all leaves (default since round 2: -9 % / -19 % traversal iterations on the
    // 29k-PMT detector, bit-identical hits); CHROMA_B200_TREE=solids builds solids first, then one
    // subtree per solid (faster to build, the round-1 default)
    const char* tree_env = getenv("CHROMA_B200_TREE");
    const bool single_level = !(tree_env && strcmp(tree_env, "solids") == 0);
    int rc = build_native_tree(leaves, single_level ? nullptr : d->solid_id, out.native, do_split ? &split : nullptr);
    if (rc != CB_OK) return rc;
    out.native.resize(out.native.size() + 16, Entry{0, 0, 0, 0});      // the batched sibling fetch may read past the last child
    return CB_OK;
}

static int prepare_tree(const CbGeometryDesc* d, PreparedTree& out)
{
    const char* tree_env = getenv("CHROMA_B200_TREE");
    const bool use_reference_tree = tree_env && strcmp(tree_env, "reference") == 0;
    if (use_reference_tree) {
        // the traversal kernels fetch at most 8 children per node (the native tree's bound); the reference
        // format allows 15 (bvh/grid.py caps groups at MAX_CHILD): such a tree cannot be walked as it is
        for (uint64_t i = 0; i < d->nnodes; i++)
            if ((d->nodes[4 * i + 3] >> 28) > 8)
                return fail(CB_ERR_UNSUPPORTED, "CHROMA_B200_TREE=reference: node %llu has %u children, the traversal "
                            "handles at most 8 (unset CHROMA_B200_TREE to traverse the engine's own tree)",
                            (unsigned long long)i, d->nodes[4 * i + 3] >> 28);
    }
    const char* dir = getenv("CHROMA_B200_TREE_CACHE");
    if (!dir || !*dir || d->ntriangles < 50000) return prepare_tree_compute(d, use_reference_tree, out);

    // key: everything the result depends on
    uint64_t key = hash_words(0x6368726f6d61ull, d->nodes, d->nnodes * 16);
    const char* split_env = getenv("CHROMA_B200_LEAF_SPLIT");
    const bool by_solid = tree_env && strcmp(tree_env, "solids") == 0;
    key = mix64(key, d->ntriangles);
    key = hash_words(key, tree_env ? tree_env : "", tree_env ? strlen(tree_env) : 0);
    key = hash_words(key, split_env ? split_env : "", split_env ? strlen(split_env) : 0);
    if (by_solid && d->solid_id) key = hash_words(key, d->solid_id, d->ntriangles * 4);
    if (split_env && *split_env) {
        key = hash_words(key, d->vertices, d->nvertices * 12);
        key = hash_words(key, d->triangles, d->ntriangles * 12);
    }
    key = mix64(key, 3);                       // format / builder version
    char name[64];
    snprintf(name, sizeof(name), "/tree_%016llx.bin", (unsigned long long)key);
    mkdir(dir, 0777);
    const std::string path = std::string(dir) + name, lock = path + ".lock";
    if (prepared_tree_load(path, key, d->ntriangles, out)) return CB_OK;
    const int fd = open(lock.c_str(), O_CREAT | O_EXCL | O_WRONLY, 0666);
    if (fd < 0) {
        // another process of this node is building it: wait for the file (or for the builder to give up)
        for (int waited = 0; waited < 6000; waited++) {
            std::this_thread::sleep_for(std::chrono::milliseconds(100));
            if (access(path.c_str(), R_OK) == 0 && prepared_tree_load(path, key, d->ntriangles, out)) return CB_OK;
            if (access(lock.c_str(), F_OK) != 0 && access(path.c_str(), R_OK) != 0) break;
        }
        return prepare_tree_compute(d, use_reference_tree, out);
    }
    close(fd);
    const int rc = prepare_tree_compute(d, use_reference_tree, out);
    if (rc == CB_OK) prepared_tree_store(path, key, d->ntriangles, out);
    remove(lock.c_str());
    return rc;
}

} // namespace cb

using namespace cb;

extern "C" {

int cb_geometry_create(const CbGeometryDesc* d, cb_geom_t* out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    if (!d || !out) return fail(CB_ERR_INVALID, "cb_geometry_create: null argument");
    if (d->nwireplanes < 0 || (d->nwireplanes > 0 && !d->wireplanes))
        return fail(CB_ERR_INVALID, "cb_geometry_create: nwireplanes without a wireplanes array");
    for (int i = 0; i < d->nwireplanes; i++) {
        const CbWirePlane& wp = d->wireplanes[i];
        if (wp.material_inner_index < 0 || wp.material_inner_index >= d->nmaterials || wp.material_outer_index < 0 ||
            wp.material_outer_index >= d->nmaterials || wp.surface_index >= d->nsurfaces)
            return fail(CB_ERR_INVALID, "wire plane %d: material/surface index out of range", i);
    }
    if (!d->vertices || !d->triangles || !d->material_codes || !d->nodes || d->nnodes == 0)
        return fail(CB_ERR_INVALID, "cb_geometry_create: vertices/triangles/material_codes/nodes are required");
    if (d->nmaterials <= 0 || d->nmaterials > 127 || d->nsurfaces < 0 || d->nsurfaces > 127)
        return fail(CB_ERR_INVALID, "material/surface count out of range (8-bit signed codes, SURVEY App. A-7)");
    if (d->ntriangles >= (1ull << 28)) return fail(CB_ERR_INVALID, "too many triangles for 28-bit child ids");
    for (uint64_t t = 0; t < 3 * d->ntriangles; t++)
        if (d->triangles[t] >= d->nvertices) return fail(CB_ERR_INVALID, "triangle %llu references a vertex out of range", (unsigned long long)(t / 3));
    for (uint64_t i = 0; i < d->nnodes; i++) {
        uint32_t w = d->nodes[4 * i + 3];
        uint32_t n = w >> 28, c = w & 0x0FFFFFFFu;
        if (n == 0) { if (c >= d->ntriangles && i != 0) return fail(CB_ERR_INVALID, "BVH leaf %llu references triangle %u out of range", (unsigned long long)i, c); }
        else if ((uint64_t)c + n > d->nnodes) return fail(CB_ERR_INVALID, "BVH node %llu children out of range", (unsigned long long)i);
    }

    Geometry* g = new Geometry();
    int rc;
    uint64_t total = 0;
#define UP(field, src, count) if ((rc = upload(&g->field, src, count, total)) != CB_OK) { free_geometry(g); return rc; }
    g->nvertices = d->nvertices; g->ntriangles = d->ntriangles; g->nnodes = d->nnodes; g->table_floats = d->table_floats;
    UP(vertices, d->vertices, 3 * d->nvertices);
    UP(triangles, d->triangles, 3 * d->ntriangles);
    UP(material_codes, d->material_codes, d->ntriangles);
    UP(colors, d->colors, d->colors ? d->ntriangles : 0);
    UP(solid_id, d->solid_id, d->solid_id ? d->ntriangles : 0);
    // node array padded so the batched sibling fetch never leaves the allocation
    {
        uint64_t bytes = (d->nnodes + 16) * sizeof(uint4);
        cudaError_t e = cudaMalloc((void**)&g->nodes, bytes);
        if (e != cudaSuccess) { free_geometry(g); return cuda_fail(e, "cudaMalloc(nodes)"); }
        if ((e = cudaMemset(g->nodes, 0, bytes)) != cudaSuccess ||
            (e = cudaMemcpy(g->nodes, d->nodes, d->nnodes * sizeof(uint4), cudaMemcpyHostToDevice)) != cudaSuccess) {
            free_geometry(g);
            return cuda_fail(e, "upload(nodes)");
        }
        total += bytes;
    }
    // table pool padded by 4 floats: interp_property may read fp[n] at the exact
    // upper edge; pool size rounded up to 16 B for the bulk copy
    {
        uint64_t nf = ((d->table_floats + 4 + 3) / 4) * 4;
        std::vector<float> pool(nf, 0.0f);
        if (d->table_floats) memcpy(pool.data(), d->table_pool, d->table_floats * sizeof(float));
        UP(tables, pool.data(), nf);
    }
    UP(materials, d->materials, (uint64_t)d->nmaterials);
    UP(surfaces, d->surfaces, (uint64_t)d->nsurfaces);
    {
        std::vector<WireFrame> frames;
        for (int i = 0; i < d->nwireplanes; i++) frames.push_back(wire_frame(d->wireplanes[i]));
        UP(wireframes, frames.data(), (uint64_t)d->nwireplanes);
    }
#undef UP

    // host-side preparation (rank, leaf boxes, the engine's tree): from the cache, or computed (and stored)
    PreparedTree prep;
    if ((rc = prepare_tree(d, prep)) != CB_OK) { free_geometry(g); return rc; }
    uint32_t* d_rank = nullptr;
    uint32_t* d_leafbox = nullptr;
    uint64_t scratch = 0;
    if ((rc = upload(&d_rank, prep.rank.data(), d->ntriangles, scratch)) != CB_OK) { free_geometry(g); return rc; }
    if ((rc = upload(&d_leafbox, prep.leafbox.data(), 3 * d->ntriangles, scratch)) != CB_OK) { cudaFree(d_rank); free_geometry(g); return rc; }
    {
        cudaError_t e = cudaMalloc((void**)&g->tri64, std::max<uint64_t>(d->ntriangles, 1) * 64);
        if (e != cudaSuccess) { cudaFree(d_rank); cudaFree(d_leafbox); free_geometry(g); return cuda_fail(e, "cudaMalloc(tri64)"); }
        total += d->ntriangles * 64;
        if (d->ntriangles) {
            pack_triangles_kernel<<<(unsigned)((d->ntriangles + 255) / 256), 256, 0, ctx().stream>>>(
                g->vertices, g->triangles, g->material_codes, d_rank, d_leafbox, d->ntriangles, g->tri64);
            e = stream_wait(ctx().stream);
            if (e != cudaSuccess) { cudaFree(d_rank); cudaFree(d_leafbox); free_geometry(g); return cuda_fail(e, "pack_triangles"); }
        }
        cudaFree(d_rank);
        cudaFree(d_leafbox);
    }
    const uint32_t* root_entry = d->nodes;
    std::vector<Entry>& native = prep.native;
    if (!native.empty()) {
        cudaError_t e = cudaMalloc((void**)&g->native_nodes, native.size() * sizeof(Entry));
        if (e != cudaSuccess) { free_geometry(g); return cuda_fail(e, "cudaMalloc(native nodes)"); }
        e = cudaMemcpy(g->native_nodes, native.data(), native.size() * sizeof(Entry), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { free_geometry(g); return cuda_fail(e, "upload(native nodes)"); }
        total += native.size() * sizeof(Entry);
        g->nnative = native.size();
        root_entry = reinterpret_cast<const uint32_t*>(native.data());
    }

    DevGeometry& v = g->dev;
    v.nodes = g->native_nodes ? g->native_nodes : g->nodes; v.tri64 = g->tri64; v.tables = g->tables;
    v.ref_nodes = g->nodes; v.ref_root_w = d->nodes[3];
    v.materials = g->materials; v.surfaces = g->surfaces;
    v.world_origin = make_float3(d->world_origin[0], d->world_origin[1], d->world_origin[2]);
    v.world_scale = d->world_scale;
    v.wavelength_n = d->wavelength_n; v.wavelength_start = d->wavelength_start; v.wavelength_step = d->wavelength_step;
    v.time_n = d->time_n; v.time_start = d->time_start; v.time_step = d->time_step;
    v.root_x = root_entry[0]; v.root_y = root_entry[1]; v.root_z = root_entry[2]; v.root_w = root_entry[3];
    v.ref_root_x = d->nodes[0]; v.ref_root_y = d->nodes[1]; v.ref_root_z = d->nodes[2];
    v.nmaterials = d->nmaterials; v.nsurfaces = d->nsurfaces;
    v.wireframes = d->nwireplanes ? g->wireframes : nullptr; v.nwireplanes = d->nwireplanes;
    // Stage the leading tables of the pool (the host lays the wavelength tables out first, the long
    // time CDFs last) into shared memory, up to 48 KB.  The cut falls on a table boundary: Tables::at()
    // decides shared vs global from a table's START offset, so no table may straddle it.
    {
        std::vector<std::pair<int64_t, int64_t>> spans;      // (offset, length) of every table in the pool
        auto add = [&](int32_t off, int64_t len) { if (off >= 0 && len > 0) spans.push_back({off, len}); };
        const int64_t W = d->wavelength_n;
        for (int i = 0; i < d->nmaterials; i++) {
            const CbMaterial& m = d->materials[i];
            add(m.refractive_index, W); add(m.absorption_length, W); add(m.scattering_length, W);
            const int64_t nc = std::max(m.num_comp, 0);
            add(m.comp_reemission_prob, nc * W); add(m.comp_reemission_wvl_cdf, nc * W);
            add(m.comp_absorption_length, nc * W); add(m.comp_reemission_time_cdf, nc * (int64_t)d->time_n);
        }
        for (int i = 0; i < d->nsurfaces; i++) {
            const CbSurface& sf = d->surfaces[i];
            if (sf.model < 0) continue;
            add(sf.detect, W); add(sf.absorb, W); add(sf.reemit, W); add(sf.reflect_diffuse, W);
            add(sf.reflect_specular, W); add(sf.eta, W); add(sf.k, W); add(sf.reemission_cdf, W);
            const int64_t nd = std::max(sf.dichroic_nangles, 0), na = std::max(sf.angular_nangles, 0);
            add(sf.dichroic_angles, nd); add(sf.dichroic_reflect, nd * W); add(sf.dichroic_transmit, nd * W);
            add(sf.angular_angles, na); add(sf.angular_transmit, na); add(sf.angular_reflect_specular, na);
            add(sf.angular_reflect_diffuse, na);
        }
        for (auto& sp : spans)
            if ((uint64_t)(sp.first + sp.second) > d->table_floats) {
                free_geometry(g);
                return fail(CB_ERR_INVALID, "table at pool offset %lld (%lld floats) exceeds the table pool (%llu floats)",
                            (long long)sp.first, (long long)sp.second, (unsigned long long)d->table_floats);
            }
        std::sort(spans.begin(), spans.end());
        const int64_t LIMIT = 12288 - 4;          // 48 KB less the one-past-the-end read of interp_property
        int64_t cut = 0, reach = 0;                 // reach: end of the tables that start before `cut`
        for (auto& sp : spans) {
            if (sp.first >= reach && reach <= LIMIT) cut = reach;    // a boundary nothing straddles
            reach = std::max(reach, sp.first + sp.second);
        }
        if (reach <= LIMIT) cut = reach;
        v.smem_floats = (uint32_t)cut;
        // the copy covers one float beyond the cut (fp[n] at the exact upper edge) and is a 16-byte multiple;
        // the pool itself is padded by 4 floats
        g->smem_table_bytes = cut ? (uint32_t)(((cut + 1 + 3) / 4) * 16) : 0u;
        v.smem_bytes = g->smem_table_bytes;
    }
    g->device_bytes = total;
    *out = geoms().add(g);
    return CB_OK;
}

int cb_geometry_destroy(cb_geom_t h)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    Geometry* g = geoms().take(h);
    if (!g) return fail(CB_ERR_INVALID, "cb_geometry_destroy: bad handle");
    stream_wait(ctx().stream);
    if (ctx().l2_window_base == (g->native_nodes ? (const void*)g->native_nodes : (const void*)g->nodes)) {
        ctx().l2_window_base = nullptr;      // a later allocation may reuse the address
        ctx().l2_window_bytes = ~(size_t)0;
    }
    free_geometry(g);
    return CB_OK;
}

int cb_geometry_info(cb_geom_t h, CbGeometryInfo* info)
{
    CB_REQUIRE_INIT();
    Geometry* g = geoms().get(h);
    if (!g || !info) return fail(CB_ERR_INVALID, "cb_geometry_info: bad handle");
    memset(info, 0, sizeof(*info));
    info->vertices = g->vertices; info->triangles = g->triangles; info->material_codes = g->material_codes;
    info->colors = g->colors; info->solid_id_map = g->solid_id; info->nodes = g->nodes;
    info->solid_id_to_channel_index = g->solid_to_channel;
    info->time_cdf_x = g->time_cdf_x; info->time_cdf_y = g->time_cdf_y;
    info->charge_cdf_x = g->charge_cdf_x; info->charge_cdf_y = g->charge_cdf_y;
    info->nvertices = g->nvertices; info->ntriangles = g->ntriangles; info->nnodes = g->nnodes;
    info->nchannels = g->nchannels; info->device_bytes = g->device_bytes;
    info->max_stack_depth = CB_PSTACK + CB_PLSTACK;
    return CB_OK;
}

int cb_detector_attach(cb_geom_t h, const int32_t* solid_id_to_channel_index, uint64_t nsolids,
                       int32_t nchannels, const float* time_cdf_x, const float* time_cdf_y,
                       int32_t time_cdf_len, const float* charge_cdf_x, const float* charge_cdf_y,
                       int32_t charge_cdf_len, float charge_unit)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    Geometry* g = geoms().get(h);
    if (!g) return fail(CB_ERR_INVALID, "cb_detector_attach: bad handle");
    if (!solid_id_to_channel_index || time_cdf_len < 2 || charge_cdf_len < 2)
        return fail(CB_ERR_INVALID, "cb_detector_attach: channel map and CDFs (len >= 2) required");
    if (!g->solid_id && g->ntriangles) return fail(CB_ERR_INVALID, "cb_detector_attach: geometry has no solid_id map");
    int rc;
    uint64_t total = 0;
    cudaFree(g->solid_to_channel); cudaFree(g->time_cdf_x); cudaFree(g->time_cdf_y);
    cudaFree(g->charge_cdf_x); cudaFree(g->charge_cdf_y);
    if ((rc = upload(&g->solid_to_channel, solid_id_to_channel_index, nsolids, total))) return rc;
    if ((rc = upload(&g->time_cdf_x, time_cdf_x, (uint64_t)time_cdf_len, total))) return rc;
    if ((rc = upload(&g->time_cdf_y, time_cdf_y, (uint64_t)time_cdf_len, total))) return rc;
    if ((rc = upload(&g->charge_cdf_x, charge_cdf_x, (uint64_t)charge_cdf_len, total))) return rc;
    if ((rc = upload(&g->charge_cdf_y, charge_cdf_y, (uint64_t)charge_cdf_len, total))) return rc;
    g->nsolids = nsolids; g->nchannels = nchannels;
    g->time_cdf_len = time_cdf_len; g->charge_cdf_len = charge_cdf_len; g->charge_unit = charge_unit;
    g->device_bytes += total;
    return CB_OK;
}

} // extern "C"

// =====================================================================
// BVH construction in the reference node format
// =====================================================================
namespace cb {

__device__ __forceinline__ unsigned long long spread3_16(uint32_t input)
{
    unsigned long long x = input;
    x = (x | (x << 16)) & 0x00000000FF0000FFull;
    x = (x | (x << 8)) & 0x000000F00F00F00Full;
    x = (x | (x << 4)) & 0x00000C30C30C30C3ull;
    x = (x | (x << 2)) & 0x0000249249249249ull;
    return x;
}
__device__ __forceinline__ uint32_t quantize(float v, float origin, float scale)
{
    return (uint32_t)((v - origin) / scale);   // truncation
}

// per triangle: 16-bit quantised AABB padded by one quantum and the 48-bit
// Morton code of the centroid (behaviour of bvh.cu:148-203)
__global__ void __launch_bounds__(256)
make_leaves_kernel(const float* __restrict__ vertices, const uint32_t* __restrict__ triangles,
                   uint64_t ntriangles, float3 origin, float scale, uint4* __restrict__ leaves,
                   unsigned long long* __restrict__ codes, uint32_t* __restrict__ ids)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntriangles) return;
    float3 a = ld3(vertices, triangles[3 * t]);
    float3 b = ld3(vertices, triangles[3 * t + 1]);
    float3 c = ld3(vertices, triangles[3 * t + 2]);
    float3 lo = f3(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z));
    float3 hi = f3(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z));
    float3 ce = (a + b + c) / 3.0f;
    uint32_t lx = quantize(lo.x, origin.x, scale), ly = quantize(lo.y, origin.y, scale), lz = quantize(lo.z, origin.z, scale);
    if (lx > 0) lx--;
    if (ly > 0) ly--;
    if (lz > 0) lz--;
    uint32_t ux = quantize(hi.x, origin.x, scale) + 1, uy = quantize(hi.y, origin.y, scale) + 1, uz = quantize(hi.z, origin.z, scale) + 1;
    uint32_t cx = quantize(ce.x, origin.x, scale), cy = quantize(ce.y, origin.y, scale), cz = quantize(ce.z, origin.z, scale);
    codes[t] = spread3_16(cx) | (spread3_16(cy) << 1) | (spread3_16(cz) << 2);
    ids[t] = (uint32_t)t;
    leaves[t] = make_uint4(lx | (ux << 16), ly | (uy << 16), lz | (uz << 16), (uint32_t)t);
}

struct BvhResult {
    std::vector<uint32_t> nodes;      // 4 words per node
    std::vector<uint64_t> layer_offsets;
    float origin[3]; float scale;
    bool valid = false;
};
static BvhResult g_bvh;

// ---------------------------------------------------------------- layers on the device
// The reference builds its layers on the host from arrays it copies back after every kernel
// (bvh/grid.py:27-95: NumPy ediff1d / argwhere per layer, make_parents_detailed on the GPU, .get()).
// Here a layer never leaves the device: run boundaries of the (shifted) Morton codes, group ids by a
// scan, groups cut into runs of at most 15 children, one parent per run; the layers are concatenated
// root first and single-child chains collapsed, and the finished node array is read back once.
constexpr uint32_t BVH_MAX_CHILD = 15;

__global__ void __launch_bounds__(256)
count_runs_kernel(const unsigned long long* __restrict__ codes, uint64_t n, int shift, unsigned long long* __restrict__ total)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool head = i < n && (i == 0 || (codes[i] >> shift) != (codes[i - 1] >> shift));
    const unsigned m = __ballot_sync(0xffffffffu, head);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(total, (unsigned long long)__popc(m));
}
__global__ void __launch_bounds__(256)
mark_runs_kernel(const unsigned long long* __restrict__ codes, uint64_t n, int shift, uint32_t* __restrict__ flag)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || (codes[i] >> shift) != (codes[i - 1] >> shift)) ? 1u : 0u;
}
// start[group] = index of the group's first node (flag marks group heads; gid = exclusive scan of flag)
__global__ void __launch_bounds__(256)
group_starts_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ gid_excl, uint64_t n, uint32_t* __restrict__ start)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) start[gid_excl[i]] = (uint32_t)i;
}
// a node heads a parent when it heads its group or sits a multiple of 15 behind the group's head
__global__ void __launch_bounds__(256)
mark_parents_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ gid_excl, const uint32_t* __restrict__ start,
                    uint64_t n, uint32_t* __restrict__ head)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = flag[i] ? gid_excl[i] : gid_excl[i] - 1;       // exclusive scan: a head's own id, else the previous head's
    head[i] = (flag[i] || ((uint32_t)i - start[g]) % BVH_MAX_CHILD == 0) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
parent_firsts_kernel(const uint32_t* __restrict__ head, const uint32_t* __restrict__ pid_excl, uint64_t n, uint32_t* __restrict__ first)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && head[i]) first[pid_excl[i]] = (uint32_t)i;
}
// parent p bounds children first[p] .. first[p+1]-1 (behaviour of make_parents_detailed, bvh.cu:269-308)
__global__ void __launch_bounds__(256)
make_parent_layer_kernel(const uint4* __restrict__ child, const unsigned long long* __restrict__ child_codes, int shift,
                         const uint32_t* __restrict__ first, uint64_t nparents, uint64_t nchildren, uint4* __restrict__ parent,
                         unsigned long long* __restrict__ parent_codes)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nparents) return;
    const uint32_t lo_i = first[p], hi_i = (p + 1 < nparents) ? first[p + 1] : (uint32_t)nchildren;
    uint32_t lo[3] = {0xFFFFu, 0xFFFFu, 0xFFFFu}, hi[3] = {0u, 0u, 0u};
    for (uint32_t i = lo_i; i < hi_i; i++) {
        const uint4 c = child[i];
        const uint32_t w[3] = {c.x, c.y, c.z};
        for (int a = 0; a < 3; a++) { lo[a] = min(lo[a], w[a] & 0xFFFFu); hi[a] = max(hi[a], w[a] >> 16); }
    }
    parent[p] = make_uint4(lo[0] | (hi[0] << 16), lo[1] | (hi[1] << 16), lo[2] | (hi[2] << 16), ((hi_i - lo_i) << 28) | lo_i);
    parent_codes[p] = child_codes[lo_i] >> shift;
}
__global__ void __launch_bounds__(256)
shift_codes_kernel(unsigned long long* codes, uint64_t n, int shift)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) codes[i] >>= shift;
}
__global__ void __launch_bounds__(256)
gather_leaves_kernel(const uint4* __restrict__ leaves, const uint32_t* __restrict__ ids, uint64_t n, uint4* __restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = leaves[ids[i]];
}
// layer -> its place in the root-first array, child ids moved by the start of the layer below (copy_and_offset, bvh.cu:364-384)
__global__ void __launch_bounds__(256)
place_layer_kernel(const uint4* __restrict__ src, uint64_t n, uint32_t child_offset, uint4* __restrict__ dst)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 v = src[i];
    if (child_offset) v.w = (v.w & 0xF0000000u) | ((v.w & 0x0FFFFFFFu) + child_offset);
    dst[i] = v;
}
// nodes with a single child become that child (collapse_child, bvh.cu:530-543); layers are done bottom up
__global__ void __launch_bounds__(256)
collapse_layer_kernel(uint4* nodes, uint64_t start, uint64_t end)
{
    const uint64_t i = start + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const uint32_t w = nodes[i].w;
    if ((w >> 28) == 1) nodes[i] = nodes[w & 0x0FFFFFFFu];
}

struct DeviceLayer { uint4* nodes; uint64_t n; };

// leaves sorted by Morton code (device) -> finished node array on the host, layer offsets
static int build_layers_device(uint4* d_sorted_leaves, unsigned long long* d_codes, uint64_t ntriangles, int target_degree,
                               BvhResult& out)
{
    Context& c = ctx();
    cudaStream_t st = c.stream;
    std::vector<DeviceLayer> layers;                 // leaf first
    std::vector<void*> owned;
    uint32_t *d_flag = nullptr, *d_gid = nullptr, *d_start = nullptr, *d_head = nullptr, *d_pid = nullptr, *d_first = nullptr,
             *d_scan = nullptr;
    unsigned long long *d_total = nullptr, *d_pcodes = nullptr;
    uint4* d_final = nullptr;
    auto cleanup = [&]() {
        for (void* p : owned) cudaFree(p);
        cudaFree(d_flag); cudaFree(d_gid); cudaFree(d_start); cudaFree(d_head); cudaFree(d_pid); cudaFree(d_first);
        cudaFree(d_scan); cudaFree(d_total); cudaFree(d_pcodes); cudaFree(d_codes); cudaFree(d_final);      // both code buffers are ours
    };
#define BL(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #call); } } while (0)
    const uint64_t cap = ntriangles;
    BL(cudaMalloc(&d_flag, cap * 4)); BL(cudaMalloc(&d_gid, cap * 4)); BL(cudaMalloc(&d_start, (cap + 1) * 4));
    BL(cudaMalloc(&d_head, cap * 4)); BL(cudaMalloc(&d_pid, cap * 4)); BL(cudaMalloc(&d_first, (cap + 1) * 4));
    BL(cudaMalloc(&d_scan, exclusive_scan_scratch_words(cap) * 4)); BL(cudaMalloc(&d_total, 8));
    BL(cudaMalloc(&d_pcodes, cap * 8));
    auto grid = [](uint64_t n) { return (unsigned)((n + 255) / 256); };
    auto count_runs = [&](uint64_t n, int shift, uint64_t& runs) -> int {
        BL(cudaMemsetAsync(d_total, 0, 8, st));
        count_runs_kernel<<<grid(n), 256, 0, st>>>(d_codes, n, shift, d_total);
        unsigned long long h = 0;
        BL(cudaMemcpyAsync(&h, d_total, 8, cudaMemcpyDeviceToHost, st));
        BL(stream_wait(st));
        runs = h;
        return CB_OK;
    };
    layers.push_back({d_sorted_leaves, ntriangles});
    uint64_t n = ntriangles;
    unsigned long long* codes = d_codes;
    while (n > 1) {
        // drop low Morton bits until the mean fan-out reaches the target (bvh/grid.py:41-45)
        int shift = 0;
        uint64_t runs = 0;
        int rc = count_runs(n, 0, runs);
        if (rc) return rc;
        while ((double)n / (double)runs < (double)target_degree && runs > 1) {
            shift++;
            if ((rc = count_runs(n, shift, runs))) return rc;
        }
        mark_runs_kernel<<<grid(n), 256, 0, st>>>(codes, n, shift, d_flag);
        exclusive_scan(d_flag, d_gid, n, d_scan, st);
        group_starts_kernel<<<grid(n), 256, 0, st>>>(d_flag, d_gid, n, d_start);
        mark_parents_kernel<<<grid(n), 256, 0, st>>>(d_flag, d_gid, d_start, n, d_head);
        exclusive_scan(d_head, d_pid, n, d_scan, st);
        uint32_t last_pid = 0, last_head = 0;
        BL(cudaMemcpyAsync(&last_pid, d_pid + (n - 1), 4, cudaMemcpyDeviceToHost, st));
        BL(cudaMemcpyAsync(&last_head, d_head + (n - 1), 4, cudaMemcpyDeviceToHost, st));
        BL(stream_wait(st));
        const uint64_t nparents = (uint64_t)last_pid + last_head;
        parent_firsts_kernel<<<grid(n), 256, 0, st>>>(d_head, d_pid, n, d_first);
        uint4* d_parent = nullptr;
        BL(cudaMalloc(&d_parent, nparents * sizeof(uint4)));
        owned.push_back(d_parent);
        make_parent_layer_kernel<<<grid(nparents), 256, 0, st>>>(layers.back().nodes, codes, shift, d_first, nparents, n, d_parent,
                                                                 d_pcodes);
        BL(cudaGetLastError());
        // the parents' codes become the next round's codes (d_codes and d_pcodes swap roles)
        std::swap(d_codes, d_pcodes);
        codes = d_codes;
        layers.push_back({d_parent, nparents});
        n = nparents;
    }
    // root first
    const size_t nl = layers.size();
    std::vector<uint64_t> bounds(nl + 1, 0);
    for (size_t l = 0; l < nl; l++) bounds[l + 1] = bounds[l] + layers[nl - 1 - l].n;
    if (bounds[nl] >= (1ull << 28)) { cleanup(); return fail(CB_ERR_INVALID, "cb_bvh_build: %llu nodes exceed the 28-bit child field", (unsigned long long)bounds[nl]); }
    BL(cudaMalloc(&d_final, bounds[nl] * sizeof(uint4)));
    for (size_t l = 0; l < nl; l++) {
        const DeviceLayer& L = layers[nl - 1 - l];
        const uint32_t child_offset = (l + 1 < nl) ? (uint32_t)bounds[l + 1] : 0u;     // leaves point at triangles
        place_layer_kernel<<<grid(L.n), 256, 0, st>>>(L.nodes, L.n, child_offset, d_final + bounds[l]);
    }
    for (size_t l = nl - 1; l-- > 0;)                       // every layer but the leaves, bottom up
        collapse_layer_kernel<<<grid(bounds[l + 1] - bounds[l]), 256, 0, st>>>(d_final, bounds[l], bounds[l + 1]);
    BL(cudaGetLastError());
    out.nodes.resize(bounds[nl] * 4);
    BL(cudaMemcpyAsync(out.nodes.data(), d_final, bounds[nl] * sizeof(uint4), cudaMemcpyDeviceToHost, st));
    BL(stream_wait(st));
#undef BL
    out.layer_offsets.assign(bounds.begin(), bounds.end() - 1);
    cleanup();
    return CB_OK;
}

} // namespace cb

// Host-only: the engine's traversal tree for a reference-format tree (test / tooling hook;
// needs no GPU).  Two-call protocol: out_nodes == NULL returns the entry count only.
static int native_tree_two_call(const uint32_t* ref_nodes, uint64_t nnodes, uint64_t ntriangles,
                                const uint32_t* solid_id, const SplitInput* split, uint32_t* out_nodes,
                                uint64_t* out_count)
{
    static std::vector<Entry> cached;
    if (!ref_nodes || nnodes == 0 || !out_count) return fail(CB_ERR_INVALID, "cb_native_tree_build: bad arguments");
    if (out_nodes && !cached.empty()) {
        memcpy(out_nodes, cached.data(), cached.size() * sizeof(Entry));
        *out_count = cached.size();
        cached.clear();
        cached.shrink_to_fit();
        return CB_OK;
    }
    std::vector<uint32_t> rank;
    reference_test_rank(ref_nodes, nnodes, ntriangles, rank);
    std::vector<Entry> leaves;
    collect_reference_leaves(ref_nodes, nnodes, ntriangles, rank, leaves);
    int rc = build_native_tree(leaves, solid_id, cached, split);
    if (rc != CB_OK) return rc;
    *out_count = cached.size();
    if (out_nodes) {
        memcpy(out_nodes, cached.data(), cached.size() * sizeof(Entry));
        cached.clear();
    }
    return CB_OK;
}

extern "C" int cb_native_tree_build(const uint32_t* ref_nodes, uint64_t nnodes, uint64_t ntriangles,
                                    const uint32_t* solid_id, uint32_t* out_nodes, uint64_t* out_count)
{
    return native_tree_two_call(ref_nodes, nnodes, ntriangles, solid_id, nullptr, out_nodes, out_count);
}

extern "C" int cb_native_tree_build_split(const uint32_t* ref_nodes, uint64_t nnodes, uint64_t ntriangles,
                                          const uint32_t* solid_id, const float* vertices,
                                          const uint32_t* triangles, const float world_origin[3],
                                          float world_scale, int32_t max_pieces, int32_t min_extent,
                                          float min_ratio, uint32_t* out_nodes, uint64_t* out_count)
{
    if (!vertices || !triangles || !world_origin || !(world_scale > 0.0f))
        return fail(CB_ERR_INVALID, "cb_native_tree_build_split: mesh and world grid are required");
    SplitInput s;
    s.vertices = vertices; s.triangles = triangles; s.scale = world_scale;
    for (int a = 0; a < 3; a++) s.origin[a] = world_origin[a];
    s.max_pieces = max_pieces; s.min_extent = min_extent; s.min_ratio = min_ratio;
    return native_tree_two_call(ref_nodes, nnodes, ntriangles, solid_id, &s, out_nodes, out_count);
}

extern "C" int cb_bvh_build(const float* vertices, uint64_t nvertices, const uint32_t* triangles,
                            uint64_t ntriangles, int32_t target_degree, float world_origin_out[3],
                            float* world_scale_out, uint32_t* nodes_out, uint64_t* nnodes_out,
                            uint64_t* layer_offsets_out, int32_t* nlayers_out)
{
    CB_REQUIRE_INIT();
    CB_SERIALISE();
    if (!vertices || !triangles || nvertices == 0 || ntriangles == 0)
        return fail(CB_ERR_INVALID, "cb_bvh_build: empty mesh");
    if (ntriangles >= (1ull << 28)) return fail(CB_ERR_INVALID, "cb_bvh_build: too many triangles");
    if (target_degree < 2) target_degree = 3;
    if (nodes_out && g_bvh.valid) {
        // second call of the two-call protocol: hand over the cached result
        memcpy(nodes_out, g_bvh.nodes.data(), g_bvh.nodes.size() * 4);
        if (layer_offsets_out) memcpy(layer_offsets_out, g_bvh.layer_offsets.data(), g_bvh.layer_offsets.size() * 8);
        if (nnodes_out) *nnodes_out = g_bvh.nodes.size() / 4;
        if (nlayers_out) *nlayers_out = (int32_t)g_bvh.layer_offsets.size();
        if (world_origin_out) memcpy(world_origin_out, g_bvh.origin, 12);
        if (world_scale_out) *world_scale_out = g_bvh.scale;
        g_bvh = BvhResult();
        return CB_OK;
    }
    // world coordinates (behaviour of gpu/bvh.py:42-47)
    float lo[3] = {vertices[0], vertices[1], vertices[2]}, hi[3] = {vertices[0], vertices[1], vertices[2]};
    for (uint64_t i = 1; i < nvertices; i++)
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], vertices[3 * i + a]);
            hi[a] = std::max(hi[a], vertices[3 * i + a]);
        }
    float extent = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    float scale = (float)((double)extent / 65534.0);

    Context& c = ctx();
    float* d_v = nullptr; uint32_t* d_t = nullptr; uint4* d_leaves = nullptr;
    unsigned long long *d_codes = nullptr, *d_codes2 = nullptr; uint32_t *d_ids = nullptr, *d_ids2 = nullptr;
    uint32_t* d_scratch = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_v); cudaFree(d_t); cudaFree(d_leaves); cudaFree(d_codes);
        cudaFree(d_codes2); cudaFree(d_ids); cudaFree(d_ids2); cudaFree(d_scratch);
    };
#define BV(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #call); } } while (0)
    BV(cudaMalloc(&d_v, nvertices * 12)); BV(cudaMalloc(&d_t, ntriangles * 12));
    BV(cudaMalloc(&d_leaves, ntriangles * 16)); BV(cudaMalloc(&d_codes, ntriangles * 8));
    BV(cudaMalloc(&d_codes2, ntriangles * 8)); BV(cudaMalloc(&d_ids, ntriangles * 4)); BV(cudaMalloc(&d_ids2, ntriangles * 4));
    BV(cudaMalloc(&d_scratch, radix_sort_scratch_words(ntriangles) * 4));
    BV(cudaMemcpyAsync(d_v, vertices, nvertices * 12, cudaMemcpyHostToDevice, c.stream));
    BV(cudaMemcpyAsync(d_t, triangles, ntriangles * 12, cudaMemcpyHostToDevice, c.stream));
    make_leaves_kernel<<<(unsigned)((ntriangles + 255) / 256), 256, 0, c.stream>>>(
        d_v, d_t, ntriangles, make_float3(lo[0], lo[1], lo[2]), scale, d_leaves, d_codes, d_ids);
    BV(cudaGetLastError());
    // stable LSD radix sort of (morton, triangle id), 6 passes of 8 bits (sort.cuh): ties keep ascending triangle order
    const bool in_alt = radix_sort_pairs<unsigned long long>(d_codes, d_codes2, d_ids, d_ids2, ntriangles, 48, d_scratch, c.stream);
    BV(cudaGetLastError());
    const unsigned long long* sorted_codes = in_alt ? d_codes2 : d_codes;
    const uint32_t* sorted_ids = in_alt ? d_ids2 : d_ids;
    // leaves into Morton order, then every layer above them, on the device; one read-back of the finished tree
    uint4* d_sorted_leaves = nullptr;
    BV(cudaMalloc(&d_sorted_leaves, ntriangles * 16));
    gather_leaves_kernel<<<(unsigned)((ntriangles + 255) / 256), 256, 0, c.stream>>>(d_leaves, sorted_ids, ntriangles, d_sorted_leaves);
    BV(cudaGetLastError());
    unsigned long long* d_level_codes = nullptr;      // build_layers_device swaps and frees code arrays of its own
    BV(cudaMalloc(&d_level_codes, ntriangles * 8));
    BV(cudaMemcpyAsync(d_level_codes, sorted_codes, ntriangles * 8, cudaMemcpyDeviceToDevice, c.stream));
    BV(stream_wait(c.stream));
#undef BV
    cleanup();
    g_bvh = BvhResult();
    {
        const int rc = build_layers_device(d_sorted_leaves, d_level_codes, ntriangles, target_degree, g_bvh);
        cudaFree(d_sorted_leaves);
        if (rc != CB_OK) { g_bvh = BvhResult(); return rc; }
    }
    memcpy(g_bvh.origin, lo, 12);
    g_bvh.scale = scale;
    g_bvh.valid = true;
    if (nnodes_out) *nnodes_out = g_bvh.nodes.size() / 4;
    if (nlayers_out) *nlayers_out = (int32_t)g_bvh.layer_offsets.size();
    if (world_origin_out) memcpy(world_origin_out, lo, 12);
    if (world_scale_out) *world_scale_out = scale;
    if (nodes_out) {
        memcpy(nodes_out, g_bvh.nodes.data(), g_bvh.nodes.size() * 4);
        if (layer_offsets_out) memcpy(layer_offsets_out, g_bvh.layer_offsets.data(), g_bvh.layer_offsets.size() * 8);
        g_bvh = BvhResult();
    }
    return CB_OK;
}
