/*
 * chroma_b200.h -- C ABI of libchroma_b200.so, the B200-native (sm_100a) photon
 * transport engine that sits behind Chroma's Python API.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes only, returns
 * 0 on success or a negative CbStatus, and leaves a message retrievable with
 * cb_last_error().  The reference binds its device code through PyCUDA
 * (compile-at-run-time + driver API); each entry below names the reference
 * interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - "device pointer" arguments are CUDA device addresses obtained from
 *     cb_malloc() (or any other allocator of the same context).
 *   - host arrays are borrowed for the duration of the call only.
 *   - all calls are synchronous on return unless stated otherwise (the
 *     reference synchronises at the end of propagate/acquire/end_acquire:
 *     chroma/gpu/photon.py:290, chroma/gpu/daq.py:92,99).
 *   - one library instance drives one CUDA device (cb_init); one process per
 *     GPU is the multi-GPU model.
 */
#ifndef CHROMA_B200_H
#define CHROMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_ABI_VERSION 4

typedef enum {
    CB_OK = 0,
    CB_ERR_CUDA = -1,
    CB_ERR_INVALID = -2,
    CB_ERR_NOMEM = -3,
    CB_ERR_UNSUPPORTED = -4,
    CB_ERR_NCCL = -5
} CbStatus;

typedef uint64_t cb_geom_t;
typedef uint64_t cb_rng_t;
typedef uint64_t cb_daq_t;

/* Surface models, chroma/cuda/geometry_types.h:22 */
enum { CB_SURFACE_DEFAULT = 0, CB_SURFACE_COMPLEX = 1, CB_SURFACE_WLS = 2,
       CB_SURFACE_DICHROIC = 3, CB_SURFACE_ANGULAR = 4 };

/* History bits, chroma/cuda/photon.h:53-68 (NAN_ABORT is bit 15 in the kernel). */
enum {
    CB_NO_HIT = 0x1, CB_BULK_ABSORB = 0x2, CB_SURFACE_DETECT = 0x4,
    CB_SURFACE_ABSORB = 0x8, CB_RAYLEIGH_SCATTER = 0x10, CB_REFLECT_DIFFUSE = 0x20,
    CB_REFLECT_SPECULAR = 0x40, CB_SURFACE_REEMIT = 0x80, CB_SURFACE_TRANSMIT = 0x100,
    CB_BULK_REEMIT = 0x200, CB_CHERENKOV = 0x400, CB_SCINTILLATION = 0x800,
    CB_NAN_ABORT = 0x8000
};

/*
 * Optical tables.  All wavelength/time tables live in ONE float pool
 * (`table_pool`); materials and surfaces reference it by float offset.  This
 * replaces the per-table device allocations + pointer structs assembled by
 * make_gpu_struct (chroma/gpu/geometry.py:44-341, chroma/gpu/tools.py:207-229;
 * device structs chroma/cuda/geometry_types.h:4-79).  Offsets < 0 mean
 * "absent".  Per-component tables are stored as [num_comp][n] blocks.
 */
typedef struct {
    int32_t refractive_index;       /* [wavelength_n] */
    int32_t absorption_length;      /* [wavelength_n] */
    int32_t scattering_length;      /* [wavelength_n] */
    int32_t num_comp;
    int32_t comp_reemission_prob;      /* [num_comp][wavelength_n] */
    int32_t comp_reemission_wvl_cdf;   /* [num_comp][wavelength_n] */
    int32_t comp_reemission_time_cdf;  /* [num_comp][time_n]       */
    int32_t comp_absorption_length;    /* [num_comp][wavelength_n] */
} CbMaterial;

typedef struct {
    int32_t detect, absorb, reemit, reflect_diffuse, reflect_specular, eta, k,
            reemission_cdf;         /* each [wavelength_n] */
    int32_t model;                  /* CB_SURFACE_*; -1 = null surface slot */
    int32_t transmissive;
    float   thickness;
    /* DichroicProps (geometry_types.h:24-30): angles[n], reflect[n][W], transmit[n][W] */
    int32_t dichroic_nangles, dichroic_angles, dichroic_reflect, dichroic_transmit;
    /* AngularProps (geometry_types.h:32-39): 4 arrays of [n] */
    int32_t angular_nangles, angular_angles, angular_transmit,
            angular_reflect_specular, angular_reflect_diffuse;
} CbSurface;

/* Analytic wire plane: a periodic row of parallel cylinders of radius `radius`, axes along
 * `u`, centres at v0 + k*pitch along `v`, clipped to [umin,umax] x [vmin,vmax]
 * (= struct WirePlane, chroma/cuda/geometry_types.h:42-58; filled like
 * chroma/gpu/geometry.py:343-387).  The photon step picks the nearer of the mesh hit
 * and the wire hit (chroma/cuda/photon.h:96-300). */
typedef struct {
    float origin[3], u[3], v[3];
    float pitch, radius, umin, umax, vmin, vmax, v0;
    int32_t surface_index;          /* planes with surface_index < 0 never win (photon.h:272) */
    int32_t material_outer_index, material_inner_index;
    uint32_t color;
} CbWirePlane;

/*
 * Flattened geometry handed to the engine; mirrors what GPUGeometry.__init__
 * uploads (chroma/gpu/geometry.py:389-520) and `struct Geometry`
 * (chroma/cuda/geometry_types.h:124-139).  `nodes` is the reference uint4
 * packing (x,y,z = lo16 | hi16<<16, w = nchild<<28 | child; root = node 0);
 * the library derives its own traversal layout from it.
 */
typedef struct {
    const float*    vertices;        uint64_t nvertices;   /* [nvertices][3]  */
    const uint32_t* triangles;       uint64_t ntriangles;  /* [ntriangles][3] */
    const uint32_t* material_codes;  /* [ntriangles] m1<<24 | m2<<16 | surf<<8 */
    const uint32_t* solid_id;        /* [ntriangles] or NULL */
    const uint32_t* colors;          /* [ntriangles] or NULL */
    const uint32_t* nodes;           uint64_t nnodes;      /* [nnodes][4] */
    float world_origin[3];
    float world_scale;
    const float*      table_pool;    uint64_t table_floats;
    const CbMaterial* materials;     int32_t nmaterials;
    const CbSurface*  surfaces;      int32_t nsurfaces;
    int32_t wavelength_n; float wavelength_start, wavelength_step;
    int32_t time_n;       float time_start, time_step;
    int32_t nwireplanes;
    const CbWirePlane* wireplanes;   /* [nwireplanes] or NULL */
} CbGeometryDesc;

/* Device-side views of what the library uploaded (for the GPUGeometry mirror). */
typedef struct {
    void* vertices; void* triangles; void* material_codes; void* colors;
    void* solid_id_map; void* nodes;           /* reference-layout copies */
    void* solid_id_to_channel_index; void* time_cdf_x; void* time_cdf_y;
    void* charge_cdf_x; void* charge_cdf_y;    /* NULL until cb_detector_attach */
    uint64_t nvertices, ntriangles, nnodes;
    int32_t  nchannels;
    uint64_t device_bytes;                     /* total bytes held by this geometry */
    uint32_t max_stack_depth;                  /* deepest traversal stack the tree can need */
} CbGeometryInfo;

/* Photon bank: struct-of-arrays of DEVICE pointers, same arrays as GPUPhotons
 * (chroma/gpu/photon.py:46-62).  pos/dir/pol are packed float3 (12 B). */
typedef struct {
    float*    pos; float* dir; float* pol;
    float*    wavelengths; float* t;
    int32_t*  last_hit_triangles;
    uint32_t* flags;
    float*    weights;
    uint32_t* evidx;
    uint64_t  n;
} CbPhotonBank;

typedef struct {
    uint64_t photons;        /* photons processed */
    uint64_t steps;          /* total propagation steps taken */
    uint64_t nodes_visited;  /* BVH node boxes tested (0 unless stats build) */
    uint64_t tris_tested;    /* triangle tests        (0 unless stats build) */
    uint64_t rays_resolved;  /* rays redone in reference visit order (0 unless stats build) */
    uint32_t launches;       /* kernels launched by this call */
    float    kernel_ms;      /* device time of those kernels (CUDA events) */
    float    intersect0_ms;  /* device time of the first step's traversal kernel (0 if not launched) */
    uint64_t intersect0_rays;/* rays that kernel traced */
    /* device time per kernel class (CUDA events around every launch on the library stream) and
     * the work each class did, for the per-kernel rooflines of bench.py */
    float    intersect_ms;   /* all wavefront traversal launches */
    float    physics_ms;     /* all wavefront physics launches */
    float    tail_ms;        /* persistent tail launches */
    uint64_t intersect_rays; /* rays traced by the wavefront traversal launches */
    uint64_t physics_steps;  /* photon steps taken by the wavefront physics launches */
    uint64_t tail_photons;   /* photons handed to the tail */
    uint64_t tail_steps;     /* photon steps taken by the tail */
} CbPropagateStats;

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

/* ---- runtime --------------------------------------------------------- */
/* replaces chroma/gpu/tools.py:182-203 create_cuda_context */
int         cb_init(int device);
int         cb_device_count(void);
int         cb_abi_version(void);
const char* cb_last_error(void);
int         cb_synchronize(void);
int         cb_sm_count(void);
/* PCI bus id of the bound device ("0000:1b:00.0"), e.g. to find the same GPU in NVML */
int         cb_device_pci_bus_id(char* out, int32_t len);

/* replaces pycuda.driver.mem_alloc / memcpy_* / GPUArray.fill */
int cb_malloc(uint64_t bytes, void** dptr);
int cb_free(void* dptr);
int cb_memcpy_h2d(void* dptr, const void* hptr, uint64_t bytes);
int cb_memcpy_d2h(void* hptr, const void* dptr, uint64_t bytes);
int cb_memcpy_d2d(void* dst, const void* src, uint64_t bytes);
int cb_memset32(void* dptr, uint32_t value, uint64_t count);
int cb_host_alloc(uint64_t bytes, void** hptr);   /* pinned */
int cb_host_alloc_flags(uint64_t bytes, int32_t write_combined, void** hptr);   /* pinned; optionally write-combined (upload-only buffers) */
int cb_host_free(void* hptr);
int cb_mem_info(uint64_t* free_bytes, uint64_t* total_bytes);

/* device timers (CUDA events on the library stream) for bench.py */
int cb_timer_start(void);
int cb_timer_stop(float* ms);
int cb_flush_l2(void);     /* writes a buffer larger than L2 */

/* ---- geometry / detector --------------------------------------------- */
/* replaces GPUGeometry.__init__ (chroma/gpu/geometry.py:14-526) */
int cb_geometry_create(const CbGeometryDesc* desc, cb_geom_t* out);
int cb_geometry_destroy(cb_geom_t g);
int cb_geometry_info(cb_geom_t g, CbGeometryInfo* info);
/* replaces GPUDetector.__init__ (chroma/gpu/detector.py:15-39), struct Detector (cuda/detector.h:4-22) */
int cb_detector_attach(cb_geom_t g, const int32_t* solid_id_to_channel_index,
                       uint64_t nsolids, int32_t nchannels,
                       const float* time_cdf_x, const float* time_cdf_y, int32_t time_cdf_len,
                       const float* charge_cdf_x, const float* charge_cdf_y, int32_t charge_cdf_len,
                       float charge_unit);

/* ---- BVH construction ------------------------------------------------ */
/* replaces make_recursive_grid_bvh (chroma/bvh/grid.py:11-95) and its kernels
 * make_leaves / make_parents_detailed / copy_and_offset / collapse_child
 * (chroma/cuda/bvh.cu:148,269,364,530).  Two-call protocol: with nodes_out ==
 * NULL only *nnodes_out / *nlayers_out are written. */
int cb_bvh_build(const float* vertices, uint64_t nvertices,
                 const uint32_t* triangles, uint64_t ntriangles,
                 int32_t target_degree,
                 float world_origin_out[3], float* world_scale_out,
                 uint32_t* nodes_out, uint64_t* nnodes_out,
                 uint64_t* layer_offsets_out, int32_t* nlayers_out);

/* The engine's own traversal tree (SAH, <= 8 children, two-level by solid) over the
 * leaves of a reference-format tree; same entry packing, root at entry 0.  Host
 * only (no GPU needed); cb_geometry_create does this internally.  Two-call
 * protocol like cb_bvh_build.  No reference counterpart (the reference traverses
 * its build tree directly, chroma/cuda/mesh.h:45-126). */
int cb_native_tree_build(const uint32_t* ref_nodes, uint64_t nnodes, uint64_t ntriangles,
                         const uint32_t* solid_id, uint32_t* out_nodes, uint64_t* out_count);
/* Same, with leaf splitting: a triangle whose box is much larger than the triangle (long
 * and oblique to the axes) is referenced by up to max_pieces leaves, each bounding the part
 * of the triangle inside one cell of its leaf box (boxes on the world grid of
 * world_origin / world_scale).  min_extent: shortest box side (grid quanta) worth splitting;
 * min_ratio: split while the box surface exceeds min_ratio x the surface of a tight box.
 * cb_geometry_create does the same when CHROMA_B200_LEAF_SPLIT=max_pieces[,min_extent[,min_ratio]]
 * is set (default: one leaf per triangle).  NOT exact with respect to the reference: a ray that grazes a
 * sliver can get the reference's (float32-spurious) hit on it only through the sliver's full leaf box;
 * measured rate 0 of 2 M random rays, 1 of 40,000 rays aimed at corners / edges. */
int cb_native_tree_build_split(const uint32_t* ref_nodes, uint64_t nnodes, uint64_t ntriangles,
                               const uint32_t* solid_id, const float* vertices,
                               const uint32_t* triangles, const float world_origin[3],
                               float world_scale, int32_t max_pieces, int32_t min_extent,
                               float min_ratio, uint32_t* out_nodes, uint64_t* out_count);

/* ---- RNG ---------------------------------------------------------------- */
/* replaces get_rng_states / init_rng (chroma/gpu/tools.py:117-145,
 * chroma/cuda/random.h:60-70): state i == curand_init(seed, i, offset). */
int cb_rng_create(uint64_t n, uint64_t seed, uint64_t offset, cb_rng_t* out);
int cb_rng_destroy(cb_rng_t r);
int cb_rng_size(cb_rng_t r, uint64_t* n);
/* Same with a stream base: state i == curand_init(seed, first_stream + i, offset).  With
 * first_stream = the global index of a rank's first photon, stream id == global photon index,
 * and results do not depend on how the photons are partitioned over GPUs.  The
 * reference has one device and always starts at stream 0 (chroma/cuda/random.h:60-70). */
int cb_rng_create_streams(uint64_t n, uint64_t seed, uint64_t first_stream, uint64_t offset, cb_rng_t* out);
/* Non-owning window [first, first+count) of a pool, usable wherever a pool is (propagate,
 * acquire): one event's slice of a run-level pool.  Destroy it before its parent. */
int cb_rng_view(cb_rng_t parent, uint64_t first, uint64_t count, cb_rng_t* out);
/* test hooks: state words {d, v0..v4} and fill_uniform (chroma/cuda/random.h:72-82) */
int cb_rng_download(cb_rng_t r, uint64_t first, uint64_t count, uint32_t* out6);
int cb_rng_fill_uniform(cb_rng_t r, uint64_t n, float low, float high, float* d_out);

/* ---- ray intersection -------------------------------------------------- */
/* replaces distance_to_mesh / intersect_mesh (chroma/cuda/mesh.h:45-155); also
 * returns the triangle index (-1 = miss; distance untouched on miss, as in the
 * reference).  d_last_hit may be NULL. All pointers are device pointers. */
int cb_intersect(cb_geom_t g, const float* d_origins, const float* d_directions,
                 const int32_t* d_last_hit, uint64_t n,
                 int32_t* d_triangle_out, float* d_distance_out);

/* ---- propagation -------------------------------------------------------- */
/* replaces GPUPhotons.propagate + the propagate kernel
 * (chroma/gpu/photon.py:227-290, chroma/cuda/propagate.cu:254-366).
 * nthreads_per_block*max_blocks is NOT a launch shape here; it only has to
 * match the rng pool contract (pool >= n gives photon i <-> stream i). */
int cb_propagate(const CbPhotonBank* bank, cb_geom_t g, cb_rng_t rng,
                 int32_t nthreads_per_block, int32_t max_blocks, int32_t max_steps,
                 int32_t use_weights, int32_t scatter_first, CbPropagateStats* stats);

/* ---- photon bank utilities (chroma/cuda/propagate.cu:29-251) ------------ */
/* GPUPhotons.__init__'s uploads (chroma/gpu/photon.py:46-62: nine gpuarray.to_gpu calls) as one call: `host`
 * holds HOST pointers (page-locked for full speed); a NULL t / last_hit_triangles / flags / weights / evidx is
 * filled with the constructor's default instead (0, -1, 0, 1.0f, evidx_value) by the copy engine. */
int cb_photon_bank_upload(const CbPhotonBank* dst, const CbPhotonBank* host, uint64_t n, uint32_t evidx_value);
int cb_photon_duplicate(const CbPhotonBank* bank, uint64_t nphotons, int32_t ncopies);
int cb_count_photons(const CbPhotonBank* bank, uint64_t first, uint64_t n,
                     uint32_t target_flag, uint32_t* count_out);
int cb_copy_photons(const CbPhotonBank* src, uint64_t first, uint64_t n,
                    uint32_t target_flag, const CbPhotonBank* dst, uint32_t* count_out);
int cb_count_photon_hits(const CbPhotonBank* bank, uint64_t first, uint64_t n,
                         uint32_t target_flag, cb_geom_t g, uint32_t* count_out);
int cb_copy_photon_hits(const CbPhotonBank* src, uint64_t first, uint64_t n,
                        uint32_t target_flag, cb_geom_t g, const CbPhotonBank* dst,
                        int32_t* d_channels_out, uint32_t* count_out);
int cb_copy_photon_queue(const CbPhotonBank* src, const uint32_t* d_queue, uint64_t n,
                         const CbPhotonBank* dst);
/* get_flat_hits without a host round trip (chroma/gpu/photon.py:141-209 reads the hit count back between
 * its two kernels): count, scan and scatter are only ENQUEUED on the library stream.  d_block receives the ten
 * hit arrays back to back, each as long as the number of hits H (words per hit: pos 3, dir 3, pol 3,
 * wavelength, t, last_hit_triangle, flags, weight, evidx, channel = 16; array k starts at word offset[k] * max(H, 1));
 * it must hold 16 * n words.  H goes to d_count_out (device).  Read both after cb_event_wait on an event
 * recorded behind this call. */
int cb_copy_photon_hits_async(const CbPhotonBank* src, uint64_t first, uint64_t n, uint32_t target_flag,
                              cb_geom_t g, uint32_t* d_block, uint32_t* d_count_out);

/* ---- completion events for the *_async entry points ---------------------- */
typedef uint64_t cb_event_t;
int cb_event_create(cb_event_t* out);
int cb_event_record(cb_event_t e);      /* marks "everything enqueued on the library stream so far" */
int cb_event_wait(cb_event_t e);        /* host wait; naps between polls, never spins on a core */
int cb_event_destroy(cb_event_t e);

/* ---- DAQ (chroma/gpu/daq.py:37-101, chroma/cuda/daq.cu) ----------------- */
int cb_daq_create(cb_geom_t g, int32_t ndaq, cb_daq_t* out);
int cb_daq_destroy(cb_daq_t d);
int cb_daq_begin_acquire(cb_daq_t d);
int cb_daq_acquire(cb_daq_t d, const CbPhotonBank* bank, cb_rng_t rng,
                   int32_t nthreads_per_block, int32_t max_blocks,
                   uint64_t start_photon, uint64_t nphotons, float weight);
int cb_daq_end_acquire(cb_daq_t d);
/* begin_acquire (if `begin`), acquire and the float conversion of end_acquire (if `finalize`) of ONE
 * acquisition (chroma/gpu/daq.py:53-101), enqueued on the library stream without a host wait: the
 * simulation pipeline issues it right behind cb_propagate and moves on to the next event; results are
 * valid after cb_event_wait on an event recorded behind this call. */
int cb_daq_acquire_async(cb_daq_t d, const CbPhotonBank* bank, cb_rng_t rng,
                         int32_t nthreads_per_block, int32_t max_blocks,
                         uint64_t start_photon, uint64_t nphotons, float weight,
                         int32_t begin, int32_t finalize);
/* device pointers of earliest_time(float), q(float), flags(u32), and the raw
 * integer accumulators time_int(u32), q_int(u32); each [nchannels*ndaq] */
int cb_daq_pointers(cb_daq_t d, void** t, void** q, void** flags,
                    void** time_int, void** q_int, uint64_t* count);
/* finalise from integer accumulators only (after a cross-GPU reduction) */
int cb_daq_finalize(cb_daq_t d);
/* dst <- dst (+) src on the device, channel by channel: MIN of the time words, SUM of the integer
 * charges, OR of the histories -- what the atomics of run_daq would have produced had src's photons been
 * acquired into dst (chroma/cuda/daq.cu:73-75).  Folds per-event acquisitions into run-level accumulators. */
int cb_daq_fold(cb_daq_t dst, cb_daq_t src);
/* the same, enqueued only (ordered on the library stream; no host wait, does not take the library lock) */
int cb_daq_fold_async(cb_daq_t dst, cb_daq_t src);

/* ---- multi-GPU: one process per GPU, photon banks partitioned, geometry replicated ------
 * The reference has no multi-GPU path; what it does with atomics on the per-channel arrays of
 * ONE device (atomicMin / atomicAdd / atomicOr, chroma/cuda/daq.cu:73-75, 143-145) becomes one
 * reduction over NVLink once photons are sharded.
 *   cb_comm_unique_id : rank 0 creates the 128-byte NCCL id; the caller carries it to the other
 *                       ranks (any side channel: torch.distributed, MPI, a file)
 *   cb_comm_init      : every rank joins (collective)
 *   cb_daq_allreduce  : earliest_time_int -> MIN, channel_q_int -> SUM (uint32 wrap-around like
 *                       atomicAdd), channel_history -> OR (NCCL has no OR: the bits of the history
 *                       word travel as per-bit counters packed into the SUM buffer), ONE grouped pair
 *                       of ncclAllReduce on the library stream, the float conversion of
 *                       cb_daq_finalize fused behind it; no host copy.  Result on every rank.
 * NCCL is loaded at cb_comm_init time (libnccl.so.2); single-GPU use does not need it. */
#define CB_COMM_ID_BYTES 128
int cb_comm_unique_id(void* id_out);
int cb_comm_init(int32_t nranks, int32_t rank, const void* id);
int cb_comm_destroy(void);
int cb_comm_size(int32_t* nranks, int32_t* rank);     /* 1, 0 without a communicator */
int cb_daq_allreduce(cb_daq_t d);
/* test hook: the same pack / SUM / MIN / unpack + finalise for n accumulators on ONE device standing in
 * for n ranks; result in daqs[0] */
int cb_daq_reduce_local(const cb_daq_t* daqs, int32_t n);
/* host waits (stream / event synchronisation inside the entry points): 0 = spin (lowest latency,
 * one busy core per waiting thread), 1 = block (yield the core; for many ranks per host).
 * Default: CHROMA_B200_SYNC=spin|block, else spin. */
int cb_set_blocking_sync(int32_t on);

/* ---- host-side mesh preparation (no GPU needed) --------------------------
 * Vertex de-duplication of Geometry.flatten / Mesh.remove_duplicate_vertices
 * (chroma/geometry.py:59-69 = np.unique over vertex rows + inverse): unique rows
 * in lexicographic float order into unique_out [<= n][3], inverse_out[i] = row of
 * vertex i.  All host threads.                                               */
int cb_unique_vertices(const float* vertices, uint64_t n, float* unique_out,
                       uint32_t* inverse_out, uint64_t* nunique_out);

/* ---- PDF / likelihood accumulators (chroma/gpu/pdf.py, chroma/cuda/pdf.cu) -
 * All pointers are DEVICE pointers ([nchannels] unless stated); `t` / `q` are the
 * arrays a DAQ acquisition leaves behind (cb_daq_pointers), ndaq copies one after
 * the other.  Each call accumulates ONE acquisition into the caller's arrays.   */
/* GPUPDF.add_hits_to_pdf (gpu/pdf.py:201-217, bin_hits pdf.cu:9-32):
 * pdf is [nchannels][tbins][qbins], row major                                   */
int cb_pdf_bin_hits(int32_t nchannels, const float* q, const float* t, uint32_t* hitcount,
                    int32_t tbins, float tmin, float tmax, int32_t qbins, float qmin, float qmax,
                    uint32_t* pdf);
/* GPUKernelPDF.accumulate_moments (gpu/pdf.py:44-61, pdf.cu:223-266)            */
int cb_pdf_accumulate_moments(int32_t time_only, int32_t nchannels, const float* mc_time,
                              const float* mc_charge, float tmin, float tmax, float qmin, float qmax,
                              uint32_t* mom0, float* t_mom1, float* t_mom2, float* q_mom1, float* q_mom2);
/* GPUKernelPDF.accumulate_kernel (gpu/pdf.py:140-160, pdf.cu:271-368)           */
int cb_pdf_accumulate_kernel_eval(int32_t time_only, int32_t nchannels, const uint32_t* event_hit,
                                  const float* event_time, const float* event_charge,
                                  const float* mc_time, const float* mc_charge,
                                  float tmin, float tmax, float qmin, float qmax,
                                  const float* inv_time_bandwidths, const float* inv_charge_bandwidths,
                                  uint32_t* hitcount, float* time_pdf_values, float* charge_pdf_values);
/* GPUPDF.accumulate_pdf_eval (gpu/pdf.py:297-330): accumulate_bincount (pdf.cu:34-96)
 * + accumulate_nearest_neighbor_block (pdf.cu:152-219) in ONE launch, without the
 * work-queue array and the host synchronisation between them.  mc_time is
 * [ndaq][nchannels]; nearest_mc is [nhit][min_bin_content], kept sorted ascending,
 * unused slots > 1e8; map_hit_offset_to_channel_id is [nhit].                   */
int cb_pdf_accumulate_eval(int32_t nchannels, int32_t ndaq, int32_t nhit, const uint32_t* event_hit,
                           const float* event_time, const float* mc_time, uint32_t* hitcount,
                           uint32_t* bincount, float min_twidth, float tmin, float tmax,
                           int32_t min_bin_content, const uint32_t* map_hit_offset_to_channel_id,
                           float* nearest_mc);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* CHROMA_B200_H */
