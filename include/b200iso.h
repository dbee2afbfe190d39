/* b200iso.h -- C ABI of libb200iso.so, the B200 (sm_100a) isosurface extractor that stands behind
 * Meshing.jl's `isosurface(sdf::AbstractArray{T,3}, method, X, Y, Z) -> (vertices, faces)`.
 *
 * The reference has no FFI layer: its boundary is Julia multiple dispatch on the `method` type
 * (src/marching_cubes.jl:27, src/marching_tetrahedra.jl:129, forwarder src/isosurface.jl:30-32).
 * This header is what a Julia shim `ccall`s instead of running those two loops (INTEGRATION.md shows
 * the binding).  Plain pointers and sizes only; no exceptions cross the boundary; every function
 * returns 0 on success or a negative B200ISO_E* code, with a message in b200iso_last_error().
 *
 * Conventions (identical to the reference):
 *   - the field is column-major, x contiguous: sample (x,y,z) (0-based) is sdf[x + ldx*(y + ny*z)]
 *     (Julia `sdf[x+1,y+1,z+1]`; ldx == nx for a dense Array, > nx for a padded x-slab);
 *   - voxels are visited x-outermost, z-innermost (src/marching_cubes.jl:40,
 *     src/marching_tetrahedra.jl:144); vertex and face order follow that scan exactly;
 *   - vertices are xyz triples (Vector{NTuple{3,T}} layout), T = Float32 or Float64 by the
 *     reference's promote_type rule (src/marching_cubes.jl:31, src/marching_tetrahedra.jl:131);
 *   - faces are triples of 1-based Int64 vertex indices (Vector{NTuple{3,Int}}).
 * The library never retains or frees caller memory.  A handle is not thread-safe; use one per thread.
 */
#ifndef B200ISO_H
#define B200ISO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200iso_handle b200iso_handle;

/* method type: replaces dispatch on MarchingCubes / MarchingTetrahedra (src/algorithmtypes.jl:23-40) */
enum { B200ISO_MC = 0, B200ISO_MT = 1 };
/* where a caller pointer lives */
enum { B200ISO_HOST = 0, B200ISO_DEVICE = 1 };
/* element type of the endpoints of X, Y, Z: Int (default -1:1 => LinRange{Float64} and no Float64
 * promotion of the vertex type), Float32, Float64 (src/marching_cubes.jl:31,36-38) */
enum { B200ISO_RANGE_INT = 0, B200ISO_RANGE_F32 = 1, B200ISO_RANGE_F64 = 2 };

enum {
  B200ISO_OK = 0,
  B200ISO_EINVAL = -1,   /* bad argument */
  B200ISO_ECUDA = -2,    /* CUDA runtime error (message has the CUDA string) */
  B200ISO_ENOMEM = -3,   /* device or host allocation failed */
  B200ISO_ESTATE = -4,   /* call order violated (e.g. generate before count) */
  B200ISO_ECAPACITY = -5 /* caller buffer too small for the mesh */
};

/* The `method` object and the positional X, Y, Z of the reference call, flattened.
 * iso/eps carry the VALUE; *_is_f32 carries typeof(method.iso)/typeof(method.eps)
 * (MarchingCubes(iso=0f0) vs the default iso=0.0::Float64, src/algorithmtypes.jl:23-25,37-40).
 * Only first(X), last(X) etc. matter (src/marching_cubes.jl:36-38). */
typedef struct b200iso_params {
  int32_t algo;       /* B200ISO_MC | B200ISO_MT */
  int32_t iso_is_f32; /* 1: typeof(iso) == Float32, 0: Float64 */
  int32_t eps_is_f32; /* MT only */
  int32_t range_kind; /* B200ISO_RANGE_* */
  double iso;
  double eps;         /* MT only (default 1e-3) */
  double x0, x1, y0, y1, z0, z1;
  /* x-slab sharding (MC): the field passed to count is the slab of samples
   * [x_offset, x_offset + nx) of a volume with nx_global samples along x, and x0/x1 are the endpoints of
   * the WHOLE volume, so that slab vertices get exactly the coordinates of the unsharded call
   * (LinRange(first(X), last(X), nx_global)[x_offset + i]).  Both 0 for an unsharded call. */
  int64_t x_offset;
  int64_t nx_global;
  /* element type of the field: 0 = Float32 (the fast path: TMA / 128-bit classify), 1 = Float64 (`sdf` then
   * points to doubles; vertices are Float64, promote_type(..., Float64)).  Integer / Float16 fields are not
   * accepted. */
  int32_t field_is_f64;
  /* Marching Tetrahedra x-slab sharding: 1 if the slab's first voxel row (samples x_offset, x_offset+1) is the
   * GHOST row, i.e. the last voxel row of the previous slab.  MT vertices are shared between voxels and created by
   * the owner voxel (the first one in scan order that touches the edge); faces of a slab's first own row reference
   * vertices owned by that previous row.  The ghost row is counted (so those references resolve locally) but not
   * emitted, totals exclude it, and face indices are vertex_base - ghost_vertices + local prefix -- so that the
   * stitched slabs equal the unsharded mesh.  Every MT slab with x_offset > 0 carries one; 0 otherwise. */
  int32_t x_ghost;
} b200iso_params;

/* ---- lifetime ------------------------------------------------------------------------------------ */
/* Creates a handle bound to CUDA device `device`; it owns a stream and all device scratch. */
int b200iso_create(b200iso_handle** out, int device);
int b200iso_destroy(b200iso_handle* h);
/* Thread-local message of the last failing call. */
const char* b200iso_last_error(void);
/* Library/ABI version (major*1000 + minor). */
int b200iso_version(void);
/* Run on the caller's CUDA stream (a cudaStream_t, used as is: NULL is CUDA's legacy default stream).
 * b200iso_use_own_stream switches back to the stream the handle created for itself. */
int b200iso_set_stream(b200iso_handle* h, void* cuda_stream);
int b200iso_use_own_stream(b200iso_handle* h);
/* Which classify (sign-pack) kernel a count uses.  mode: -1 = automatic (TMA-staged on big 16-byte aligned Float32
 * fields, per-lane loads otherwise; the default), 1 = TMA whenever the field can be described by a tensor map
 * (Float32, base 16-byte aligned, ldx % 4 == 0), 0 = never TMA.  Results are identical; the parity tests use this
 * to run every shape through both kernels.  The environment variable B200ISO_TMA=0|1 sets the initial mode.
 * b200iso_classify_path reports what the last count actually ran (B200ISO_CLASSIFY_*, -1 before the first). */
enum { B200ISO_CLASSIFY_LDG128 = 0, B200ISO_CLASSIFY_TMA = 1, B200ISO_CLASSIFY_SCALAR = 2, B200ISO_CLASSIFY_F64 = 3 };
int b200iso_set_classify_mode(b200iso_handle* h, int mode);
int b200iso_classify_path(b200iso_handle* h);
/* On the TMA classify path every classify CTA carries `warps` extra warps that count the generate blocks (Marching
 * Cubes or Marching Tetrahedra) from the finished rows of the bit-field while the field still streams (the classify
 * kernel is bound by HBM and leaves most issue slots idle); the count kernel behind it only takes what they did not get
 * to.  0 switches this off (count kernel only), default 6, at most 8 (6 for Marching Tetrahedra; 4 are used on classify
 * kernels of fewer than 8192 tasks); results are identical.
 * Environment: B200ISO_RIDE=<warps>; B200ISO_RIDE_MIN_TASKS=<n> (default 6144: smaller classify kernels end before
 * the rows they complete can be followed) lets tests send small grids through the counting warps.
 * b200iso_ride_claimed: how many generate blocks the riding warps counted in the last count (synchronises). */
int b200iso_set_ride_warps(b200iso_handle* h, int warps);
int64_t b200iso_ride_claimed(b200iso_handle* h);

/* ---- the drop-in pair: replaces the body of isosurface(sdf, method, X, Y, Z) ---------------------------
 * b200iso_count   : classify + count + scan.  `sdf` is Float32 (Float64 if p->field_is_f64), host or device (mem), nx*ny*nz samples with
 *                   leading dimension ldx.  Reports the mesh size and the vertex element type
 *                   (vert_is_f64: 0 => Float32 triples, 1 => Float64 triples) so the caller can allocate
 *                   `Vector{NTuple{3,T}}(undef, nverts)` / `Vector{NTuple{3,Int}}(undef, nfaces)`.
 * b200iso_generate: writes 3*nverts vertex scalars and 3*nfaces Int64 indices into caller memory
 *                   (host or device).  `vertex_base` is added to every face index: 0 for a whole
 *                   volume, the global index base of the slab's first vertex for an x-slab shard. */
int b200iso_count(b200iso_handle* h, const b200iso_params* p, const void* sdf, int mem, int64_t nx, int64_t ny,
                  int64_t nz, int64_t ldx, int64_t* nverts, int64_t* nfaces, int* vert_is_f64);
int b200iso_generate(b200iso_handle* h, void* verts, int64_t* faces, int mem, int64_t vertex_base);

/* ---- asynchronous device-resident form (no host synchronisation; for pipelines and sharding) ---------
 * b200iso_count_async   : enqueues classify/count/scan on the handle's stream; the totals
 *                         {nverts, nfaces} are written to totals_dev (device int64[2]) -- the buffer a
 *                         sharded caller hands to its NCCL all-gather.
 * b200iso_generate_async: enqueues generate into device buffers of capacity vcap vertices / fcap faces
 *                         (elements beyond capacity are dropped; check the totals afterwards).
 *                         The face index base is vertex_base + (vertex_base_dev ? *vertex_base_dev : 0),
 *                         read on the device when the kernel runs.
 * b200iso_totals        : synchronises the stream and returns the totals of the last count. */
int b200iso_count_async(b200iso_handle* h, const b200iso_params* p, const void* sdf_dev, int64_t nx, int64_t ny,
                        int64_t nz, int64_t ldx, int64_t* totals_dev);
int b200iso_generate_async(b200iso_handle* h, void* verts_dev, int64_t vcap, int64_t* faces_dev, int64_t fcap,
                           const int64_t* vertex_base_dev, int64_t vertex_base);
int b200iso_totals(b200iso_handle* h, int64_t* nverts, int64_t* nfaces, int* vert_is_f64);

/* ---- one-enqueue form (device-resident, capacity known up front) ----------------------------------------
 * b200iso_extract_async: the whole isosurface() in one enqueue -- classify, count + scan, generate back to
 *                        back on the handle's stream -- into device buffers of capacity
 *                        vcap vertices / fcap faces.  Elements beyond capacity are dropped, the true totals
 *                        are written to totals_dev (device int64[2], may be NULL) and kept for
 *                        b200iso_totals, so a caller that guessed too small re-allocates and calls again.
 *                        Face indices get vertex_base + (vertex_base_dev ? *vertex_base_dev : 0).
 * b200iso_add_vertex_base_async: adds *vertex_base_dev to the first min(totals_dev[1], fcap) faces -- the
 *                        sharded fix-up when the slab's global vertex base (from the all-gather of the
 *                        slabs' totals) becomes known only after the slab was extracted. */
int b200iso_extract_async(b200iso_handle* h, const b200iso_params* p, const void* sdf_dev, int64_t nx, int64_t ny,
                          int64_t nz, int64_t ldx, void* verts_dev, int64_t vcap, int64_t* faces_dev, int64_t fcap,
                          const int64_t* vertex_base_dev, int64_t vertex_base, int64_t* totals_dev);
int b200iso_add_vertex_base_async(b200iso_handle* h, int64_t* faces_dev, int64_t fcap, const int64_t* totals_dev,
                                  const int64_t* vertex_base_dev);

/* ---- sharded path: exchange of the slabs' totals over NVLink peer memory ------------------------------------
 * The sharded path's only exchange is 16 bytes per rank (nverts, nfaces of its x-slab).  Instead of
 * an NCCL all-gather it can run as plain peer-memory stores inside the handle's stream:
 * b200iso_set_peer_exchange: slots[r] = device pointer, valid on THIS device, of rank r's exchange buffer
 *                       (B200ISO_PEER_BYTES bytes, zeroed once, mapped into every rank: CUDA IPC or torch symmetric
 *                       memory; slots[rank] is this rank's own buffer).  world <= 1 or slots == NULL switches it off.
 * b200iso_exchange_async : after a count: publishes this rank's totals into every rank's buffer (release.sys P2P
 *                       stores over NVLink), waits for everyone's in the local buffer and writes
 *                       bases_dev[0..3] = {vertex base, face base of this rank, total nverts, total nfaces};
 *                       all_dev (may be NULL) receives every rank's (nverts, nfaces).  Pass bases_dev as
 *                       vertex_base_dev to b200iso_generate_async.  Every rank must call it once per count (the
 *                       calls are matched by an epoch counter).
 * b200iso_set_peer_timeout: how long a rank waits for its peers inside b200iso_exchange_async: seconds (default 60:
 *                       ordinary rank skew -- first-call allocations, data loading -- must not trip it); <= 0 waits
 *                       for ever, like an NCCL collective.  After a time-out the generate kernels enqueued behind the
 *                       exchange emit nothing, b200iso_totals returns B200ISO_ESTATE, and the exchange must be re-armed
 *                       on ALL ranks (b200iso_set_peer_exchange with re-zeroed buffers) before it is used again. */
#define B200ISO_PEER_MAX 16
#define B200ISO_PEER_BYTES (2 * B200ISO_PEER_MAX * 4 * 8)
int b200iso_set_peer_exchange(b200iso_handle* h, int rank, int world, void* const* slots);
int b200iso_set_peer_timeout(b200iso_handle* h, double seconds);
int b200iso_exchange_async(b200iso_handle* h, int64_t* bases_dev, int64_t* all_dev);

/* ---- one-shot host form (host arrays in, host arrays out, capacity known up front) ----------------------
 * b200iso_extract_host: the whole isosurface() of a HOST field into caller-owned HOST arrays of capacity vcap
 *                       vertices / fcap faces, as an x-slab software pipeline over three streams: strided H2D of
 *                       slab k+1, the kernels of slab k and the D2H of slab k-1's mesh (straight into its final
 *                       offset of verts/faces) overlap, so the call costs about the H2D time of the field instead
 *                       of H2D + kernels + D2H (src/marching_cubes.jl:40 scans x outermost, so every x-slab's
 *                       mesh is a contiguous piece of the output; MT slabs carry a ghost row).  Results are
 *                       byte-identical to b200iso_count + b200iso_generate.  nverts/nfaces receive the true
 *                       totals; if they exceed the capacities nothing useful is in the arrays and
 *                       B200ISO_ECAPACITY is returned -- re-allocate and call again (vcap = fcap = 0 is a pure
 *                       count).  Pinned (cudaHostAlloc / cudaHostRegister) arrays get the full PCIe rate;
 *                       pageable ones work at the driver's staged-copy rate.  A rank's slab of a sharded volume
 *                       (x_offset / nx_global / x_ghost set) is accepted; its face indices are slab-local
 *                       (add the slab's vertex base afterwards).  Synchronous: the mesh is in the arrays on return. */
int b200iso_extract_host(b200iso_handle* h, const b200iso_params* p, const void* sdf, int64_t nx, int64_t ny,
                         int64_t nz, int64_t ldx, void* verts, int64_t vcap, int64_t* faces, int64_t fcap,
                         int64_t* nverts, int64_t* nfaces, int* vert_is_f64);
/* b200iso_extract_host_resident: finishes a b200iso_extract_host call that returned B200ISO_ECAPACITY (or was a pure
 *                       count, vcap = fcap = 0) WITHOUT uploading the field again: the x-slabs are still resident on
 *                       the device, only the kernels (milliseconds) and the device->host copy of the mesh run.  This
 *                       is how the drop-in `isosurface(::Array)` works: one b200iso_extract_host into
 *                       arrays sized by a guess -- the previous call's totals or a surface-area estimate; the mesh
 *                       then streams out while the field still streams in -- and, only if the guess was short, this
 *                       call into arrays of the exact size.  B200ISO_ESTATE if nothing is resident (any other
 *                       host-field call on the handle in between invalidates the slabs). */
int b200iso_extract_host_resident(b200iso_handle* h, void* verts, int64_t vcap, int64_t* faces, int64_t fcap,
                                  int64_t* nverts, int64_t* nfaces);

/* ---- mesh consumers: what callers run next on the mesh (not part of Meshing.jl v0.7.0) ------------
 * b200iso_vertex_normals_async: unit normals of `nverts` vertices (device, Float32 or Float64 triples) from the gradient
 *                       of the field (device, same conventions as count: nx*ny*nz samples, leading dimension ldx; a
 *                       slab with x_offset / nx_global): trilinear blend of central differences, pointing towards
 *                       increasing field values (outwards for "negative inside"); float[3*nverts]; (0,0,0) where the
 *                       gradient vanishes.  Works for Marching Cubes and Marching Tetrahedra vertices alike.
 * b200iso_vertex_keys_async: after a Marching Cubes count, the grid-edge key of every vertex in output order
 *                       (3 * linear index of the edge's lower end node in the whole volume + axis), Int64[kcap].
 *                       Meshing.jl's Marching Cubes repeats the vertex of an edge in every voxel that touches it
 *                       (src/algorithmtypes.jl:18-19: "vertices may be repeated"); equal keys <=> same geometric vertex.
 * b200iso_weld        : the indexed (welded) form of a Marching Cubes mesh: keeps the first occurrence of every key in
 *                       output order, writes the kept vertices to verts_out_dev (capacity nverts) and the renumbered
 *                       faces (1-based into the welded vertices) to faces_out_dev (3*nfaces); `vertex_base` is what the
 *                       faces' indices carry on top of 1-based local ones (0 for an unsharded mesh).  Returns the
 *                       number of welded vertices.  Synchronous; all pointers are device pointers.
 * b200iso_write_ply / b200iso_write_stl: binary little-endian PLY (optional per-vertex normals) / binary STL of a mesh
 *                       in HOST memory (1-based Int64 faces as returned by the extraction). */
int b200iso_vertex_normals_async(b200iso_handle* h, const b200iso_params* p, const void* sdf_dev, int64_t nx, int64_t ny,
                                 int64_t nz, int64_t ldx, const void* verts_dev, int64_t nverts, int vert_is_f64,
                                 float* normals_dev);
int b200iso_vertex_keys_async(b200iso_handle* h, int64_t* keys_dev, int64_t kcap);
int b200iso_weld(b200iso_handle* h, const int64_t* keys_dev, const void* verts_dev, int64_t nverts, int vert_is_f64,
                 const int64_t* faces_dev, int64_t nfaces, int64_t vertex_base, void* verts_out_dev,
                 int64_t* faces_out_dev, int64_t* nwelded);
int b200iso_write_ply(const char* path, const void* verts, int64_t nverts, int vert_is_f64, const float* normals,
                      const int64_t* faces, int64_t nfaces);
int b200iso_write_stl(const char* path, const void* verts, int64_t nverts, int vert_is_f64, const int64_t* faces,
                      int64_t nfaces);

/* ---- parity / introspection ----------------------------------------------------------------------------
 * Per-voxel case index (_get_cubeindex, src/common.jl:10-20; corner order of the counted algo) for the
 * last counted field, (nx-1)(ny-1)(nz-1) bytes in scan-rank order ((x*(ny-1)+y)*(nz-1)+z). */
int b200iso_case_indices(b200iso_handle* h, uint8_t* out, int mem);
/* Device milliseconds of the stages of the last count/generate pair, measured with CUDA events on the
 * handle's stream: ms[0] classify (sign-pack), ms[1] count+scan, ms[2] generate, ms[3] host->device copy,
 * ms[4] device->host copy.  n = number of floats the caller provides (<= 5 are written).
 * Timing must be switched on first (it adds event records to the stream). */
int b200iso_enable_timing(b200iso_handle* h, int on);
int b200iso_timings(b200iso_handle* h, float* ms, int n);
/* Number of kernels launched by this handle since creation (bench.py reports it as gpu_launches). */
int64_t b200iso_launch_count(b200iso_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* B200ISO_H */
