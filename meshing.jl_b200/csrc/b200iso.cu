// b200iso.cu -- C ABI (include/b200iso.h) and host orchestration of the sm_100a isosurface kernels.
// Replaces the bodies of isosurface(::MarchingCubes) (src/marching_cubes.jl:27-64) and
// isosurface(::MarchingTetrahedra) (src/marching_tetrahedra.jl:129-163) of the reference.
#include "../../include/b200iso.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <thread>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <limits>

#include "iso_kernels.cuh"
#include "mt_kernels.cuh"
#include "count_kernel.cuh"
#include "signpack_tma.cuh"
#include "host_pipeline.h"
#include "mesh_consumers.cuh"
#include <cub/device/device_scan.cuh>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(e_ == cudaErrorMemoryAllocation ? B200ISO_ENOMEM : B200ISO_ECUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e_), __FILE__, __LINE__);                                    \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  int reserve(size_t n) {
    if (n <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr, cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(B200ISO_ENOMEM, "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e));
    }
    cap = n;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr, cap = 0;
  }
};

// Entry points run on the handle's device and put the caller's current device back afterwards (a host process that
// tracks the current device -- torch, CUDA.jl -- must not find it moved by a library call).
struct DeviceGuard {
  int prev = -1, dev;
  explicit DeviceGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// smallest float t such that for every float s: (double)s < iso  <=>  s < t   (strict compare,
// src/common.jl:11-18, after Julia's promotion of both sides to Float64).
float threshold_for(double iso, bool iso_is_f32) {
  if (iso_is_f32) return (float)iso;
  if (std::isnan(iso)) return std::numeric_limits<float>::quiet_NaN();
  float f = (float)iso;  // round to nearest
  if ((double)f < iso) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
  return f;
}

}  // namespace

struct b200iso_handle {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  DevBuf<uint32_t> bits;
  DevBuf<uint32_t> celloff;      // MT: per-cell vertex prefix inside its block
  // scan state, cleared by block 0 of the classify kernel: [0] = ticket, [2 ..) = look-back chain of the scan blocks
  DevBuf<unsigned long long> chain;
  DevBuf<unsigned long long> status;  // MT: inclusive (vertex, face) prefix of every generate block (FLAG_INC | value)
  DevBuf<uint32_t> recs, nrecs;       // active-voxel records of every generate block (REC_CAP each) and their number
  DevBuf<unsigned long long> woff;    // MC: exclusive (vertex, face) prefix of every generate block; MT: raw block totals
  DevBuf<double> coords;
  DevBuf<uint8_t> cases;             // b200iso_case_indices(HOST): device scratch, kept across calls
  DevBuf<unsigned long long> weld_tab;  // mesh consumers: hash table (keys | representatives), slots, flags, new indices, cub scratch
  DevBuf<unsigned int> weld_u32;
  DevBuf<unsigned char> weld_tmp;
  // counting warps inside the TMA classify kernel (signpack_tma.cuh): header | next_x[nbi] | rows_done[ny]
  DevBuf<unsigned int> ride;
  long long ride_ny = -1, ride_tpr = -1, ride_nbi = -1;  // shape the cumulative row counters belong to
  unsigned int ride_step = 0;
  // counting warps ride only in classify kernels of at least this many tasks (env B200ISO_RIDE_MIN_TASKS: tests).  Measured on
  // x-slabs of the 1024^3 gyroid (profiles/r2_tail_and_small_slab.txt): 257 planes (6144 tasks) 0.443 -> 0.427 ms per step with
  // 4 warps (0.433 with 6); 129 planes (4096 tasks) 0.263 -> 0.268: the rows complete too late for the warps to get much done.
  long long ride_min_tasks = 6144;
  int ride_stop = iso::TM_WARPS * iso::SP_ZW * 3 / 4;  // counting warps stop claiming at 3/4 of the CTA's classify work (env B200ISO_RIDE_STOP: z-words of 64)
  int ride_warps = 6;  // counting warps per TMA classify CTA (0 = separate count kernel only; b200iso_set_ride_warps, env B200ISO_RIDE)
  // grid coordinates are a function of the call's shape and ranges only: recomputed when those change
  struct CoordsKey {
    long long nx = -1, ny = -1, nz = -1, xoff = 0, nxg = 0;
    double r[6] = {0, 0, 0, 0, 0, 0};
    int f32 = 0;
    bool operator==(const CoordsKey& o) const {
      return nx == o.nx && ny == o.ny && nz == o.nz && xoff == o.xoff && nxg == o.nxg && f32 == o.f32 && memcmp(r, o.r, sizeof(r)) == 0;
    }
  } coords_key;
  DevBuf<unsigned char> field;   // staging of a host field
  DevBuf<unsigned char> vstage;  // staging of vertices for host output
  DevBuf<long long> fstage;
  DevBuf<unsigned char> vstage1;  // second staging set: the slab pipeline of b200iso_extract_host ping-pongs
  DevBuf<long long> fstage1;
  hostpipe::Pool pool;             // copy lanes (streams, pinned chunks, events) of the HOST paths
  std::vector<cudaEvent_t> slab_ev;  // b200iso_extract_host: slab k generated; [n-2] fork, [n-1] join
  // the x-slab sub-problems of the last b200iso_extract_host, resident in `field` (b200iso_extract_host_resident)
  struct HostSlab {
    b200iso_params prm;
    const void* dev = nullptr;
    int64_t nx = 0, ldx = 0;
  };
  std::vector<HostSlab> hs;
  int64_t hs_ny = 0, hs_nz = 0, hs_nv = 0, hs_nf = 0;
  int hs_vf64 = 0;
  long long* totals_dev = nullptr;   // device int64[4]: nverts, nfaces, peer-exchange error flag, spare
  // sharded path: peer exchange of the slab totals (b200iso_set_peer_exchange)
  iso::PeerSlots peers{};
  int peer_rank = 0, peer_world = 0;
  long long peer_epoch = 0;
  double peer_timeout_s = 60.0;  // b200iso_set_peer_timeout; <= 0 waits for ever (like NCCL)
  long long* totals_host = nullptr;  // pinned int64[2]
  // last counted problem
  bool counted = false, totals_known = false;
  b200iso_params prm{};
  iso::Grid grid{};
  const void* sdf_dev = nullptr;
  long long* totals_out = nullptr;
  long long nblocks = 0;
  int vert_is_f64 = 0;
  long long nverts = 0, nfaces = 0;
  // timing: a ring of per-step event sets, read back (averaged) only when b200iso_timings is called
  static constexpr int NSLOT = 256;
  enum { E_C0, E_C1, E_C2, E_G0, E_G1, E_H0, E_H1, E_D0, E_D1, E_N };
  bool timing = false;
  cudaEvent_t* ev = nullptr;          // [NSLOT][E_N], created lazily
  unsigned char* ev_set = nullptr;    // which events of a slot were recorded
  long long step = 0;                 // steps (count calls) since timing was enabled
  int64_t launches = 0;
  int tma_mode = -1;  // classify staging: -1 = TMA boxes on big fields (default), 1 = TMA whenever a tensor map exists, 0 = per-lane loads (b200iso_set_classify_mode; env B200ISO_TMA)
  int classify_path = -1;  // what the last count ran: B200ISO_CLASSIFY_*
  int rec(int which) {
    if (!timing) return 0;
    const int slot = (int)((step > 0 ? step - 1 : 0) % NSLOT);
    cudaError_t e = cudaEventRecord(ev[slot * E_N + which], stream);
    if (e != cudaSuccess) return fail(B200ISO_ECUDA, "cudaEventRecord: %s", cudaGetErrorString(e));
    ev_set[slot * E_N + which] = 1;
    return 0;
  }
  void begin_step() {
    if (!timing) return;
    ++step;
    const int slot = (int)((step - 1) % NSLOT);
    for (int i = 0; i < E_N; ++i) ev_set[slot * E_N + i] = 0;
  }
};

namespace {

int vertex_is_f64(const b200iso_params& p) {
  // float(promote_type(eltype(X), eltype(Y), eltype(Z), Float32, typeof(iso)[, typeof(eps)]))
  return p.field_is_f64 || (!p.iso_is_f32) || p.range_kind == B200ISO_RANGE_F64 || (p.algo == B200ISO_MT && !p.eps_is_f32);
}

int check_params(const b200iso_params* p, int64_t nx, int64_t ny, int64_t nz, int64_t ldx) {
  if (!p) return fail(B200ISO_EINVAL, "params is NULL");
  if (p->algo != B200ISO_MC && p->algo != B200ISO_MT) return fail(B200ISO_EINVAL, "unknown algo %d", p->algo);
  if (p->range_kind < 0 || p->range_kind > 2) return fail(B200ISO_EINVAL, "unknown range_kind %d", p->range_kind);
  if (nx < 0 || ny < 0 || nz < 0) return fail(B200ISO_EINVAL, "negative dimension");
  if (ldx < nx) return fail(B200ISO_EINVAL, "ldx (%lld) < nx (%lld)", (long long)ldx, (long long)nx);
  if (nx > 65536 || ny > 65536 || nz > 65536) return fail(B200ISO_EINVAL, "dimension larger than 65536 is not supported");
  if (p->nx_global != 0 || p->x_offset != 0 || p->x_ghost != 0) {
    if (p->x_offset < 0 || p->nx_global < p->x_offset + nx) return fail(B200ISO_EINVAL, "slab [x_offset, x_offset+nx) outside nx_global");
    if (p->x_ghost != 0 && p->x_ghost != 1) return fail(B200ISO_EINVAL, "x_ghost must be 0 or 1");
    if (p->x_ghost && (p->algo != B200ISO_MT || p->x_offset == 0)) return fail(B200ISO_EINVAL, "x_ghost is for MT slabs that do not start at x = 0");
    if (p->algo == B200ISO_MT && p->x_offset > 0 && !p->x_ghost) return fail(B200ISO_EINVAL, "an MT slab with x_offset > 0 needs its ghost row (x_ghost = 1)");
  }
  return 0;
}

// classify -> count -> scan (-> grid coordinates when the shape or the ranges changed) on the handle's stream
int enqueue_count(b200iso_handle* h, const b200iso_params* p, const void* sdf_dev, int64_t nx, int64_t ny, int64_t nz,
                  int64_t ldx, long long* totals_out, bool step_begun = false) {
  cudaStream_t st = h->stream;
  h->prm = *p;
  h->sdf_dev = sdf_dev;
  h->vert_is_f64 = vertex_is_f64(*p);
  h->counted = false, h->totals_known = false;
  iso::Grid& g = h->grid;
  iso::grid_setup(g, nx, ny, nz, ldx);
  g.xoff = (int)p->x_offset;
  g.ghost = p->algo == B200ISO_MT && p->x_ghost != 0 ? 1 : 0;
  h->nblocks = (nx > 1 && ny > 1 && nz > 1) ? (long long)(nx - 1) * g.blocks_per_row : 0;

  if (h->nblocks >= (1ll << 31)) return fail(B200ISO_EINVAL, "grid too large: %lld generate blocks (limit 2^31)", h->nblocks);
  if (h->nblocks == 0) {  // a dimension < 2: zero voxels, empty mesh (src/marching_cubes.jl:40)
    CU(cudaMemsetAsync(h->totals_dev, 0, 2 * sizeof(long long), st));
    if (totals_out) CU(cudaMemsetAsync(totals_out, 0, 2 * sizeof(long long), st));
    h->counted = true;
    return 0;
  }
  const bool mt = p->algo == B200ISO_MT;
  const size_t nbits = (size_t)nx * ny * g.W;
  const long long per = (long long)iso::SC_THREADS * iso::SC_PER, nsb = (h->nblocks + per - 1) / per;  // scan blocks
  if (int rc = h->bits.reserve(nbits)) return rc;
  if (int rc = h->chain.reserve((size_t)nsb * 2 + 4)) return rc;
  if (int rc = h->woff.reserve((size_t)h->nblocks * 2)) return rc;
  if (mt) {
    if (int rc = h->status.reserve((size_t)h->nblocks * 2)) return rc;
    if (int rc = h->celloff.reserve(nbits)) return rc;
  }
  g.rec_cap = iso::rec_cap_for(h->nblocks);
  if (int rc = h->recs.reserve((size_t)h->nblocks * g.rec_cap + 2)) return rc;
  if (int rc = h->nrecs.reserve((size_t)h->nblocks)) return rc;
  if (int rc = h->coords.reserve((size_t)(nx + ny + nz))) return rc;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(h->chain.p);
  unsigned long long* chain = h->chain.p + 2;
  unsigned long long* ghost_words = chain + 2 * nsb;  // MT slabs: inclusive prefix at the end of the ghost row
  const int nclear = (int)(nsb * 2 + 4);

  if (!step_begun) h->begin_step();
  if (int rc = h->rec(b200iso_handle::E_C0)) return rc;
  // (1) sign-pack; its block 0 also clears the scan state (ticket + chain) -- no memset nodes in the step
  bool ride = false;  // the classify kernel counted (most of) the generate blocks itself
  {
    const bool f64 = p->field_is_f64 != 0;
    const bool vec = !f64 && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(sdf_dev) & 15) == 0);
    const float* sdf_f = reinterpret_cast<const float*>(sdf_dev);
    const int nxseg = (int)((nx + iso::SP_XSEG - 1) / iso::SP_XSEG);
    const int nzc = (g.W + iso::SP_ZW - 1) / iso::SP_ZW;
    const long long ntasks = (long long)nxseg * ny * nzc;
    const unsigned nb = (unsigned)((ntasks + iso::SP_WARPS - 1) / iso::SP_WARPS);
    const float thr = threshold_for(p->iso, p->iso_is_f32 != 0);
    CUtensorMap tmap;
    // (measured: TMA wins on big fields -- 0.656 vs 0.694 ms at 1024^3 -- and loses a few % when the grid is under two waves)
    const bool tma_wanted = h->tma_mode == 1 || (h->tma_mode < 0 && ntasks >= 4096);
    if (vec && tma_wanted && iso::make_field_tmap(&tmap, sdf_f, nx, ny, nz, ldx)) {
      // (x-slabs with their halo plane have 128 k + 1 samples: the last 128-wide segment then holds one valid column.
      // Tail tasks for those columns were tried three ways -- lanes along z + ballot, lane per row with scalar loads,
      // a second tensor map with 4 x * 32 y boxes -- and none beat the plain out-of-bounds box: the TMA engine does
      // not fetch the out-of-bounds part, and a 129-plane slab is bound by its 1.7 waves of CTAs, not by that segment.)
      const long long tpr = (long long)nxseg * nzc;  // classify tasks per sample row
      iso::CountRide cr{};
      cr.g = g;
      // (only when the classify kernel runs for many waves of CTAs: on a 129-plane slab -- 4096 tasks, 3.5 waves -- the
      // rows complete too late for the riding warps to get anything done, and they only cost: 0.139 vs 0.122 ms)
      if (h->ride_warps > 0 && ntasks >= h->ride_min_tasks) {
        const long long nbi = g.blocks_per_row;
        const size_t words = (size_t)iso::RIDE_HDR + (size_t)nbi + (size_t)ny;
        if (int rc = h->ride.reserve(words)) return rc;
        if (h->ride_ny != ny || h->ride_tpr != tpr || h->ride_nbi != nbi ||
            (unsigned long long)(h->ride_step + 2) * (unsigned long long)tpr >= (1ull << 31)) {
          CU(cudaMemsetAsync(h->ride.p, 0, words * sizeof(unsigned int), st));
          h->ride_ny = ny, h->ride_tpr = tpr, h->ride_nbi = nbi, h->ride_step = 0;
        }
        cr.recs = h->recs.p, cr.nrecs = h->nrecs.p;
        cr.nbi = (int)nbi, cr.woff = h->woff.p, cr.head = h->ride.p, cr.next_x = h->ride.p + iso::RIDE_HDR;
        cr.rows_done = cr.next_x + nbi;
        cr.target = ++h->ride_step * (unsigned int)tpr;
        cr.stop_at = h->ride_stop;
        cr.celloff = mt ? h->celloff.p : nullptr;
        ride = true;
      }
      const int rw = ntasks < 8192 ? std::min(h->ride_warps, 4) : h->ride_warps;  // (short kernels: fewer counting warps, see ride_min_tasks)
      const unsigned tb = (unsigned)((ntasks + iso::TM_WARPS - 1) / iso::TM_WARPS);
      if (mt && ride)
        iso::signpack_tma_kernel<true><<<tb, (iso::TM_WARPS + std::min(rw, iso::TM_CNT_WARPS_MT)) * 32, iso::TM_SMEM, st>>>(
            tmap, h->bits.p, g.nx, g.ny, g.nz, g.W, thr, nxseg, nzc, ntasks, h->chain.p, nclear, cr);
      else
        iso::signpack_tma_kernel<false><<<tb, (iso::TM_WARPS + (ride ? rw : 0)) * 32, iso::TM_SMEM, st>>>(
            tmap, h->bits.p, g.nx, g.ny, g.nz, g.W, thr, nxseg, nzc, ntasks, h->chain.p, nclear, cr);
      h->classify_path = B200ISO_CLASSIFY_TMA;
    } else if (vec) {
      iso::signpack_kernel<true, float><<<nb, iso::SP_WARPS * 32, 0, st>>>(sdf_f, h->bits.p, g.nx, g.ny, g.nz, g.ldx, g.W, thr, nxseg, ntasks, h->chain.p, nclear);
      h->classify_path = B200ISO_CLASSIFY_LDG128;
    } else if (!f64) {
      iso::signpack_kernel<false, float><<<nb, iso::SP_WARPS * 32, 0, st>>>(sdf_f, h->bits.p, g.nx, g.ny, g.nz, g.ldx, g.W, thr, nxseg, ntasks, h->chain.p, nclear);
      h->classify_path = B200ISO_CLASSIFY_SCALAR;
    } else {  // Float64 field: Float64 < promote(iso) compares exactly in Float64
      iso::signpack_kernel<false, double><<<nb, iso::SP_WARPS * 32, 0, st>>>(reinterpret_cast<const double*>(sdf_dev), h->bits.p, g.nx, g.ny,
                                                                           g.nz, g.ldx, g.W, p->iso_is_f32 ? (double)(float)p->iso : p->iso,
                                                                           nxseg, ntasks, h->chain.p, nclear);
      h->classify_path = B200ISO_CLASSIFY_F64;
    }
    CU(cudaGetLastError());
    h->launches++;
  }
  if (int rc = h->rec(b200iso_handle::E_C1)) return rc;
  // (2) count (raw totals per generate block), then a light single-pass decoupled look-back scan over them
  h->totals_out = totals_out;
  if (!mt) {
    // after a classify with counting warps only the last y-blocks are left: a quarter of the full grid, grid-stride
    // (correct for any remainder), instead of thousands of blocks that look at the queue and leave
    unsigned ncb = (unsigned)((h->nblocks + iso::WC_THREADS / 32 - 1) / (iso::WC_THREADS / 32));
    if (ride) ncb = std::max(1u, std::min(ncb, std::max(148u * 4u, ncb / 4)));
    static_assert(iso::RIDE_HDR_WORDS == iso::RIDE_HDR, "ride buffer layout");
    iso::mc_count_chunks_kernel<false><<<ncb, iso::WC_THREADS, 0, st>>>(h->bits.p, g, h->nblocks, h->woff.p, h->recs.p, h->nrecs.p,
                                                                        ride ? h->ride.p : nullptr, nullptr);
    CU(cudaGetLastError());
    iso::mc_scan_chunks_kernel<<<(unsigned)nsb, iso::SC_THREADS, 0, st>>>(h->woff.p, h->nblocks, chain, ticket, nsb, h->totals_dev, totals_out,
                                                                          ride ? h->ride.p : nullptr, ride ? iso::RIDE_HDR + g.blocks_per_row : 0);
  } else {
    if (ride) {  // what the counting warps left (the y-blocks of the last rows), a warp per generate block
      unsigned ncb = (unsigned)((h->nblocks + iso::WC_THREADS / 32 - 1) / (iso::WC_THREADS / 32));
      ncb = std::max(1u, std::min(ncb, std::max(148u * 4u, ncb / 4)));
      iso::mc_count_chunks_kernel<true><<<ncb, iso::WC_THREADS, 0, st>>>(h->bits.p, g, h->nblocks, h->woff.p, h->recs.p, h->nrecs.p, h->ride.p,
                                                                         h->celloff.p);
    } else {
      iso::mt_count_kernel<<<(unsigned)h->nblocks, iso::CB_THREADS, 0, st>>>(h->bits.p, g, h->celloff.p, h->woff.p, h->recs.p, h->nrecs.p);
    }
    CU(cudaGetLastError());
    iso::mt_scan_blocks_kernel<<<(unsigned)nsb, iso::SC_THREADS, 0, st>>>(h->woff.p, h->nblocks, h->status.p, chain, ticket, nsb,
                                                                          g.ghost ? (long long)g.blocks_per_row - 1 : -1LL, ghost_words, h->totals_dev, totals_out,
                                                                          ride ? h->ride.p : nullptr, ride ? iso::RIDE_HDR + g.blocks_per_row : 0);
  }
  CU(cudaGetLastError());
  h->launches += 2;
  // grid coordinates (LinRange), consumed by generate: a function of the shape and the ranges only
  {
    b200iso_handle::CoordsKey key;
    key.nx = nx, key.ny = ny, key.nz = nz, key.xoff = p->x_offset, key.nxg = p->nx_global > 0 ? p->nx_global : nx;
    key.r[0] = p->x0, key.r[1] = p->x1, key.r[2] = p->y0, key.r[3] = p->y1, key.r[4] = p->z0, key.r[5] = p->z1;
    key.f32 = p->range_kind == B200ISO_RANGE_F32;
    if (!(key == h->coords_key)) {
      const int n = (int)(nx + ny + nz);
      iso::coords_kernel<<<(n + 255) / 256, 256, 0, st>>>(h->coords.p, g.nx, g.ny, g.nz, p->x0, p->x1, p->y0, p->y1, p->z0, p->z1, key.f32,
                                                          (int)key.xoff, (int)key.nxg);
      CU(cudaGetLastError());
      h->launches++;
      h->coords_key = key;
    }
  }
  if (int rc = h->rec(b200iso_handle::E_C2)) return rc;
  h->counted = true;
  return 0;
}

int enqueue_generate(b200iso_handle* h, void* verts_dev, int64_t vcap, int64_t* faces_dev, int64_t fcap,
                     const int64_t* vertex_base_dev, int64_t vertex_base) {
  if (!h->counted) return fail(B200ISO_ESTATE, "generate called before count");
  cudaStream_t st = h->stream;
  if (int rc = h->rec(b200iso_handle::E_G0)) return rc;
  if (h->nblocks > 0) {
    const b200iso_params& p = h->prm;
    iso::GenArgs a{};
    a.sdf = h->sdf_dev, a.bits = h->bits.p, a.status = h->status.p, a.woff = h->woff.p, a.coords = h->coords.p;
    a.verts = verts_dev, a.faces = (long long*)faces_dev, a.vcap = vcap, a.fcap = fcap;
    a.vbase_dev = (const long long*)vertex_base_dev, a.vbase = vertex_base;
    a.iso_d = p.iso, a.iso_f = (float)p.iso, a.eps_d = p.eps, a.eps_f = (float)p.eps;
    a.iso_is_f32 = p.iso_is_f32, a.eps_is_f32 = p.eps_is_f32, a.p_is_f32 = p.range_kind == B200ISO_RANGE_F32;
    a.sdf_vec = !p.field_is_f64 && h->grid.ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(h->sdf_dev) & 15) == 0;
    a.recs = h->recs.p, a.nrecs = h->nrecs.p;
    a.nblocks = h->nblocks, a.totals_a = h->totals_dev, a.abort_flag = h->totals_dev + 2;
    const unsigned nb = (unsigned)h->nblocks;
    const bool pf32 = p.range_kind == B200ISO_RANGE_F32;
    if (p.algo == B200ISO_MC) {
      if (p.field_is_f64) iso::mc_generate_kernel<3, double><<<nb, iso::CB_THREADS, 0, st>>>(a, h->grid);
      else if (!p.iso_is_f32) iso::mc_generate_kernel<2, double><<<nb, iso::CB_THREADS, 0, st>>>(a, h->grid);
      else if (pf32) iso::mc_generate_kernel<1, float><<<nb, iso::CB_THREADS, 0, st>>>(a, h->grid);
      else if (h->vert_is_f64) iso::mc_generate_kernel<0, double><<<nb, iso::CB_THREADS, 0, st>>>(a, h->grid);
      else iso::mc_generate_kernel<0, float><<<nb, iso::CB_THREADS, 0, st>>>(a, h->grid);
    } else {
      if (int rc = iso::launch_mt_generate(a, h->grid, p, h->vert_is_f64, h->celloff.p, nb, st)) return fail(B200ISO_EINVAL, "MT launch failed (%d)", rc);
    }
    CU(cudaGetLastError());
    h->launches++;
  }
  if (int rc = h->rec(b200iso_handle::E_G1)) return rc;
  return 0;
}

int fetch_totals(b200iso_handle* h) {
  if (!h->counted) return fail(B200ISO_ESTATE, "no counted field");
  if (!h->totals_known) {
    CU(cudaMemcpyAsync(h->totals_host, h->totals_dev, 3 * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (h->totals_host[2] != 0) {
      cudaMemsetAsync(h->totals_dev + 2, 0, sizeof(long long), h->stream);
      return fail(B200ISO_ESTATE, "peer exchange timed out: a rank did not publish its totals (re-arm with b200iso_set_peer_exchange on zeroed buffers)");
    }
    h->nverts = h->totals_host[0], h->nfaces = h->totals_host[1];
    h->totals_known = true;
  }
  return 0;
}

}  // namespace

extern "C" {

int b200iso_version(void) { return 2000; }
const char* b200iso_last_error(void) { return g_err; }

static int create_impl(b200iso_handle* h) {
  if (const char* e = getenv("B200ISO_TMA")) h->tma_mode = atoi(e) != 0 ? 1 : 0;
  if (const char* e = getenv("B200ISO_RIDE_STOP")) h->ride_stop = atoi(e);
  if (const char* e = getenv("B200ISO_RIDE_MIN_TASKS")) h->ride_min_tasks = std::max(1ll, atoll(e));
  if (const char* e = getenv("B200ISO_RIDE")) h->ride_warps = std::max(0, std::min(iso::TM_CNT_WARPS_MAX, atoi(e)));
  // per device: the TMA classify kernel needs more than the default 48 KB of dynamic shared memory
  CU(cudaFuncSetAttribute(iso::signpack_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iso::TM_SMEM));
  CU(cudaFuncSetAttribute(iso::signpack_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iso::TM_SMEM));
  CU(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  CU(cudaMalloc((void**)&h->totals_dev, 4 * sizeof(long long)));
  CU(cudaMemset(h->totals_dev, 0, 4 * sizeof(long long)));
  CU(cudaMallocHost((void**)&h->totals_host, 4 * sizeof(long long)));
  return 0;
}

int b200iso_create(b200iso_handle** out, int device) {
  if (!out) return fail(B200ISO_EINVAL, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(B200ISO_EINVAL, "device %d out of range (%d devices)", device, ndev);
  DeviceGuard guard(device);
  b200iso_handle* h = new (std::nothrow) b200iso_handle();
  if (!h) return fail(B200ISO_ENOMEM, "out of host memory");
  h->device = device;
  h->pool.device = device;
  if (int rc = create_impl(h)) {  // (the message of the failing call stays in g_err)
    b200iso_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

int b200iso_destroy(b200iso_handle* h) {
  if (!h) return 0;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();  // (the caller's stream may already be gone: do not touch h->stream)
  h->vstage1.release(), h->fstage1.release();
  for (cudaEvent_t e : h->slab_ev) cudaEventDestroy(e);
  h->pool.release();
  h->bits.release(), h->celloff.release(), h->woff.release(), h->chain.release(), h->status.release(), h->coords.release(), h->field.release(), h->vstage.release(), h->fstage.release();
  h->cases.release(), h->ride.release(), h->recs.release(), h->nrecs.release();
  h->weld_tab.release(), h->weld_u32.release(), h->weld_tmp.release();
  if (h->totals_dev) cudaFree(h->totals_dev);
  if (h->totals_host) cudaFreeHost(h->totals_host);
  if (h->ev) {
    for (int i = 0; i < b200iso_handle::NSLOT * b200iso_handle::E_N; ++i) cudaEventDestroy(h->ev[i]);
    delete[] h->ev;
    delete[] h->ev_set;
  }
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  cudaGetLastError();
  delete h;
  return 0;
}

int b200iso_set_stream(b200iso_handle* h, void* cuda_stream) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  h->stream = (cudaStream_t)cuda_stream;
  h->coords_key = b200iso_handle::CoordsKey();  // the cached grid coordinates were produced on the previous stream
  return 0;
}

int b200iso_use_own_stream(b200iso_handle* h) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  h->stream = h->own_stream;
  h->coords_key = b200iso_handle::CoordsKey();
  return 0;
}

int b200iso_set_classify_mode(b200iso_handle* h, int mode) {
  if (!h || mode < -1 || mode > 1) return fail(B200ISO_EINVAL, "bad handle or classify mode");
  h->tma_mode = mode;
  return 0;
}

int b200iso_classify_path(b200iso_handle* h) { return h ? h->classify_path : -1; }

int b200iso_set_ride_warps(b200iso_handle* h, int warps) {
  if (!h || warps < 0 || warps > iso::TM_CNT_WARPS_MAX) return fail(B200ISO_EINVAL, "bad handle or warp count (0..%d)", iso::TM_CNT_WARPS_MAX);
  h->ride_warps = warps;
  return 0;
}

int64_t b200iso_ride_claimed(b200iso_handle* h) {
  if (!h || !h->ride.p) return 0;
  DeviceGuard guard(h->device);
  unsigned int v = 0;
  if (cudaStreamSynchronize(h->stream) != cudaSuccess || cudaMemcpy(&v, h->ride.p + 2, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return (int64_t)v;
}

int b200iso_set_peer_timeout(b200iso_handle* h, double seconds) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  h->peer_timeout_s = seconds;
  return 0;
}

int b200iso_count_async(b200iso_handle* h, const b200iso_params* p, const void* sdf_dev, int64_t nx, int64_t ny,
                        int64_t nz, int64_t ldx, int64_t* totals_dev) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (int rc = check_params(p, nx, ny, nz, ldx)) return rc;
  if (!sdf_dev && nx * ny * nz > 0) return fail(B200ISO_EINVAL, "sdf is NULL");
  DeviceGuard guard(h->device);
  return enqueue_count(h, p, sdf_dev, nx, ny, nz, ldx, (long long*)totals_dev);
}

int b200iso_generate_async(b200iso_handle* h, void* verts_dev, int64_t vcap, int64_t* faces_dev, int64_t fcap,
                           const int64_t* vertex_base_dev, int64_t vertex_base) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (vcap < 0 || fcap < 0) return fail(B200ISO_EINVAL, "negative capacity");
  if ((vcap > 0 && !verts_dev) || (fcap > 0 && !faces_dev)) return fail(B200ISO_EINVAL, "output pointer is NULL");
  DeviceGuard guard(h->device);
  return enqueue_generate(h, verts_dev, vcap, faces_dev, fcap, vertex_base_dev, vertex_base);
}

int b200iso_extract_async(b200iso_handle* h, const b200iso_params* p, const void* sdf_dev, int64_t nx, int64_t ny, int64_t nz,
                          int64_t ldx, void* verts_dev, int64_t vcap, int64_t* faces_dev, int64_t fcap,
                          const int64_t* vertex_base_dev, int64_t vertex_base, int64_t* totals_dev) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (int rc = check_params(p, nx, ny, nz, ldx)) return rc;
  if (!sdf_dev && nx * ny * nz > 0) return fail(B200ISO_EINVAL, "sdf is NULL");
  if (vcap < 0 || fcap < 0) return fail(B200ISO_EINVAL, "negative capacity");
  if ((vcap > 0 && !verts_dev) || (fcap > 0 && !faces_dev)) return fail(B200ISO_EINVAL, "output pointer is NULL");
  DeviceGuard guard(h->device);
  if (int rc = enqueue_count(h, p, sdf_dev, nx, ny, nz, ldx, (long long*)totals_dev)) return rc;
  return enqueue_generate(h, verts_dev, vcap, faces_dev, fcap, vertex_base_dev, vertex_base);
}

int b200iso_add_vertex_base_async(b200iso_handle* h, int64_t* faces_dev, int64_t fcap, const int64_t* totals_dev,
                                  const int64_t* vertex_base_dev) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (!faces_dev || !totals_dev || !vertex_base_dev) return fail(B200ISO_EINVAL, "NULL argument");
  if ((reinterpret_cast<uintptr_t>(faces_dev) & 15) != 0) return fail(B200ISO_EINVAL, "faces must be 16-byte aligned");
  DeviceGuard guard(h->device);
  iso::add_base_kernel<<<148 * 8, 256, 0, h->stream>>>((long long*)faces_dev, fcap, (const long long*)totals_dev,
                                                       (const long long*)vertex_base_dev);
  CU(cudaGetLastError());
  h->launches++;
  return 0;
}

int b200iso_set_peer_exchange(b200iso_handle* h, int rank, int world, void* const* slots) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (world <= 1 || !slots) {
    h->peer_world = 0;
    return 0;
  }
  if (world > iso::PEER_MAX || rank < 0 || rank >= world) return fail(B200ISO_EINVAL, "bad rank/world %d/%d (at most %d ranks)", rank, world, iso::PEER_MAX);
  for (int r = 0; r < world; ++r) {
    if (!slots[r]) return fail(B200ISO_EINVAL, "slots[%d] is NULL", r);
    h->peers.slot[r] = (long long*)slots[r];
  }
  h->peer_rank = rank, h->peer_world = world, h->peer_epoch = 0;
  DeviceGuard guard(h->device);
  CU(cudaMemsetAsync(h->totals_dev + 2, 0, sizeof(long long), h->stream));  // re-armed: a previous time-out is forgotten
  return 0;
}

int b200iso_exchange_async(b200iso_handle* h, int64_t* bases_dev, int64_t* all_dev) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (!h->counted) return fail(B200ISO_ESTATE, "exchange called before count");
  if (h->peer_world <= 1) return fail(B200ISO_ESTATE, "no peer exchange configured (b200iso_set_peer_exchange)");
  if (!bases_dev) return fail(B200ISO_EINVAL, "bases_dev is NULL");
  DeviceGuard guard(h->device);
  const long long epoch = ++h->peer_epoch;
  const unsigned long long timeout_ns = h->peer_timeout_s > 0 ? (unsigned long long)(h->peer_timeout_s * 1e9) : 0ull;
  iso::peer_exchange_kernel<<<1, 32, 0, h->stream>>>(h->peers, h->peer_world, h->peer_rank, epoch, h->totals_dev, (long long*)bases_dev,
                                                     (long long*)all_dev, h->totals_dev + 2, timeout_ns);
  CU(cudaGetLastError());
  h->launches += 1;
  h->totals_known = false;  // the next b200iso_totals re-reads the totals together with the exchange's error flag
  return 0;
}

int b200iso_totals(b200iso_handle* h, int64_t* nverts, int64_t* nfaces, int* vert_is_f64) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  DeviceGuard guard(h->device);
  if (int rc = fetch_totals(h)) return rc;
  if (nverts) *nverts = h->nverts;
  if (nfaces) *nfaces = h->nfaces;
  if (vert_is_f64) *vert_is_f64 = h->vert_is_f64;
  return 0;
}

static int b200iso_count_impl(b200iso_handle* h, const b200iso_params* p, const void* sdf, int mem, int64_t nx, int64_t ny, int64_t nz,
                  int64_t ldx, int64_t* nverts, int64_t* nfaces, int* vert_is_f64) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (int rc = check_params(p, nx, ny, nz, ldx)) return rc;
  if (mem != B200ISO_HOST && mem != B200ISO_DEVICE) return fail(B200ISO_EINVAL, "bad mem kind %d", mem);
  if (!sdf && nx * ny * nz > 0) return fail(B200ISO_EINVAL, "sdf is NULL");
  DeviceGuard guard(h->device);
  const void* dev = sdf;
  int64_t dldx = ldx;
  const size_t esz = p->field_is_f64 ? 8 : 4;
  const bool staged = mem == B200ISO_HOST && nx * ny * nz > 0;
  if (staged) {
    h->hs.clear();  // (the staging buffer is about to be overwritten: nothing resident for b200iso_extract_host_resident)
    // stage with a 16-byte aligned leading dimension so the 128-bit load path always applies
    dldx = (nx + 3) / 4 * 4;
    if (int rc = h->field.reserve((size_t)dldx * ny * nz * esz)) return rc;
    h->begin_step();
    if (int rc = h->rec(b200iso_handle::E_H0)) return rc;
    if (!hostpipe::is_pinned(sdf)) {
      // pageable caller array (the usual Julia Array): worker threads stage it through pinned chunks
      if (h->slab_ev.empty()) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->slab_ev.push_back(e);
      }
      CU(cudaEventRecord(h->slab_ev[0], h->stream));
      hostpipe::Slab whole;
      whole.dst = h->field.p, whole.dpitch = (size_t)dldx * esz, whole.x0 = 0, whole.x1 = nx;
      hostpipe::Uploader up;
      CU(up.start(&h->pool, {whole}, (const unsigned char*)sdf, (size_t)ldx * esz, esz, (size_t)ny * nz, h->slab_ev[0]));
      CU(up.wait_slab(0, h->stream));
    } else if (dldx == ldx)
      CU(cudaMemcpyAsync(h->field.p, sdf, (size_t)ldx * ny * nz * esz, cudaMemcpyHostToDevice, h->stream));
    else
      CU(cudaMemcpy2DAsync(h->field.p, (size_t)dldx * esz, sdf, (size_t)ldx * esz, (size_t)nx * esz,
                           (size_t)ny * nz, cudaMemcpyHostToDevice, h->stream));
    if (int rc = h->rec(b200iso_handle::E_H1)) return rc;
    dev = h->field.p;
  }
  if (int rc = enqueue_count(h, p, dev, nx, ny, nz, dldx, nullptr, staged)) return rc;
  if (int rc = fetch_totals(h)) return rc;
  if (nverts) *nverts = h->nverts;
  if (nfaces) *nfaces = h->nfaces;
  if (vert_is_f64) *vert_is_f64 = h->vert_is_f64;
  return 0;
}

int b200iso_count(b200iso_handle* h, const b200iso_params* p, const void* sdf, int mem, int64_t nx, int64_t ny, int64_t nz,
                  int64_t ldx, int64_t* nverts, int64_t* nfaces, int* vert_is_f64) {
  try {
    return b200iso_count_impl(h, p, sdf, mem, nx, ny, nz, ldx, nverts, nfaces, vert_is_f64);
  } catch (const std::exception& e) {  // (std::bad_alloc, std::system_error from a thread): nothing may cross the C ABI
    return fail(B200ISO_ENOMEM, "b200iso_count: %s", e.what());
  } catch (...) {
    return fail(B200ISO_ENOMEM, "b200iso_count: unknown C++ exception");
  }
}

static int b200iso_generate_impl(b200iso_handle* h, void* verts, int64_t* faces, int mem, int64_t vertex_base) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (mem != B200ISO_HOST && mem != B200ISO_DEVICE) return fail(B200ISO_EINVAL, "bad mem kind %d", mem);
  DeviceGuard guard(h->device);
  if (int rc = fetch_totals(h)) return rc;
  const size_t vsz = h->vert_is_f64 ? 8 : 4;
  if ((h->nverts > 0 && !verts) || (h->nfaces > 0 && !faces)) return fail(B200ISO_EINVAL, "output pointer is NULL");
  if (mem == B200ISO_DEVICE) {
    if (int rc = enqueue_generate(h, verts, h->nverts, faces, h->nfaces, nullptr, vertex_base)) return rc;
    CU(cudaStreamSynchronize(h->stream));
  } else {
    if (int rc = h->vstage.reserve((size_t)h->nverts * 3 * vsz + 16)) return rc;
    if (int rc = h->fstage.reserve((size_t)h->nfaces * 3 + 2)) return rc;
    if (int rc = enqueue_generate(h, h->vstage.p, h->nverts, (int64_t*)h->fstage.p, h->nfaces, nullptr, vertex_base)) return rc;
    if (int rc = h->rec(b200iso_handle::E_D0)) return rc;
    if (!(hostpipe::is_pinned(verts) && hostpipe::is_pinned(faces)) && (h->nverts || h->nfaces)) {
      // pageable caller arrays: worker threads drain the staging through pinned chunks
      if (h->slab_ev.empty()) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->slab_ev.push_back(e);
      }
      CU(cudaEventRecord(h->slab_ev[0], h->stream));
      hostpipe::advise_huge(verts, (size_t)h->nverts * 3 * vsz);
      hostpipe::advise_huge(faces, (size_t)h->nfaces * 3 * sizeof(int64_t));
      hostpipe::Downloader down;
      CU(down.start(&h->pool, 1, false, h->slab_ev[0]));
      hostpipe::Downloader::Job j;
      j.src[0] = h->vstage.p, j.dst[0] = (unsigned char*)verts, j.bytes[0] = (size_t)h->nverts * 3 * vsz;
      j.src[1] = (const unsigned char*)h->fstage.p, j.dst[1] = (unsigned char*)faces, j.bytes[1] = (size_t)h->nfaces * 3 * sizeof(int64_t);
      j.ready = h->slab_ev[0];
      CU(down.push(0, j));
      CU(down.finish());
    } else {
      if (h->nverts) CU(cudaMemcpyAsync(verts, h->vstage.p, (size_t)h->nverts * 3 * vsz, cudaMemcpyDeviceToHost, h->stream));
      if (h->nfaces) CU(cudaMemcpyAsync(faces, h->fstage.p, (size_t)h->nfaces * 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    }
    if (int rc = h->rec(b200iso_handle::E_D1)) return rc;
    CU(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

int b200iso_generate(b200iso_handle* h, void* verts, int64_t* faces, int mem, int64_t vertex_base) {
  try {
    return b200iso_generate_impl(h, verts, faces, mem, vertex_base);
  } catch (const std::exception& e) {  // (std::bad_alloc, std::system_error from a thread): nothing may cross the C ABI
    return fail(B200ISO_ENOMEM, "b200iso_generate: %s", e.what());
  } catch (...) {
    return fail(B200ISO_ENOMEM, "b200iso_generate: unknown C++ exception");
  }
}

// One-shot host form: x-slab software pipeline over three streams.  x is the scan-outermost
// axis, so the mesh of voxel rows [a, b) is a contiguous piece of the output and needs sample planes [a, b] only:
//   in_stream : strided (2-D) H2D copies of the slabs, all enqueued up front (>= 256 B rows run at full PCIe rate)
//   h->stream : per slab classify -> count/scan -> (16-byte totals read-back) -> generate into a staging set
//   out_stream: D2H of slab k's vertices/faces straight into their final offsets of the caller's arrays,
//               overlapping the H2D of the slabs behind it (PCIe is full duplex)
// Every slab is the sharded sub-problem of api.isosurface_slab (x_offset / nx_global / vertex base; Marching
// Tetrahedra slabs carry their ghost row), so the concatenation is byte-identical to the unsharded mesh.
// Per-slab pass of the one-shot host forms: count (after the slab has arrived, when `up` is given), totals, generate into
// a staging set, D2H into the final offsets of the caller's arrays.  The slab decomposition (h->hs) stays in the handle
// with the slabs resident in h->field, so that a call whose capacities were too small is finished by
// b200iso_extract_host_resident without uploading the field again.
static int host_slab_pass(b200iso_handle* h, hostpipe::Uploader* up, void* verts, int64_t vcap, int64_t* faces, int64_t fcap, int64_t* nverts,
                          int64_t* nfaces) {
  cudaStream_t st = h->stream;
  const int S = (int)h->hs.size();
  const size_t vsz = h->hs_vf64 ? 8 : 4;
  const cudaEvent_t ev_fork = h->slab_ev[S];
  hostpipe::Downloader down;
  const bool out_pinned = hostpipe::is_pinned(verts) && hostpipe::is_pinned(faces);
  if (!out_pinned) {  // fresh pageable result arrays: first touched by the download workers
    hostpipe::advise_huge(verts, (size_t)std::max<int64_t>(0, vcap) * 3 * vsz);
    hostpipe::advise_huge(faces, (size_t)std::max<int64_t>(0, fcap) * 3 * sizeof(int64_t));
  }
  CU(down.start(&h->pool, S, out_pinned, ev_fork));
  int64_t cv = 0, cf = 0;
  int njobs = 0;
  bool overflow = false;
  for (int k = 0; k < S; ++k) {
    const b200iso_handle::HostSlab& hk = h->hs[k];
    if (hk.nx < 2) continue;
    if (up) {
      CU(up->wait_slab(k, st));
      if (k == S - 1)
        if (int rc = h->rec(b200iso_handle::E_H1)) return rc;
    }
    if (int rc = enqueue_count(h, &hk.prm, hk.dev, hk.nx, h->hs_ny, h->hs_nz, hk.ldx, nullptr, true)) return rc;
    if (int rc = fetch_totals(h)) return rc;
    const int64_t nv = h->nverts, nf = h->nfaces;
    if (cv + nv > vcap || cf + nf > fcap) overflow = true;
    if (!overflow && (nv > 0 || nf > 0)) {
      DevBuf<unsigned char>& vs = (njobs & 1) ? h->vstage1 : h->vstage;
      DevBuf<long long>& fs = (njobs & 1) ? h->fstage1 : h->fstage;
      if (njobs >= 2) CU(down.wait_job(njobs - 2));  // this staging set's previous D2H
      if (int rc = vs.reserve((size_t)nv * 3 * vsz + 16)) return rc;
      if (int rc = fs.reserve((size_t)nf * 3 + 2)) return rc;
      if (int rc = enqueue_generate(h, vs.p, nv, (int64_t*)fs.p, nf, nullptr, cv)) return rc;
      CU(cudaEventRecord(h->slab_ev[k], st));
      hostpipe::Downloader::Job j;
      j.src[0] = vs.p, j.dst[0] = (unsigned char*)verts + (size_t)cv * 3 * vsz, j.bytes[0] = (size_t)nv * 3 * vsz;
      j.src[1] = (const unsigned char*)fs.p, j.dst[1] = (unsigned char*)(faces + (size_t)cf * 3), j.bytes[1] = (size_t)nf * 3 * sizeof(int64_t);
      j.ready = h->slab_ev[k];
      CU(down.push(njobs++, j));
    }
    cv += nv, cf += nf;
  }
  if (up) up->join();
  CU(down.finish());  // the mesh is in the caller's arrays
  CU(cudaStreamSynchronize(st));
  h->counted = false;  // the handle holds the last slab only: not a state b200iso_generate may continue from
  if (nverts) *nverts = cv;
  if (nfaces) *nfaces = cf;
  h->hs_nv = cv, h->hs_nf = cf;
  if (overflow) return fail(B200ISO_ECAPACITY, "mesh has %lld vertices / %lld faces, capacity is %lld / %lld", (long long)cv, (long long)cf, (long long)vcap, (long long)fcap);
  return 0;
}

static int b200iso_extract_host_impl(b200iso_handle* h, const b200iso_params* p, const void* sdf, int64_t nx, int64_t ny, int64_t nz,
                         int64_t ldx, void* verts, int64_t vcap, int64_t* faces, int64_t fcap, int64_t* nverts,
                         int64_t* nfaces, int* vert_is_f64) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (int rc = check_params(p, nx, ny, nz, ldx)) return rc;
  if (!sdf && nx * ny * nz > 0) return fail(B200ISO_EINVAL, "sdf is NULL");
  if (vcap < 0 || fcap < 0) return fail(B200ISO_EINVAL, "negative capacity");
  if ((vcap > 0 && !verts) || (fcap > 0 && !faces)) return fail(B200ISO_EINVAL, "output pointer is NULL");
  DeviceGuard guard(h->device);
  const int f64 = vertex_is_f64(*p);
  if (vert_is_f64) *vert_is_f64 = f64;
  if (nverts) *nverts = 0;
  if (nfaces) *nfaces = 0;
  h->hs.clear();
  h->hs_nv = h->hs_nf = 0;
  if (nx < 2 || ny < 2 || nz < 2) return 0;  // zero voxels, empty mesh
  const size_t esz = p->field_is_f64 ? 8 : 4;
  const bool mt = p->algo == B200ISO_MT;
  // slabs of >= 256 samples, at most 8, and only on big fields.  Measured at 1024^3 (pinned arrays, PCIe 5 x16):
  // 4 slabs 88.4 ms, 8 slabs 90.9, 16 slabs 98.8, 32 slabs 158 -- against 96.3 ms for count + generate; alone the
  // strided copies hold 55.6 GB/s down to 256-byte rows, but next to the D2H traffic the narrow ones fall behind.
  const int64_t nvx = nx - 1;
  int S = (size_t)nx * ny * nz * esz < ((size_t)32 << 20) ? 1 : (int)std::max<int64_t>(1, std::min<int64_t>(8, nx / 256));
  if (const char* e = getenv("B200ISO_HOST_SLABS")) S = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(64, nvx), atoi(e)));
  // voxel-row boundaries; the first sample of every slab (its ghost row for MT) is 16-byte aligned in the staging
  std::vector<int64_t> bound(S + 1);
  for (int k = 0; k <= S; ++k) {
    int64_t b = nvx * k / S;
    if (k > 0 && k < S) b = std::min(nvx, b / 4 * 4 + (mt ? 1 : 0));
    bound[k] = b;
  }
  for (int k = 1; k <= S; ++k) bound[k] = std::max(bound[k], bound[k - 1]);
  while (h->slab_ev.size() < (size_t)S + 2) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->slab_ev.push_back(e);
  }
  // device staging: every slab compact (row pitch = its padded width), holding sample planes [a - lo, b]
  std::vector<hostpipe::Slab> slabs(S);
  size_t field_bytes = 0;
  for (int k = 0; k < S; ++k) {
    const int64_t a = bound[k], b = bound[k + 1];
    const int lo = a > 0 && mt ? 1 : 0;  // an MT slab starts with its ghost row
    hostpipe::Slab& sb = slabs[k];
    sb.x0 = a - lo, sb.x1 = b > a ? b + 1 : sb.x0;
    sb.dpitch = (size_t)((sb.x1 - sb.x0 + 3) / 4 * 4) * esz;
    field_bytes = (field_bytes + 255) / 256 * 256;
    sb.dst = (unsigned char*)field_bytes;  // offset for now
    field_bytes += sb.dpitch * (size_t)ny * nz;
  }
  if (int rc = h->field.reserve(field_bytes)) return rc;
  for (hostpipe::Slab& sb : slabs) sb.dst = h->field.p + (size_t)sb.dst;
  // the sub-problems (a caller's own slab of a sharded volume keeps its x_offset / nx_global / ghost row: sub-slab 0 inherits them)
  h->hs.resize(S);
  h->hs_ny = ny, h->hs_nz = nz, h->hs_vf64 = f64;
  for (int k = 0; k < S; ++k) {
    const int64_t a = bound[k], b = bound[k + 1];
    b200iso_handle::HostSlab& hk = h->hs[k];
    const int ghost = a > 0 ? (mt ? 1 : 0) : p->x_ghost;
    hk.prm = *p;
    hk.prm.x_offset = p->x_offset + a - (a > 0 ? ghost : 0), hk.prm.nx_global = p->nx_global > 0 ? p->nx_global : nx, hk.prm.x_ghost = ghost;
    const int lo = a > 0 ? ghost : 0;  // sample planes below voxel row a that the sub-slab starts with
    hk.nx = b > a ? b - a + 1 + lo : 0;
    hk.dev = slabs[k].dst, hk.ldx = (int64_t)(slabs[k].dpitch / esz);
  }
  cudaStream_t st = h->stream;
  const cudaEvent_t ev_fork = h->slab_ev[S];
  h->begin_step();
  if (int rc = h->rec(b200iso_handle::E_H0)) return rc;
  CU(cudaEventRecord(ev_fork, st));  // the copy lanes start after whatever the caller's stream holds
  // H2D, enqueued by the uploader's worker threads (the enqueue of a million-row 2-D copy keeps its thread busy
  // for about the copy's duration, and this thread has the kernels and the D2H of the earlier slabs to enqueue).
  hostpipe::Uploader up;
  CU(up.start(&h->pool, slabs, (const unsigned char*)sdf, (size_t)ldx * esz, esz, (size_t)ny * nz, ev_fork));
  const int rc = host_slab_pass(h, &up, verts, vcap, faces, fcap, nverts, nfaces);
  up.join();
  if (rc != 0 && rc != B200ISO_ECAPACITY) h->hs.clear();
  return rc;
}

static int b200iso_extract_host_resident_impl(b200iso_handle* h, void* verts, int64_t vcap, int64_t* faces, int64_t fcap, int64_t* nverts,
                                              int64_t* nfaces) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (h->hs.empty()) return fail(B200ISO_ESTATE, "no host field resident: call b200iso_extract_host first");
  if (vcap < 0 || fcap < 0) return fail(B200ISO_EINVAL, "negative capacity");
  if ((vcap > 0 && !verts) || (fcap > 0 && !faces)) return fail(B200ISO_EINVAL, "output pointer is NULL");
  DeviceGuard guard(h->device);
  const int S = (int)h->hs.size();
  h->begin_step();
  CU(cudaEventRecord(h->slab_ev[S], h->stream));
  return host_slab_pass(h, nullptr, verts, vcap, faces, fcap, nverts, nfaces);
}

int b200iso_extract_host_resident(b200iso_handle* h, void* verts, int64_t vcap, int64_t* faces, int64_t fcap, int64_t* nverts, int64_t* nfaces) {
  try {
    return b200iso_extract_host_resident_impl(h, verts, vcap, faces, fcap, nverts, nfaces);
  } catch (const std::exception& e) {
    return fail(B200ISO_ENOMEM, "b200iso_extract_host_resident: %s", e.what());
  } catch (...) {
    return fail(B200ISO_ENOMEM, "b200iso_extract_host_resident: unknown C++ exception");
  }
}

int b200iso_extract_host(b200iso_handle* h, const b200iso_params* p, const void* sdf, int64_t nx, int64_t ny, int64_t nz,
                         int64_t ldx, void* verts, int64_t vcap, int64_t* faces, int64_t fcap, int64_t* nverts,
                         int64_t* nfaces, int* vert_is_f64) {
  try {
    return b200iso_extract_host_impl(h, p, sdf, nx, ny, nz, ldx, verts, vcap, faces, fcap, nverts, nfaces, vert_is_f64);
  } catch (const std::exception& e) {  // (std::bad_alloc, std::system_error from a thread): nothing may cross the C ABI
    return fail(B200ISO_ENOMEM, "b200iso_extract_host: %s", e.what());
  } catch (...) {
    return fail(B200ISO_ENOMEM, "b200iso_extract_host: unknown C++ exception");
  }
}

int b200iso_case_indices(b200iso_handle* h, uint8_t* out, int mem) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (!h->counted) return fail(B200ISO_ESTATE, "no counted field");
  DeviceGuard guard(h->device);
  const iso::Grid& g = h->grid;
  if (h->nblocks == 0) return 0;
  const long long nvox = (long long)(g.nx - 1) * (g.ny - 1) * (g.nz - 1);
  if (!out) return fail(B200ISO_EINVAL, "out is NULL");
  uint8_t* dev = out;
  if (mem == B200ISO_HOST) {
    if (int rc = h->cases.reserve((size_t)nvox)) return rc;
    dev = h->cases.p;
  }
  const unsigned nb = (unsigned)((nvox + 255) / 256);
  if (h->prm.algo == B200ISO_MC) iso::case_kernel<0><<<nb, 256, 0, h->stream>>>(h->bits.p, g, dev, nvox);
  else iso::case_kernel<1><<<nb, 256, 0, h->stream>>>(h->bits.p, g, dev, nvox);
  cudaError_t e = cudaGetLastError();
  h->launches++;
  if (e == cudaSuccess && mem == B200ISO_HOST) e = cudaMemcpyAsync(out, dev, (size_t)nvox, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(B200ISO_ECUDA, "case_indices: %s", cudaGetErrorString(e));
  return 0;
}

// ---- mesh consumers (csrc/mesh_consumers.cuh) ---------------------------------------------------------
int b200iso_vertex_normals_async(b200iso_handle* h, const b200iso_params* p, const void* sdf_dev, int64_t nx, int64_t ny, int64_t nz,
                                 int64_t ldx, const void* verts_dev, int64_t nverts, int vert_is_f64, float* normals_dev) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (int rc = check_params(p, nx, ny, nz, ldx)) return rc;
  if (nverts < 0 || (nverts > 0 && (!sdf_dev || !verts_dev || !normals_dev))) return fail(B200ISO_EINVAL, "NULL argument");
  if (nx < 1 || ny < 1 || nz < 1) return nverts == 0 ? 0 : fail(B200ISO_EINVAL, "empty field");
  if (nverts == 0) return 0;
  DeviceGuard guard(h->device);
  iso::NormalArgs a{};
  a.sdf = sdf_dev, a.field_is_f64 = p->field_is_f64, a.nx = (int)nx, a.ny = (int)ny, a.nz = (int)nz, a.ldx = ldx, a.plane = ldx * ny;
  a.x0 = p->x0, a.x1 = p->x1, a.y0 = p->y0, a.y1 = p->y1, a.z0 = p->z0, a.z1 = p->z1;
  if (p->range_kind == B200ISO_RANGE_F32) a.x0 = (float)a.x0, a.x1 = (float)a.x1, a.y0 = (float)a.y0, a.y1 = (float)a.y1, a.z0 = (float)a.z0, a.z1 = (float)a.z1;
  a.x_offset = p->x_offset, a.nx_global = p->nx_global > 0 ? p->nx_global : nx;
  a.verts = verts_dev, a.vert_is_f64 = vert_is_f64, a.nverts = nverts, a.normals = normals_dev;
  const unsigned nb = (unsigned)((nverts + 255) / 256);
  if (p->field_is_f64) {
    if (vert_is_f64) iso::vertex_normals_kernel<double, double><<<nb, 256, 0, h->stream>>>(a);
    else iso::vertex_normals_kernel<double, float><<<nb, 256, 0, h->stream>>>(a);
  } else {
    if (vert_is_f64) iso::vertex_normals_kernel<float, double><<<nb, 256, 0, h->stream>>>(a);
    else iso::vertex_normals_kernel<float, float><<<nb, 256, 0, h->stream>>>(a);
  }
  CU(cudaGetLastError());
  h->launches++;
  return 0;
}

int b200iso_vertex_keys_async(b200iso_handle* h, int64_t* keys_dev, int64_t kcap) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (!h->counted) return fail(B200ISO_ESTATE, "vertex keys need a counted field (b200iso_count / b200iso_count_async first)");
  if (h->prm.algo != B200ISO_MC) return fail(B200ISO_EINVAL, "vertex keys are for Marching Cubes (Marching Tetrahedra vertices are already shared)");
  if (kcap < 0 || (kcap > 0 && !keys_dev)) return fail(B200ISO_EINVAL, "bad keys buffer");
  if (h->nblocks == 0) return 0;
  DeviceGuard guard(h->device);
  iso::GenArgs a{};
  a.bits = h->bits.p, a.woff = h->woff.p, a.coords = h->coords.p, a.recs = h->recs.p, a.nrecs = h->nrecs.p;
  a.verts = keys_dev, a.vcap = kcap, a.fcap = 0, a.nblocks = h->nblocks, a.totals_a = h->totals_dev, a.abort_flag = nullptr;
  a.key_nx_global = h->prm.nx_global > 0 ? h->prm.nx_global : h->grid.nx;
  iso::mc_generate_kernel<0, float, true><<<(unsigned)h->nblocks, iso::CB_THREADS, 0, h->stream>>>(a, h->grid);
  CU(cudaGetLastError());
  h->launches++;
  return 0;
}

static int b200iso_weld_impl(b200iso_handle* h, const int64_t* keys_dev, const void* verts_dev, int64_t nverts, int vert_is_f64,
                             const int64_t* faces_dev, int64_t nfaces, int64_t vertex_base, void* verts_out_dev, int64_t* faces_out_dev,
                             int64_t* nwelded) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  if (nverts < 0 || nfaces < 0 || nverts >= (1ll << 31)) return fail(B200ISO_EINVAL, "bad sizes (at most 2^31 - 1 vertices)");
  if (nwelded) *nwelded = 0;
  if (nverts == 0) return 0;
  if (!keys_dev || !verts_dev || !verts_out_dev || (nfaces > 0 && (!faces_dev || !faces_out_dev))) return fail(B200ISO_EINVAL, "NULL argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = h->stream;
  unsigned long long slots = 1;
  while (slots < 2ull * (unsigned long long)nverts) slots <<= 1;
  if (int rc = h->weld_tab.reserve((size_t)slots * 2)) return rc;
  if (int rc = h->weld_u32.reserve((size_t)nverts * 3 + 4)) return rc;
  unsigned long long* tab_key = h->weld_tab.p;
  long long* tab_min = reinterpret_cast<long long*>(h->weld_tab.p + slots);
  unsigned int *slot_of = h->weld_u32.p, *keep = slot_of + nverts, *newidx = keep + nverts;
  CU(cudaMemsetAsync(tab_key, 0, slots * sizeof(unsigned long long), st));
  CU(cudaMemsetAsync(tab_min, 0xff, slots * sizeof(long long), st));  // (unsigned max: atomicMin on the unsigned view)
  const unsigned nb = (unsigned)((nverts + 255) / 256);
  iso::weld_insert_kernel<<<nb, 256, 0, st>>>((const long long*)keys_dev, nverts, tab_key, tab_min, slots - 1, slot_of);
  iso::weld_flag_kernel<<<nb, 256, 0, st>>>(slot_of, tab_min, nverts, keep);
  CU(cudaGetLastError());
  size_t tmp_bytes = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, keep, newidx, (int)nverts, st));
  if (int rc = h->weld_tmp.reserve(tmp_bytes + 16)) return rc;
  CU(cub::DeviceScan::ExclusiveSum(h->weld_tmp.p, tmp_bytes, keep, newidx, (int)nverts, st));
  if (vert_is_f64) iso::weld_compact_kernel<double><<<nb, 256, 0, st>>>((const double*)verts_dev, keep, newidx, nverts, (double*)verts_out_dev);
  else iso::weld_compact_kernel<float><<<nb, 256, 0, st>>>((const float*)verts_dev, keep, newidx, nverts, (float*)verts_out_dev);
  if (nfaces > 0)
    iso::weld_faces_kernel<<<(unsigned)((nfaces * 3 + 255) / 256), 256, 0, st>>>((const long long*)faces_dev, nfaces * 3, vertex_base, slot_of, tab_min,
                                                                                newidx, (long long*)faces_out_dev);
  CU(cudaGetLastError());
  h->launches += 5;
  unsigned int last[2] = {0, 0};
  CU(cudaMemcpyAsync(&last[0], newidx + nverts - 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(&last[1], keep + nverts - 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (nwelded) *nwelded = (int64_t)last[0] + last[1];
  return 0;
}

int b200iso_weld(b200iso_handle* h, const int64_t* keys_dev, const void* verts_dev, int64_t nverts, int vert_is_f64, const int64_t* faces_dev,
                 int64_t nfaces, int64_t vertex_base, void* verts_out_dev, int64_t* faces_out_dev, int64_t* nwelded) {
  try {
    return b200iso_weld_impl(h, keys_dev, verts_dev, nverts, vert_is_f64, faces_dev, nfaces, vertex_base, verts_out_dev, faces_out_dev, nwelded);
  } catch (const std::exception& e) {
    return fail(B200ISO_ENOMEM, "b200iso_weld: %s", e.what());
  } catch (...) {
    return fail(B200ISO_ENOMEM, "b200iso_weld: unknown C++ exception");
  }
}

// Binary little-endian PLY: vertex x y z (float or double) [+ nx ny nz float], face = uchar 3 + 3 x int32 (0-based).
int b200iso_write_ply(const char* path, const void* verts, int64_t nverts, int vert_is_f64, const float* normals, const int64_t* faces,
                      int64_t nfaces) {
  if (!path || nverts < 0 || nfaces < 0 || (nverts > 0 && !verts) || (nfaces > 0 && !faces)) return fail(B200ISO_EINVAL, "bad argument");
  if (nverts >= (1ll << 31)) return fail(B200ISO_EINVAL, "PLY faces are written with 32-bit indices: at most 2^31 - 1 vertices");
  FILE* f = fopen(path, "wb");
  if (!f) return fail(B200ISO_EINVAL, "cannot open %s for writing", path);
  const char* vt = vert_is_f64 ? "double" : "float";
  fprintf(f, "ply\nformat binary_little_endian 1.0\ncomment written by libb200iso\nelement vertex %lld\nproperty %s x\nproperty %s y\nproperty %s z\n",
          (long long)nverts, vt, vt, vt);
  if (normals) fprintf(f, "property float nx\nproperty float ny\nproperty float nz\n");
  fprintf(f, "element face %lld\nproperty list uchar int vertex_indices\nend_header\n", (long long)nfaces);
  const size_t vsz = vert_is_f64 ? 8 : 4;
  bool ok = true;
  if (!normals) {
    ok = fwrite(verts, 3 * vsz, (size_t)nverts, f) == (size_t)nverts;
  } else {
    std::vector<unsigned char> row(3 * vsz + 12);
    for (int64_t i = 0; i < nverts && ok; ++i) {
      memcpy(row.data(), (const unsigned char*)verts + (size_t)i * 3 * vsz, 3 * vsz);
      memcpy(row.data() + 3 * vsz, normals + 3 * i, 12);
      ok = fwrite(row.data(), row.size(), 1, f) == 1;
    }
  }
  std::vector<unsigned char> buf;
  buf.reserve((size_t)13 * 4096);
  for (int64_t i = 0; i < nfaces && ok; ++i) {
    const int32_t t[3] = {(int32_t)(faces[3 * i] - 1), (int32_t)(faces[3 * i + 1] - 1), (int32_t)(faces[3 * i + 2] - 1)};
    buf.push_back(3);
    buf.insert(buf.end(), (const unsigned char*)t, (const unsigned char*)t + 12);
    if (buf.size() >= (size_t)13 * 4096 || i == nfaces - 1) {
      ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
      buf.clear();
    }
  }
  ok = (fclose(f) == 0) && ok;
  return ok ? 0 : fail(B200ISO_EINVAL, "short write to %s", path);
}

// Binary STL: 80-byte header, uint32 count, per triangle: float normal[3] (of the triangle), 3 x float vertex[3], uint16 0.
int b200iso_write_stl(const char* path, const void* verts, int64_t nverts, int vert_is_f64, const int64_t* faces, int64_t nfaces) {
  if (!path || nverts < 0 || nfaces < 0 || (nverts > 0 && !verts) || (nfaces > 0 && !faces)) return fail(B200ISO_EINVAL, "bad argument");
  if (nfaces >= (1ll << 32)) return fail(B200ISO_EINVAL, "binary STL holds at most 2^32 - 1 triangles");
  FILE* f = fopen(path, "wb");
  if (!f) return fail(B200ISO_EINVAL, "cannot open %s for writing", path);
  char hdr[80];
  memset(hdr, 0, sizeof(hdr));
  snprintf(hdr, sizeof(hdr), "binary STL written by libb200iso");
  const uint32_t n32 = (uint32_t)nfaces;
  bool ok = fwrite(hdr, 1, 80, f) == 80 && fwrite(&n32, 4, 1, f) == 1;
  auto vtx = [&](int64_t idx, float out[3]) {
    for (int q = 0; q < 3; ++q) out[q] = vert_is_f64 ? (float)((const double*)verts)[3 * idx + q] : ((const float*)verts)[3 * idx + q];
  };
  std::vector<unsigned char> buf;
  buf.reserve((size_t)50 * 4096);
  for (int64_t i = 0; i < nfaces && ok; ++i) {
    float rec[12];
    for (int c = 0; c < 3; ++c) {
      const int64_t idx = faces[3 * i + c] - 1;
      if (idx < 0 || idx >= nverts) {
        fclose(f);
        return fail(B200ISO_EINVAL, "face %lld references vertex %lld of %lld", (long long)i, (long long)(idx + 1), (long long)nverts);
      }
      vtx(idx, rec + 3 + 3 * c);
    }
    const float u[3] = {rec[6] - rec[3], rec[7] - rec[4], rec[8] - rec[5]}, w[3] = {rec[9] - rec[3], rec[10] - rec[4], rec[11] - rec[5]};
    float n[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
    const float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int q = 0; q < 3; ++q) rec[q] = len > 0 ? n[q] / len : 0.f;
    const uint16_t attr = 0;
    buf.insert(buf.end(), (const unsigned char*)rec, (const unsigned char*)rec + 48);
    buf.insert(buf.end(), (const unsigned char*)&attr, (const unsigned char*)&attr + 2);
    if (buf.size() >= (size_t)50 * 4096 || i == nfaces - 1) {
      ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
      buf.clear();
    }
  }
  ok = (fclose(f) == 0) && ok;
  return ok ? 0 : fail(B200ISO_EINVAL, "short write to %s", path);
}

int b200iso_enable_timing(b200iso_handle* h, int on) {
  if (!h) return fail(B200ISO_EINVAL, "handle is NULL");
  DeviceGuard guard(h->device);
  if (on && !h->ev) {
    const int n = b200iso_handle::NSLOT * b200iso_handle::E_N;
    h->ev = new cudaEvent_t[n];
    h->ev_set = new unsigned char[n]();
    for (int i = 0; i < n; ++i) CU(cudaEventCreate(&h->ev[i]));
  }
  h->timing = on != 0;
  h->step = 0;
  return 0;
}

int b200iso_timings(b200iso_handle* h, float* ms, int n) {
  if (!h || !ms) return fail(B200ISO_EINVAL, "NULL argument");
  using H = b200iso_handle;
  double sum[5] = {0, 0, 0, 0, 0};
  long long cnt[5] = {0, 0, 0, 0, 0};
  if (h->ev && h->step > 0) {
    DeviceGuard guard(h->device);
    CU(cudaStreamSynchronize(h->stream));
    const long long nsteps = h->step < H::NSLOT ? h->step : H::NSLOT;
    static const int span[5][2] = {{H::E_C0, H::E_C1}, {H::E_C1, H::E_C2}, {H::E_G0, H::E_G1}, {H::E_H0, H::E_H1}, {H::E_D0, H::E_D1}};
    for (long long s = 0; s < nsteps; ++s)
      for (int k = 0; k < 5; ++k) {
        const int a = (int)s * H::E_N + span[k][0], b = (int)s * H::E_N + span[k][1];
        if (!h->ev_set[a] || !h->ev_set[b]) continue;
        float t = 0;
        if (cudaEventElapsedTime(&t, h->ev[a], h->ev[b]) == cudaSuccess) sum[k] += t, cnt[k]++;
      }
    cudaGetLastError();
  }
  for (int i = 0; i < n && i < 5; ++i) ms[i] = cnt[i] ? (float)(sum[i] / cnt[i]) : 0.f;
  return 0;
}

int64_t b200iso_launch_count(b200iso_handle* h) { return h ? h->launches : 0; }

}  // extern "C"
