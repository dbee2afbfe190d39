// signpack_tma.cuh -- classify (sign-pack) with TMA staging: cp.async.bulk.tensor 3-D boxes + mbarrier pipeline.
//
// Same task decomposition and output as signpack_kernel (iso_kernels.cuh): a warp owns (128-sample x-segment, y,
// 512-z chunk).  Instead of per-lane 128-bit loads, lane 0 of every warp keeps TM_STAGES TMA boxes in flight
// (box = 128 x * 1 y * TM_BZ z Float32 = 4 KB, landing dense in the warp's slice of shared memory, completion on
// an mbarrier with expect_tx); the 32 lanes then read conflict-free LDS.128 rows and build the z-packed words.
// Out-of-range samples (x >= nx, z >= nz) are filled with NaN by the tensor map (NaN < iso is false, like the
// reference's compare), so no lane-side bounds logic is needed.  Each warp runs its own pipeline: no block barrier.
#pragma once
#include <cuda.h>

#include "count_kernel.cuh"

namespace iso {

#ifndef ISO_TM_WARPS
#define ISO_TM_WARPS 4
#endif
#ifndef ISO_TM_BZ
#define ISO_TM_BZ 8
#endif
#ifndef ISO_TM_STAGES
#define ISO_TM_STAGES 4
#endif
constexpr int TM_WARPS = ISO_TM_WARPS;    // warps per CTA, each with a private pipeline
constexpr int TM_BZ = ISO_TM_BZ;          // z-planes per TMA box
constexpr int TM_STAGES = ISO_TM_STAGES;  // boxes in flight per warp
constexpr int TM_BOX_FLOATS = SP_XSEG * TM_BZ;                 // 1024 floats = 4 KB
constexpr int TM_BOXES = SP_ZW * 32 / TM_BZ;                   // boxes per warp task
constexpr size_t TM_SMEM_PER_WARP = (size_t)TM_STAGES * TM_BOX_FLOATS * 4 + (size_t)SP_ZW * SP_XSEG * 4;
constexpr size_t TM_SMEM = TM_WARPS * TM_SMEM_PER_WARP + 128;  // + alignment slack

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// ---- counting warps riding along in the classify CTAs ------------------------------------------------------
// The classify warps are bound by HBM and leave three quarters of the SM's issue slots idle; the Marching Cubes count
// (0.17 ms as a kernel of its own at 1024^3: issue-bound, reads only the bit-field) fits in them.  Every classify
// CTA therefore carries a few extra warps that count generate blocks as soon as the bit-field rows they need are
// complete.  The generate blocks of one "y-block" bi (block-in-row index: the same rows for every x) form one queue:
//   - classify tasks are numbered y-major (all x-segments and z-chunks of row y, then row y + 1, ...), so rows
//     complete in order while the kernel runs; a finished task bumps rows_done[y] (release);
//   - a counting warp reads the current y-block (cur_bi), checks rows_done[] of its rows against the step's target
//     (acquire), takes the next TM_CNT_BATCH values of x with ONE fetch-add on that y-block's counter (a
//     compare-and-swap on a single queue head makes a thousand warps fight for the same item: measured 300 claims
//     per step) and counts them from L2 (ld.global.cg: L1 is not coherent); an exhausted y-block advances cur_bi;
//   - counting warps poll only while their CTA's classify warps still stream and leave with them (after the batch
//     in hand), so they hold the CTA's shared memory against the next classify CTA only for that long.
// Whatever is left when the kernel ends (the y-blocks of the last rows) is counted by mc_count_chunks_kernel, which
// skips x < next_x[bi] -- the result never depends on how much was claimed here.
// rows_done[] counts cumulatively over the steps (target = step number x tasks per row): no reset between steps;
// cur_bi, next_x[] and the statistics word are zeroed by the scan kernel at the end of the step.
constexpr int TM_CNT_WARPS_MAX = 8;  // counting warps per CTA: a launch parameter (blockDim), 4 by default
constexpr unsigned TM_CNT_BATCH = 1;  // generate blocks per claim (more per claim = longer overhang past the classify warps)
// layout of the ride buffer (unsigned int): [0] cur_bi, [1] claimed (statistics), [32 .. 32 + nbi) next_x, then rows_done[ny]
constexpr int RIDE_HDR = 32;
struct CountRide {
  Grid g;                      // geometry of the count (same as generate)
  int nbi;                     // y-blocks = generate blocks per voxel x-row; 0 = no counting warps
  unsigned long long* woff;    // raw (vertex, face) pair per generate block
  uint32_t* recs;              // active-voxel records per generate block (REC_CAP each) ...
  uint32_t* nrecs;             // ... and their number
  unsigned int* head;          // [0] cur_bi, [1] claimed
  unsigned int* next_x;        // [nbi]
  unsigned int* rows_done;     // [ny] finished classify tasks per sample row, cumulative over steps
  unsigned int target;         // rows_done[y] >= target  <=>  row y is complete in this step
  int stop_at;                 // counting warps stop claiming once the CTA's classify warps have finished this many z-words
  uint32_t* celloff;           // Marching Tetrahedra count (signpack_tma_kernel<true>): in-block vertex prefix of every cell
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// cta_prog: z-words finished by the CTA's classify warps (ZW per warp, see the kernel).  A counting warp that is still counting
// when they are through keeps the CTA -- and its 96 KB of shared memory -- from the next classify CTA (measured:
// classify 0.66 -> 0.72 ms with unbounded claiming), so claiming stops when the classify warps near their end.
template <bool MT>
__device__ __forceinline__ void count_ride(const uint32_t* __restrict__ bits, const CountRide& cr, const volatile int* cta_prog,
                                           const uint8_t* nf_s) {
  const int lane = threadIdx.x & 31;
  const Grid& g = cr.g;
  const unsigned nxv = (unsigned)(g.nx - 1);
  unsigned ready_bi = 0xffffffffu;  // y-block this warp has already seen complete
  while (true) {
    if (*cta_prog >= cr.stop_at) return;  // this CTA's classify warps are nearly through: leave the rest to later CTAs
    unsigned int bi = 0;
    if (lane == 0) bi = ld_relaxed_u32(cr.head);
    bi = __shfl_sync(0xffffffffu, bi, 0);
    if (bi >= (unsigned)cr.nbi) return;
    if (bi != ready_bi) {
      // sample rows the y-block needs: the voxel rows of its quad-cells plus one
      const int y_lo = (int)fast_div(bi * CB_THREADS, g.wq_mul, g.wq_sh);
      int q_hi = (int)bi * CB_THREADS + CB_THREADS - 1;
      if (q_hi > g.quads_per_row - 1) q_hi = g.quads_per_row - 1;
      const int y_hi = (int)fast_div((unsigned)q_hi, g.wq_mul, g.wq_sh) + 1;
      bool ready = true;
      for (int y = y_lo + lane; y <= y_hi; y += 32) ready = ready && (int)(ld_acquire_u32(cr.rows_done + y) - cr.target) >= 0;
      if (!__all_sync(0xffffffffu, ready)) {
        __nanosleep(4000);
        continue;
      }
      ready_bi = bi;
    }
    unsigned int x0 = 0;
    if (lane == 0) x0 = atomicAdd(cr.next_x + bi, TM_CNT_BATCH);
    x0 = __shfl_sync(0xffffffffu, x0, 0);
    if (x0 >= nxv) {  // this y-block is handed out: move on
      if (lane == 0) atomicMax(cr.head, bi + 1u);
      continue;
    }
    const unsigned n = min(TM_CNT_BATCH, nxv - x0);
    for (unsigned k = 0; k < n; ++k) {
      const long long chunk = (long long)(x0 + k) * g.blocks_per_row + bi;
      uint32_t nv, nf;
      if constexpr (MT) mt_count_chunk<true>(bits, g, chunk, nf_s, cr.celloff, cr.recs, cr.nrecs, nv, nf);
      else mc_count_chunk<true>(bits, g, chunk, nf_s, cr.recs, cr.nrecs, nv, nf);
      if (lane == 0) cr.woff[2 * chunk] = nv, cr.woff[2 * chunk + 1] = nf;
    }
    if (lane == 0) atomicAdd(cr.head + 1, n);
  }
}

// MT: the counting warps run the Marching Tetrahedra count, which needs more registers than the classify warps' 80: that
// instantiation is bounded to 6 counting warps and two CTAs per SM (the classify pipeline is sized for two).
constexpr int TM_CNT_WARPS_MT = 6;
// (Tasks of 8 or 4 z-words instead of 16 -- more, shorter tasks, so that the counting warps also find rows to follow on
// 512^3 grids and 129-plane slabs -- were measured and lost: 512^3 classify 0.093 -> 0.128 ms for 0.021 ms less count.)
template <bool MT>
__global__ void __launch_bounds__((TM_WARPS + (MT ? TM_CNT_WARPS_MT : TM_CNT_WARPS_MAX)) * 32, MT ? 2 : 0)
signpack_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint32_t* __restrict__ bits, int nx, int ny, int nz, int W,
                    float thresh, int nxseg, int nzc, long long ntasks, unsigned long long* __restrict__ clear, int nclear,
                    const __grid_constant__ CountRide cr) {
  extern __shared__ __align__(128) unsigned char tm_smem[];
  __shared__ __align__(8) uint64_t full[TM_WARPS][TM_STAGES];
  __shared__ uint8_t nf_s[256];
  constexpr int ZW = SP_ZW, NBOXES = TM_BOXES;  // z-words and boxes per task
  __shared__ int cta_prog;  // z-words finished by the classify warps of this CTA (ZW each)
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // block 0 resets the scan state (ticket + look-back chain) of the count/scan kernels that follow in the stream
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < nclear; i += blockDim.x) clear[i] = 0ull;
  if (threadIdx.x == 0) cta_prog = 0;
  if (cr.nbi > 0)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) nf_s[i] = MT ? ISO_MT_NF[i] : (uint8_t)((ISO_MC_VERTS[i] >> 52) & 7);
  __syncthreads();
  if (wib >= TM_WARPS) {  // ---- counting warps ----
    if (cr.nbi > 0) count_ride<MT>(bits, cr, &cta_prog, nf_s);
    return;
  }
  // ---- classify warps ----
  const long long task = (long long)blockIdx.x * TM_WARPS + wib;
  if (task >= ntasks) {
    if (lane == 0) atomicAdd(&cta_prog, ZW);
    return;
  }
  unsigned char* base = tm_smem + (128 - (smem_u32(tm_smem) & 127)) % 128 + (size_t)wib * TM_SMEM_PER_WARP;
  float* box = reinterpret_cast<float*>(base);                                                      // [TM_STAGES][TM_BZ][128]
  uint32_t* wstage = reinterpret_cast<uint32_t*>(base + (size_t)TM_STAGES * TM_BOX_FLOATS * 4);      // [ZW][128]
  // y-major task order: rows complete one after the other while the kernel runs (the counting warps follow them)
  const int xseg = (int)(task % nxseg);
  const long long t2 = task / nxseg;
  const int zc = (int)(t2 % nzc);
  const int y = (int)(t2 / nzc);
  const int x0 = xseg * SP_XSEG, z0 = zc * ZW * 32;
  // boxes of this task that contain at least one valid z-plane
  const int nbox = min(NBOXES, (nz - z0 + TM_BZ - 1) / TM_BZ);

  if (lane == 0) {
    for (int s = 0; s < TM_STAGES; ++s) mbar_init(&full[wib][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int s = 0; s < TM_STAGES && s < nbox; ++s) {
      mbar_expect_tx(&full[wib][s], TM_BOX_FLOATS * 4);
      tma_load_3d(box + s * TM_BOX_FLOATS, &tmap, x0, y, z0 + s * TM_BZ, &full[wib][s]);
    }
  }
  __syncwarp();

  uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
  for (int b = 0; b < NBOXES; ++b) {
    const int s = b % TM_STAGES;
    if (b < nbox) {
      mbar_wait(&full[wib][s], (uint32_t)((b / TM_STAGES) & 1));
      const float4* row = reinterpret_cast<const float4*>(box + s * TM_BOX_FLOATS) + lane;
      const int sh = (b * TM_BZ) & 31;
#pragma unroll
      for (int k = 0; k < TM_BZ; ++k) {
        const float4 v = row[k * (SP_XSEG / 4)];
        w0 |= (v.x < thresh) ? (1u << (sh + k)) : 0u;
        w1 |= (v.y < thresh) ? (1u << (sh + k)) : 0u;
        w2 |= (v.z < thresh) ? (1u << (sh + k)) : 0u;
        w3 |= (v.w < thresh) ? (1u << (sh + k)) : 0u;
      }
      __syncwarp();  // every lane has read the stage: it may be overwritten
      if (lane == 0 && b + TM_STAGES < nbox) {
        mbar_expect_tx(&full[wib][s], TM_BOX_FLOATS * 4);
        tma_load_3d(box + s * TM_BOX_FLOATS, &tmap, x0, y, z0 + (b + TM_STAGES) * TM_BZ, &full[wib][s]);
      }
    }
    if (((b + 1) * TM_BZ) % 32 == 0) {  // a z-word is complete
      *reinterpret_cast<uint4*>(&wstage[(b * TM_BZ / 32) * SP_XSEG + lane * 4]) = make_uint4(w0, w1, w2, w3);
      w0 = w1 = w2 = w3 = 0;
      if (lane == 0) atomicAdd(&cta_prog, 1);
    }
  }
  // write-out, identical to signpack_kernel<true>: column j of this lane, ZW words as 16-byte stores
  uint4 r[ZW];
#pragma unroll
  for (int zw = 0; zw < ZW; ++zw) r[zw] = *reinterpret_cast<const uint4*>(&wstage[zw * SP_XSEG + lane * 4]);
  const int wofs = zc * ZW;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = x0 + lane * 4 + j;
    if (x < nx) {
      uint32_t* dst = bits + ((long long)x * ny + y) * W + wofs;
      uint32_t c[ZW];
#pragma unroll
      for (int zw = 0; zw < ZW; ++zw) c[zw] = j == 0 ? r[zw].x : j == 1 ? r[zw].y : j == 2 ? r[zw].z : r[zw].w;
#pragma unroll
      for (int c4 = 0; c4 < ZW; c4 += 4)
        if (wofs + c4 < W) *reinterpret_cast<uint4*>(dst + c4) = make_uint4(c[c4], c[c4 + 1], c[c4 + 2], c[c4 + 3]);
    }
  }
  // release: this task's words are visible before the row counter moves
  __threadfence();
  __syncwarp();
  if (lane == 0) {
    if (cr.nbi > 0) atomicAdd(cr.rows_done + y, 1u);
  }
}

// Host side: tensor map over the field (x innermost), encoded through the driver entry point (no libcuda link).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
  }
  return fn;
}

// returns false when the field cannot be described by a tensor map (caller falls back to the LDG kernel)
inline bool make_field_tmap(CUtensorMap* map, const float* sdf, long long nx, long long ny, long long nz, long long ldx) {
  EncodeTiledFn fn = tma_encode_fn();
  if (!fn) return false;
  if ((reinterpret_cast<uintptr_t>(sdf) & 15) != 0 || ldx % 4 != 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
  const cuuint64_t strides[2] = {(cuuint64_t)ldx * 4, (cuuint64_t)ldx * ny * 4};
  const cuuint32_t box[3] = {(cuuint32_t)SP_XSEG, 1, (cuuint32_t)TM_BZ};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (strides[1] >= (1ull << 40)) return false;
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(sdf), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA) == CUDA_SUCCESS;
}

}  // namespace iso
