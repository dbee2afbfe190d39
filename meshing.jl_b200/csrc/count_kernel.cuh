// count_kernel.cuh -- (2) count + single-pass decoupled look-back scan.  ALGO 0 = MC, 1 = MT.
//
// Same block/thread mapping as generate (iso_kernels.cuh, thread_map): a block owns whole voxel columns,
// lanes run across 32 adjacent columns, so empty regions are skipped per warp.  Blocks are numbered in the
// reference's scan order (x, then y, then z) through an atomic ticket and publish their totals to the
// look-back chain.
//   MC: vertices = crossed cube edges (12 masked popcounts per word), faces = table per active voxel
//   MT: vertices = owned crossed edges (7+ masked popcounts per word), faces = table per active voxel;
//       additionally writes celloff[cell] = vertices created in this block before the cell, which the
//       generate kernel uses to resolve vertex ids through owner voxels of other blocks.
#pragma once
#include "iso_kernels.cuh"
#include "mt_kernels.cuh"

namespace iso {

template <int ALGO>
__global__ void __launch_bounds__(CB_THREADS)
count_kernel(const uint32_t* __restrict__ bits, Grid g, unsigned long long* status, unsigned int* ticket,
             long long nblocks, long long* totals_a, long long* totals_b, uint32_t* __restrict__ celloff) {
  __shared__ uint8_t nf_s[256];
  __shared__ uint32_t s_val[CB_THREADS], s_w[CB_THREADS / 32];
  __shared__ uint32_t red_v[CB_THREADS / 32], red_f[CB_THREADS / 32];
  __shared__ long long sb;
  if (threadIdx.x == 0) sb = atomicAdd(ticket, 1u);
  nf_s[threadIdx.x] = ALGO == 0 ? (uint8_t)((ISO_MC_VERTS[threadIdx.x] >> 52) & 7) : ISO_MT_NF[threadIdx.x];
  __syncthreads();
  const long long b = sb;
  const TMap tm = thread_map(g, b);
  const int fxy = (tm.x == 0 ? 1 : 0) | (tm.y == 0 ? 2 : 0);
  uint32_t nv = 0, nf = 0;
  if (tm.live) {
    for (int zq = tm.zq_lo; zq < tm.zq_hi; ++zq) {
      Quad q;
      if (!load_quad(bits, g, tm.x, tm.y, zq, q)) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t mm = active_mask(q, i);
        if (mm) {
          nv += ALGO == 0 ? mc_nverts_masked(q, i, q.vm[i]) : mt_owned_masked(q, i, q.vm[i], fxy, zq == 0 && i == 0);
          while (mm) {
            const int k = __ffs(mm) - 1;
            mm &= mm - 1;
            nf += nf_s[case_of<ALGO>(q, i, k)];
          }
        }
      }
    }
  }
  if (ALGO == 1) {
    // in-block exclusive vertex prefix of every cell, in scan order (a thread's cells are consecutive)
    uint32_t tot;
    uint32_t run = block_excl_scan_ord(nv, tm.ord, s_val, s_w, tot);
    if (tm.live) {
      for (int zq = tm.zq_lo; zq < tm.zq_hi; ++zq) {
        Quad q;
        uint32_t cv[4] = {0, 0, 0, 0};
        if (load_quad(bits, g, tm.x, tm.y, zq, q)) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (active_mask(q, i)) cv[i] = mt_owned_masked(q, i, q.vm[i], fxy, zq == 0 && i == 0);
        }
        uint4 o;
        o.x = run, o.y = run + cv[0], o.z = o.y + cv[1], o.w = o.z + cv[2];
        run = o.w + cv[3];
        *reinterpret_cast<uint4*>(celloff + (long long)tm.x * g.row_words + (long long)tm.y * g.W + zq * 4) = o;
      }
    }
  }
  // block totals
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nv += __shfl_xor_sync(0xffffffffu, nv, o);
    nf += __shfl_xor_sync(0xffffffffu, nf, o);
  }
  if ((threadIdx.x & 31) == 0) red_v[threadIdx.x >> 5] = nv, red_f[threadIdx.x >> 5] = nf;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned long long av = 0, af = 0;
#pragma unroll
    for (int w = 0; w < CB_THREADS / 32; ++w) av += red_v[w], af += red_f[w];
    unsigned long long ev, ef;
    lookback(status, b, av, af, ev, ef);
    if (b == nblocks - 1 && threadIdx.x == 0) {
      totals_a[0] = (long long)(ev + av), totals_a[1] = (long long)(ef + af);
      if (totals_b) totals_b[0] = (long long)(ev + av), totals_b[1] = (long long)(ef + af);
    }
  }
}

}  // namespace iso
