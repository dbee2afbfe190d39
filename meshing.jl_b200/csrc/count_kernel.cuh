// count_kernel.cuh -- (2) count + single-pass decoupled look-back scan.  ALGO 0 = MC, 1 = MT.
//
// One thread per quad-cell (128 z-consecutive voxels of one (x,y) voxel column), 256 quad-cells per block,
// blocks numbered in the reference's scan order (x, then y, then z) through an atomic ticket.
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
  __shared__ unsigned long long red_s[CB_THREADS / 32];
  __shared__ long long sb;
  if (threadIdx.x == 0) sb = atomicAdd(ticket, 1u);
  nf_s[threadIdx.x] = ALGO == 0 ? (uint8_t)((ISO_MC_VERTS[threadIdx.x] >> 52) & 7) : ISO_MT_NF[threadIdx.x];
  __syncthreads();
  const long long b = sb;
  int x, quad0;
  block_coords(g, b, x, quad0);
  const int qr = quad0 + threadIdx.x;
  uint32_t nv = 0, nf = 0;
  uint32_t cv[4] = {0, 0, 0, 0};
  int y = 0, zq = 0;
  const bool live = qr < g.quads_per_row;
  if (live) {
    y = qr / g.Wq, zq = qr - y * g.Wq;
    Quad q;
    load_quad(bits, g, x, y, zq, q);
    const int fxy = (x == 0 ? 1 : 0) | (y == 0 ? 2 : 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t mm = active_mask(q, i);
      if (mm) {
        cv[i] = ALGO == 0 ? mc_nverts_masked(q, i, q.vm[i]) : mt_owned_masked(q, i, q.vm[i], fxy, zq == 0 && i == 0);
        nv += cv[i];
        while (mm) {
          const int k = __ffs(mm) - 1;
          mm &= mm - 1;
          nf += nf_s[case_of<ALGO>(q, i, k)];
        }
      }
    }
  }
  // block totals (and, for MT, the exclusive prefix of every cell inside the block)
  unsigned long long total;
  const unsigned long long excl = block_excl_scan((unsigned long long)nv | ((unsigned long long)nf << 32), red_s, total);
  if (ALGO == 1 && live) {
    const uint32_t e0 = (uint32_t)excl;
    uint4 o;
    o.x = e0, o.y = e0 + cv[0], o.z = o.y + cv[1], o.w = o.z + cv[2];
    *reinterpret_cast<uint4*>(celloff + (long long)x * g.row_words + (long long)y * g.W + zq * 4) = o;
  }
  if (threadIdx.x < 32) {
    const unsigned long long av = total & 0xffffffffull, af = total >> 32;
    unsigned long long ev, ef;
    lookback(status, b, av, af, ev, ef);
    if (b == nblocks - 1 && threadIdx.x == 0) {
      totals_a[0] = (long long)(ev + av), totals_a[1] = (long long)(ef + af);
      if (totals_b) totals_b[0] = (long long)(ev + av), totals_b[1] = (long long)(ef + af);
    }
  }
}

}  // namespace iso
