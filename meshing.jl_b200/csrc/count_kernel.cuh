// count_kernel.cuh -- (2) count + single-pass decoupled look-back scan.
//
// Counting and scanning are separate kernels (a look-back chain over the counting blocks themselves makes every
// block wait for the slowest one ahead of it: measured 0.233 -> 0.180 ms on the gyroid, 0.355 -> 0.189 ms on the
// clustered multi-sphere/torus field):
//   MC: mc_count_chunks_kernel (a warp per generate block: raw (vertex, face) pair)  -> mc_scan_chunks_kernel
//       vertices = crossed cube edges (12 masked popcounts per word), faces = table per active voxel
//   MT: mt_count_kernel (block-cooperative, same thread mapping as generate)         -> mt_scan_blocks_kernel
//       vertices = owned crossed edges (7+ masked popcounts per word), faces = table per active voxel;
//       additionally writes celloff[cell] = vertices created in this block before the cell, which the
//       generate kernel uses to resolve vertex ids through owner voxels of other blocks.
#pragma once
#include "iso_kernels.cuh"
#include "mt_kernels.cuh"

namespace iso {

__global__ void __launch_bounds__(CB_THREADS)
mt_count_kernel(const uint32_t* __restrict__ bits, Grid g, uint32_t* __restrict__ celloff, unsigned long long* __restrict__ raw,
                uint32_t* __restrict__ recs, uint32_t* __restrict__ nrecs) {
  __shared__ uint8_t nf_s[256];
  __shared__ uint32_t s_w[CB_THREADS / 32];
  __shared__ uint32_t red_v[CB_THREADS / 32], red_f[CB_THREADS / 32];
  for (int i = threadIdx.x; i < 256; i += CB_THREADS) nf_s[i] = ISO_MT_NF[i];
  __syncthreads();
  const unsigned b = blockIdx.x;
  const TMap tm = thread_map(g, b);
  uint32_t nv = 0, nf = 0;
  uint32_t cv[4] = {0, 0, 0, 0};
  // Like the Marching Cubes count, this one leaves the block's active-voxel records for generate (scan order, 4 bytes:
  // case | voxel-in-quad-cell << 8 | quad-cell-in-block << 15, up to REC_CAP per block) and their number.
  Quad q;
  uint32_t mm4[4] = {0, 0, 0, 0}, na = 0;
  if (tm.live && load_quad(bits, g, tm.x, tm.y, tm.zq, q)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mm4[i] = active_mask(q, i), na += __popc(mm4[i]);
  }
  uint32_t blk_na;
  uint32_t pos = block_excl_scan_u32(na, s_w, blk_na);
  if (na) {
    const int fxy = ((tm.x + g.xoff) == 0 ? 1 : 0) | (tm.y == 0 ? 2 : 0);  // low-boundary flags use GLOBAL x
    uint32_t* rec = recs + (unsigned long long)b * (unsigned)g.rec_cap;
    const uint32_t qtag = (uint32_t)threadIdx.x << 15;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t mm = mm4[i];
      if (mm) {
        cv[i] = mt_owned_masked(q, i, q.vm[i], fxy, tm.zq == 0 && i == 0);
        nv += cv[i];
        while (mm) {
          const int k = __ffs(mm) - 1;
          mm &= mm - 1;
          const uint32_t c = case_of<1>(q, i, k);
          nf += nf_s[c];
          if (pos < (uint32_t)g.rec_cap) rec[pos] = c | ((uint32_t)(i * 32 + k) << 8) | qtag;
          ++pos;
        }
      }
    }
  }
  if (threadIdx.x == 0) nrecs[b] = blk_na;
  {
    // in-block exclusive vertex prefix of every cell (thread order == scan order)
    uint32_t tot;
    const uint32_t e0 = block_excl_scan_u32(nv, s_w, tot);
    if (tm.live) {
      uint4 o;
      o.x = e0, o.y = e0 + cv[0], o.z = o.y + cv[1], o.w = o.z + cv[2];
      *reinterpret_cast<uint4*>(celloff + (long long)tm.x * g.row_words + (long long)tm.y * g.W + tm.zq * 4) = o;
    }
  }
  // block totals -> raw[b]; mt_scan_blocks_kernel scans them afterwards
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nv += __shfl_xor_sync(0xffffffffu, nv, o);
    nf += __shfl_xor_sync(0xffffffffu, nf, o);
  }
  if ((threadIdx.x & 31) == 0) red_v[threadIdx.x >> 5] = nv, red_f[threadIdx.x >> 5] = nf;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long av = 0, af = 0;
#pragma unroll
    for (int w = 0; w < CB_THREADS / 32; ++w) av += red_v[w], af += red_f[w];
    raw[2 * (unsigned long long)b] = av, raw[2 * (unsigned long long)b + 1] = af;
  }
}

// ------------------------------------------------------------------------------------------------------
// Count, warp-autonomous: a warp walks one generate block's quad-cells (lane <-> quad-cell, 32 at a time) and stores
// the block's raw (vertex, face) pair -- no chain, no ticket, no barrier.  The Marching Cubes form is mc_count_chunk
// (iso_kernels.cuh); this is the Marching Tetrahedra form of mt_count_kernel above -- same records, same celloff, same
// totals -- used by the counting warps inside the TMA classify kernel and by the kernel that takes what they left.
template <bool CG>
__device__ __forceinline__ void mt_count_chunk(const uint32_t* __restrict__ bits, const Grid& g, long long chunk, const uint8_t* nf_s,
                                               uint32_t* __restrict__ celloff, uint32_t* __restrict__ recs, uint32_t* __restrict__ nrecs,
                                               uint32_t& nv_out, uint32_t& nf_out) {
  const int lane = threadIdx.x & 31;
  uint32_t nf = 0, base = 0, vbase = 0;  // records / vertices of the block before this group of 32 quad-cells (uniform)
  const int x = (int)(chunk / g.blocks_per_row);
  const int q_lo = (int)(chunk - (long long)x * g.blocks_per_row) * CB_THREADS;
  uint32_t* rec = recs + chunk * g.rec_cap;
  const int fx = (x + g.xoff) == 0 ? 1 : 0;  // low-boundary flags use GLOBAL x
  for (int q0 = q_lo; q0 < q_lo + CB_THREADS && q0 < g.quads_per_row; q0 += 32) {
    const int qr = q0 + lane;
    const bool live = qr < g.quads_per_row;
    const int y = (int)fast_div((unsigned)qr, g.wq_mul, g.wq_sh), zq = qr - y * g.Wq;
    Quad q;
    const bool act = live && load_quad<CG>(bits, g, x, y, zq, q);
    uint32_t mm[4] = {0, 0, 0, 0}, na = 0, cv[4] = {0, 0, 0, 0}, mynv = 0, e0 = vbase;
    if (act) {
#pragma unroll
      for (int i = 0; i < 4; ++i) mm[i] = active_mask(q, i), na += __popc(mm[i]);
    }
    if (__any_sync(0xffffffffu, na != 0)) {  // (uniform)
      uint32_t inc = na;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      uint32_t pos = base + inc - na;
      base += __shfl_sync(0xffffffffu, inc, 31);
      if (na) {
        const int fxy = fx | (y == 0 ? 2 : 0);
        const uint32_t qtag = (uint32_t)(qr - q_lo) << 15;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t m = mm[i];
          if (m) {
            cv[i] = mt_owned_masked(q, i, q.vm[i], fxy, zq == 0 && i == 0);
            mynv += cv[i];
            while (m) {
              const int k = __ffs(m) - 1;
              m &= m - 1;
              const uint32_t c = case_of<1>(q, i, k);
              nf += nf_s[c];
              if (pos < (uint32_t)g.rec_cap) rec[pos] = c | ((uint32_t)(i * 32 + k) << 8) | qtag;
              ++pos;
            }
          }
        }
      }
      // in-block exclusive vertex prefix of every cell (lane order == scan order)
      uint32_t vinc = mynv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, vinc, o);
        if (lane >= o) vinc += t;
      }
      e0 = vbase + vinc - mynv;
      vbase += __shfl_sync(0xffffffffu, vinc, 31);
    }
    if (live) {
      uint4 o;
      o.x = e0, o.y = e0 + cv[0], o.z = o.y + cv[1], o.w = o.z + cv[2];
      *reinterpret_cast<uint4*>(celloff + (long long)x * g.row_words + (long long)y * g.W + zq * 4) = o;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nf += __shfl_xor_sync(0xffffffffu, nf, o);
  if (lane == 0) nrecs[chunk] = base;
  nv_out = vbase, nf_out = nf;
}

constexpr int WC_THREADS = 256;  // 8 warps = 8 generate blocks per counting block

// ride: the counting warps inside the TMA classify kernel (signpack_tma.cuh) have already counted every y-block
// below ride[0] (cur_bi) and, of the y-blocks from there on, the generate blocks x < next_x[bi]; this kernel takes the
// rest with a grid-stride loop (everything, starting at item 0, when ride == nullptr).  Items are numbered
// y-block-major: item = bi * nxv + x.
constexpr int RIDE_HDR_WORDS = 32;  // == RIDE_HDR of signpack_tma.cuh: cur_bi, statistics, then next_x[]
// MT = true: the Marching Tetrahedra count (celloff is written as well), what the counting warps left of it.
template <bool MT>
__global__ void __launch_bounds__(WC_THREADS, 4)
mc_count_chunks_kernel(const uint32_t* __restrict__ bits, Grid g, long long nchunks, unsigned long long* __restrict__ woff,
                       uint32_t* __restrict__ recs, uint32_t* __restrict__ nrecs, const unsigned int* __restrict__ ride,
                       uint32_t* __restrict__ celloff) {
  __shared__ uint8_t nf_s[256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nxv = g.nx - 1;
  const long long first = ride ? (long long)__ldg(ride) * nxv : 0ll;
  if (first + (long long)blockIdx.x * (WC_THREADS / 32) >= nchunks) return;  // (uniform over the block)
  nf_s[threadIdx.x] = MT ? ISO_MT_NF[threadIdx.x] : (uint8_t)((ISO_MC_VERTS[threadIdx.x] >> 52) & 7);
  __syncthreads();
  for (long long item = first + (long long)blockIdx.x * (WC_THREADS / 32) + w; item < nchunks; item += (long long)gridDim.x * (WC_THREADS / 32)) {
    const long long bi = item / nxv, x = item - bi * nxv;
    if (ride && x < (long long)__ldg(ride + RIDE_HDR_WORDS + bi)) continue;  // counted inside the classify kernel
    const long long chunk = x * g.blocks_per_row + bi;
    uint32_t nv, nf;
    if constexpr (MT) mt_count_chunk<false>(bits, g, chunk, nf_s, celloff, recs, nrecs, nv, nf);
    else mc_count_chunk<false>(bits, g, chunk, nf_s, recs, nrecs, nv, nf);
    if (lane == 0) woff[2 * chunk] = nv, woff[2 * chunk + 1] = nf;
  }
}

constexpr int SC_THREADS = 256, SC_PER = 4;  // scan block: 1024 (vertex, face) pairs
// In-place exclusive scan of the nchunks pairs in woff: block-local scan (shuffles), decoupled look-back across the
// scan blocks (ticketed), totals from the last block.
__global__ void __launch_bounds__(SC_THREADS)
mc_scan_chunks_kernel(unsigned long long* __restrict__ woff, long long nchunks, unsigned long long* status, unsigned int* ticket,
                      long long nsb, long long* totals_a, long long* totals_b, unsigned int* ride, int nride) {
  __shared__ unsigned long long wsum_v[SC_THREADS / 32], wsum_f[SC_THREADS / 32];
  __shared__ unsigned long long base_s[2];
  __shared__ unsigned sb;
  // the queues of the counting warps (cur_bi, next_x[]) are empty again for the next step: everything that reads them
  // ran before this kernel.  ride[1] (statistics: blocks counted inside classify) moves to ride[2] first.
  if (ride && blockIdx.x == 0) {
    if (threadIdx.x == 0) ride[2] = ride[1];
    __syncthreads();
    for (int i = threadIdx.x; i < nride; i += SC_THREADS)
      if (i != 2) ride[i] = 0u;
  }
  if (threadIdx.x == 0) sb = atomicAdd(ticket, 1u);
  __syncthreads();
  const long long b = sb;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i0 = (b * SC_THREADS + threadIdx.x) * SC_PER;
  unsigned long long v[SC_PER], f[SC_PER], tv = 0, tf = 0;
#pragma unroll
  for (int k = 0; k < SC_PER; ++k) {
    const bool in = i0 + k < nchunks;
    v[k] = in ? woff[2 * (i0 + k)] : 0ull, f[k] = in ? woff[2 * (i0 + k) + 1] : 0ull;
    tv += v[k], tf += f[k];
  }
  unsigned long long iv = tv, jf = tf;  // inclusive over the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long a = __shfl_up_sync(0xffffffffu, iv, o), c = __shfl_up_sync(0xffffffffu, jf, o);
    if (lane >= o) iv += a, jf += c;
  }
  if (lane == 31) wsum_v[w] = iv, wsum_f[w] = jf;
  __syncthreads();
  unsigned long long wbv = 0, wbf = 0, av = 0, af = 0;
#pragma unroll
  for (int k = 0; k < SC_THREADS / 32; ++k) {
    if (k < w) wbv += wsum_v[k], wbf += wsum_f[k];
    av += wsum_v[k], af += wsum_f[k];
  }
  if (w == 0) {
    unsigned long long ev, ef;
    lookback(status, b, av, af, ev, ef);
    if (lane == 0) {
      base_s[0] = ev, base_s[1] = ef;
      if (b == nsb - 1) {
        totals_a[0] = (long long)(ev + av), totals_a[1] = (long long)(ef + af);
        if (totals_b) totals_b[0] = (long long)(ev + av), totals_b[1] = (long long)(ef + af);
      }
    }
  }
  __syncthreads();
  unsigned long long pv = base_s[0] + wbv + (iv - tv), pf = base_s[1] + wbf + (jf - tf);
#pragma unroll
  for (int k = 0; k < SC_PER; ++k) {
    if (i0 + k < nchunks) woff[2 * (i0 + k)] = pv, woff[2 * (i0 + k) + 1] = pf;
    pv += v[k], pf += f[k];
  }
}


// MT: scans the raw block totals into the INCLUSIVE prefixes mt_generate_kernel reads (`status`, FLAG_INC | value per
// word, one pair per counting block); `chain` is the look-back state of the scan blocks themselves.  The last scan
// block writes the totals, minus the ghost row of an MT slab (its prefix ends at block blocks_per_row - 1).
__global__ void __launch_bounds__(SC_THREADS)
mt_scan_blocks_kernel(const unsigned long long* __restrict__ raw, long long nblocks, unsigned long long* status,
                      unsigned long long* chain, unsigned int* ticket, long long nsb, long long ghost_block,
                      unsigned long long* ghost_words, long long* totals_a, long long* totals_b, unsigned int* ride, int nride) {
  __shared__ unsigned long long wsum_v[SC_THREADS / 32], wsum_f[SC_THREADS / 32];
  __shared__ unsigned long long base_s[2];
  __shared__ unsigned sb;
  if (ride && blockIdx.x == 0) {  // the counting warps' queues are empty again for the next step (as in mc_scan_chunks_kernel)
    if (threadIdx.x == 0) ride[2] = ride[1];
    __syncthreads();
    for (int i = threadIdx.x; i < nride; i += SC_THREADS)
      if (i != 2) ride[i] = 0u;
  }
  if (threadIdx.x == 0) sb = atomicAdd(ticket, 1u);
  __syncthreads();
  const long long b = sb;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i0 = (b * SC_THREADS + threadIdx.x) * SC_PER;
  unsigned long long v[SC_PER], f[SC_PER], tv = 0, tf = 0;
#pragma unroll
  for (int k = 0; k < SC_PER; ++k) {
    const bool in = i0 + k < nblocks;
    v[k] = in ? raw[2 * (i0 + k)] : 0ull, f[k] = in ? raw[2 * (i0 + k) + 1] : 0ull;
    tv += v[k], tf += f[k];
  }
  unsigned long long iv = tv, jf = tf;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long a = __shfl_up_sync(0xffffffffu, iv, o), c = __shfl_up_sync(0xffffffffu, jf, o);
    if (lane >= o) iv += a, jf += c;
  }
  if (lane == 31) wsum_v[w] = iv, wsum_f[w] = jf;
  __syncthreads();
  unsigned long long wbv = 0, wbf = 0, av = 0, af = 0;
#pragma unroll
  for (int k = 0; k < SC_THREADS / 32; ++k) {
    if (k < w) wbv += wsum_v[k], wbf += wsum_f[k];
    av += wsum_v[k], af += wsum_f[k];
  }
  if (w == 0) {
    unsigned long long ev, ef;
    lookback(chain, b, av, af, ev, ef);
    if (lane == 0) base_s[0] = ev, base_s[1] = ef;
  }
  __syncthreads();
  unsigned long long pv = base_s[0] + wbv + (iv - tv), pf = base_s[1] + wbf + (jf - tf);
#pragma unroll
  for (int k = 0; k < SC_PER; ++k) {
    pv += v[k], pf += f[k];
    if (i0 + k < nblocks) {
      st_relaxed(status + 2 * (i0 + k), FLAG_INC | pv);
      st_relaxed(status + 2 * (i0 + k) + 1, FLAG_INC | pf);
      if (i0 + k == ghost_block) {  // (ghost_words are part of the scan state the classify kernel cleared)
        st_relaxed(ghost_words, FLAG_INC | pv);
        st_relaxed(ghost_words + 1, FLAG_INC | pf);
      }
    }
  }
  if (b == nsb - 1 && threadIdx.x == 0) {
    unsigned long long gv = 0, gf = 0;
    if (ghost_block >= 0) {  // written by a scan block with a lower (or this) ticket: it is running or done
      unsigned long long sv, sf;
      do {
        sv = ld_relaxed(ghost_words), sf = ld_relaxed(ghost_words + 1);
      } while ((sv >> 62) != 2 || (sf >> 62) != 2);
      gv = sv & VAL_MASK, gf = sf & VAL_MASK;
    }
    totals_a[0] = (long long)(base_s[0] + av - gv), totals_a[1] = (long long)(base_s[1] + af - gf);
    if (totals_b) totals_b[0] = totals_a[0], totals_b[1] = totals_a[1];
  }
}

}  // namespace iso
