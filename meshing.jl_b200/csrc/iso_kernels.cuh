// iso_kernels.cuh -- sm_100a kernels of the isosurface hot path (see DESIGN.md for the data layout).
//
// Pipeline (replaces the sequential sweeps of src/marching_cubes.jl:40-62 and
// src/marching_tetrahedra.jl:144-160 of the reference):
//
//   signpack   field (4 B/sample, read ONCE, 128-bit coalesced streaming loads)
//              -> sign bit-field, 1 bit/sample = [sample < iso], packed along z:
//                 bits[(x*ny + y)*W + zw] bit k  <->  sample (x, y, 32*zw + k)
//              z is the innermost axis of the reference's scan, so one 32-bit word holds 32 consecutive
//              voxels of the output order and every per-column prefix becomes popcount arithmetic.
//   count      per quad-cell (128 z-consecutive voxels of one (x,y) column) active mask + vertex/face
//              counts from the bit-field only; block aggregate; single-pass decoupled look-back scan
//              over blocks in scan order -> exclusive vertex/face offsets per block.
//   generate   same partition; re-derives the in-block offsets, compacts the active voxels of the
//              block, then runs dense thread-per-vertex / thread-per-face emission so that consecutive
//              threads write consecutive output elements (the only place the field is touched again:
//              2 samples per vertex, gathered through L1/L2).
//
// All floating-point arithmetic that reaches the output uses explicit round-to-nearest intrinsics
// (no FMA contraction), mirroring the reference operation by operation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ISO_TABLE_QUAL static __device__
#include "iso_tables.h"

namespace iso {

// ------------------------------------------------------------------------------------------------------
// problem geometry and the block/thread -> voxel mapping shared by count / generate
//
// Unit of work: the quad-cell = 4 consecutive z-words = 128 z-consecutive voxels of one voxel column (x,y).
// The quad-cells of one voxel x-row, ordered (y, z), are in the reference's scan order; thread t of a block
// takes quad-cell (first of the block) + t, so a block owns one contiguous range of the output order,
// thread order == scan order, and the 16-byte loads of consecutive lanes are contiguous in memory.
struct Grid {
  int nx, ny, nz;        // samples
  int W;                 // words per sample column (multiple of 4, >= ceil(nz/32))
  int Wq;                // W / 4 : quad-cells per column
  int quads_per_row;     // (ny-1) * Wq : quad-cells of one voxel x-row, in scan order (y, z)
  int blocks_per_row;    // ceil(quads_per_row / CB_THREADS)
  long long row_words;   // ny * W : words between sample column (x,y) and (x+1,y)
  long long ldx;         // field leading dimension (elements)
  long long plane;       // ldx * ny
  int xoff;              // global index of the slab's first sample plane (x-slab sharding; 0 otherwise)
  int ghost;             // MT sharding: voxel row 0 belongs to the previous slab (counted, not emitted)
  int rec_cap;           // active-voxel records kept per generate block (REC_CAP_MIN .. REC_CAP_MAX, set by the host)
  // exact division by Wq and blocks_per_row without the ~45-instruction integer divide (every thread of every
  // count/generate block maps itself to (x, y, zq) with them): q = (umulhi(n, M) + n) >> s  (Granlund-Montgomery)
  unsigned wq_mul, wq_sh, bpr_mul, bpr_sh;
};

// n / d for the divisor behind (mul, sh); exact for n < 2^31 (the sum cannot wrap there: umulhi(n, M) < n)
__host__ __device__ __forceinline__ unsigned fast_div(unsigned n, unsigned mul, unsigned sh) {
#ifdef __CUDA_ARCH__
  return (__umulhi(n, mul) + n) >> sh;
#else
  return (unsigned)((((unsigned long long)n * mul) >> 32) + n) >> sh;
#endif
}
inline void fast_div_setup(unsigned d, unsigned& mul, unsigned& sh) {
  sh = 0;
  while ((1ull << sh) < d) ++sh;  // ceil(log2 d)
  mul = (unsigned)((((1ull << sh) - d) << 32) / d + 1);
}

#ifndef ISO_CB_THREADS
#define ISO_CB_THREADS 128
#endif
#ifndef ISO_GEN_MINB
#define ISO_GEN_MINB (1024 / ISO_CB_THREADS)
#endif
constexpr int CB_THREADS = ISO_CB_THREADS;  // threads (= quad-cells) per count/generate block

inline void grid_setup(Grid& g, long long nx, long long ny, long long nz, long long ldx) {
  g.nx = (int)nx, g.ny = (int)ny, g.nz = (int)nz;
  g.ldx = ldx, g.plane = ldx * ny;
  const int words = (int)((nz + 31) / 32);
  g.W = (words + 3) / 4 * 4;
  if (g.W == 0) g.W = 4;
  g.Wq = g.W / 4;
  g.row_words = ny * g.W;
  g.quads_per_row = (int)((ny > 0 ? ny - 1 : 0) * g.Wq);
  g.blocks_per_row = (g.quads_per_row + CB_THREADS - 1) / CB_THREADS;
  fast_div_setup((unsigned)g.Wq, g.wq_mul, g.wq_sh);
  fast_div_setup((unsigned)(g.blocks_per_row > 0 ? g.blocks_per_row : 1), g.bpr_mul, g.bpr_sh);
}

// what one thread of block b works on
struct TMap {
  int x, y, zq;  // quad-cell zq of voxel column (x, y)
  bool live;
};

__device__ __forceinline__ TMap thread_map(const Grid& g, unsigned b) {
  TMap m;
  m.x = (int)fast_div(b, g.bpr_mul, g.bpr_sh);
  const int qr = (int)(b - (unsigned)m.x * (unsigned)g.blocks_per_row) * CB_THREADS + (int)threadIdx.x;
  m.y = (int)fast_div((unsigned)qr, g.wq_mul, g.wq_sh);
  m.zq = qr - m.y * g.Wq;
  m.live = qr < g.quads_per_row;
  return m;
}

// the same mapping with plain integer divides (the MT kernels are register-bound: the extra multiplier operands
// cost them spills -- measured 0.612 -> 0.623 ms -- so they keep this form)
__device__ __forceinline__ TMap thread_map_div(const Grid& g, unsigned b) {
  TMap m;
  m.x = (int)(b / (unsigned)g.blocks_per_row);
  const int qr = (int)(b - (unsigned)m.x * (unsigned)g.blocks_per_row) * CB_THREADS + (int)threadIdx.x;
  m.y = qr / g.Wq;
  m.zq = qr - m.y * g.Wq;
  m.live = qr < g.quads_per_row;
  return m;
}

// ------------------------------------------------------------------------------------------------------
// (1) signpack
#ifndef ISO_SP_WARPS
#define ISO_SP_WARPS 4
#endif
constexpr int SP_WARPS = ISO_SP_WARPS;
constexpr int SP_XSEG = 128;   // samples of one row handled by a warp (one float4 per lane)
#ifndef ISO_SP_ZW
#define ISO_SP_ZW 16
#endif
constexpr int SP_ZW = ISO_SP_ZW;  // z-words per warp task (16 words = 64 contiguous bytes per column; 8 measured 6% slower: 32-byte scattered writes)

__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 v;
#ifndef ISO_SP_L2HINT
#define ISO_SP_L2HINT "L2::128B"
#endif
  asm volatile("ld.global.nc.L1::no_allocate." ISO_SP_L2HINT ".v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream_f1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// VEC: rows are 16-byte aligned (ldx % 4 == 0, base aligned): lane owns columns 4*lane..4*lane+3.
// !VEC: scalar loads, lane owns columns lane, lane+32, lane+64, lane+96 of the segment.
// T = field element type: Float32 (both paths) or Float64 (scalar path only; the compare is then exact in Float64).
template <bool VEC, typename T = float>
__global__ void __launch_bounds__(SP_WARPS * 32)
signpack_kernel(const T* __restrict__ sdf, uint32_t* __restrict__ bits, int nx, int ny, int nz, long long ldx,
                int W, T thresh, int nxseg, long long ntasks, unsigned long long* __restrict__ clear, int nclear) {
  __shared__ __align__(16) uint32_t stage[SP_WARPS][SP_ZW][SP_XSEG];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // block 0 resets the scan state (ticket + look-back chain) of the count/scan kernels that follow in the stream
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < nclear; i += blockDim.x) clear[i] = 0ull;
  const long long task = (long long)blockIdx.x * SP_WARPS + wib;
  if (task >= ntasks) return;
  const int xseg = (int)(task % nxseg);
  const long long t2 = task / nxseg;
  const int y = (int)(t2 % ny);
  const int zc = (int)(t2 / ny);
  const long long plane = ldx * ny;
  const int xbase = xseg * SP_XSEG;
  const int xs = VEC ? xbase + lane * 4 : xbase + lane;
  const T* rowp = sdf + (long long)y * ldx + xs;
  const float qnan = __int_as_float(0x7fc00000);

#pragma unroll 1
  for (int zw = 0; zw < SP_ZW; ++zw) {
    const int zword = zc * SP_ZW + zw;
    const int zb = zword * 32;
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    if (zword < W && zb < nz) {
      const T* p = rowp + (long long)zb * plane;
      const bool full = zb + 32 <= nz;
      if (VEC && sizeof(T) == 4) {
        const bool xin = xs < nx;
#ifndef ISO_SP_U
#define ISO_SP_U 8
#endif
#pragma unroll
        for (int kb = 0; kb < 32; kb += ISO_SP_U) {
          float4 v[ISO_SP_U];
#pragma unroll
          for (int u = 0; u < ISO_SP_U; ++u) {
            const bool ok = xin && (full || zb + kb + u < nz);
            v[u] = ok ? ldg_stream_f4(reinterpret_cast<const float*>(p + (long long)(kb + u) * plane)) : make_float4(qnan, qnan, qnan, qnan);
          }
#pragma unroll
          for (int u = 0; u < ISO_SP_U; ++u) {
            w0 |= (v[u].x < thresh) ? (1u << (kb + u)) : 0u;
            w1 |= (v[u].y < thresh) ? (1u << (kb + u)) : 0u;
            w2 |= (v[u].z < thresh) ? (1u << (kb + u)) : 0u;
            w3 |= (v[u].w < thresh) ? (1u << (kb + u)) : 0u;
          }
        }
      } else {
        const bool in0 = xs < nx, in1 = xs + 32 < nx, in2 = xs + 64 < nx, in3 = xs + 96 < nx;
#pragma unroll
        for (int kb = 0; kb < 32; kb += 4) {
          T v[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool zok = full || zb + kb + u < nz;
            const T* q = p + (long long)(kb + u) * plane;
            v[u][0] = (zok && in0) ? __ldg(q) : (T)qnan;
            v[u][1] = (zok && in1) ? __ldg(q + 32) : (T)qnan;
            v[u][2] = (zok && in2) ? __ldg(q + 64) : (T)qnan;
            v[u][3] = (zok && in3) ? __ldg(q + 96) : (T)qnan;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            w0 |= (v[u][0] < thresh) ? (1u << (kb + u)) : 0u;
            w1 |= (v[u][1] < thresh) ? (1u << (kb + u)) : 0u;
            w2 |= (v[u][2] < thresh) ? (1u << (kb + u)) : 0u;
            w3 |= (v[u][3] < thresh) ? (1u << (kb + u)) : 0u;
          }
        }
      }
    }
    *reinterpret_cast<uint4*>(&stage[wib][zw][lane * 4]) = make_uint4(w0, w1, w2, w3);
  }
  // each lane reads back only what it wrote: column j of this lane, words 0..7 -> two 16-byte stores
  uint4 r[SP_ZW];
#pragma unroll
  for (int zw = 0; zw < SP_ZW; ++zw) r[zw] = *reinterpret_cast<const uint4*>(&stage[wib][zw][lane * 4]);
  const int wofs = zc * SP_ZW;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = VEC ? xs + j : xs + 32 * j;
    if (x < nx) {
      uint32_t* dst = bits + ((long long)x * ny + y) * W + wofs;
      uint32_t c[SP_ZW];
#pragma unroll
      for (int zw = 0; zw < SP_ZW; ++zw) c[zw] = j == 0 ? r[zw].x : j == 1 ? r[zw].y : j == 2 ? r[zw].z : r[zw].w;
#pragma unroll
      for (int c4 = 0; c4 < SP_ZW; c4 += 4)
        if (wofs + c4 < W) *reinterpret_cast<uint4*>(dst + c4) = make_uint4(c[c4], c[c4 + 1], c[c4 + 2], c[c4 + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// quad-cell: 4 consecutive z-words (128 voxels) of one voxel column (x, y).
// s?? = sign words of the 4 sample columns at z, t?? = the same shifted to z+1.
// Column naming by (dx,dy): 00 = (x,y), 10 = (x+1,y), 11 = (x+1,y+1), 01 = (x,y+1).
struct Quad {
  uint32_t s00[4], s10[4], s11[4], s01[4];
  uint32_t t00[4], t10[4], t11[4], t01[4];
  uint32_t vm[4];  // valid-voxel mask (z < nz-1)
};

__device__ __forceinline__ void shift_up(const uint4 a, uint32_t nxt, uint32_t* s, uint32_t* t) {
  s[0] = a.x, s[1] = a.y, s[2] = a.z, s[3] = a.w;
  t[0] = __funnelshift_r(a.x, a.y, 1);
  t[1] = __funnelshift_r(a.y, a.z, 1);
  t[2] = __funnelshift_r(a.z, a.w, 1);
  t[3] = __funnelshift_r(a.w, nxt, 1);
}

// Loads the quad-cell (x, y, zq) from the bit-field.  Returns false -- without building the shifted words --
// when no sign changes anywhere in its 4x(128+1) samples (cheap test on the raw words): then no voxel of
// the quad-cell is active.
// CG = true: loads that bypass L1 (ld.global.cg).  The counting warps that run INSIDE the classify kernel read
// bit-field rows other CTAs of the same launch have just written; L1 is not coherent and a 128-byte line can hold
// words of a row that is not complete yet, so those reads must come from L2.
template <bool CG>
__device__ __forceinline__ uint4 ld_bits4(const uint32_t* p) {
  return CG ? __ldcg(reinterpret_cast<const uint4*>(p)) : __ldg(reinterpret_cast<const uint4*>(p));
}
template <bool CG>
__device__ __forceinline__ uint32_t ld_bits1(const uint32_t* p) {
  return CG ? __ldcg(p) : __ldg(p);
}
template <bool CG = false>
__device__ __forceinline__ bool load_quad(const uint32_t* __restrict__ bits, const Grid& g, int x, int y, int zq, Quad& q) {
  const uint32_t* c00 = bits + (long long)x * g.row_words + (long long)y * g.W + zq * 4;
  const uint32_t* c10 = c00 + g.row_words;
  const bool more = zq + 1 < g.Wq;
  const uint4 a00 = ld_bits4<CG>(c00);
  const uint4 a01 = ld_bits4<CG>(c00 + g.W);
  const uint4 a10 = ld_bits4<CG>(c10);
  const uint4 a11 = ld_bits4<CG>(c10 + g.W);
  const uint32_t n00 = more ? ld_bits1<CG>(c00 + 4) : 0u, n01 = more ? ld_bits1<CG>(c00 + g.W + 4) : 0u;
  const uint32_t n10 = more ? ld_bits1<CG>(c10 + 4) : 0u, n11 = more ? ld_bits1<CG>(c10 + g.W + 4) : 0u;
  // bits of samples beyond nz are 0, so an all-ones run that reaches the padding reads as "mixed" here;
  // that only costs the slow path, the exact valid-mask is applied below.
  const uint32_t nany = (n00 | n01 | n10 | n11) & 1u;
  const uint32_t nall = more ? ((n00 & n01 & n10 & n11 & 1u) ? 0xffffffffu : 0u) : 0xffffffffu;
  const uint32_t any = a00.x | a00.y | a00.z | a00.w | a01.x | a01.y | a01.z | a01.w | a10.x | a10.y | a10.z | a10.w |
                       a11.x | a11.y | a11.z | a11.w | nany;
  const uint32_t all = a00.x & a00.y & a00.z & a00.w & a01.x & a01.y & a01.z & a01.w & a10.x & a10.y & a10.z & a10.w &
                       a11.x & a11.y & a11.z & a11.w & nall;
  if (any == 0u || all == 0xffffffffu) return false;
  shift_up(a00, n00, q.s00, q.t00);
  shift_up(a01, n01, q.s01, q.t01);
  shift_up(a10, n10, q.s10, q.t10);
  shift_up(a11, n11, q.s11, q.t11);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rem = g.nz - 1 - (zq * 4 + i) * 32;  // voxels z .. valid iff z < nz-1
    q.vm[i] = rem >= 32 ? 0xffffffffu : rem <= 0 ? 0u : ((1u << rem) - 1u);
  }
  return true;
}

__device__ __forceinline__ uint32_t active_mask(const Quad& q, int i) {
  const uint32_t any = q.s00[i] | q.s10[i] | q.s11[i] | q.s01[i] | q.t00[i] | q.t10[i] | q.t11[i] | q.t01[i];
  const uint32_t all = q.s00[i] & q.s10[i] & q.s11[i] & q.s01[i] & q.t00[i] & q.t10[i] & q.t11[i] & q.t01[i];
  return any & ~all & q.vm[i];
}

// Case index of voxel bit k of cell i.  MC corner order (src/marching_cubes.jl:42-49):
// 1 (0,0,0) 2 (1,0,0) 3 (1,1,0) 4 (0,1,0) then z+1.  MT (src/marching_tetrahedra.jl:146-153):
// 1 (0,0,0) 2 (0,1,0) 3 (1,1,0) 4 (1,0,0) then z+1.
template <int ALGO>
__device__ __forceinline__ uint32_t case_of(const Quad& q, int i, int k) {
  const uint32_t b00 = (q.s00[i] >> k) & 1u, b10 = (q.s10[i] >> k) & 1u, b11 = (q.s11[i] >> k) & 1u, b01 = (q.s01[i] >> k) & 1u;
  const uint32_t u00 = (q.t00[i] >> k) & 1u, u10 = (q.t10[i] >> k) & 1u, u11 = (q.t11[i] >> k) & 1u, u01 = (q.t01[i] >> k) & 1u;
  if (ALGO == 0) return b00 | b10 << 1 | b11 << 2 | b01 << 3 | u00 << 4 | u10 << 5 | u11 << 6 | u01 << 7;
  return b00 | b01 << 1 | b11 << 2 | b10 << 3 | u00 << 4 | u01 << 5 | u11 << 6 | u10 << 7;
}

// MC: vertices of a voxel = its sign-changing cube edges (popcount(edge_table[c]) == nverts, App. B)
// -> the per-cell vertex total is 12 masked popcounts.
__device__ __forceinline__ uint32_t mc_nverts_masked(const Quad& q, int i, uint32_t mask) {
  uint32_t n = __popc((q.s00[i] ^ q.s10[i]) & mask) + __popc((q.s10[i] ^ q.s11[i]) & mask) +
               __popc((q.s11[i] ^ q.s01[i]) & mask) + __popc((q.s01[i] ^ q.s00[i]) & mask);
  n += __popc((q.t00[i] ^ q.t10[i]) & mask) + __popc((q.t10[i] ^ q.t11[i]) & mask) +
       __popc((q.t11[i] ^ q.t01[i]) & mask) + __popc((q.t01[i] ^ q.t00[i]) & mask);
  n += __popc((q.s00[i] ^ q.t00[i]) & mask) + __popc((q.s10[i] ^ q.t10[i]) & mask) +
       __popc((q.s11[i] ^ q.t11[i]) & mask) + __popc((q.s01[i] ^ q.t01[i]) & mask);
  return n;
}


// Marching Cubes count of one generate block ("chunk": CB_THREADS consecutive quad-cells of one voxel x-row), by ONE
// warp: lane <-> quad-cell, 32 at a time.  vertices = crossed cube edges (12 masked popcounts per word), faces = table
// look-up per active voxel (nf_s: 256-byte table in shared memory).  Returns the chunk's totals in every lane.
// While it is there the count also leaves the block's ACTIVE-VOXEL RECORDS for generate (scan order, 4 bytes each:
// case | voxel-in-quad-cell << 8 | quad-cell-in-block << 15), up to REC_CAP per block, and the number of active voxels
// in nrecs[chunk].  generate then starts from the records instead of re-deriving them from the bit-field (its
// thread mapping, quad-cell loads, active masks, block scan and case extraction were a third of its instructions);
// a block with more than REC_CAP active voxels (dense fields) takes generate's own front end instead.
// Records per generate block: 512 (2 KB: 3 % of the block's 16384 voxels; the 1024^3 gyroid averages 155 per block, and
// the record array is then as large as the bit-field) on big grids, up to 4096 on small ones, where a block of a
// surface-like field holds more active voxels (16384 * 3.8 % = 620 at 256^3) and the array is small anyway.
constexpr int REC_CAP_MIN = 512, REC_CAP_MAX = 4096;
inline int rec_cap_for(long long nblocks) {
  const long long budget = (64ll << 20) / 4;  // records that fit 64 MB
  long long cap = nblocks > 0 ? budget / nblocks : REC_CAP_MAX;
  cap = cap / 256 * 256;
  return (int)(cap < REC_CAP_MIN ? REC_CAP_MIN : cap > REC_CAP_MAX ? REC_CAP_MAX : cap);
}
template <bool CG>
__device__ __forceinline__ void mc_count_chunk(const uint32_t* __restrict__ bits, const Grid& g, long long chunk, const uint8_t* nf_s,
                                               uint32_t* __restrict__ recs, uint32_t* __restrict__ nrecs, uint32_t& nv_out, uint32_t& nf_out) {
  const int lane = threadIdx.x & 31;
  uint32_t nv = 0, nf = 0, base = 0;  // base: active voxels of the block before this group of 32 quad-cells (uniform)
  const int x = (int)(chunk / g.blocks_per_row);
  const int q_lo = (int)(chunk - (long long)x * g.blocks_per_row) * CB_THREADS;
  uint32_t* rec = recs + chunk * g.rec_cap;
  for (int q0 = q_lo; q0 < q_lo + CB_THREADS && q0 < g.quads_per_row; q0 += 32) {
    const int qr = q0 + lane;
    const int y = (int)fast_div((unsigned)qr, g.wq_mul, g.wq_sh), zq = qr - y * g.Wq;
    Quad q;
    const bool act = qr < g.quads_per_row && load_quad<CG>(bits, g, x, y, zq, q);
    uint32_t mm[4] = {0, 0, 0, 0}, na = 0;
    if (act) {
#pragma unroll
      for (int i = 0; i < 4; ++i) mm[i] = active_mask(q, i), na += __popc(mm[i]);
    }
    if (!__any_sync(0xffffffffu, na != 0)) continue;  // (uniform)
    // position of this lane's first record: exclusive scan over the lanes (lane order == scan order)
    uint32_t inc = na;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    uint32_t pos = base + inc - na;
    base += __shfl_sync(0xffffffffu, inc, 31);
    if (na) {
      const uint32_t qtag = (uint32_t)(qr - q_lo) << 15;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t m = mm[i];
        if (m) {
          nv += mc_nverts_masked(q, i, q.vm[i]);
          while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t c = case_of<0>(q, i, k);
            nf += nf_s[c];
            if (pos < (uint32_t)g.rec_cap) rec[pos] = c | ((uint32_t)(i * 32 + k) << 8) | qtag;
            ++pos;
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nv += __shfl_xor_sync(0xffffffffu, nv, o);
    nf += __shfl_xor_sync(0xffffffffu, nf, o);
  }
  if (lane == 0) nrecs[chunk] = base;
  nv_out = nv, nf_out = nf;
}

// ---- block scans (256 threads) ---------------------------------------------------------------------------
// exclusive scan in thread order; s_w: 8 words of shared scratch.  Contains two barriers.
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* s_w, uint32_t& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // s_w may still be read by the previous scan
  if (lane == 31) s_w[w] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < CB_THREADS / 32; ++i) {
    const uint32_t t = s_w[i];
    if (i < w) base += t;
    tot += t;
  }
  total = tot;
  return base + inc - v;
}

// the same over 64-bit values (two packed 32-bit scans in one pass); s_w: CB_THREADS / 32 words
__device__ __forceinline__ unsigned long long block_excl_scan_u64(unsigned long long v, unsigned long long* s_w, unsigned long long& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) s_w[w] = inc;
  __syncthreads();
  unsigned long long base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < CB_THREADS / 32; ++i) {
    const unsigned long long t = s_w[i];
    if (i < w) base += t;
    tot += t;
  }
  total = tot;
  return base + inc - v;
}

// ------------------------------------------------------------------------------------------------------
// decoupled look-back scan state: one 16-byte entry per block, {nverts, nfaces}, each word
// [flag:2 | value:62]; flag 0 = not yet published, 1 = block aggregate, 2 = inclusive prefix.
constexpr unsigned long long FLAG_AGG = 1ull << 62, FLAG_INC = 2ull << 62, VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Executed by warp 0 of block b (b = ticket): publishes the aggregate, looks back over predecessor
// windows of 32 blocks with ballot/shuffle, publishes the inclusive prefix, returns the exclusive prefix.
__device__ __forceinline__ void lookback(unsigned long long* status, long long b, unsigned long long agg_v,
                                         unsigned long long agg_f, unsigned long long& excl_v, unsigned long long& excl_f) {
  const int lane = threadIdx.x & 31;
  if (b == 0) {
    if (lane == 0) {
      st_relaxed(status + 0, FLAG_INC | agg_v);
      st_relaxed(status + 1, FLAG_INC | agg_f);
    }
    excl_v = excl_f = 0;
    return;
  }
  if (lane == 0) {
    st_relaxed(status + 2 * b, FLAG_AGG | agg_v);
    st_relaxed(status + 2 * b + 1, FLAG_AGG | agg_f);
  }
  unsigned long long sum_v = 0, sum_f = 0;
  long long j0 = b - 1;  // nearest predecessor of this window
  while (true) {
    const long long j = j0 - lane;
    unsigned long long sv = FLAG_INC, sf = FLAG_INC;  // lanes before block 0 read as "inclusive 0"
    if (j >= 0) {
      do {
        sv = ld_relaxed(status + 2 * j);
        sf = ld_relaxed(status + 2 * j + 1);
      } while ((sv >> 62) == 0 || (sf >> 62) == 0 || (sv >> 62) != (sf >> 62));
    }
    const unsigned inc = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
    // lanes at or before the first inclusive entry contribute
    const int stop = inc ? __ffs(inc) - 1 : 31;
    unsigned long long cv = lane <= stop ? (sv & VAL_MASK) : 0, cf = lane <= stop ? (sf & VAL_MASK) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cv += __shfl_xor_sync(0xffffffffu, cv, o);
      cf += __shfl_xor_sync(0xffffffffu, cf, o);
    }
    sum_v += cv, sum_f += cf;
    if (inc) break;
    j0 -= 32;
  }
  if (lane == 0) {
    st_relaxed(status + 2 * b, FLAG_INC | (sum_v + agg_v));
    st_relaxed(status + 2 * b + 1, FLAG_INC | (sum_f + agg_f));
  }
  excl_v = sum_v, excl_f = sum_f;
}

// ------------------------------------------------------------------------------------------------------
// Sharded path: exchange of the slabs' totals over NVLink peer memory (instead of a 16-byte NCCL all-gather).
// Every rank owns an exchange buffer int64[2][PEER_MAX][4] that all ranks have mapped (CUDA IPC / symmetric memory);
// slot [epoch & 1][r] of it is written by rank r only: {nverts, nfaces, 0, epoch}, the epoch word last with
// release.sys.  Two parities suffice: rank r can publish epoch e + 2 only after its generate of epoch e + 1, which
// waited for everyone's count of epoch e + 1, which (stream order) follows everyone's gather of epoch e.
constexpr int PEER_MAX = 16;
struct PeerSlots {
  long long* slot[PEER_MAX];
};
__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// One warp.  Lane r stores this rank's totals into rank r's buffer (a P2P store over NVLink for r != rank), then waits
// for rank r's totals of this epoch in the LOCAL buffer; bases[0..1] = exclusive prefix of (nverts, nfaces) over the
// ranks below this one, bases[2..3] = grand totals; optionally all pairs to `all`.  Every rank publishes before it
// waits, so the ranks cannot block each other.  timeout_ns == 0 waits for ever (like an NCCL collective); otherwise
// a peer that never arrives sets *err = 1 -- generate kernels behind this one then emit nothing -- and zero bases.
__global__ void peer_exchange_kernel(PeerSlots ps, int world, int rank, long long epoch, const long long* __restrict__ totals,
                                     long long* __restrict__ bases, long long* __restrict__ all, long long* err,
                                     unsigned long long timeout_ns) {
  const int r = threadIdx.x;
  long long nv = 0, nf = 0;
  bool ok = true;
  if (r < world) {
    long long* s = ps.slot[r] + ((epoch & 1) * PEER_MAX + rank) * 4;
    s[0] = totals[0], s[1] = totals[1], s[2] = 0;
    st_release_sys(s + 3, epoch);
    const long long* m = ps.slot[rank] + ((epoch & 1) * PEER_MAX + r) * 4;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(m + 3) != epoch) {
      __nanosleep(200);
      if (timeout_ns && global_timer_ns() - t0 > timeout_ns) {
        ok = false;
        break;
      }
    }
    if (ok) nv = m[0], nf = m[1];
  }
  const bool all_ok = __all_sync(0xffffffffu, ok);
  if (!all_ok) nv = nf = 0;
  if (r < world && all) all[2 * r] = nv, all[2 * r + 1] = nf;
  long long bv = r < rank ? nv : 0, bf = r < rank ? nf : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bv += __shfl_xor_sync(0xffffffffu, bv, o), bf += __shfl_xor_sync(0xffffffffu, bf, o);
    nv += __shfl_xor_sync(0xffffffffu, nv, o), nf += __shfl_xor_sync(0xffffffffu, nf, o);
  }
  if (r == 0) {
    bases[0] = bv, bases[1] = bf, bases[2] = nv, bases[3] = nf;
    if (!all_ok) *err = 1;
  }
}

// ------------------------------------------------------------------------------------------------------
// coordinates: LinRange(first, last, n)[i] = P((1-t)*a + t*b), t = i/(n-1) in Float64 (Julia Base lerpi,
// SURVEY.md A2).  Stored as doubles (a Float32 value is exact in a double).
__global__ void coords_kernel(double* out, int nx, int ny, int nz, double x0, double x1, double y0, double y1, double z0,
                              double z1, int f32, int x_offset, int nx_global) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx + ny + nz) return;
  int n, k;
  double a, b;
  if (i < nx) n = nx_global, k = i + x_offset, a = x0, b = x1;  // slab of a larger volume: global index
  else if (i < nx + ny) n = ny, k = i - nx, a = y0, b = y1;
  else n = nz, k = i - nx - ny, a = z0, b = z1;
  if (f32) a = (double)(float)a, b = (double)(float)b;
  const int d = n - 1 > 1 ? n - 1 : 1;
  const double t = __ddiv_rn((double)k, (double)d);
  const double v = __dadd_rn(__dmul_rn(__dsub_rn(1.0, t), a), __dmul_rn(t, b));
  out[i] = f32 ? (double)__double2float_rn(v) : v;
}

// ------------------------------------------------------------------------------------------------------
// per-voxel case indices in scan-rank order (parity output)
template <int ALGO>
__global__ void case_kernel(const uint32_t* __restrict__ bits, Grid g, uint8_t* __restrict__ out, long long nvox) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nvox) return;
  const int nzv = g.nz - 1, nyv = g.ny - 1;
  const int z = (int)(r % nzv);
  const long long r2 = r / nzv;
  const int y = (int)(r2 % nyv), x = (int)(r2 / nyv);
  auto bit = [&](int dx, int dy, int dz) -> uint32_t {
    const int zz = z + dz;
    const uint32_t w = __ldg(bits + (long long)(x + dx) * g.row_words + (long long)(y + dy) * g.W + (zz >> 5));
    return (w >> (zz & 31)) & 1u;
  };
  uint32_t c;
  if (ALGO == 0)
    c = bit(0, 0, 0) | bit(1, 0, 0) << 1 | bit(1, 1, 0) << 2 | bit(0, 1, 0) << 3 | bit(0, 0, 1) << 4 | bit(1, 0, 1) << 5 |
        bit(1, 1, 1) << 6 | bit(0, 1, 1) << 7;
  else
    c = bit(0, 0, 0) | bit(0, 1, 0) << 1 | bit(1, 1, 0) << 2 | bit(1, 0, 0) << 3 | bit(0, 0, 1) << 4 | bit(0, 1, 1) << 5 |
        bit(1, 1, 1) << 6 | bit(1, 0, 1) << 7;
  out[r] = (uint8_t)c;
}

// ------------------------------------------------------------------------------------------------------
// (3) generate, Marching Cubes.
// MODE selects the arithmetic types of vertex_interp (src/marching_cubes.jl:100-104):
//   0: iso::Float32, points Float64 (Int or Float64 ranges): mu Float32, position Float64
//   1: iso::Float32, points Float32: everything Float32
//   2: iso::Float64: mu Float64, position Float64 (Float32 points are exact in Float64)
// V = vertex element type (float / double), the reference's float(FT).
struct GenArgs {
  const void* sdf;  // Float32 field (Float64 for the *_f64 instantiations)
  const uint32_t* bits;
  const unsigned long long* status;  // MT: inclusive (vertex, face) prefix of every generate block
  const unsigned long long* woff;    // MC: exclusive (vertex, face) prefix of every generate block
  const double* coords;              // xp | yp | zp
  void* verts;
  long long* faces;
  long long vcap, fcap;
  const long long* vbase_dev;
  long long vbase;
  double iso_d;
  float iso_f;
  float eps_f;
  double eps_d;
  int iso_is_f32, eps_is_f32, p_is_f32;  // typeof(iso), typeof(eps), eltype of the points (ranges)
  int sdf_vec;                           // Float32 field, base 16-byte aligned, ldx % 4 == 0: aligned pair loads
  const uint32_t* recs;          // MC: active-voxel records of every generate block (REC_CAP each), written by the count
  const uint32_t* nrecs;         // MC: active voxels per generate block
  long long key_nx_global;       // KEYS instantiation: samples along x of the whole volume
  long long nblocks;
  const long long* totals_a;     // device totals {nverts, nfaces} of the count
  const long long* abort_flag;   // != 0: a peer exchange ahead of this kernel failed -- emit nothing
};

// Float64 field (MODE 3): everything is Float64 (src/marching_cubes.jl:100-104 with T = Float64).
__device__ __forceinline__ void mc_interp_f64(const GenArgs& a, double va, double vb, const double pa[3], const double pb[3],
                                              double out[3]) {
  const double iso = a.iso_is_f32 ? (double)a.iso_f : a.iso_d;
  const double mu = __ddiv_rn(__dsub_rn(iso, va), __dsub_rn(vb, va));
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const double d = a.p_is_f32 ? (double)__fsub_rn((float)pb[q], (float)pa[q]) : __dsub_rn(pb[q], pa[q]);
    out[q] = __dadd_rn(pa[q], __dmul_rn(mu, d));
  }
}

template <int MODE>
__device__ __forceinline__ void mc_interp(const GenArgs& a, float va, float vb, const double pa[3], const double pb[3],
                                          double out[3]) {
  const float den = __fsub_rn(vb, va);  // valp2 - valp1 in the field type
  if (MODE == 0) {
    const float mu = __fdiv_rn(__fsub_rn(a.iso_f, va), den);
    const double mud = (double)mu;
#pragma unroll
    for (int q = 0; q < 3; ++q) out[q] = __dadd_rn(pa[q], __dmul_rn(mud, __dsub_rn(pb[q], pa[q])));
  } else if (MODE == 1) {
    const float mu = __fdiv_rn(__fsub_rn(a.iso_f, va), den);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float fa = (float)pa[q], fb = (float)pb[q];
      out[q] = (double)__fadd_rn(fa, __fmul_rn(mu, __fsub_rn(fb, fa)));
    }
  } else {
    const double mu = __ddiv_rn(__dsub_rn(a.iso_d, (double)va), (double)den);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      // p2 .- p1 is evaluated in the points' own type before the promotion to Float64
      const double d = a.p_is_f32 ? (double)__fsub_rn((float)pb[q], (float)pa[q]) : __dsub_rn(pb[q], pa[q]);
      out[q] = __dadd_rn(pa[q], __dmul_rn(mu, d));
    }
  }
}
constexpr int GEN_NB = 2 * CB_THREADS;  // active voxels per dense round (two records per thread)
#ifndef ISO_GEN_SPLIT
#define ISO_GEN_SPLIT 1
#endif
// Which two records of a round a thread takes: t and t + 128 (SPLIT) or 2t and 2t + 1.  A typical block has ~155
// records: with pairs only threads 0..77 gather samples while the fourth warp waits at the barrier; split, every warp
// holds a record per lane and the few records beyond 128 fall to the first warp.
constexpr bool GEN_SPLIT = ISO_GEN_SPLIT != 0;
constexpr int GEN_MAXV = GEN_NB * 12;   // vertices of a round (MC: <= 12 per voxel)
constexpr int GEN_MAXF = GEN_NB * 5;    // faces of a round (MC: <= 5 per voxel)

// Pushes the (y,z | case) records of this thread's active voxels whose position in the block's scan order
// falls in [lo, hi).  `a0` = position of the thread's first active voxel, q = its (already loaded) quad-cell.
// A record is one 8-byte store: .x = y | z << 16, .y = case index.
template <int ALGO>
__device__ __forceinline__ void push_records(const Quad& q, const TMap& tm, uint32_t a0, uint32_t lo, uint32_t hi, uint2* rec, int stride) {
  uint32_t idx = a0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t mm = active_mask(q, i);
    if (idx + __popc(mm) <= lo || idx >= hi) {  // whole cell outside the window
      idx += __popc(mm);
      continue;
    }
    while (mm) {
      const int k = __ffs(mm) - 1;
      mm &= mm - 1;
      if (idx >= lo && idx < hi)
        rec[(idx - lo) * stride] = make_uint2((uint32_t)tm.y | ((uint32_t)((tm.zq * 4 + i) * 32 + k) << 16), case_of<ALGO>(q, i, k));
      ++idx;
    }
  }
}

// loads the thread's quad-cell (kept in registers for the first window) and counts its active voxels
__device__ __forceinline__ uint32_t count_active(const uint32_t* __restrict__ bits, const Grid& g, const TMap& tm, Quad& q) {
  uint32_t na = 0;
  if (tm.live && load_quad(bits, g, tm.x, tm.y, tm.zq, q)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) na += __popc(active_mask(q, i));
  }
  return na;
}

// samples (x, x+1) of a row whose element x - (x & 3) is 16-byte aligned; ph = x & 3
__device__ __forceinline__ float2 pair_load(const float* p, int ph) {
  if (ph == 0 || ph == 2) return __ldg(reinterpret_cast<const float2*>(p));
  if (ph == 1) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p - 1));
    return make_float2(v.y, v.z);
  }
  const float2 lo = __ldg(reinterpret_cast<const float2*>(p - 1)), hi = __ldg(reinterpret_cast<const float2*>(p + 1));
  return make_float2(lo.y, hi.x);
}

// The four samples of one z-plane of a voxel: (x,y) (x+1,y) (x,y+1) (x+1,y+1).
template <class T>
struct Plane {
  T a00, a10, a01, a11;
};
template <class T>
__device__ __forceinline__ Plane<T> load_plane(const T* p, long long ldx, int ph, bool vec) {
  Plane<T> r;
  if constexpr (sizeof(T) == 4) {
    if (vec) {
      // Float32 field with 16-byte aligned rows: the (x, x+1) pair of a row is ONE aligned load (two when it
      // straddles a 16-byte boundary); x is uniform over the block, so the switch does not diverge
      const float2 q0 = pair_load(reinterpret_cast<const float*>(p), ph), q1 = pair_load(reinterpret_cast<const float*>(p) + ldx, ph);
      r.a00 = q0.x, r.a10 = q0.y, r.a01 = q1.x, r.a11 = q1.y;
      return r;
    }
  }
  r.a00 = __ldg(p), r.a10 = __ldg(p + 1), r.a01 = __ldg(p + ldx), r.a11 = __ldg(p + ldx + 1);
  return r;
}
template <class T>
__device__ __forceinline__ Plane<T> shfl_up_plane(const Plane<T>& v) {
  Plane<T> r;
  r.a00 = __shfl_up_sync(0xffffffffu, v.a00, 1), r.a10 = __shfl_up_sync(0xffffffffu, v.a10, 1);
  r.a01 = __shfl_up_sync(0xffffffffu, v.a01, 1), r.a11 = __shfl_up_sync(0xffffffffu, v.a11, 1);
  return r;
}

template <int MODE>
struct FieldOf {
  using type = float;
};
template <>
struct FieldOf<3> {
  using type = double;
};

// Three consecutive output scalars of element `gi` (a vertex: 3 V; a face: 3 Int64) as one 2-wide and one 1-wide
// store instead of three: elements are 12 / 24 bytes, so the 2-wide half is the one that is naturally aligned.
// Fewer store requests = fewer L1 line look-ups (the kernel is bound by those).
template <class S>
__device__ __forceinline__ void store3(S* base, long long gi, S v0, S v1, S v2, bool wide_ok) {
  S* o = base + 3 * gi;
  if (wide_ok) {
    struct alignas(2 * sizeof(S)) S2 {
      S a, b;
    };
    // (selects, not branches: even and odd elements of a warp issue the same two store instructions)
    const bool odd = (gi & 1) != 0;
    *reinterpret_cast<S2*>(o + (odd ? 1 : 0)) = S2{odd ? v1 : v0, odd ? v2 : v1};
    o[odd ? 0 : 2] = odd ? v0 : v2;
  } else {
    o[0] = v0, o[1] = v1, o[2] = v2;
  }
}

// generate after the count (two-phase ABI: the caller sizes its arrays in between; the async form runs them back to back).
//
// Shared-memory record of an active voxel (16 bytes, read by the vertex threads with ONE 128-bit load):
//   .x = y | z << 16          .y = local vertex offset | local face offset << 16
//   .z/.w = the voxel's ordered edge list (nibbles 0..7 / 8..11 of ISO_MC_VERTS[case])
// plus its packed face list (ISO_MC_FACES[case]) in recf[].  The per-case tables are read once per voxel (B1b) --
// not once per vertex and once per face: each of those was a 32-address gather into a 2 KB table, i.e. up to 16 L1
// line look-ups per warp instruction, and the kernel is bound by L1 line look-ups.
// KEYS = true (mesh consumers, mesh_consumers.cuh): instead of the vertices, write every vertex's grid-edge key
// (3 * linear index of the edge's lower end node in the WHOLE volume + axis) as Int64 into a.verts; no faces.
template <int MODE, typename V, bool KEYS = false>
__global__ void __launch_bounds__(CB_THREADS, ISO_GEN_MINB)
mc_generate_kernel(GenArgs a, Grid g) {
  __shared__ uint32_t s_w[CB_THREADS / 32];
  __shared__ unsigned long long s_w64[CB_THREADS / 32];
  __shared__ __align__(16) uint4 rec[GEN_NB];
  __shared__ unsigned long long recf[GEN_NB];
  using T = typename FieldOf<MODE>::type;  // Float32, or Float64 for MODE 3
  __shared__ __align__(16) T corner[GEN_NB][8];
  __shared__ uint8_t owner_v[GEN_MAXV], owner_f[GEN_MAXF];

  const int tid = threadIdx.x;
  const unsigned b = blockIdx.x;
  if (a.abort_flag && *a.abort_flag) return;  // the exchange that was to deliver the vertex base failed
  // The count left the block's number of active voxels and (unless there are more than REC_CAP of them: dense fields,
  // which take this kernel's own front end, A + B1a below) their records in scan order.  A block without active voxels
  // leaves on that one word, before touching anything else; on sparse fields (a few shapes in a big volume) that is
  // most blocks.  The block's prefix (woff) is not needed before the first store: its loads overlap the record loads.
  const uint32_t nrec = __ldg(a.nrecs + b);
  if (nrec == 0) return;
  unsigned long long bv = a.woff[2 * (unsigned long long)b], bf = a.woff[2 * (unsigned long long)b + 1];
  const bool from_recs = nrec <= (uint32_t)g.rec_cap;  // (uniform over the block)
  TMap tm;
  tm.x = (int)fast_div(b, g.bpr_mul, g.bpr_sh);
  const uint32_t q_lo = (b - (unsigned)tm.x * (unsigned)g.blocks_per_row) * CB_THREADS;  // first quad-cell of the block in its x-row
  uint32_t tna = 0, my_a0 = 0, blk_na = nrec;
  if (!from_recs) {
    // ---- A: active voxels per thread, exclusive scan (thread order == scan order) ----
    tm = thread_map(g, b);
    Quad q;
    tna = count_active(a.bits, g, tm, q);
    my_a0 = block_excl_scan_u32(tna, s_w, blk_na);
    // ---- B1a: records (position, case) of the first window's voxels, in scan order, from the quad-cell in registers
    // (dead afterwards); later windows reload it
    if (tna && my_a0 < (uint32_t)GEN_NB) push_records<0>(q, tm, my_a0, 0, min((uint32_t)GEN_NB, blk_na), reinterpret_cast<uint2*>(rec), 2);
  }
  if (blk_na == 0) return;  // (cannot happen after the test above; kept as a guard)

  const long long vbase = a.vbase + (a.vbase_dev ? *a.vbase_dev : 0);
  const unsigned yofs = (unsigned)g.nx, zofs = (unsigned)(g.nx + g.ny);  // coords = xp | yp | zp
  const double x0d = __ldg(a.coords + tm.x), x1d = __ldg(a.coords + tm.x + 1);
  V* verts = reinterpret_cast<V*>(a.verts);
  const int x = tm.x;
  const int lane = tid & 31;
  const bool vwide = (reinterpret_cast<uintptr_t>(a.verts) & (2 * sizeof(V) - 1)) == 0;
  const bool fwide = (reinterpret_cast<uintptr_t>(a.faces) & 15) == 0;

  for (uint32_t lo = 0; lo < blk_na; lo += GEN_NB) {
    const uint32_t hi = min(lo + (uint32_t)GEN_NB, blk_na);
    const uint32_t cnt = hi - lo;
    if (!from_recs) {
      if (lo > 0 && tna && my_a0 < hi && my_a0 + tna > lo) {
        Quad q2;
        load_quad(a.bits, g, tm.x, tm.y, tm.zq, q2);
        push_records<0>(q2, tm, my_a0, lo, hi, reinterpret_cast<uint2*>(rec), 2);
      }
      __syncthreads();
    }
    // ---- B1b: two records per thread: tables -> counts -> scan -> owner maps; gather the corner samples.
    // z-adjacent records (same column, z + 1: adjacent in scan order) share a sample plane: the lower plane of the
    // second is the upper plane of the first -- taken from registers (in-thread pair) or from the lane below
    // (shuffle) instead of being gathered again.
    const uint32_t r0 = GEN_SPLIT ? (uint32_t)tid : 2u * tid, r1 = GEN_SPLIT ? (uint32_t)tid + CB_THREADS : r0 + 1;
    uint32_t nvf0 = 0, nvf1 = 0;  // vertices | faces << 16 of the two records
    uint32_t yz0 = 0xffffffffu, yz1 = 0xfffffffeu;
    uint2 rc0 = make_uint2(0, 0), rc1 = make_uint2(0, 0);  // (y | z << 16, case) of the two records
    if (from_recs) {
      // the count's records: case | voxel-in-quad-cell << 8 | quad-cell-in-block << 15
      const uint32_t* rp = a.recs + (unsigned long long)b * (unsigned)g.rec_cap + lo;
      uint2 w = make_uint2(0, 0);
      if constexpr (GEN_SPLIT) {
        if (r0 < cnt) w.x = __ldg(rp + r0);
        if (r1 < cnt) w.y = __ldg(rp + r1);
      } else {
        if (r0 < cnt) w = __ldg(reinterpret_cast<const uint2*>(rp + r0));  // (one aligned 8-byte load per thread)
      }
      auto decode = [&](uint32_t wd) {
        const uint32_t qr = q_lo + (wd >> 15), y = fast_div(qr, g.wq_mul, g.wq_sh), zq = qr - y * (uint32_t)g.Wq;
        return make_uint2(y | ((zq * 128u + ((wd >> 8) & 127u)) << 16), wd & 0xffu);
      };
      rc0 = decode(w.x), rc1 = decode(w.y);
      if (r0 < cnt) rec[r0].x = rc0.x;
      if (r1 < cnt) rec[r1].x = rc1.x;
    } else {
      if (r0 < cnt) rc0 = *reinterpret_cast<const uint2*>(&rec[r0]);
      if (r1 < cnt) rc1 = *reinterpret_cast<const uint2*>(&rec[r1]);
    }
    if (r0 < cnt) {
      const uint2 rc = rc0;
      yz0 = rc.x;
      const unsigned long long tv = __ldg(&ISO_MC_VERTS[rc.y]);
      recf[r0] = __ldg(&ISO_MC_FACES[rc.y]);
      *reinterpret_cast<uint2*>(&rec[r0].z) = make_uint2((uint32_t)tv, (uint32_t)(tv >> 32) & 0xffffu);
      nvf0 = (uint32_t)((tv >> 48) & 15) | ((uint32_t)((tv >> 52) & 7) << 16);
    }
    if (r1 < cnt) {
      const uint2 rc = rc1;
      yz1 = rc.x;
      const unsigned long long tv = __ldg(&ISO_MC_VERTS[rc.y]);
      recf[r1] = __ldg(&ISO_MC_FACES[rc.y]);
      *reinterpret_cast<uint2*>(&rec[r1].z) = make_uint2((uint32_t)tv, (uint32_t)(tv >> 32) & 0xffffu);
      nvf1 = (uint32_t)((tv >> 48) & 15) | ((uint32_t)((tv >> 52) & 7) << 16);
    }
    if constexpr (!KEYS) {
      const bool vec = sizeof(T) == 4 && a.sdf_vec;
      const int ph = x & 3;
      const T* fld = reinterpret_cast<const T*>(a.sdf) + x;
      // pairs: r0 continues the z-run of the lane below's r1, r1 continues r0's; split: each continues the lane below's
      const uint32_t pyz1 = __shfl_up_sync(0xffffffffu, yz1, 1);
      const uint32_t pyz0 = GEN_SPLIT ? __shfl_up_sync(0xffffffffu, yz0, 1) : 0u;
      const bool adj0 = lane > 0 && r0 < cnt && yz0 == (GEN_SPLIT ? pyz0 : pyz1) + 0x10000u;
      const bool adj1 = r1 < cnt && (GEN_SPLIT ? (lane > 0 && yz1 == pyz1 + 0x10000u) : yz1 == yz0 + 0x10000u);
      Plane<T> L0{}, U0{}, L1{}, U1{};
      if (r0 < cnt) {
        const T* p = fld + g.ldx * (long long)(yz0 & 0xffffu) + g.plane * (long long)(yz0 >> 16);
        if (!adj0) L0 = load_plane<T>(p, g.ldx, ph, vec);
        U0 = load_plane<T>(p + g.plane, g.ldx, ph, vec);
      }
      if (r1 < cnt) {
        const T* p = fld + g.ldx * (long long)(yz1 & 0xffffu) + g.plane * (long long)(yz1 >> 16);
        if (!adj1) L1 = load_plane<T>(p, g.ldx, ph, vec);
        U1 = load_plane<T>(p + g.plane, g.ldx, ph, vec);
      }
      if constexpr (GEN_SPLIT) {
        const Plane<T> below0 = shfl_up_plane<T>(U0);  // (all lanes take part)
        if (adj0) L0 = below0;
        if ((uint32_t)(tid & ~31) + CB_THREADS < cnt) {  // (uniform over the warp: some lane has a second record)
          const Plane<T> below1 = shfl_up_plane<T>(U1);
          if (adj1) L1 = below1;
        }
      } else {
        const Plane<T> below = shfl_up_plane<T>(U1);  // (all lanes take part)
        if (adj0) L0 = below;
        if (adj1) L1 = U0;
      }
      // MC corner order (src/marching_cubes.jl:42-49): 0 (0,0,0) 1 (1,0,0) 2 (1,1,0) 3 (0,1,0), then the same at z + 1.
      // A thread's two records are 4 chunks of 16 bytes; chunk c of thread t lives at chunk c ^ ((t >> 1) & 3) of the
      // thread's 64 bytes, so that the 8 lanes of a quarter-warp hit 8 different bank groups (unswizzled: 4-way conflicts).
      {
        if constexpr (sizeof(T) == 4 && GEN_SPLIT) {
          // a record is 2 chunks of 16 bytes; chunk c of record r lives at chunk c ^ ((r >> 2) & 1) of its 32 bytes:
          // the 8 lanes of a quarter-warp (rows 32 bytes apart) then hit 8 different bank groups
          const int sw = (tid >> 2) & 1;  // (the same for r0 and r1 = r0 + 128)
          if (r0 < cnt) {
            float4* crow = reinterpret_cast<float4*>(&corner[r0][0]);
            crow[0 ^ sw] = make_float4(L0.a00, L0.a10, L0.a11, L0.a01);
            crow[1 ^ sw] = make_float4(U0.a00, U0.a10, U0.a11, U0.a01);
          }
          if (r1 < cnt) {
            float4* crow = reinterpret_cast<float4*>(&corner[r1][0]);
            crow[0 ^ sw] = make_float4(L1.a00, L1.a10, L1.a11, L1.a01);
            crow[1 ^ sw] = make_float4(U1.a00, U1.a10, U1.a11, U1.a01);
          }
        } else if constexpr (sizeof(T) == 4) {
          float4* crow = reinterpret_cast<float4*>(&corner[r0][0]);  // (T = double: 8 chunks of 16 bytes, same rule on pairs)
          const int sw = (tid >> 1) & 3;
          if (r0 < cnt) {
            crow[0 ^ sw] = make_float4(L0.a00, L0.a10, L0.a11, L0.a01);
            crow[1 ^ sw] = make_float4(U0.a00, U0.a10, U0.a11, U0.a01);
          }
          if (r1 < cnt) {
            crow[2 ^ sw] = make_float4(L1.a00, L1.a10, L1.a11, L1.a01);
            crow[3 ^ sw] = make_float4(U1.a00, U1.a10, U1.a11, U1.a01);
          }
        } else {
          if (r0 < cnt) {
            corner[r0][0] = L0.a00, corner[r0][1] = L0.a10, corner[r0][2] = L0.a11, corner[r0][3] = L0.a01;
            corner[r0][4] = U0.a00, corner[r0][5] = U0.a10, corner[r0][6] = U0.a11, corner[r0][7] = U0.a01;
          }
          if (r1 < cnt) {
            corner[r1][0] = L1.a00, corner[r1][1] = L1.a10, corner[r1][2] = L1.a11, corner[r1][3] = L1.a01;
            corner[r1][4] = U1.a00, corner[r1][5] = U1.a10, corner[r1][6] = U1.a11, corner[r1][7] = U1.a01;
          }
        }
      }
    }
    // local offsets in record order: pairs -- one scan of the pair sums; split -- records 0..127 then 128..255, both
    // scans in one 64-bit pass
    uint32_t wtot, ex0, ex1;
    if constexpr (GEN_SPLIT) {
      unsigned long long tot64;
      const unsigned long long ex64 = block_excl_scan_u64((unsigned long long)nvf0 | ((unsigned long long)nvf1 << 32), s_w64, tot64);
      ex0 = (uint32_t)ex64, ex1 = (uint32_t)tot64 + (uint32_t)(ex64 >> 32);
      wtot = (uint32_t)tot64 + (uint32_t)(tot64 >> 32);
    } else {
      ex0 = block_excl_scan_u32(nvf0 + nvf1, s_w, wtot);
      ex1 = ex0 + nvf0;
    }
    {
      const uint32_t nv0 = nvf0 & 0xffffu, nf0 = nvf0 >> 16, nv1 = nvf1 & 0xffffu, nf1 = nvf1 >> 16;
      if (r0 < cnt) {
        rec[r0].y = ex0;  // local vertex offset | local face offset << 16
        const uint32_t v0 = ex0 & 0xffffu, f0 = ex0 >> 16;
        for (uint32_t i = 0; i < nv0; ++i) owner_v[v0 + i] = (uint8_t)r0;
        for (uint32_t i = 0; i < nf0; ++i) owner_f[f0 + i] = (uint8_t)r0;
      }
      if (r1 < cnt) {
        rec[r1].y = ex1;
        const uint32_t v1 = ex1 & 0xffffu, f1 = ex1 >> 16;
        for (uint32_t i = 0; i < nv1; ++i) owner_v[v1 + i] = (uint8_t)r1;
        for (uint32_t i = 0; i < nf1; ++i) owner_f[f1 + i] = (uint8_t)r1;
      }
    }
    const uint32_t nvr = wtot & 0xffffu, nfr = wtot >> 16;
    __syncthreads();

    // capacity guard hoisted: the whole window fits in the output buffers in all but the overflow case
    const bool vfits = (long long)bv + nvr <= a.vcap, ffits = (long long)bf + nfr <= a.fcap;
    // ---- B2: thread per vertex (vertex_interp, src/marching_cubes.jl:100-104) ----
    for (uint32_t k = tid; k < nvr; k += CB_THREADS) {
      const uint32_t s = owner_v[k];  // record index 0..255
      const uint4 r = rec[s];
      const uint32_t which = k - (r.y & 0xffffu);
      const uint32_t e = ((which < 8 ? r.z >> (4 * which) : r.w >> (4 * which - 32)) & 15u);
      // _mc_edge_list (src/lut/mc.jl:606-608), 0-based: edge e runs from corner e & 7 to corner
      // (e & 4) | ((e + 1) & 3) for the 8 in-plane edges and to corner e - 4 for the 4 vertical ones
      const uint32_t ca = e & 7u, cb = e < 8u ? ((e & 4u) | ((e + 1u) & 3u)) : e - 4u;
      // MC corner offsets (dx | dy<<1 | dz<<2) for corners 0..7: 0,1,3,2,4,5,7,6
      const uint32_t oa = (0x67542310u >> (4 * ca)) & 7u, ob = (0x67542310u >> (4 * cb)) & 7u;
      if constexpr (KEYS) {
        // the edge's end nodes differ in exactly one axis: oa & ob is the lower one, oa ^ ob the axis bit
        const uint32_t lo_n = oa & ob, ax = oa ^ ob;
        const long long nxg = a.key_nx_global, node = (long long)(x + g.xoff + (lo_n & 1u)) +
                              nxg * ((long long)((r.x & 0xffffu) + ((lo_n >> 1) & 1u)) + (long long)g.ny * (long long)((r.x >> 16) + (lo_n >> 2)));
        if ((long long)bv + k < a.vcap) reinterpret_cast<long long*>(a.verts)[(long long)bv + k] = 3 * node + (ax == 1u ? 0 : ax == 2u ? 1 : 2);
        continue;
      }
      T va, vb;
      if constexpr (sizeof(T) == 4) {  // (chunk swizzle of the corner store, see B1b)
        if constexpr (GEN_SPLIT) {
          const T* cp = &corner[s][0];
          const uint32_t sw = (s >> 2) & 1u;
          va = cp[(((ca >> 2) ^ sw) << 2) + (ca & 3u)], vb = cp[(((cb >> 2) ^ sw) << 2) + (cb & 3u)];
        } else {
          const T* cp = &corner[s & ~1u][0];
          const uint32_t sw = (s >> 2) & 3u, hb = 2u * (s & 1u);
          va = cp[(((hb + (ca >> 2)) ^ sw) << 2) + (ca & 3u)], vb = cp[(((hb + (cb >> 2)) ^ sw) << 2) + (cb & 3u)];
        }
      } else {
        va = corner[s][ca], vb = corner[s][cb];
      }
      const unsigned vy = yofs + (r.x & 0xffffu), vz = zofs + (r.x >> 16);
      double pa[3], pb[3];
      pa[0] = (oa & 1u) ? x1d : x0d;
      pa[1] = __ldg(a.coords + (vy + ((oa >> 1) & 1u)));
      pa[2] = __ldg(a.coords + (vz + (oa >> 2)));
      const uint32_t df = oa ^ ob;  // the one axis along which the edge runs
      pb[0] = (ob & 1u) ? x1d : x0d;
      pb[1] = pa[1], pb[2] = pa[2];
      if (df & 2u) pb[1] = __ldg(a.coords + (vy + ((ob >> 1) & 1u)));
      if (df & 4u) pb[2] = __ldg(a.coords + (vz + (ob >> 2)));
      double p[3];
      if constexpr (MODE == 3) mc_interp_f64(a, va, vb, pa, pb, p);
      else mc_interp<MODE>(a, va, vb, pa, pb, p);
      if (vfits || (long long)bv + k < a.vcap) store3<V>(verts, (long long)bv + k, (V)p[0], (V)p[1], (V)p[2], vwide);
    }
    // ---- B3: thread per face ----
    for (uint32_t k = tid; !KEYS && k < nfr; k += CB_THREADS) {
      const uint32_t s = owner_f[k];
      const uint32_t ry = rec[s].y;
      const uint32_t fi = k - (ry >> 16);
      const uint32_t tri = (uint32_t)(recf[s] >> (12 * fi)) & 0xfffu;
      const long long fct = vbase + (long long)bv + (ry & 0xffffu) + 1;  // 1-based index of the voxel's first vertex
      if (ffits || (long long)bf + k < a.fcap)
        store3<long long>(a.faces, (long long)bf + k, fct + (tri & 15u), fct + ((tri >> 4) & 15u), fct + ((tri >> 8) & 15u), fwide);
    }
    bv += nvr, bf += nfr;  // next window continues where this one ended
    __syncthreads();
  }
}

// faces[0 .. 3*min(totals[1], fcap)) += *base   (sharded fix-up after the all-gather of the slab totals)
__global__ void add_base_kernel(long long* __restrict__ faces, long long fcap, const long long* __restrict__ totals,
                                const long long* __restrict__ base) {
  const long long add = *base;
  const long long nf = totals[1] < fcap ? totals[1] : fcap;
  const long long n2 = nf * 3 / 2;  // pairs of indices
  if (add == 0) return;
  longlong2* f2 = reinterpret_cast<longlong2*>(faces);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    longlong2 v = f2[i];
    v.x += add, v.y += add;
    f2[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && ((nf * 3) & 1)) faces[nf * 3 - 1] += add;
}

}  // namespace iso
