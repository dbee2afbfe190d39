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
// problem geometry shared by count / generate / case kernels
struct Grid {
  int nx, ny, nz;        // samples
  int W;                 // words per sample column (multiple of 4, >= ceil(nz/32))
  int Wq;                // W / 4 : quad-cells per column
  int quads_per_row;     // (ny-1) * Wq : quad-cells of one voxel x-row, in scan order (y, z)
  int blocks_per_row;    // ceil(quads_per_row / CB_THREADS)
  long long row_words;   // ny * W : words between sample column (x,y) and (x+1,y)
  long long ldx;         // field leading dimension (elements)
  long long plane;       // ldx * ny
};

constexpr int CB_THREADS = 256;            // threads per count/generate block
constexpr int CB_CELLS = CB_THREADS * 4;   // 32-voxel cells per block

// ------------------------------------------------------------------------------------------------------
// (1) signpack
constexpr int SP_WARPS = 4;
constexpr int SP_XSEG = 128;   // samples of one row handled by a warp (one float4 per lane)
constexpr int SP_ZW = 8;       // z-words per warp task (8 words = one 32-byte sector per column)

__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];"
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
template <bool VEC>
__global__ void __launch_bounds__(SP_WARPS * 32)
signpack_kernel(const float* __restrict__ sdf, uint32_t* __restrict__ bits, int nx, int ny, int nz, long long ldx,
                int W, float thresh, int nxseg, long long ntasks) {
  __shared__ __align__(16) uint32_t stage[SP_WARPS][SP_ZW][SP_XSEG];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long task = (long long)blockIdx.x * SP_WARPS + wib;
  if (task >= ntasks) return;
  const int xseg = (int)(task % nxseg);
  const long long t2 = task / nxseg;
  const int y = (int)(t2 % ny);
  const int zc = (int)(t2 / ny);
  const long long plane = ldx * ny;
  const int xbase = xseg * SP_XSEG;
  const int xs = VEC ? xbase + lane * 4 : xbase + lane;
  const float* rowp = sdf + (long long)y * ldx + xs;
  const float qnan = __int_as_float(0x7fc00000);

#pragma unroll 1
  for (int zw = 0; zw < SP_ZW; ++zw) {
    const int zword = zc * SP_ZW + zw;
    const int zb = zword * 32;
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    if (zword < W && zb < nz) {
      const float* p = rowp + (long long)zb * plane;
      const bool full = zb + 32 <= nz;
      if (VEC) {
        const bool xin = xs < nx;
#pragma unroll
        for (int kb = 0; kb < 32; kb += 8) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const bool ok = xin && (full || zb + kb + u < nz);
            v[u] = ok ? ldg_stream_f4(p + (long long)(kb + u) * plane) : make_float4(qnan, qnan, qnan, qnan);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
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
          float v[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool zok = full || zb + kb + u < nz;
            const float* q = p + (long long)(kb + u) * plane;
            v[u][0] = (zok && in0) ? ldg_stream_f1(q) : qnan;
            v[u][1] = (zok && in1) ? ldg_stream_f1(q + 32) : qnan;
            v[u][2] = (zok && in2) ? ldg_stream_f1(q + 64) : qnan;
            v[u][3] = (zok && in3) ? ldg_stream_f1(q + 96) : qnan;
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
      if (wofs < W) *reinterpret_cast<uint4*>(dst) = make_uint4(c[0], c[1], c[2], c[3]);
      if (wofs + 4 < W) *reinterpret_cast<uint4*>(dst + 4) = make_uint4(c[4], c[5], c[6], c[7]);
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

// Loads the quad-cell (x, y, zq) from the bit-field.
__device__ __forceinline__ void load_quad(const uint32_t* __restrict__ bits, const Grid& g, int x, int y, int zq, Quad& q) {
  const uint32_t* c00 = bits + (long long)x * g.row_words + (long long)y * g.W + zq * 4;
  const uint32_t* c10 = c00 + g.row_words;
  const bool more = zq + 1 < g.Wq;
  const uint4 a00 = __ldg(reinterpret_cast<const uint4*>(c00));
  const uint4 a01 = __ldg(reinterpret_cast<const uint4*>(c00 + g.W));
  const uint4 a10 = __ldg(reinterpret_cast<const uint4*>(c10));
  const uint4 a11 = __ldg(reinterpret_cast<const uint4*>(c10 + g.W));
  const uint32_t n00 = more ? __ldg(c00 + 4) : 0u, n01 = more ? __ldg(c00 + g.W + 4) : 0u;
  const uint32_t n10 = more ? __ldg(c10 + 4) : 0u, n11 = more ? __ldg(c10 + g.W + 4) : 0u;
  shift_up(a00, n00, q.s00, q.t00);
  shift_up(a01, n01, q.s01, q.t01);
  shift_up(a10, n10, q.s10, q.t10);
  shift_up(a11, n11, q.s11, q.t11);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rem = g.nz - 1 - (zq * 4 + i) * 32;  // voxels z .. valid iff z < nz-1
    q.vm[i] = rem >= 32 ? 0xffffffffu : rem <= 0 ? 0u : ((1u << rem) - 1u);
  }
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
// -> the per-cell vertex total and every in-cell prefix is 12 masked popcounts.
__device__ __forceinline__ uint32_t mc_nverts_masked(const Quad& q, int i, uint32_t mask) {
  uint32_t n = __popc((q.s00[i] ^ q.s10[i]) & mask) + __popc((q.s10[i] ^ q.s11[i]) & mask) +
               __popc((q.s11[i] ^ q.s01[i]) & mask) + __popc((q.s01[i] ^ q.s00[i]) & mask);
  n += __popc((q.t00[i] ^ q.t10[i]) & mask) + __popc((q.t10[i] ^ q.t11[i]) & mask) +
       __popc((q.t11[i] ^ q.t01[i]) & mask) + __popc((q.t01[i] ^ q.t00[i]) & mask);
  n += __popc((q.s00[i] ^ q.t00[i]) & mask) + __popc((q.s10[i] ^ q.t10[i]) & mask) +
       __popc((q.s11[i] ^ q.t11[i]) & mask) + __popc((q.s01[i] ^ q.t01[i]) & mask);
  return n;
}

// block -> (x, first quad of the row chunk)
__device__ __forceinline__ void block_coords(const Grid& g, long long b, int& x, int& quad0) {
  x = (int)(b / g.blocks_per_row);
  quad0 = (int)(b - (long long)x * g.blocks_per_row) * CB_THREADS;
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
  const float* sdf;
  const uint32_t* bits;
  const unsigned long long* status;  // inclusive prefixes per block (after count_kernel)
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
};

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

// packed block-scan element: nverts bits 0..23, nfaces bits 24..43, active voxels bits 44..59
__device__ __forceinline__ unsigned long long pack3(uint32_t nv, uint32_t nf, uint32_t na) {
  return (unsigned long long)nv | ((unsigned long long)nf << 24) | ((unsigned long long)na << 44);
}
#define ISO_PK_V(p) ((uint32_t)((p) & 0xffffffull))
#define ISO_PK_F(p) ((uint32_t)(((p) >> 24) & 0xfffffull))
#define ISO_PK_A(p) ((uint32_t)(((p) >> 44) & 0xffffull))

// exclusive block scan of one packed value per thread; returns the exclusive prefix, `total` = block sum
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long* warp_s,
                                                              unsigned long long& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_s[w] = inc;
  __syncthreads();
  unsigned long long base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < CB_THREADS / 32; ++i) {
    const unsigned long long t = warp_s[i];
    if (i < w) base += t;
    tot += t;
  }
  total = tot;
  return base + inc - v;
}

constexpr int GEN_NB = CB_THREADS;         // active voxels per dense round (one per thread)
constexpr int GEN_MAXV = GEN_NB * 12;      // vertices of a round (MC: <= 12 per voxel)
constexpr int GEN_MAXF = GEN_NB * 5;       // faces of a round (MC: <= 5 per voxel)

template <int MODE, typename V>
__global__ void __launch_bounds__(CB_THREADS)
mc_generate_kernel(GenArgs a, Grid g) {
  __shared__ unsigned long long tabV[256], tabF[256];
  __shared__ unsigned long long warp_s[CB_THREADS / 32];
  __shared__ uint32_t rec_yz[GEN_NB], rec_vc[GEN_NB], rec_f[GEN_NB];
  __shared__ uint8_t owner_v[GEN_MAXV], owner_f[GEN_MAXF];
  __shared__ uint8_t edge_c[12];

  const int tid = threadIdx.x;
  tabV[tid] = ISO_MC_VERTS[tid];
  tabF[tid] = ISO_MC_FACES[tid];
  if (tid < 12) edge_c[tid] = ISO_MC_EDGE_CORNERS[tid];

  const long long b = blockIdx.x;
  int x, quad0;
  block_coords(g, b, x, quad0);
  const int qr = quad0 + tid;
  const bool live = qr < g.quads_per_row;
  int y = 0, zq = 0;
  uint32_t m[4] = {0, 0, 0, 0};
  uint32_t tnv = 0, tnf = 0, tna = 0;
  __syncthreads();  // tables visible
  if (live) {
    y = qr / g.Wq, zq = qr - y * g.Wq;
    Quad q;
    load_quad(a.bits, g, x, y, zq, q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      m[i] = active_mask(q, i);
      if (m[i]) {
        tnv += mc_nverts_masked(q, i, q.vm[i]);
        tna += __popc(m[i]);
        uint32_t mm = m[i];
        while (mm) {
          const int k = __ffs(mm) - 1;
          mm &= mm - 1;
          tnf += (uint32_t)((tabV[case_of<0>(q, i, k)] >> 52) & 7);
        }
      }
    }
  }
  unsigned long long total;
  const unsigned long long excl = block_excl_scan(pack3(tnv, tnf, tna), warp_s, total);
  const uint32_t blk_na = ISO_PK_A(total);
  if (blk_na == 0) return;  // uniform

  // exclusive prefixes of this block in the whole mesh
  unsigned long long bv = 0, bf = 0;
  if (b > 0) {
    bv = a.status[2 * (b - 1)] & VAL_MASK;
    bf = a.status[2 * (b - 1) + 1] & VAL_MASK;
  }
  const long long vbase = a.vbase + (a.vbase_dev ? *a.vbase_dev : 0);
  const double* xp = a.coords;
  const double* yp = a.coords + g.nx;
  const double* zp = a.coords + g.nx + g.ny;
  V* verts = reinterpret_cast<V*>(a.verts);

  const uint32_t my_a0 = ISO_PK_A(excl), my_v0 = ISO_PK_V(excl), my_f0 = ISO_PK_F(excl);

  for (uint32_t lo = 0; lo < blk_na; lo += GEN_NB) {
    const uint32_t hi = min(lo + (uint32_t)GEN_NB, blk_na);
    // ---- B1a: owners of the quad-cells push the records of their voxels that fall in [lo, hi) ----
    if (tna && my_a0 < hi && my_a0 + tna > lo) {
      Quad q;
      load_quad(a.bits, g, x, y, zq, q);
      uint32_t idx = my_a0, v = my_v0, f = my_f0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t mm = m[i];
        while (mm) {
          const int k = __ffs(mm) - 1;
          mm &= mm - 1;
          const uint32_t c = case_of<0>(q, i, k);
          const unsigned long long tv = tabV[c];
          if (idx >= lo && idx < hi) {
            const uint32_t s = idx - lo;
            rec_yz[s] = (uint32_t)y | ((uint32_t)((zq * 4 + i) * 32 + k) << 16);
            rec_vc[s] = v | (c << 24);
            rec_f[s] = f;
          }
          v += (uint32_t)((tv >> 48) & 15);
          f += (uint32_t)((tv >> 52) & 7);
          ++idx;
        }
      }
    }
    __syncthreads();
    const uint32_t cnt = hi - lo;
    const uint32_t rv0 = rec_vc[0] & 0xffffffu, rf0 = rec_f[0];
    // ---- B1b: thread per voxel: expand owner maps ----
    uint32_t last_v = 0, last_f = 0;
    if ((uint32_t)tid < cnt) {
      const uint32_t vc = rec_vc[tid];
      const unsigned long long tv = tabV[vc >> 24];
      const uint32_t nv = (uint32_t)((tv >> 48) & 15), nf = (uint32_t)((tv >> 52) & 7);
      const uint32_t v0 = (vc & 0xffffffu) - rv0, f0 = rec_f[tid] - rf0;
      for (uint32_t i = 0; i < nv; ++i) owner_v[v0 + i] = (uint8_t)tid;
      for (uint32_t i = 0; i < nf; ++i) owner_f[f0 + i] = (uint8_t)tid;
      last_v = v0 + nv, last_f = f0 + nf;
    }
    // round totals = end of the last voxel of the round
    __shared__ uint32_t round_nv, round_nf;
    if ((uint32_t)tid == cnt - 1) round_nv = last_v, round_nf = last_f;
    __syncthreads();
    const uint32_t nvr = round_nv, nfr = round_nf;
    const long long gv0 = (long long)bv + rv0;  // index of the round's first vertex in this slab's buffer
    const long long gf0 = (long long)bf + rf0;

    // ---- B2: thread per vertex ----
    for (uint32_t k = tid; k < nvr; k += CB_THREADS) {
      const uint32_t s = owner_v[k];
      const uint32_t vc = rec_vc[s], yz = rec_yz[s];
      const uint32_t c = vc >> 24;
      const uint32_t which = k - ((vc & 0xffffffu) - rv0);
      const uint32_t e = (uint32_t)(tabV[c] >> (4 * which)) & 15u;
      const uint32_t cc = edge_c[e];
      // MC corner offsets (dx | dy<<1 | dz<<2) for corners 0..7: 0,1,3,2,4,5,7,6
      const uint32_t oa = (0x67542310u >> (4 * (cc & 15u))) & 7u, ob = (0x67542310u >> (4 * (cc >> 4))) & 7u;
      const int vy = (int)(yz & 0xffffu), vz = (int)(yz >> 16);
      const int ax = x + (int)(oa & 1u), ay = vy + (int)((oa >> 1) & 1u), az = vz + (int)(oa >> 2);
      const int bx = x + (int)(ob & 1u), by = vy + (int)((ob >> 1) & 1u), bz = vz + (int)(ob >> 2);
      const float va = __ldg(a.sdf + ax + g.ldx * ay + g.plane * az);
      const float vb = __ldg(a.sdf + bx + g.ldx * by + g.plane * bz);
      const double pa[3] = {__ldg(xp + ax), __ldg(yp + ay), __ldg(zp + az)};
      const double pb[3] = {__ldg(xp + bx), __ldg(yp + by), __ldg(zp + bz)};
      double p[3];
      mc_interp<MODE>(a, va, vb, pa, pb, p);
      const long long gi = gv0 + k;
      if (gi < a.vcap) {
        V* o = verts + 3 * gi;
        o[0] = (V)p[0], o[1] = (V)p[1], o[2] = (V)p[2];
      }
    }
    // ---- B3: thread per face ----
    for (uint32_t k = tid; k < nfr; k += CB_THREADS) {
      const uint32_t s = owner_f[k];
      const uint32_t vc = rec_vc[s];
      const uint32_t fi = k - (rec_f[s] - rf0);
      const uint32_t tri = (uint32_t)(tabF[vc >> 24] >> (12 * fi)) & 0xfffu;
      const long long fct = vbase + (long long)bv + (vc & 0xffffffu) + 1;  // 1-based index of the voxel's first vertex
      const long long gi = gf0 + k;
      if (gi < a.fcap) {
        long long* o = a.faces + 3 * gi;
        o[0] = fct + (tri & 15u), o[1] = fct + ((tri >> 4) & 15u), o[2] = fct + ((tri >> 8) & 15u);
      }
    }
    __syncthreads();
  }
}

}  // namespace iso
