// mt_kernels.cuh -- Marching Tetrahedra on the sign bit-field (replaces src/marching_tetrahedra.jl:11-163).
//
// The reference de-duplicates vertices with a Dict keyed on a global edge id (getVertId, :67-84) while it
// sweeps the voxels sequentially.  The Dict is never iterated, so the result is fully determined by
// "the first voxel in scan order that touches an edge creates its vertex" -- the OWNER rule
// (SURVEY.md Appendix A7): the owner of an edge is the lexicographically smallest in-bounds voxel that
// contains it, i.e. the voxel shifted by -1 on every axis on which both edge endpoints have offset 0.
// Interior voxels own exactly the 7 edges that end in corner 7 = (1,1,1); voxels with index 0 on an axis
// own the extra edges lying in that low face.  Hence, without any hash table:
//   vertex order  = voxels in scan order, inside a voxel its owned crossed edges in first-appearance
//                   order of its face list (table ISO_MT_OWNED[case][lowflags]),
//   vertex count  = per 32-voxel word a handful of masked popcounts (mt_owned_masked),
//   face indices  = offset of the owner voxel + rank of the edge in the owner's list.
#pragma once
#include "../../include/b200iso.h"
#include "iso_kernels.cuh"

namespace iso {

// Crossing words of the 19 voxel edges of cell i (bit k: edge of voxel k is crossed).  Corner positions
// (src/lut/mt.jl:23-51): word of corner (dx,dy,dz) is s<dx><dy> for dz = 0, t<dx><dy> for dz = 1.
#define MT_E1(q, i) ((q).s00[i] ^ (q).s01[i])   // (0,0,0)-(0,1,0)  low axes x,z
#define MT_E2(q, i) ((q).s01[i] ^ (q).s11[i])   // (0,1,0)-(1,1,0)  low z
#define MT_E3(q, i) ((q).s10[i] ^ (q).s11[i])   // (1,0,0)-(1,1,0)  low z
#define MT_E4(q, i) ((q).s00[i] ^ (q).s10[i])   // (0,0,0)-(1,0,0)  low y,z
#define MT_E5(q, i) ((q).t00[i] ^ (q).t01[i])   // (0,0,1)-(0,1,1)  low x
#define MT_E6(q, i) ((q).t01[i] ^ (q).t11[i])   // (0,1,1)-(1,1,1)
#define MT_E7(q, i) ((q).t10[i] ^ (q).t11[i])   // (1,0,1)-(1,1,1)
#define MT_E8(q, i) ((q).t00[i] ^ (q).t10[i])   // (0,0,1)-(1,0,1)  low y
#define MT_E9(q, i) ((q).s00[i] ^ (q).t00[i])   // (0,0,0)-(0,0,1)  low x,y
#define MT_E10(q, i) ((q).s01[i] ^ (q).t01[i])  // (0,1,0)-(0,1,1)  low x
#define MT_E11(q, i) ((q).s11[i] ^ (q).t11[i])  // (1,1,0)-(1,1,1)
#define MT_E12(q, i) ((q).s10[i] ^ (q).t10[i])  // (1,0,0)-(1,0,1)  low y
#define MT_E13(q, i) ((q).s00[i] ^ (q).s11[i])  // (0,0,0)-(1,1,0)  low z
#define MT_E14(q, i) ((q).s00[i] ^ (q).t10[i])  // (0,0,0)-(1,0,1)  low y
#define MT_E15(q, i) ((q).s00[i] ^ (q).t01[i])  // (0,0,0)-(0,1,1)  low x
#define MT_E16(q, i) ((q).t00[i] ^ (q).t11[i])  // (0,0,1)-(1,1,1)
#define MT_E17(q, i) ((q).s01[i] ^ (q).t11[i])  // (0,1,0)-(1,1,1)
#define MT_E18(q, i) ((q).s10[i] ^ (q).t11[i])  // (1,0,0)-(1,1,1)
#define MT_E19(q, i) ((q).s00[i] ^ (q).t11[i])  // (0,0,0)-(1,1,1)

// Number of vertices created by the voxels of cell i selected by `mask` (owned AND crossed edges).
// fxy: bit0 = column has x == 0, bit1 = y == 0; first_word: bit 0 of this cell is the voxel z == 0.
__device__ __forceinline__ uint32_t mt_owned_masked(const Quad& q, int i, uint32_t mask, int fxy, bool first_word) {
  uint32_t n = __popc(MT_E6(q, i) & mask) + __popc(MT_E7(q, i) & mask) + __popc(MT_E11(q, i) & mask) +
               __popc(MT_E16(q, i) & mask) + __popc(MT_E17(q, i) & mask) + __popc(MT_E18(q, i) & mask) +
               __popc(MT_E19(q, i) & mask);
  if (fxy & 1) n += __popc(MT_E5(q, i) & mask) + __popc(MT_E10(q, i) & mask) + __popc(MT_E15(q, i) & mask);
  if (fxy & 2) n += __popc(MT_E8(q, i) & mask) + __popc(MT_E12(q, i) & mask) + __popc(MT_E14(q, i) & mask);
  if (fxy == 3) n += __popc(MT_E9(q, i) & mask);
  if (first_word) {
    const uint32_t mz = mask & 1u;
    n += __popc(MT_E2(q, i) & mz) + __popc(MT_E3(q, i) & mz) + __popc(MT_E13(q, i) & mz);
    if (fxy & 1) n += __popc(MT_E1(q, i) & mz);
    if (fxy & 2) n += __popc(MT_E4(q, i) & mz);
  }
  return n;
}

// single 32-voxel cell (x, y, zw) for owner look-ups outside the block's own quad-cells
__device__ __forceinline__ void load_cell(const uint32_t* __restrict__ bits, const Grid& g, int x, int y, int zw, Quad& q) {
  const uint32_t* c00 = bits + (long long)x * g.row_words + (long long)y * g.W + zw;
  const uint32_t* c10 = c00 + g.row_words;
  const bool more = zw + 1 < g.W;
  const uint32_t a00 = __ldg(c00), a01 = __ldg(c00 + g.W), a10 = __ldg(c10), a11 = __ldg(c10 + g.W);
  const uint32_t n00 = more ? __ldg(c00 + 1) : 0u, n01 = more ? __ldg(c00 + g.W + 1) : 0u;
  const uint32_t n10 = more ? __ldg(c10 + 1) : 0u, n11 = more ? __ldg(c10 + g.W + 1) : 0u;
  q.s00[0] = a00, q.s01[0] = a01, q.s10[0] = a10, q.s11[0] = a11;
  q.t00[0] = __funnelshift_r(a00, n00, 1), q.t01[0] = __funnelshift_r(a01, n01, 1);
  q.t10[0] = __funnelshift_r(a10, n10, 1), q.t11[0] = __funnelshift_r(a11, n11, 1);
  const int rem = g.nz - 1 - zw * 32;
  q.vm[0] = rem >= 32 ? 0xffffffffu : rem <= 0 ? 0u : ((1u << rem) - 1u);
}

// Base.max / Base.min on floats: NaN if either argument is NaN; equal arguments resolve by sign.
template <class F>
__device__ __forceinline__ F jl_max(F x, F y) {
  if (x != x || y != y) return x + y;
  if (x == y) return signbit(x) ? y : x;
  return x > y ? x : y;
}
template <class F>
__device__ __forceinline__ F jl_min(F x, F y) {
  if (x != x || y != y) return x + y;
  if (x == y) return signbit(x) ? x : y;
  return x < y ? x : y;
}

constexpr int MTG_NB = CB_THREADS;       // active voxels per dense round
constexpr int MTG_MAXV = MTG_NB * 13;    // <= 13 owned edges per voxel (boundary corner voxel)
constexpr int MTG_MAXF = MTG_NB * 12;    // <= 12 faces per voxel
constexpr int MTG_EDGES = 13;            // <= 13 crossed edges per voxel

// A32: promote_type(typeof(iso), typeof(eps)) == Float32 (vertPos weights in Float32), else Float64.
// P32: points (ranges) are Float32.  V: vertex element type.  T: field element type (Float64 implies !A32).
// per-case tables through L1 straight from global memory, like the MC kernel (generate 0.613 -> 0.606 ms at 512^3,
// 11.8 instead of 24.6 KB of shared memory per block); -DISO_MT_SMEM_TABLES restores the per-block staging
#ifndef ISO_MT_SMEM_TABLES
#define MT_OWN0_S(i) __ldg(&ISO_MT_OWNED[(i) * 8])
#define MT_FACES_S(i) __ldg(&ISO_MT_FACES[i])
#define MT_CROSS_S(i) __ldg(&ISO_MT_CROSS[i])
#define MT_CLIST_S(i) __ldg(&ISO_MT_CROSSLIST[i])
#define MT_RANK0_S(i) __ldg(&ISO_MT_RANK0[i])
#define MT_NOWN0_S(i) __ldg(&ISO_MT_NOWN[(i) * 8])
#define MT_NF_S(i) __ldg(&ISO_MT_NF[i])
#else
#define MT_OWN0_S(i) own0_s[i]
#define MT_FACES_S(i) faces_s[i]
#define MT_CROSS_S(i) cross_s[i]
#define MT_CLIST_S(i) clist_s[i]
#define MT_RANK0_S(i) rank0_s[i]
#define MT_NOWN0_S(i) nown0_s[i]
#define MT_NF_S(i) nf_s[i]
#endif
template <bool A32, bool P32, typename V, typename T = float>
#ifndef ISO_MT_MINB
#define ISO_MT_MINB 8
#endif
__global__ void __launch_bounds__(CB_THREADS, ISO_MT_MINB)
mt_generate_kernel(GenArgs a, Grid g, const uint32_t* __restrict__ celloff) {
#ifdef ISO_MT_SMEM_TABLES
  __shared__ unsigned long long own0_s[256];   // ISO_MT_OWNED[c][flags = 0]
  __shared__ unsigned long long faces_s[768];  // ISO_MT_FACES
  __shared__ uint32_t cross_s[256];
  __shared__ unsigned long long clist_s[256];  // ISO_MT_CROSSLIST
  __shared__ uint32_t rank0_s[256];            // ISO_MT_RANK0
  __shared__ uint8_t nown0_s[256], nf_s[256];
#endif
  __shared__ uint32_t emask_s[8];               // edges (bit e - 1) by the axes on which their owner is the previous voxel
  __shared__ int32_t nb_base[MTG_NB * 8];       // per (voxel, lower neighbour): id of the neighbour's first vertex, relative to bv ...
  __shared__ uint16_t nb_info[MTG_NB * 8];      // ... and its case | low-boundary flags << 8
  __shared__ uint8_t slot_s[20];               // ISO_MT_OWN_SLOT
  __shared__ uint16_t einfo_s[20];
  __shared__ uint8_t eshift_s[160];
  __shared__ uint32_t s_w[CB_THREADS / 32];
  __shared__ uint2 rec_yc[MTG_NB];  // .x = y | z << 16, .y = case index
  __shared__ uint32_t rec_vc[MTG_NB], rec_f[MTG_NB];
  __shared__ int32_t evid[MTG_NB * MTG_EDGES];  // vertex id of each crossed edge, relative to the block's first vertex
  __shared__ uint8_t owner_v[MTG_MAXV];
  __shared__ uint8_t owner_f[MTG_MAXF];

  const int tid = threadIdx.x;

  const unsigned b = blockIdx.x;
  if (a.abort_flag && *a.abort_flag) return;  // the exchange that was to deliver the vertex base failed
  const TMap tm = thread_map_div(g, b);
  const int x = tm.x;
  if (g.ghost && x == 0) return;  // ghost row of an MT slab: its faces and vertices belong to the previous slab
  {
    // inclusive (vertex, face) prefixes of every block are in `status`: a block that adds no face has no active
    // voxel (each one emits >= 1 face) -- leave before touching the bit-field (most blocks of a sparse field).
    const unsigned long long f1 = a.status[2 * (unsigned long long)b + 1] & VAL_MASK;
    const unsigned long long f0 = b > 0 ? (a.status[2 * (unsigned long long)(b - 1) + 1] & VAL_MASK) : 0ull;
    if (f0 == f1) return;
  }
  // per-case tables into shared memory (after the early exits: empty blocks do not pay for it)
#ifdef ISO_MT_SMEM_TABLES
  for (int i = tid; i < 256; i += CB_THREADS) {
    own0_s[i] = ISO_MT_OWNED[i * 8];
    nown0_s[i] = ISO_MT_NOWN[i * 8];
    nf_s[i] = ISO_MT_NF[i];
    cross_s[i] = ISO_MT_CROSS[i];
    clist_s[i] = ISO_MT_CROSSLIST[i];
    rank0_s[i] = ISO_MT_RANK0[i];
  }
  for (int i = tid; i < 768; i += CB_THREADS) faces_s[i] = ISO_MT_FACES[i];
#endif
  for (int i = tid; i < 20; i += CB_THREADS) einfo_s[i] = ISO_MT_EDGE_INFO[i], slot_s[i] = ISO_MT_OWN_SLOT[i];
  for (int i = tid; i < 160; i += CB_THREADS) eshift_s[i] = ISO_MT_EDGE_SHIFT[i];
  if (tid < 8) {
    uint32_t m = 0;
    for (int e = 1; e <= 19; ++e)
      if (((ISO_MT_EDGE_INFO[e] >> 6) & 7) == tid) m |= 1u << (e - 1);
    emask_s[tid] = m;
  }
  // The count left the block's active-voxel records (scan order) unless there are more than REC_CAP of them: then
  // (dense fields) this block derives them from the bit-field itself (A + B1a).
  const uint32_t nrec = __ldg(a.nrecs + b);
  const bool from_recs = nrec <= (uint32_t)g.rec_cap;  // (uniform over the block)
  const uint32_t q_lo = (b - (unsigned)x * (unsigned)g.blocks_per_row) * CB_THREADS;  // first quad-cell of the block in its x-row
  uint32_t tna = 0, my_a0 = 0, blk_na = nrec;
  if (!from_recs) {
    // ---- A: active voxels per thread, exclusive scan (thread order == scan order) ----
    Quad qc;  // (not kept: the MT kernel is register-bound, the push reloads the quad-cell through L1)
    tna = count_active(a.bits, g, tm, qc);
    my_a0 = block_excl_scan_u32(tna, s_w, blk_na);
  }
  if (blk_na == 0) return;

  unsigned long long bv = 0, bf = 0;
  if (b > 0) {
    bv = a.status[2 * (unsigned long long)(b - 1)] & VAL_MASK;
    bf = a.status[2 * (unsigned long long)(b - 1) + 1] & VAL_MASK;
  }
  const long long vbase = a.vbase + (a.vbase_dev ? *a.vbase_dev : 0);
  const double* xp = a.coords;
  const double* yp = a.coords + g.nx;
  const double* zp = a.coords + g.nx + g.ny;
  V* verts = reinterpret_cast<V*>(a.verts);
  const int fx = (x + g.xoff) == 0 ? 1 : 0;  // low-boundary flag in GLOBAL coordinates
  // MT sharding: vertices/faces of the ghost row precede this slab's output; shift positions and ids by them
  unsigned long long gshift_v = 0, gshift_f = 0;
  if (g.ghost) {
    gshift_v = a.status[2 * (unsigned long long)(g.blocks_per_row - 1)] & VAL_MASK;
    gshift_f = a.status[2 * (unsigned long long)(g.blocks_per_row - 1) + 1] & VAL_MASK;
  }
  uint32_t wv = 0;  // vertices of the block emitted by previous windows

  auto owned_word = [&](uint32_t c, int flags) -> unsigned long long {
    return flags == 0 ? MT_OWN0_S(c) : __ldg(&ISO_MT_OWNED[c * 8 + flags]);
  };
  auto nown_of = [&](uint32_t c, int flags) -> uint32_t {
    return flags == 0 ? (uint32_t)MT_NOWN0_S(c) : (uint32_t)__ldg(&ISO_MT_NOWN[c * 8 + flags]);
  };

  for (uint32_t lo = 0; lo < blk_na; lo += MTG_NB) {
    const uint32_t hi = min(lo + (uint32_t)MTG_NB, blk_na);
    const uint32_t cnt = hi - lo;
    // ---- B1a: records (position, case) of the window's voxels, in scan order ----
    if (from_recs) {
      if ((uint32_t)tid < cnt) {  // the count's record: case | voxel-in-quad-cell << 8 | quad-cell-in-block << 15
        const uint32_t wd = __ldg(a.recs + (unsigned long long)b * (unsigned)g.rec_cap + lo + tid);
        const uint32_t qr = q_lo + (wd >> 15), y = qr / (uint32_t)g.Wq, zq = qr - y * (uint32_t)g.Wq;
        rec_yc[tid] = make_uint2(y | ((zq * 128u + ((wd >> 8) & 127u)) << 16), wd & 0xffu);
      }
    } else if (tna && my_a0 < hi && my_a0 + tna > lo) {
      Quad qp;
      load_quad(a.bits, g, tm.x, tm.y, tm.zq, qp);
      push_records<1>(qp, tm, my_a0, lo, hi, rec_yc, 1);
    }
    __syncthreads();
    // ---- B1b: thread per voxel: counts -> scan -> owner maps ----
    uint32_t nv = 0, nf = 0;
    if ((uint32_t)tid < cnt) {
      const uint32_t c = rec_yc[tid].y, yz = rec_yc[tid].x;
      const int flags = fx | ((yz & 0xffffu) == 0 ? 2 : 0) | ((yz >> 16) == 0 ? 4 : 0);
      nv = nown_of(c, flags), nf = MT_NF_S(c);
    }
    uint32_t wtot;
    const uint32_t ex = block_excl_scan_u32(nv | (nf << 16), s_w, wtot);
    if ((uint32_t)tid < cnt) {
      const uint32_t v0 = ex & 0xffffu, f0 = ex >> 16;
      rec_vc[tid] = (wv + v0) | (rec_yc[tid].y << 24);  // vertex offset relative to the block's first vertex
      rec_f[tid] = f0;
      for (uint32_t i = 0; i < nv; ++i) owner_v[v0 + i] = (uint8_t)tid;
      for (uint32_t i = 0; i < nf; ++i) owner_f[f0 + i] = (uint8_t)tid;
    }
    __syncthreads();
    const uint32_t rv0 = wv, rf0 = 0;
    // ---- B1c: vertex id of every crossed edge, through the voxel that OWNS the edge (mt_kernels.cuh header).
    // The owner of a foreign edge is one of the 7 lower neighbours (shift sh = 1..7: -1 along x / y / z for bits 0 / 1 / 2);
    // its case and the id of its first vertex come from the bit-field, the block prefixes and celloff -- the expensive
    // part -- and several crossed edges usually share one owner.  So (1) a thread per (voxel, neighbour) resolves each
    // NEEDED neighbour once, then (2) a thread per (voxel, crossed edge) only looks the edge's rank up in the owner's list.
    for (uint32_t it = tid; it < cnt * 8; it += CB_THREADS) {
      const uint32_t s = it >> 3;
      const int sh = (int)(it & 7u);
      const uint32_t vc = rec_vc[s], yz = rec_yc[s].x;
      const uint32_t cm = MT_CROSS_S(vc >> 24);
      const int vy = (int)(yz & 0xffffu), vz = (int)(yz >> 16);
      const int flags = fx | (vy == 0 ? 2 : 0) | (vz == 0 ? 4 : 0);
      // crossed edges whose owner is neighbour sh: low-axes L with L & ~flags == sh, i.e. L = sh | m for m within flags
      uint32_t mine = 0;
      if (sh != 0 && (sh & flags) == 0) {
        mine = emask_s[sh];
        if (flags & 1) mine |= emask_s[sh | 1];
        if (flags & 2) mine |= emask_s[sh | 2];
        if (flags & 4) mine |= emask_s[sh | 4];
        if ((flags & 3) == 3) mine |= emask_s[sh | 3];
        if ((flags & 5) == 5) mine |= emask_s[sh | 5];
        if ((flags & 6) == 6) mine |= emask_s[sh | 6];
        // (flags == 7 leaves sh == 0 only)
      }
      if ((mine & cm) == 0) continue;
      const int ox = x - (sh & 1), oy = vy - ((sh >> 1) & 1), oz = vz - (sh >> 2);
      const int oflags = ((ox + g.xoff) == 0 ? 1 : 0) | (oy == 0 ? 2 : 0) | (oz == 0 ? 4 : 0);
      Quad q;
      load_cell(a.bits, g, ox, oy, oz >> 5, q);
      const int k = oz & 31;
      const uint32_t oc = case_of<1>(q, 0, k);
      // vertices created before the owner voxel: block prefix + cell prefix + in-cell prefix
      const uint32_t below = q.vm[0] & ((1u << k) - 1u);
      const uint32_t incell = mt_owned_masked(q, 0, below, oflags & 3, (oz >> 5) == 0);
      const long long ob = (long long)ox * g.blocks_per_row + (oy * g.Wq + (oz >> 7)) / CB_THREADS;  // the owner's block
      const unsigned long long obv = ob > 0 ? (a.status[2 * (ob - 1)] & VAL_MASK) : 0ull;
      const uint32_t co = __ldg(celloff + (long long)ox * g.row_words + (long long)oy * g.W + (oz >> 5));
      nb_base[it] = (int32_t)((long long)(obv + co + incell) - (long long)bv);
      nb_info[it] = (uint16_t)(oc | ((uint32_t)oflags << 8));
    }
    __syncthreads();
    for (uint32_t it = tid; it < cnt * MTG_EDGES; it += CB_THREADS) {
      const uint32_t s = it / MTG_EDGES, j = it - s * MTG_EDGES;
      const uint32_t vc = rec_vc[s], yz = rec_yc[s].x;
      const uint32_t c = vc >> 24;
      const uint32_t cm = MT_CROSS_S(c);
      if (j >= (uint32_t)__popc(cm)) continue;
      // j-th crossed edge (ascending id), 1..19: from the per-case list; a 13th is the highest crossed edge
      int e = j < 12 ? (int)((MT_CLIST_S(c) >> (5 * j)) & 31u) : 32 - __clz(cm);
      const int vy = (int)(yz & 0xffffu), vz = (int)(yz >> 16);
      int flags = fx | (vy == 0 ? 2 : 0) | (vz == 0 ? 4 : 0);
      const int sh = ((einfo_s[e] >> 6) & 7) & ~flags;  // axes on which the owner is the previous voxel
      uint32_t oc = c;
      int32_t base = (int32_t)(vc & 0xffffffu);
      if (sh != 0) {  // foreign edge: the owner's case / flags / first vertex, and the edge's name in the owner's frame
        const uint32_t info = nb_info[s * 8 + sh];
        base = nb_base[s * 8 + sh];
        oc = info & 0xffu, flags = (int)(info >> 8);
        e = eshift_s[e * 8 + sh];
      }
      int r = 0;
      if (flags == 0) {  // interior owner: rank of an interior-owned edge from the per-case table
        r = (int)((MT_RANK0_S(oc) >> (3 * slot_s[e])) & 7u);
      } else {
        const unsigned long long ow = owned_word(oc, flags);
        while (((ow >> (5 * r)) & 31u) != (unsigned)e) ++r;
      }
      evid[s * MTG_EDGES + j] = base + r;
    }
    __syncthreads();
    const uint32_t nvr = wtot & 0xffffu, nfr = wtot >> 16;
    const long long gv0 = (long long)bv + rv0 - (long long)gshift_v;  // position in this slab's vertex buffer
    const long long gf0 = (long long)bf - (long long)gshift_f;

    // ---- B2: thread per vertex (vertPos, src/marching_tetrahedra.jl:42-55) ----
    for (uint32_t k = tid; k < nvr; k += CB_THREADS) {
      const uint32_t s = owner_v[k];
      const uint32_t vc = rec_vc[s], yz = rec_yc[s].x;
      const uint32_t c = vc >> 24;
      const int vy = (int)(yz & 0xffffu), vz = (int)(yz >> 16);
      const int flags = fx | (vy == 0 ? 2 : 0) | (vz == 0 ? 4 : 0);
      const uint32_t which = k - ((vc & 0xffffffu) - rv0);
      const int e = (int)((owned_word(c, flags) >> (5 * which)) & 31u);
      const uint32_t info = einfo_s[e];
      const int sx = info & 1, sy = (info >> 1) & 1, sz = (info >> 2) & 1;
      const int tx = (info >> 3) & 1, ty = (info >> 4) & 1, tz = (info >> 5) & 1;
      const T* fld = reinterpret_cast<const T*>(a.sdf);
      const T srcVal = __ldg(fld + (x + sx) + g.ldx * (vy + sy) + g.plane * (vz + sz));
      const T tgtVal = __ldg(fld + (x + tx) + g.ldx * (vy + ty) + g.plane * (vz + tz));
      const double bx = __ldg(xp + x), by = __ldg(yp + vy), bz = __ldg(zp + vz);
      const double ex = __ldg(xp + x + 1), ey = __ldg(yp + vy + 1), ez = __ldg(zp + vz + 1);
      const T den = tgtVal - srcVal;  // one IEEE subtraction in the field type (-fmad=false, nothing to contract)
      double p[3];
      const double base[3] = {bx, by, bz}, endp[3] = {ex, ey, ez};
      const int c1[3] = {sx, sy, sz}, c2[3] = {tx, ty, tz};
      if constexpr (sizeof(T) == 8) {
        // Float64 field: every operand promotes to Float64 (one(T) - eps too)
        const double isod = a.iso_is_f32 ? (double)a.iso_f : a.iso_d;
        const double q = __ddiv_rn(__dsub_rn(isod, (double)srcVal), (double)den);
        const double epsd = a.eps_is_f32 ? (double)a.eps_f : a.eps_d;
        const double av = jl_min(jl_max(q, epsd), __dsub_rn(1.0, epsd));
        const double bw = __dsub_rn(1.0, av);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double w = __dadd_rn(__dmul_rn((double)c1[d], bw), __dmul_rn((double)c2[d], av));
          const double dd = P32 ? (double)__fsub_rn((float)endp[d], (float)base[d]) : __dsub_rn(endp[d], base[d]);
          p[d] = __dadd_rn(base[d], __dmul_rn(w, dd));
        }
      } else if (A32) {
        const float q = __fdiv_rn(__fsub_rn(a.iso_f, srcVal), den);
        const float av = jl_min(jl_max(q, a.eps_f), __fsub_rn(1.0f, a.eps_f));
        const float bw = __fsub_rn(1.0f, av);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float w = __fadd_rn(__fmul_rn((float)c1[d], bw), __fmul_rn((float)c2[d], av));
          if (P32) {
            const float dd = __fsub_rn((float)endp[d], (float)base[d]);
            p[d] = (double)__fadd_rn((float)base[d], __fmul_rn(w, dd));
          } else {
            const double dd = __dsub_rn(endp[d], base[d]);
            p[d] = __dadd_rn(base[d], __dmul_rn((double)w, dd));
          }
        }
      } else {
        const double q = a.iso_is_f32 ? (double)__fdiv_rn(__fsub_rn(a.iso_f, srcVal), den)
                                      : __ddiv_rn(__dsub_rn(a.iso_d, (double)srcVal), (double)den);
        const double epsd = a.eps_is_f32 ? (double)a.eps_f : a.eps_d;
        const double hi1 = a.eps_is_f32 ? (double)__fsub_rn(1.0f, a.eps_f) : __dsub_rn(1.0, a.eps_d);
        const double av = jl_min(jl_max(q, epsd), hi1);
        const double bw = __dsub_rn(1.0, av);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double w = __dadd_rn(__dmul_rn((double)c1[d], bw), __dmul_rn((double)c2[d], av));
          const double dd = P32 ? (double)__fsub_rn((float)endp[d], (float)base[d]) : __dsub_rn(endp[d], base[d]);
          p[d] = __dadd_rn(base[d], __dmul_rn(w, dd));
        }
      }
      const long long gi = gv0 + k;
      if (gi < a.vcap) {
        V* o = verts + 3 * gi;
        o[0] = (V)p[0], o[1] = (V)p[1], o[2] = (V)p[2];
      }
    }
    // ---- B3: thread per face ----
    for (uint32_t k = tid; k < nfr; k += CB_THREADS) {
      const uint32_t s = owner_f[k];
      const uint32_t c = rec_vc[s] >> 24;
      const uint32_t fi = k - (rec_f[s] - rf0);
      const uint32_t cm = MT_CROSS_S(c);
      const long long gi = gf0 + k;
      if (gi < a.fcap) {
        long long* o = a.faces + 3 * gi;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const uint32_t slot = 3 * fi + j;
          const uint32_t e = (uint32_t)(MT_FACES_S(c * 3 + slot / 12) >> (5 * (slot % 12))) & 31u;
          const uint32_t es = __popc(cm & ((1u << (e - 1)) - 1u));
          o[j] = vbase - (long long)gshift_v + (long long)bv + evid[s * MTG_EDGES + es] + 1;
        }
      }
    }
    wv += nvr, bf += nfr;
    __syncthreads();
  }
}

inline int launch_mt_generate(const GenArgs& a, const Grid& g, const b200iso_params& p, int vert_is_f64, const uint32_t* celloff,
                              unsigned nb, cudaStream_t st) {
  const bool a32 = p.iso_is_f32 && p.eps_is_f32;
  const bool p32 = p.range_kind == B200ISO_RANGE_F32;
  if (p.field_is_f64) {
    if (p32) mt_generate_kernel<false, true, double, double><<<nb, CB_THREADS, 0, st>>>(a, g, celloff);
    else mt_generate_kernel<false, false, double, double><<<nb, CB_THREADS, 0, st>>>(a, g, celloff);
  } else if (a32) {
    if (p32) mt_generate_kernel<true, true, float><<<nb, CB_THREADS, 0, st>>>(a, g, celloff);
    else if (vert_is_f64) mt_generate_kernel<true, false, double><<<nb, CB_THREADS, 0, st>>>(a, g, celloff);
    else mt_generate_kernel<true, false, float><<<nb, CB_THREADS, 0, st>>>(a, g, celloff);
  } else {
    if (p32) mt_generate_kernel<false, true, double><<<nb, CB_THREADS, 0, st>>>(a, g, celloff);
    else mt_generate_kernel<false, false, double><<<nb, CB_THREADS, 0, st>>>(a, g, celloff);
  }
  return 0;
}

}  // namespace iso
