// oracle/iso_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the isosurface hot path of Meshing.jl v0.7.0, used as the parity oracle for the
// CUDA library and as the "reference CPU path" timing stand-in (Julia is not installed in this image,
// so the reference itself cannot run).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library.  The product
// (meshing.jl_b200/) never links, imports or calls it.
//
// Parity pinning: the reference's own known-answer tests (test/runtests.jl:15-33) and its one exact
// count test ("noisy spheres", test/runtests.jl:152-172: 3466 vertices / 6928 faces, input from
// Julia's MersenneTwister(0) stream restated in oracle/julia_mt.py) are checked in
// tests/test_oracle.py.  No reference test pins vertex coordinates or face index values bit-exactly;
// for those the oracle's fidelity rests on following the cited lines below.
//
// What is restated (behaviour, not text):
//   _get_cubeindex / _no_triangles          src/common.jl:10-29
//   isosurface(::MarchingCubes) + helpers    src/marching_cubes.jl:27-120
//   isosurface(::MarchingTetrahedra) + helpers src/marching_tetrahedra.jl:11-163
//   tables                                   oracle/ref_tables.h (generated from src/lut/*.jl)
//   LinRange / lerpi, promote_type, min/max  Julia Base (not under /root/reference), SURVEY.md App. A
//
// Scan order is x outermost, z innermost (src/marching_cubes.jl:40, src/marching_tetrahedra.jl:144)
// on column-major data sdf[x + nx*(y + ny*z)].  Compile with -O2 -ffp-contract=off: every Julia
// arithmetic operation is one IEEE operation in the promoted type; C++'s usual arithmetic
// conversions (float op double -> double) coincide with Julia's promotion for Float32/Float64.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "ref_tables.h"

namespace {

// ---- Julia Base pieces -----------------------------------------------------------------------------

// LinRange{P}(a, b, n)[i] (0-based i here) == lerpi(i, max(n-1,1), a, b) = P((1-t)*a + t*b), t = i/d
// evaluated in Float64 (base/range.jl, Julia >= 1.9).
// `off`: the array handed to the oracle holds samples [off, off + nx) of a volume with n samples along this axis
// (slab-wise parity of volumes too big for the host: Marching Cubes carries no state between voxels, so the sweep
// of a slab with the WHOLE volume's coordinates is exactly that part of the whole sweep).
template <class P>
struct LinRange {
  P a, b;
  int64_t d, off;
  LinRange(double a_, double b_, int64_t n, int64_t off_ = 0) : a((P)a_), b((P)b_), d(n - 1 > 1 ? n - 1 : 1), off(off_) {}
  P operator[](int64_t i) const {
    double t = (double)(i + off) / (double)d;
    double u = (1.0 - t) * (double)a;
    double v = t * (double)b;
    return (P)(u + v);
  }
};

// Base.max / Base.min on floats: NaN if either argument is NaN (unlike fmax/fmin).
template <class F>
F jl_max(F x, F y) {
  if (std::isnan(x) || std::isnan(y)) return x + y;
  if (x == y) return std::signbit(x) ? y : x;  // max(-0.0, 0.0) == 0.0
  return x > y ? x : y;
}
template <class F>
F jl_min(F x, F y) {
  if (std::isnan(x) || std::isnan(y)) return x + y;
  if (x == y) return std::signbit(x) ? x : y;
  return x < y ? x : y;
}

template <class A, class B>
using Promote = typename std::common_type<A, B>::type;

// ---- src/common.jl:10-20 ---------------------------------------------------------------------------
template <class T, class I>
inline uint8_t get_cubeindex(const T* v, I iso) {
  uint8_t c = v[0] < iso ? 0x01 : 0x00;
  if (v[1] < iso) c |= 0x02;
  if (v[2] < iso) c |= 0x04;
  if (v[3] < iso) c |= 0x08;
  if (v[4] < iso) c |= 0x10;
  if (v[5] < iso) c |= 0x20;
  if (v[6] < iso) c |= 0x40;
  if (v[7] < iso) c |= 0x80;
  return c;
}
// src/common.jl:27-29
inline bool no_triangles(uint8_t c) { return c == 0x00 || c == 0xff; }

struct Result {
  int vert_is_f64 = 0;
  std::vector<float> v32;
  std::vector<double> v64;
  std::vector<int64_t> faces;  // 1-based, 3 per face
  int64_t nverts() const { return (int64_t)((vert_is_f64 ? v64.size() : v32.size()) / 3); }
};

template <class V>
std::vector<V>& verts_of(Result& r);
template <>
std::vector<float>& verts_of<float>(Result& r) { return r.v32; }
template <>
std::vector<double>& verts_of<double>(Result& r) { return r.v64; }

// ---- Marching Cubes: src/marching_cubes.jl:27-120 ------------------------------------------------
// One x-range [xa, xb) of the sweep; the full call uses [0, nx-1).  Face indices are relative to the
// number of vertices already in `vts` (fct = length(vts), :69).
template <class T, class I, class P, class V>
void mc_sweep(const T* sdf, int64_t nx, int64_t ny, int64_t nz, I iso, const LinRange<P>& xp,
              const LinRange<P>& yp, const LinRange<P>& zp, int64_t xa, int64_t xb, std::vector<V>& vts,
              std::vector<int64_t>& fcs) {
  const int64_t sx = 1, sy = nx, sz = nx * ny;
  for (int64_t xi = xa; xi < xb; ++xi)
    for (int64_t yi = 0; yi < ny - 1; ++yi)
      for (int64_t zi = 0; zi < nz - 1; ++zi) {
        const T* p = sdf + xi * sx + yi * sy + zi * sz;
        // corner order :42-49
        T vals[8] = {p[0], p[sx], p[sx + sy], p[sy], p[sz], p[sx + sz], p[sx + sy + sz], p[sy + sz]};
        uint8_t c = get_cubeindex(vals, iso);  // :53
        if (no_triangles(c)) continue;         // :56
        // mc_vert_points :111-120
        P X0 = xp[xi], X1 = xp[xi + 1], Y0 = yp[yi], Y1 = yp[yi + 1], Z0 = zp[zi], Z1 = zp[zi + 1];
        P pts[8][3] = {{X0, Y0, Z0}, {X1, Y0, Z0}, {X1, Y1, Z0}, {X0, Y1, Z0},
                       {X0, Y0, Z1}, {X1, Y0, Z1}, {X1, Y1, Z1}, {X0, Y1, Z1}};
        // process_mc_voxel! :67-92
        int64_t fct = (int64_t)(vts.size() / 3);
        const uint8_t* vert_to_add = REF_mc_verts[c - 1];
        for (int i = 0; i < 12; ++i) {
          uint8_t vt = vert_to_add[i];
          if (vt == 0) break;
          int e1 = REF_mc_edge_list[vt - 1][0] - 1, e2 = REF_mc_edge_list[vt - 1][1] - 1;
          // vertex_interp :100-104
          auto mu = (iso - vals[e1]) / (vals[e2] - vals[e1]);
          for (int q = 0; q < 3; ++q) {
            auto pq = pts[e1][q] + mu * (pts[e2][q] - pts[e1][q]);
            vts.push_back((V)pq);
          }
        }
        const uint8_t* offsets = REF_mc_connectivity[REF_mc_eq_mapping[c - 1] - 1];
        fcs.push_back(fct + 3);
        fcs.push_back(fct + 2);
        fcs.push_back(fct + 1);
        for (int i = 0; i < 12; i += 3) {
          if (offsets[i] == 0) break;
          fcs.push_back(fct + offsets[i + 2]);
          fcs.push_back(fct + offsets[i + 1]);
          fcs.push_back(fct + offsets[i]);
        }
      }
}

template <class T, class I, class P, class V>
void run_mc(const T* sdf, int64_t nx, int64_t ny, int64_t nz, I iso, double x0, double x1, double y0,
            double y1, double z0, double z1, int nthreads, int64_t xlo, int64_t xhi, int64_t x_off, int64_t nx_glob, Result& out) {
  LinRange<P> xp(x0, x1, nx_glob > 0 ? nx_glob : nx, x_off), yp(y0, y1, ny), zp(z0, z1, nz);
  std::vector<V>& vts = verts_of<V>(out);
  out.vert_is_f64 = std::is_same<V, double>::value;
  if (nx < 2 || ny < 2 || nz < 2) return;
  // [xlo, xhi): voxel x-planes to sweep (bench: a bounded sample of a large field, same strides as the
  // full sweep); the whole volume is [0, nx-1).
  if (xlo < 0) xlo = 0;
  if (xhi < 0 || xhi > nx - 1) xhi = nx - 1;
  if (xhi <= xlo) return;
  if (nthreads <= 1) {
    mc_sweep<T, I, P, V>(sdf, nx, ny, nz, iso, xp, yp, zp, xlo, xhi, vts, out.faces);
    return;
  }
  // x-slab threaded driver (bench CPU arm only): each thread runs the same sweep on its x-range,
  // results are concatenated in x order with face indices rebased -- byte-identical to 1 thread.
  int64_t nvx = xhi - xlo;
  if (nthreads > nvx) nthreads = (int)nvx;
  std::vector<std::vector<V>> pv(nthreads);
  std::vector<std::vector<int64_t>> pf(nthreads);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    int64_t xa = xlo + nvx * t / nthreads, xb = xlo + nvx * (t + 1) / nthreads;
    th.emplace_back([&, t, xa, xb] { mc_sweep<T, I, P, V>(sdf, nx, ny, nz, iso, xp, yp, zp, xa, xb, pv[t], pf[t]); });
  }
  for (auto& t : th) t.join();
  size_t tv = 0, tf = 0;
  for (int t = 0; t < nthreads; ++t) tv += pv[t].size(), tf += pf[t].size();
  vts.reserve(tv);
  out.faces.reserve(tf);
  for (int t = 0; t < nthreads; ++t) {
    int64_t base = (int64_t)(vts.size() / 3);
    vts.insert(vts.end(), pv[t].begin(), pv[t].end());
    for (int64_t f : pf[t]) out.faces.push_back(f + base);
  }
}

// ---- Marching Tetrahedra: src/marching_tetrahedra.jl:11-163 ----------------------------------------
// tetIx :11-18
inline uint8_t tetIx(int tIx, uint8_t cubeindex) {
  uint8_t v1 = REF_subTetsMask[tIx - 1][0], v2 = REF_subTetsMask[tIx - 1][1];
  uint8_t ix = (0x01 & cubeindex) | ((0x40 & cubeindex) >> 3);
  if (v1 & cubeindex) ix |= 0x02;
  if (v2 & cubeindex) ix |= 0x04;
  return ix;
}
// vertId :29-32 (x, y, z 1-based)
inline int64_t vertId(int e, int64_t x, int64_t y, int64_t z, int64_t nx, int64_t ny) {
  const uint8_t* d = REF_voxCrnrPos[REF_voxEdgeCrnrs[e - 1][0] - 1];
  return REF_voxEdgeDir[e - 1] + 7 * (x - 1 + d[0] + nx * (y - 1 + d[1] + ny * (z - 1 + d[2])));
}
// voxEdgeId :91-95
inline int voxEdgeId(int subTetIx, int tetEdgeIx) {
  int s = REF_subTets[subTetIx - 1][REF_tetEdgeCrnrs[tetEdgeIx - 1][0] - 1];
  int t = REF_subTets[subTetIx - 1][REF_tetEdgeCrnrs[tetEdgeIx - 1][1] - 1];
  return REF_voxEdgeIx[s - 1][t - 1];
}

template <class T, class I, class E, class P, class V>
struct MTState {
  const T* sdf;
  int64_t nx, ny, nz;
  I iso;
  E eps;
  LinRange<P> xp, yp, zp;
  std::unordered_map<int64_t, int64_t> vts;  // Dict{Int,Int} :133 -- lookup only, never iterated
  std::vector<V>* vtsAry;
  std::vector<int64_t>* fcs;

  // vertPos :42-55 ; x,y,z 1-based voxel index, pushes the converted vertex
  void vertPos(int e, int64_t x, int64_t y, int64_t z, const T* vals) {
    const uint8_t* ixs = REF_voxEdgeCrnrs[e - 1];
    T srcVal = vals[ixs[0] - 1], tgtVal = vals[ixs[1] - 1];
    auto q = (iso - srcVal) / (tgtVal - srcVal);
    using Q = decltype(q);
    using QE = Promote<Q, E>;
    using TE = Promote<T, E>;
    using A = Promote<QE, TE>;
    QE m = jl_max<QE>((QE)q, (QE)eps);
    TE hi = (TE)T(1) - (TE)eps;
    A a = jl_min<A>((A)m, (A)hi);
    auto b = (Promote<T, A>)T(1) - (Promote<T, A>)a;  // one(T) - a  :49
    const uint8_t* c1 = REF_voxCrnrPos[ixs[0] - 1];
    const uint8_t* c2 = REF_voxCrnrPos[ixs[1] - 1];
    P base[3] = {xp[x - 1], yp[y - 1], zp[z - 1]};
    P d[3] = {(P)(xp[x] - xp[x - 1]), (P)(yp[y] - yp[y - 1]), (P)(zp[z] - zp[z - 1])};
    for (int k = 0; k < 3; ++k) {
      // (c1 .* b .+ c2 .* a) : Int * Float multiplications then one add
      auto w = (decltype(b))((double)c1[k]) * b + (decltype(a))((double)c2[k]) * a;
      auto pos = base[k] + w * d[k];
      vtsAry->push_back((V)pos);
    }
  }
  // getVertId :67-84
  int64_t getVertId(int e, int64_t x, int64_t y, int64_t z, const T* vals) {
    int64_t vId = vertId(e, x, y, z, nx, ny);
    auto it = vts.find(vId);
    if (it != vts.end()) return it->second;
    vertPos(e, x, y, z, vals);
    int64_t n = (int64_t)(vtsAry->size() / 3);
    vts[vId] = n;
    return n;
  }
  // procVox :105-127
  void procVox(const T* vals, int64_t x, int64_t y, int64_t z, uint8_t cubeindex) {
    for (int i = 1; i <= 6; ++i) {
      uint8_t tIx = tetIx(i, cubeindex);
      if (tIx == 0x00 || tIx == 0x0f) continue;
      const uint8_t* e = REF_tetTri[tIx - 1];
      int64_t a = getVertId(voxEdgeId(i, e[0]), x, y, z, vals);
      int64_t b = getVertId(voxEdgeId(i, e[1]), x, y, z, vals);
      int64_t c = getVertId(voxEdgeId(i, e[2]), x, y, z, vals);
      fcs->push_back(a), fcs->push_back(b), fcs->push_back(c);
      if (e[3] == 0) continue;
      a = getVertId(voxEdgeId(i, e[3]), x, y, z, vals);
      b = getVertId(voxEdgeId(i, e[4]), x, y, z, vals);
      c = getVertId(voxEdgeId(i, e[5]), x, y, z, vals);
      fcs->push_back(a), fcs->push_back(b), fcs->push_back(c);
    }
  }
};

template <class T, class I, class E, class P, class V>
void run_mt(const T* sdf, int64_t nx, int64_t ny, int64_t nz, I iso, E eps, double x0, double x1, double y0,
            double y1, double z0, double z1, int nthreads, int64_t xlo, int64_t xhi, Result& out) {
  out.vert_is_f64 = std::is_same<V, double>::value;
  if (nx < 2 || ny < 2 || nz < 2) return;
  // [xlo, xhi): voxel x-planes to sweep (bench: a bounded sample of a large field); the whole volume is [0, nx-1)
  if (xlo < 0) xlo = 0;
  if (xhi < 0 || xhi > nx - 1) xhi = nx - 1;
  if (xhi <= xlo) return;
  auto sweep = [&](int64_t xa, int64_t xb, Result& r) {
    r.vert_is_f64 = out.vert_is_f64;
    MTState<T, I, E, P, V> st{sdf, nx, ny, nz, iso, eps, LinRange<P>(x0, x1, nx), LinRange<P>(y0, y1, ny),
                              LinRange<P>(z0, z1, nz), {}, &verts_of<V>(r), &r.faces};
    const int64_t sx = 1, sy = nx, sz = nx * ny;
    for (int64_t i = xa; i < xb; ++i)
      for (int64_t j = 0; j < ny - 1; ++j)
        for (int64_t k = 0; k < nz - 1; ++k) {
          const T* p = sdf + i * sx + j * sy + k * sz;
          // corner order :146-153
          T vals[8] = {p[0], p[sy], p[sx + sy], p[sx], p[sz], p[sy + sz], p[sx + sy + sz], p[sx + sz]};
          uint8_t c = get_cubeindex(vals, iso);
          if (no_triangles(c)) continue;
          st.procVox(vals, i + 1, j + 1, k + 1, c);
        }
  };
  if (nthreads <= 1) {
    sweep(xlo, xhi, out);
    return;
  }
  // x-slab threaded driver (bench CPU arm only, a THROUGHPUT measure): every thread sweeps its x-range with its own
  // vertex dictionary, so vertices on slab boundaries are duplicated -- not the reference's mesh, never used for parity.
  int64_t nvx = xhi - xlo;
  if (nthreads > nvx) nthreads = (int)nvx;
  std::vector<Result> parts(nthreads);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    int64_t xa = xlo + nvx * t / nthreads, xb = xlo + nvx * (t + 1) / nthreads;
    th.emplace_back([&, t, xa, xb] { sweep(xa, xb, parts[t]); });
  }
  for (auto& t : th) t.join();
  std::vector<V>& vts = verts_of<V>(out);
  for (int t = 0; t < nthreads; ++t) {
    int64_t base = (int64_t)(vts.size() / 3);
    std::vector<V>& pv = verts_of<V>(parts[t]);
    vts.insert(vts.end(), pv.begin(), pv.end());
    for (int64_t f : parts[t].faces) out.faces.push_back(f + base);
  }
}

// ---- type dispatch ------------------------------------------------------------------------------------
// range_kind: 0 = Int endpoints (LinRange{Float64}, contributes Int to promote_type),
//             1 = Float32 endpoints, 2 = Float64 endpoints
struct Args {
  int algo;  // 0 = MC, 1 = MT
  const void* sdf;
  int sdf_is_f64;
  int64_t nx, ny, nz;
  double iso;
  int iso_is_f32;
  double eps;
  int eps_is_f32;
  double x0, x1, y0, y1, z0, z1;
  int range_kind;
  int nthreads;
  int64_t xlo, xhi;  // voxel x-plane range of the sweep, -1 = whole volume (bench samples)
  int64_t x_off = 0, nx_glob = 0;  // MC only: the array is the slab [x_off, x_off + nx) of a volume with nx_glob samples along x
};

template <class T, class I, class E, class P>
void dispatch_v(const Args& a, Result& out) {
  // vertex element type: float(promote_type(range eltype, T, typeof(iso)[, typeof(eps)]))
  // src/marching_cubes.jl:31, src/marching_tetrahedra.jl:131
  bool f64 = std::is_same<T, double>::value || std::is_same<I, double>::value || a.range_kind == 2 ||
             (a.algo == 1 && std::is_same<E, double>::value);
  if (a.algo == 0) {
    if (f64)
      run_mc<T, I, P, double>((const T*)a.sdf, a.nx, a.ny, a.nz, (I)a.iso, a.x0, a.x1, a.y0, a.y1, a.z0, a.z1, a.nthreads, a.xlo, a.xhi, a.x_off, a.nx_glob, out);
    else
      run_mc<T, I, P, float>((const T*)a.sdf, a.nx, a.ny, a.nz, (I)a.iso, a.x0, a.x1, a.y0, a.y1, a.z0, a.z1, a.nthreads, a.xlo, a.xhi, a.x_off, a.nx_glob, out);
  } else {
    if (f64)
      run_mt<T, I, E, P, double>((const T*)a.sdf, a.nx, a.ny, a.nz, (I)a.iso, (E)a.eps, a.x0, a.x1, a.y0, a.y1, a.z0, a.z1, a.nthreads, a.xlo, a.xhi, out);
    else
      run_mt<T, I, E, P, float>((const T*)a.sdf, a.nx, a.ny, a.nz, (I)a.iso, (E)a.eps, a.x0, a.x1, a.y0, a.y1, a.z0, a.z1, a.nthreads, a.xlo, a.xhi, out);
  }
}
template <class T, class I, class E>
void dispatch_p(const Args& a, Result& out) {
  if (a.range_kind == 1) dispatch_v<T, I, E, float>(a, out);
  else dispatch_v<T, I, E, double>(a, out);
}
template <class T, class I>
void dispatch_e(const Args& a, Result& out) {
  if (a.algo == 1 && a.eps_is_f32) dispatch_p<T, I, float>(a, out);
  else dispatch_p<T, I, double>(a, out);
}
template <class T>
void dispatch_i(const Args& a, Result& out) {
  if (a.iso_is_f32) dispatch_e<T, float>(a, out);
  else dispatch_e<T, double>(a, out);
}

template <class T, class I>
void case_indices(int algo, const T* sdf, int64_t nx, int64_t ny, int64_t nz, I iso, uint8_t* out) {
  const int64_t sx = 1, sy = nx, sz = nx * ny;
  int64_t r = 0;
  for (int64_t i = 0; i < nx - 1; ++i)
    for (int64_t j = 0; j < ny - 1; ++j)
      for (int64_t k = 0; k < nz - 1; ++k) {
        const T* p = sdf + i * sx + j * sy + k * sz;
        if (algo == 0) {
          T vals[8] = {p[0], p[sx], p[sx + sy], p[sy], p[sz], p[sx + sz], p[sx + sy + sz], p[sy + sz]};
          out[r++] = get_cubeindex(vals, iso);
        } else {
          T vals[8] = {p[0], p[sy], p[sx + sy], p[sx], p[sz], p[sy + sz], p[sx + sy + sz], p[sx + sz]};
          out[r++] = get_cubeindex(vals, iso);
        }
      }
}

}  // namespace

extern "C" {

// Runs the restated isosurface(); returns an opaque result (free with oracle_free).
void* oracle_isosurface(int algo, const void* sdf, int sdf_is_f64, int64_t nx, int64_t ny, int64_t nz, double iso,
                        int iso_is_f32, double eps, int eps_is_f32, double x0, double x1, double y0, double y1,
                        double z0, double z1, int range_kind, int nthreads, int64_t xlo, int64_t xhi) {
  Args a{algo, sdf, sdf_is_f64, nx, ny, nz, iso, iso_is_f32, eps, eps_is_f32, x0, x1, y0, y1, z0, z1, range_kind, nthreads, xlo, xhi};
  Result* r = new Result();
  if (sdf_is_f64) dispatch_i<double>(a, *r);
  else dispatch_i<float>(a, *r);
  return r;
}
// Marching Cubes on the slab of samples [x_offset, x_offset + nx) of a volume with nx_global samples along x: vertex
// coordinates are those of the whole volume, face indices are relative to the slab's first vertex.
void* oracle_isosurface_slab(int algo, const void* sdf, int sdf_is_f64, int64_t nx, int64_t ny, int64_t nz, double iso,
                             int iso_is_f32, double eps, int eps_is_f32, double x0, double x1, double y0, double y1,
                             double z0, double z1, int range_kind, int nthreads, int64_t x_offset, int64_t nx_global) {
  if (algo != 0) return nullptr;  // MT shares vertices across voxels: no slab-wise restatement
  Args a{algo, sdf, sdf_is_f64, nx, ny, nz, iso, iso_is_f32, eps, eps_is_f32, x0, x1, y0, y1, z0, z1, range_kind, nthreads, -1, -1};
  a.x_off = x_offset, a.nx_glob = nx_global;
  Result* r = new Result();
  if (sdf_is_f64) dispatch_i<double>(a, *r);
  else dispatch_i<float>(a, *r);
  return r;
}
int64_t oracle_nverts(void* h) { return ((Result*)h)->nverts(); }
int64_t oracle_nfaces(void* h) { return (int64_t)(((Result*)h)->faces.size() / 3); }
int oracle_vert_is_f64(void* h) { return ((Result*)h)->vert_is_f64; }
void oracle_copy(void* h, void* verts, int64_t* faces) {
  Result* r = (Result*)h;
  if (verts) {
    if (r->vert_is_f64) memcpy(verts, r->v64.data(), r->v64.size() * sizeof(double));
    else memcpy(verts, r->v32.data(), r->v32.size() * sizeof(float));
  }
  if (faces) memcpy(faces, r->faces.data(), r->faces.size() * sizeof(int64_t));
}
void oracle_free(void* h) { delete (Result*)h; }

// Per-voxel case index in scan-rank order ((x*(ny-1)+y)*(nz-1)+z), algo's corner order.
void oracle_case_indices(int algo, const void* sdf, int sdf_is_f64, int64_t nx, int64_t ny, int64_t nz, double iso,
                         int iso_is_f32, uint8_t* out) {
  if (nx < 2 || ny < 2 || nz < 2) return;
  if (sdf_is_f64) {
    if (iso_is_f32) case_indices<double, float>(algo, (const double*)sdf, nx, ny, nz, (float)iso, out);
    else case_indices<double, double>(algo, (const double*)sdf, nx, ny, nz, iso, out);
  } else {
    if (iso_is_f32) case_indices<float, float>(algo, (const float*)sdf, nx, ny, nz, (float)iso, out);
    else case_indices<float, double>(algo, (const float*)sdf, nx, ny, nz, iso, out);
  }
}

// Known-answer hooks for test/runtests.jl:15-33
int oracle_get_cubeindex_f64(const double* vals, double iso) { return get_cubeindex(vals, iso); }
void oracle_vertex_interp_f64(double iso, const double* p1, const double* p2, double v1, double v2, double* out) {
  double mu = (iso - v1) / (v2 - v1);
  for (int q = 0; q < 3; ++q) out[q] = p1[q] + mu * (p2[q] - p1[q]);
}
double oracle_linrange_f64(double a, double b, int64_t n, int64_t i) { return LinRange<double>(a, b, n)[i]; }
float oracle_linrange_f32(double a, double b, int64_t n, int64_t i) { return LinRange<float>(a, b, n)[i]; }
}
