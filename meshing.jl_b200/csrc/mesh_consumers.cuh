// mesh_consumers.cuh -- what callers run next on the extracted mesh (not part of the reference):
//   * per-vertex normals from the gradient of the field (trilinear blend of central differences),
//   * welding of the Marching Cubes mesh into an indexed mesh: Meshing.jl's Marching Cubes repeats the vertex of a
//     grid edge in every voxel that touches it ("vertices may be repeated", src/algorithmtypes.jl:18-19); the welded
//     form keeps the FIRST occurrence in the reference's scan order and renumbers the faces.
// Vertex identity for the weld is the geometric grid edge (lower end node, axis), produced by the generate kernel's
// KEYS instantiation from the same records -- never coordinates: two voxels interpolate a shared edge in opposite
// directions, so their copies may differ in the last bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace iso {

struct NormalArgs {
  const void* sdf;       // Float32 or Float64 field, x contiguous
  int field_is_f64;
  int nx, ny, nz;        // samples of the array passed (a slab: nx local)
  long long ldx, plane;
  double x0, x1, y0, y1, z0, z1;  // extents of the WHOLE volume
  long long x_offset, nx_global;  // slab position along x (0 / nx for a whole volume)
  const void* verts;
  int vert_is_f64;
  long long nverts;
  float* normals;        // out: float[3 * nverts], unit length (0,0,0 where the gradient vanishes)
};

template <class T>
__device__ __forceinline__ double fld(const T* f, const NormalArgs& a, int x, int y, int z) {
  return (double)__ldg(f + x + a.ldx * (long long)y + a.plane * (long long)z);
}

// gradient at node (x, y, z): central differences, one-sided at the borders of the array; hx/hy/hz = grid spacing
template <class T>
__device__ __forceinline__ void node_gradient(const T* f, const NormalArgs& a, int x, int y, int z, double ihx, double ihy, double ihz,
                                              double g[3]) {
  const int xm = x > 0 ? x - 1 : x, xp = x < a.nx - 1 ? x + 1 : x;
  const int ym = y > 0 ? y - 1 : y, yp = y < a.ny - 1 ? y + 1 : y;
  const int zm = z > 0 ? z - 1 : z, zp = z < a.nz - 1 ? z + 1 : z;
  g[0] = xp > xm ? (fld(f, a, xp, y, z) - fld(f, a, xm, y, z)) * ihx / (double)(xp - xm) : 0.0;
  g[1] = yp > ym ? (fld(f, a, x, yp, z) - fld(f, a, x, ym, z)) * ihy / (double)(yp - ym) : 0.0;
  g[2] = zp > zm ? (fld(f, a, x, y, zp) - fld(f, a, x, y, zm)) * ihz / (double)(zp - zm) : 0.0;
}

template <class T, class V>
__global__ void __launch_bounds__(256)
vertex_normals_kernel(NormalArgs a) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nverts) return;
  const T* f = reinterpret_cast<const T*>(a.sdf);
  const V* v = reinterpret_cast<const V*>(a.verts) + 3 * i;
  const double p[3] = {(double)v[0], (double)v[1], (double)v[2]};
  const double lo[3] = {a.x0, a.y0, a.z0}, hi[3] = {a.x1, a.y1, a.z1};
  const long long n[3] = {a.nx_global, a.ny, a.nz};
  const int nl[3] = {a.nx, a.ny, a.nz};
  int c[3];
  double t[3], ih[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const double span = hi[q] - lo[q];
    const double h = n[q] > 1 && span != 0.0 ? span / (double)(n[q] - 1) : 1.0;
    ih[q] = 1.0 / h;
    double gq = (p[q] - lo[q]) * ih[q] - (q == 0 ? (double)a.x_offset : 0.0);  // continuous index in the local array
    // A vertex within 1e-4 of a grid plane takes that plane's gradients: isosurface vertices sit on grid edges, i.e. ON
    // two of the three families of planes up to the rounding of their coordinates, and the 4 or 6 nodes whose weight
    // would be ~1e-7 cost 6 loads each (measured at 1024^3, 40.6 M vertices: 3.09 ms with them, 1.91 ms without).
    const double rq = rint(gq);
    if (fabs(gq - rq) < 1e-4) gq = rq;
    if (!(gq >= 0.0)) gq = 0.0;                                              // (also catches NaN)
    if (gq > (double)(nl[q] - 1)) gq = (double)(nl[q] - 1);
    int ci = (int)gq;
    if (ci > nl[q] - 2) ci = nl[q] - 2 > 0 ? nl[q] - 2 : 0;
    c[q] = ci, t[q] = gq - (double)ci;
  }
  double g[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
    const int x = min(c[0] + dx, a.nx - 1), y = min(c[1] + dy, a.ny - 1), z = min(c[2] + dz, a.nz - 1);
    const double w = (dx ? t[0] : 1.0 - t[0]) * (dy ? t[1] : 1.0 - t[1]) * (dz ? t[2] : 1.0 - t[2]);
    if (w != 0.0) {
      double gn[3];
      node_gradient(f, a, x, y, z, ih[0], ih[1], ih[2], gn);
      g[0] += w * gn[0], g[1] += w * gn[1], g[2] += w * gn[2];
    }
  }
  const double len = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  float* o = a.normals + 3 * i;
  if (len > 0.0 && len < 1e300) o[0] = (float)(g[0] / len), o[1] = (float)(g[1] / len), o[2] = (float)(g[2] / len);
  else o[0] = o[1] = o[2] = 0.f;
}

// ---- weld -------------------------------------------------------------------------------------------------------
// Open-addressing hash table keyed by the edge key: slot -> (key + 1, smallest vertex index with that key).
__device__ __forceinline__ unsigned long long weld_hash(unsigned long long k) {
  k ^= k >> 33, k *= 0xff51afd7ed558ccdull, k ^= k >> 33, k *= 0xc4ceb9fe1a85ec53ull, k ^= k >> 33;
  return k;
}

__global__ void weld_insert_kernel(const long long* __restrict__ keys, long long n, unsigned long long* tab_key, long long* tab_min,
                                   unsigned long long mask, unsigned int* __restrict__ slot_of) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = (unsigned long long)keys[i] + 1ull;  // 0 = empty slot
  unsigned long long s = weld_hash(k) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(tab_key + s, 0ull, k);
    if (prev == 0ull || prev == k) break;
    s = (s + 1) & mask;
  }
  atomicMin(reinterpret_cast<unsigned long long*>(tab_min) + s, (unsigned long long)i);
  slot_of[i] = (unsigned int)s;
}

// keep[i] = 1 iff vertex i is the first occurrence of its edge
__global__ void weld_flag_kernel(const unsigned int* __restrict__ slot_of, const long long* __restrict__ tab_min, long long n,
                                 unsigned int* __restrict__ keep) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = tab_min[slot_of[i]] == i ? 1u : 0u;
}

// newidx = exclusive scan of keep (done by the host with cub); compacts the kept vertices and leaves remap[i]
template <class V>
__global__ void weld_compact_kernel(const V* __restrict__ verts, const unsigned int* __restrict__ keep, const unsigned int* __restrict__ newidx,
                                    long long n, V* __restrict__ verts_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  const long long j = newidx[i];
  verts_out[3 * j] = verts[3 * i], verts_out[3 * j + 1] = verts[3 * i + 1], verts_out[3 * j + 2] = verts[3 * i + 2];
}

// faces_out = 1 + new index of the representative of (faces - base - 1); base = vertex base of the faces' indices
__global__ void weld_faces_kernel(const long long* __restrict__ faces, long long n3, long long base, const unsigned int* __restrict__ slot_of,
                                  const long long* __restrict__ tab_min, const unsigned int* __restrict__ newidx,
                                  long long* __restrict__ faces_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const long long v = faces[i] - base - 1;
  faces_out[i] = (long long)newidx[tab_min[slot_of[v]]] + 1;
}

}  // namespace iso
