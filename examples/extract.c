/* examples/extract.c -- the C ABI of libb200iso.so from plain C (what a Julia ccall / cgo / JNI binding does).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/extract.c -o extract -Lmeshing.jl_b200/lib -lb200iso -lm \
 *       -Wl,-rpath,$PWD/meshing.jl_b200/lib
 *   ./extract 256            # Marching Cubes on a 256^3 sphere SDF, host arrays in and out
 *
 * Two-phase form (count -> the caller sizes its arrays -> generate), then the one-shot slab-pipelined form with
 * the now-known capacity; both must give the same mesh. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200iso.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != B200ISO_OK) {                                                     \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, b200iso_last_error());       \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 128;
  float* sdf = (float*)malloc((size_t)n * n * n * sizeof(float)); /* sdf[x + n*(y + n*z)], like a Julia Array */
  if (!sdf) return 1;
  for (int64_t z = 0; z < n; ++z)
    for (int64_t y = 0; y < n; ++y)
      for (int64_t x = 0; x < n; ++x) {
        const double px = -1.0 + 2.0 * x / (n - 1), py = -1.0 + 2.0 * y / (n - 1), pz = -1.0 + 2.0 * z / (n - 1);
        sdf[x + n * (y + n * z)] = (float)(sqrt(px * px + py * py + pz * pz) - 0.5);
      }

  b200iso_params p;
  memset(&p, 0, sizeof p);
  p.algo = B200ISO_MC;
  p.iso = 0.0, p.iso_is_f32 = 1;        /* MarchingCubes(iso=0f0) */
  p.eps = 1e-3, p.eps_is_f32 = 1;
  p.range_kind = B200ISO_RANGE_INT;     /* X = Y = Z = -1:1 */
  p.x0 = p.y0 = p.z0 = -1.0, p.x1 = p.y1 = p.z1 = 1.0;

  b200iso_handle* h = NULL;
  CHECK(b200iso_create(&h, 0));

  int64_t nv = 0, nf = 0;
  int f64 = 0;
  CHECK(b200iso_count(h, &p, sdf, B200ISO_HOST, n, n, n, n, &nv, &nf, &f64));
  float* verts = (float*)malloc((size_t)(nv > 0 ? nv : 1) * 3 * sizeof(float)); /* f64 == 0 for this call */
  int64_t* faces = (int64_t*)malloc((size_t)(nf > 0 ? nf : 1) * 3 * sizeof(int64_t));
  CHECK(b200iso_generate(h, verts, faces, B200ISO_HOST, 0));
  printf("two-phase : %lld vertices, %lld faces (vertex type %s)\n", (long long)nv, (long long)nf, f64 ? "Float64" : "Float32");

  float* verts2 = (float*)malloc((size_t)(nv > 0 ? nv : 1) * 3 * sizeof(float));
  int64_t* faces2 = (int64_t*)malloc((size_t)(nf > 0 ? nf : 1) * 3 * sizeof(int64_t));
  int64_t nv2 = 0, nf2 = 0;
  CHECK(b200iso_extract_host(h, &p, sdf, n, n, n, n, verts2, nv, faces2, nf, &nv2, &nf2, &f64));
  const int same = nv2 == nv && nf2 == nf && memcmp(verts, verts2, (size_t)nv * 12) == 0 && memcmp(faces, faces2, (size_t)nf * 24) == 0;
  printf("one-shot  : %lld vertices, %lld faces, %s\n", (long long)nv2, (long long)nf2, same ? "identical bytes" : "DIFFERENT");
  if (nf > 0) printf("first face: (%lld, %lld, %lld), 1-based\n", (long long)faces[0], (long long)faces[1], (long long)faces[2]);

  b200iso_destroy(h);
  free(sdf), free(verts), free(faces), free(verts2), free(faces2);
  return same ? 0 : 2;
}
