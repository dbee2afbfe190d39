// Development probe: PCIe H2D throughput of x-slab (strided 2D) copies vs one contiguous copy, and H2D || D2H overlap.
// Build: nvcc -O2 -o tools/bin/pcie_probe tools/pcie_probe.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main(int argc, char** argv) {
  const size_t n = argc > 1 ? atoi(argv[1]) : 1024;
  const size_t bytes = n * n * n * 4, obytes = bytes / 4;
  float *h, *d; char *ho, *dout;
  CK(cudaMallocHost(&h, bytes)); CK(cudaMalloc(&d, bytes));
  CK(cudaMallocHost(&ho, obytes)); CK(cudaMalloc(&dout, obytes));
  for (size_t i = 0; i < bytes / 4; i += 1024) h[i] = 1.f;
  cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
  cudaEvent_t e0, e1, f0, f1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&f0); cudaEventCreate(&f1);
  float ms, ms2;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0, s1)); CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s1)); CK(cudaEventRecord(e1, s1));
    CK(cudaStreamSynchronize(s1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("contiguous H2D %.1f MB: %.2f ms  %.1f GB/s\n", bytes / 1e6, ms, bytes / ms / 1e6);
  }
  CK(cudaEventRecord(e0, s1)); CK(cudaMemcpyAsync(ho, dout, obytes, cudaMemcpyDeviceToHost, s1)); CK(cudaEventRecord(e1, s1));
  CK(cudaStreamSynchronize(s1)); cudaEventElapsedTime(&ms, e0, e1);
  printf("contiguous D2H %.1f MB: %.2f ms  %.1f GB/s\n", obytes / 1e6, ms, obytes / ms / 1e6);
  for (size_t w : {32, 64, 128, 256, 512}) {
    if (w > n) continue;
    CK(cudaEventRecord(e0, s1));
    for (size_t x0 = 0; x0 < n; x0 += w)
      CK(cudaMemcpy2DAsync(d + x0, n * 4, h + x0, n * 4, w * 4, n * n, cudaMemcpyHostToDevice, s1));
    CK(cudaEventRecord(e1, s1)); CK(cudaStreamSynchronize(s1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("x-slab 2D H2D, slab width %zu samples (%zu B rows), %zu slabs: %.2f ms  %.1f GB/s\n", w, w * 4, n / w, ms, bytes / ms / 1e6);
  }
  // device-side compaction alternative: contiguous z-chunk copies are what a z-slab pipeline would use
  // bidirectional: full H2D on s1 while D2H of obytes on s2
  CK(cudaEventRecord(e0, s1)); CK(cudaEventRecord(f0, s2));
  CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s1));
  CK(cudaMemcpyAsync(ho, dout, obytes, cudaMemcpyDeviceToHost, s2));
  CK(cudaEventRecord(e1, s1)); CK(cudaEventRecord(f1, s2));
  CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); cudaEventElapsedTime(&ms2, f0, f1);
  printf("concurrent: H2D %.2f ms (%.1f GB/s)  D2H %.2f ms (%.1f GB/s)\n", ms, bytes / ms / 1e6, ms2, obytes / ms2 / 1e6);
  // 2D slabs + concurrent D2H
  {
    const size_t w = 128;
    CK(cudaEventRecord(e0, s1)); CK(cudaEventRecord(f0, s2));
    for (size_t x0 = 0; x0 < n; x0 += w)
      CK(cudaMemcpy2DAsync(d + x0, n * 4, h + x0, n * 4, w * 4, n * n, cudaMemcpyHostToDevice, s1));
    CK(cudaMemcpyAsync(ho, dout, obytes, cudaMemcpyDeviceToHost, s2));
    CK(cudaEventRecord(e1, s1)); CK(cudaEventRecord(f1, s2));
    CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); cudaEventElapsedTime(&ms2, f0, f1);
    printf("concurrent 2D(w=128): H2D %.2f ms (%.1f GB/s)  D2H %.2f ms (%.1f GB/s)\n", ms, bytes / ms / 1e6, ms2, obytes / ms2 / 1e6);
  }
  // pageable source
  {
    float* hp = (float*)malloc(bytes);
    for (size_t i = 0; i < bytes / 4; i += 1024) hp[i] = 1.f;
    CK(cudaEventRecord(e0, s1)); CK(cudaMemcpyAsync(d, hp, bytes, cudaMemcpyHostToDevice, s1)); CK(cudaEventRecord(e1, s1));
    CK(cudaStreamSynchronize(s1)); cudaEventElapsedTime(&ms, e0, e1);
    printf("pageable contiguous H2D: %.2f ms  %.1f GB/s\n", ms, bytes / ms / 1e6);
    free(hp);
  }
  return 0;
}
