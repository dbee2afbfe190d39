// copy_probe.cu -- how fast can the host stage a pageable field into pinned chunks (and a mesh back out)?
// The pageable drop-in call is bound by these copies, so this measures the choices:
//   up:   T threads, each gathers row pieces (w bytes out of a pitch-byte row) into its 8 MB pinned chunk
//         with memcpy / with streaming (non-temporal) stores;
//   down: T threads copy pinned chunks into a pageable array that is fresh (first touch), pre-touched, or fresh with
//         MADV_HUGEPAGE.
// build: nvcc -O2 -o tools/bin/copy_probe tools/copy_probe.cu      run: tools/bin/copy_probe
#include <cuda_runtime.h>
#include <emmintrin.h>
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// dst 16-byte aligned, n multiple of 16
static void copy_nt(unsigned char* dst, const unsigned char* src, size_t n) {
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    const __m128i a = _mm_loadu_si128((const __m128i*)(src + i)), b = _mm_loadu_si128((const __m128i*)(src + i + 16));
    const __m128i c = _mm_loadu_si128((const __m128i*)(src + i + 32)), d = _mm_loadu_si128((const __m128i*)(src + i + 48));
    _mm_stream_si128((__m128i*)(dst + i), a), _mm_stream_si128((__m128i*)(dst + i + 16), b);
    _mm_stream_si128((__m128i*)(dst + i + 32), c), _mm_stream_si128((__m128i*)(dst + i + 48), d);
  }
  for (; i + 16 <= n; i += 16) _mm_stream_si128((__m128i*)(dst + i), _mm_loadu_si128((const __m128i*)(src + i)));
}

constexpr size_t CHUNK = (size_t)8 << 20;

// up: rows of `pitch` bytes, piece [xo, xo + w) of each row, gathered compactly (row pitch in the chunk = w)
static size_t PF_AHEAD = 4;
static double run_up(const unsigned char* src, size_t rows, size_t pitch, size_t xo, size_t w, int T, int nt, std::vector<unsigned char*>& pin) {
  std::vector<std::thread> th;
  const double t0 = now();
  for (int t = 0; t < T; ++t)
    th.emplace_back([=, &pin]() {
      const size_t r0 = rows * t / T, r1 = rows * (t + 1) / T, per = CHUNK / w;
      int tog = 0;
      for (size_t r = r0; r < r1; r += per) {
        const size_t n = std::min(per, r1 - r);
        unsigned char* b = pin[2 * t + tog];
        if (w == pitch) {
          if (nt) copy_nt(b, src + r * pitch, n * w);
          else memcpy(b, src + r * pitch, n * w);
        } else {
          for (size_t i = 0; i < n; ++i) {
            if (nt == 2 && i + PF_AHEAD < n) {
              const unsigned char* nx = src + (r + i + PF_AHEAD) * pitch + xo;
              for (size_t o = 0; o < w + 63; o += 64) _mm_prefetch((const char*)(nx + o), _MM_HINT_NTA);
            }
            if (nt) copy_nt(b + i * w, src + (r + i) * pitch + xo, w);
            else memcpy(b + i * w, src + (r + i) * pitch + xo, w);
          }
        }
        if (nt) _mm_sfence();
        tog ^= 1;
      }
    });
  for (auto& x : th) x.join();
  return now() - t0;
}

static double run_down(unsigned char* dst, size_t bytes, int T, bool nt, std::vector<unsigned char*>& pin) {
  std::vector<std::thread> th;
  const double t0 = now();
  for (int t = 0; t < T; ++t)
    th.emplace_back([=, &pin]() {
      const size_t a = bytes * t / T / 64 * 64, b = t == T - 1 ? bytes : bytes * (t + 1) / T / 64 * 64;
      int tog = 0;
      for (size_t off = a; off < b; off += CHUNK) {
        const size_t n = std::min(CHUNK, b - off);
        if (nt) copy_nt(dst + off, pin[2 * t + tog], n / 16 * 16);
        else memcpy(dst + off, pin[2 * t + tog], n);
        tog ^= 1;
      }
      if (nt) _mm_sfence();
    });
  for (auto& x : th) x.join();
  return now() - t0;
}

int main() {
  if (FILE* f = fopen("/proc/cpuinfo", "r")) {
    char line[256];
    while (fgets(line, sizeof line, f))
      if (!strncmp(line, "model name", 10)) {
        printf("cpu: %s", line);
        break;
      }
    fclose(f);
  }
  printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
  if (FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r")) {
    char line[128];
    if (fgets(line, sizeof line, f)) printf("thp: %s", line);
    fclose(f);
  }
  const size_t n = 1024, pitch = n * 4, rows = n * n, total = pitch * rows;
  unsigned char* src = (unsigned char*)malloc(total);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < 16; ++t)
      th.emplace_back([=]() {
        for (size_t i = total * t / 16; i < total * (t + 1) / 16; i += 4096) src[i] = (unsigned char)i;
      });
    for (auto& x : th) x.join();
  }
  const int TMAX = 32;
  std::vector<unsigned char*> pin(2 * TMAX);
  for (auto& p : pin) {
    if (cudaMallocHost((void**)&p, CHUNK) != cudaSuccess) return printf("cudaMallocHost failed\n"), 1;
    memset(p, 1, CHUNK);
  }
  printf("-- up: pageable field 4.29 GB -> pinned chunks (no DMA)\n");
  for (int rep = 0; rep < 1; ++rep)
    for (int T : {12, 16, 24})
      for (int nt = 0; nt < 3; ++nt) {
        const double whole = run_up(src, rows, pitch, 0, pitch, T, nt, pin);
        double slabs = 0, slabs2 = 0;  // 4 slabs of 1040-byte pieces (the 1024^3 default); 2 slabs of 2064
        for (int k = 0; k < 4; ++k) slabs += run_up(src, rows, pitch, (size_t)k * 1008, 1040, T, nt, pin);
        for (int k = 0; k < 2; ++k) slabs2 += run_up(src, rows, pitch, (size_t)k * 2032, 2064, T, nt, pin);
        printf("T=%2d %-9s whole rows %6.1f GB/s   4 slabs of 1040-B pieces %6.1f GB/s   2 slabs of 2064-B pieces %6.1f GB/s\n", T,
               nt == 2 ? "stream+pf" : nt ? "stream" : "memcpy", total / whole / 1e9, 4.0 * 1040 * rows / slabs / 1e9, 2.0 * 2064 * rows / slabs2 / 1e9);
      }
  for (size_t pf : {1, 2, 8, 16}) {
    PF_AHEAD = pf;
    double slabs = 0;
    for (int k = 0; k < 4; ++k) slabs += run_up(src, rows, pitch, (size_t)k * 1008, 1040, 16, 2, pin);
    printf("T=16 stream+pf ahead %zu rows: 4 slabs %6.1f GB/s\n", pf, 4.0 * 1040 * rows / slabs / 1e9);
  }
  printf("-- down: pinned chunks -> pageable mesh 0.97 GB\n");
  const size_t mesh = (size_t)974 << 20;
  for (int T : {8, 12})
    for (int nt = 0; nt < 2; ++nt) {
      unsigned char* fresh = (unsigned char*)mmap(nullptr, mesh, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      const double t_fresh = run_down(fresh, mesh, T, nt, pin);
      const double t_warm = run_down(fresh, mesh, T, nt, pin);
      munmap(fresh, mesh);
      unsigned char* huge = (unsigned char*)mmap(nullptr, mesh, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      madvise(huge, mesh, MADV_HUGEPAGE);
      const double t_huge = run_down(huge, mesh, T, nt, pin);
      munmap(huge, mesh);
      printf("T=%2d %-6s fresh %6.1f GB/s (%5.1f ms)  touched %6.1f GB/s  fresh+MADV_HUGEPAGE %6.1f GB/s (%5.1f ms)\n", T, nt ? "stream" : "memcpy",
             mesh / t_fresh / 1e9, t_fresh * 1e3, mesh / t_warm / 1e9, mesh / t_huge / 1e9, t_huge * 1e3);
    }
  return 0;
}
