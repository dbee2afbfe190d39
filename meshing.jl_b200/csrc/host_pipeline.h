// host_pipeline.h -- moving a HOST field in and a HOST mesh out at PCIe speed.
//
// The reference's caller owns ordinary (pageable) Julia arrays (examples/nrrd.jl:11-21).  A cudaMemcpy from pageable
// memory is staged by the driver through one thread and reaches ~11 GB/s on this box (55.6 GB/s pinned, measured with
// tools/pcie_probe.cu), so the copy, not the kernels, would decide the call's cost.  Both directions are therefore
// staged here through pinned ring buffers by a few worker threads: each worker owns a row range of the field (or a
// byte range of the mesh), two pinned chunks and a CUDA stream; it gathers the rows of x-slab k into a chunk with
// memcpy, enqueues the chunk's 2-D H2D copy and moves on, so host copies and DMA overlap.  Pinned caller arrays skip
// the staging (one worker, direct DMA).  The workers signal per x-slab, which is what lets b200iso_extract_host
// start slab k's kernels while slab k+1 is still arriving.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__linux__)
#include <sys/mman.h>
#endif

namespace hostpipe {

// Bytes per pinned staging chunk.  Measured at 1024^3 on the B200 host, whole pageable call, 24 workers
// (profiles/r2_host_chunks.txt): 256 KB 238 ms, 512 KB 171 ms, 1 MB 124 ms (the per-copy driver cost, ~7 us under the
// workers' contention, dominates), 2 MB 112-120 ms, 3 MB 119 ms, 4 MB 130 ms, 8 MB 120-136 ms.
constexpr size_t CHUNK = (size_t)2 << 20;

// experiment knobs, read per call: B200ISO_HOST_CHUNK_KB (<= 2048) and B200ISO_HOST_STREAM (bit 0: the upload's staging
// copies use streaming stores, bit 1: the download's).  Default 2: the upload's chunks are read back by the DMA engine
// while still in the last-level cache, so plain memcpy wins there (2 MB chunks: 112-120 ms against 120 streamed); the
// download writes the caller's array once and never reads it, so it streams (pre-touched target: 76 against 47 GB/s).
inline size_t chunk_bytes() {
  if (const char* e = getenv("B200ISO_HOST_CHUNK_KB")) return std::max<size_t>(64 << 10, std::min<size_t>(CHUNK, (size_t)atoll(e) << 10));
  return CHUNK;
}
inline int stream_copies() {
  const char* e = getenv("B200ISO_HOST_STREAM");
  return e ? atoi(e) : 2;
}

// Staging copy with streaming (non-temporal) stores: for a destination that is written once and not read by this core
// again, filling it through the cache only costs a read-for-ownership of every line.  Measured on the B200 host (16 vCPU
// Xeon, tools/copy_probe.cu, profiles/r2_copy_probe.txt; copies alone, no DMA), 12-16 threads: whole 4 KB rows 43-57 GB/s
// with memcpy, 76-80 GB/s streamed; 1040-byte row pieces (4 x-slabs) 41-45 against 48-55 GB/s.
// Callers end a batch of copies with copy_fence() before anything else may read the destination.
inline void copy_stream(unsigned char* dst, const unsigned char* src, size_t n) {
#if defined(__SSE2__)
  if (n >= 256) {
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (head) memcpy(dst, src, head), dst += head, src += head, n -= head;
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
      const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
      const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
      const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
      const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
    }
    for (; i + 16 <= n; i += 16) _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)));
    dst += i, src += i, n -= i;
  }
#endif
  if (n) memcpy(dst, src, n);
}
inline void copy_staged(unsigned char* dst, const unsigned char* src, size_t n, bool stream) {
  if (stream) copy_stream(dst, src, n);
  else memcpy(dst, src, n);
}
inline void copy_fence() {
#if defined(__SSE2__)
  _mm_sfence();
#endif
}

// A freshly allocated result array is first touched by the download workers: 250 000 page faults per GB with 4 KB pages
// (measured: 22 GB/s into fresh pages against 39 GB/s with transparent huge pages, 48-76 GB/s into touched ones).  Where
// the kernel leaves huge pages to madvise (the B200 host: "always [madvise] never") ask for them on the 2 MB-aligned
// interior of the caller's pageable output.  A hint only: contents and semantics are untouched; B200ISO_NO_HUGEPAGE=1 skips it.
inline void advise_huge(void* p, size_t bytes) {
#if defined(__linux__) && defined(MADV_HUGEPAGE)
  static const bool off = getenv("B200ISO_NO_HUGEPAGE") != nullptr;
  constexpr uintptr_t H = (uintptr_t)2 << 20;
  if (off || !p || bytes < 4 * H) return;
  const uintptr_t a = (reinterpret_cast<uintptr_t>(p) + H - 1) / H * H, b = (reinterpret_cast<uintptr_t>(p) + bytes) / H * H;
  if (b > a) madvise(reinterpret_cast<void*>(a), b - a, MADV_HUGEPAGE);
#else
  (void)p, (void)bytes;
#endif
}

inline bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// Per-handle pool of worker resources (streams, pinned chunks, events); created lazily, reused by every call.
struct Pool {
  int device = 0;
  struct Lane {
    cudaStream_t stream = nullptr;
    unsigned char* buf[2] = {nullptr, nullptr};
    cudaEvent_t buf_ev[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> slab_ev;  // one per x-slab, grown on demand
  };
  std::vector<Lane> in, out;

  static cudaError_t make_lane(Lane& l, bool with_buffers) {
    cudaError_t e = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
      if (with_buffers) e = cudaMallocHost((void**)&l.buf[b], CHUNK);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l.buf_ev[b], cudaEventDisableTiming);
    }
    return e;
  }
  static void free_lane(Lane& l) {
    for (int b = 0; b < 2; ++b) {
      if (l.buf[b]) cudaFreeHost(l.buf[b]);
      if (l.buf_ev[b]) cudaEventDestroy(l.buf_ev[b]);
    }
    for (cudaEvent_t e : l.slab_ev) cudaEventDestroy(e);
    if (l.stream) cudaStreamDestroy(l.stream);
    l = Lane{};
  }
  // lanes [0, n) of `v` exist, have `slabs` slab events and (if asked) their pinned chunks
  cudaError_t ensure(std::vector<Lane>& v, int n, int slabs, bool with_buffers) {
    cudaError_t e = cudaSuccess;
    while ((int)v.size() < n && e == cudaSuccess) {
      v.emplace_back();
      e = make_lane(v.back(), with_buffers);
    }
    for (int i = 0; i < n && e == cudaSuccess; ++i) {
      for (int b = 0; b < 2 && with_buffers && e == cudaSuccess; ++b)
        if (!v[i].buf[b]) e = cudaMallocHost((void**)&v[i].buf[b], CHUNK);
      while ((int)v[i].slab_ev.size() < slabs && e == cudaSuccess) {
        cudaEvent_t ev;
        e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e == cudaSuccess) v[i].slab_ev.push_back(ev);
      }
    }
    return e;
  }
  void release() {
    for (Lane& l : in) free_lane(l);
    for (Lane& l : out) free_lane(l);
    in.clear(), out.clear();
  }
};

// waiting threads sleep instead of spinning: the cores belong to the workers that are copying
inline void nap() { std::this_thread::sleep_for(std::chrono::microseconds(20)); }

// Staging threads of the upload (the download takes half as many).  The workers spend much of their time blocked in
// cudaEventSynchronize, so more threads than cores pays: measured on the 16-vCPU B200 host at 1024^3, whole call on
// pageable arrays: 8 threads 162 ms, 16 threads 150 ms, 24 threads 140 ms (tools/host_e2e.py, profiles/r2_host_paths.txt).
inline int default_threads() {
  if (const char* e = getenv("B200ISO_HOST_THREADS")) return std::max(1, std::min(64, atoi(e)));
  unsigned hw = std::thread::hardware_concurrency();
  // one process per GPU on a shared host (torchrun exports LOCAL_WORLD_SIZE): the ranks share the cores
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) hw = std::max(2u, hw / (unsigned)std::max(1, atoi(e)));
  return (int)std::max(1u, std::min(32u, hw + hw / 2));
}

// ---- field upload: x-slab k = sample planes [x0, x1) of every (y, z) row, into its own compact device array ------
// (every slab is stored compactly -- row pitch = its own padded width -- so that a staged chunk of gathered rows
// is ONE contiguous DMA: chunk-sized 2-D copies cost ~0.4 ms of serialised driver time per call, measured.)
struct Slab {
  unsigned char* dst = nullptr;  // device base of the slab
  size_t dpitch = 0;             // device row pitch, bytes (>= width)
  int64_t x0 = 0, x1 = 0;        // sample planes [x0, x1)
};

struct Uploader {
  Pool* pool = nullptr;
  int T = 0, S = 0;
  std::vector<Slab> slabs;
  std::vector<std::thread> threads;
  std::vector<std::atomic<int>> issued;  // per lane: slabs whose copies are enqueued
  std::atomic<int> err{(int)cudaSuccess};

  // src: host field with pitch spitch bytes; rows = ny * nz.
  cudaError_t start(Pool* p, const std::vector<Slab>& sl, const unsigned char* src, size_t spitch, size_t esz, size_t rows, cudaEvent_t after) {
    pool = p, slabs = sl, S = (int)sl.size();
    const bool pinned = is_pinned(src);
    T = pinned ? 1 : default_threads();
    if (cudaError_t e = pool->ensure(pool->in, T, S, !pinned)) return e;
    issued = std::vector<std::atomic<int>>(T);
    for (auto& a : issued) a.store(0);
    for (int t = 0; t < T; ++t)
      if (cudaError_t e = cudaStreamWaitEvent(pool->in[t].stream, after, 0)) return e;
    for (int t = 0; t < T; ++t) try {
      threads.emplace_back([=]() {
        Pool::Lane& L = pool->in[t];
        cudaError_t e = cudaSetDevice(pool->device);
        const size_t r0 = rows * t / T, r1 = rows * (t + 1) / T;
        const size_t chunk = chunk_bytes();
        const bool stream = (stream_copies() & 1) != 0;
        int tog = 0;
        for (int k = 0; k < S; ++k) {
          const Slab& sb = slabs[k];
          const size_t w = (size_t)(sb.x1 - sb.x0) * esz, xo = (size_t)sb.x0 * esz, dp = sb.dpitch;
          if (e == cudaSuccess && w > 0 && r1 > r0) {
            if (pinned) {
              e = (w == spitch && w == dp) ? cudaMemcpyAsync(sb.dst + r0 * dp, src + r0 * spitch, (r1 - r0) * w, cudaMemcpyHostToDevice, L.stream)
                                           : cudaMemcpy2DAsync(sb.dst + r0 * dp, dp, src + r0 * spitch + xo, spitch, w, r1 - r0,
                                                               cudaMemcpyHostToDevice, L.stream);
            } else if (dp > chunk) {
              // a row piece longer than a chunk (millions of samples along x): the row goes up in chunk-sized segments
              for (size_t r = r0; r < r1 && e == cudaSuccess; ++r)
                for (size_t off = 0; off < w && e == cudaSuccess; off += chunk) {
                  const size_t nb = std::min(chunk, w - off);
                  e = cudaEventSynchronize(L.buf_ev[tog]);
                  if (e != cudaSuccess) break;
                  copy_staged(L.buf[tog], src + r * spitch + xo + off, nb, stream);
                  copy_fence();
                  e = cudaMemcpyAsync(sb.dst + r * dp + off, L.buf[tog], nb, cudaMemcpyHostToDevice, L.stream);
                  if (e == cudaSuccess) e = cudaEventRecord(L.buf_ev[tog], L.stream);
                  tog ^= 1;
                }
            } else {
              const size_t per = chunk / dp;  // (>= 1)
              for (size_t r = r0; r < r1 && e == cudaSuccess; r += per) {
                const size_t n = std::min(per, r1 - r);
                e = cudaEventSynchronize(L.buf_ev[tog]);  // the chunk's previous DMA has drained
                if (e != cudaSuccess) break;
                unsigned char* b = L.buf[tog];
                if (w == spitch && w == dp) copy_staged(b, src + r * spitch, n * w, stream);
                else
                  // every staged row is written over its whole pitch, pad included (the pad is never read as samples): a row
                  // that ends inside a cache line would leave the line to be flushed half-written and finished by the next
                  // row (measured: 8 slabs of 516-byte rows 345 ms per call against 140 with memcpy).  The few bytes past
                  // the piece come from the same source row or the next one; the array's last row is copied exactly.
                  for (size_t i = 0; i < n; ++i) copy_staged(b + i * dp, src + (r + i) * spitch + xo, r + i + 1 < rows ? dp : w, stream);
                copy_fence();  // the streamed lines are in memory before the DMA is told to read them
                e = cudaMemcpyAsync(sb.dst + r * dp, b, n * dp, cudaMemcpyHostToDevice, L.stream);
                if (e == cudaSuccess) e = cudaEventRecord(L.buf_ev[tog], L.stream);
                tog ^= 1;
              }
            }
          }
          if (e == cudaSuccess) e = cudaEventRecord(L.slab_ev[k], L.stream);
          if (e != cudaSuccess) err.store((int)e);
          issued[t].store(k + 1, std::memory_order_release);  // always advances: the consumer never waits forever
        }
      });
    } catch (...) {  // thread creation failed: no exception may cross the C ABI; unstarted lanes count as failed
      err.store((int)cudaErrorUnknown);
      for (int u = t; u < T; ++u) issued[u].store(S, std::memory_order_release);
      return cudaErrorUnknown;
    }
    return cudaSuccess;
  }
  // makes `stream` wait for x-slab k (blocks the host only until the slab's copies are ENQUEUED)
  cudaError_t wait_slab(int k, cudaStream_t stream) {
    for (int t = 0; t < T; ++t) {
      while (issued[t].load(std::memory_order_acquire) <= k) nap();
      if (err.load() != (int)cudaSuccess) return (cudaError_t)err.load();
      if (cudaError_t e = cudaStreamWaitEvent(stream, pool->in[t].slab_ev[k], 0)) return e;
    }
    return cudaSuccess;
  }
  void join() {
    for (auto& th : threads)
      if (th.joinable()) th.join();
    threads.clear();
  }
  ~Uploader() { join(); }
};

// ---- mesh download: job k = two device byte ranges (vertices, faces) to their final host offsets ---------------
struct Downloader {
  struct Job {
    const unsigned char* src[2] = {nullptr, nullptr};
    unsigned char* dst[2] = {nullptr, nullptr};
    size_t bytes[2] = {0, 0};
    cudaEvent_t ready = nullptr;  // recorded after the kernels that produce src
  };
  Pool* pool = nullptr;
  int T = 0;
  bool pinned = true;
  std::vector<Job> jobs;
  std::atomic<int> njobs{0};
  std::atomic<bool> closed{false};
  std::vector<std::atomic<int>> drained;  // per lane: jobs fully in host memory (pageable) / enqueued (pinned)
  std::vector<std::thread> threads;
  std::atomic<int> err{(int)cudaSuccess};

  cudaError_t start(Pool* p, int max_jobs, bool dst_pinned, cudaEvent_t after) {
    pool = p, pinned = dst_pinned;
    T = pinned ? 1 : std::max(1, default_threads() / 2);
    if (const char* e = getenv("B200ISO_HOST_DOWN_THREADS"))
      if (!pinned) T = std::max(1, std::min(64, atoi(e)));
    if (cudaError_t e = pool->ensure(pool->out, T, max_jobs, !pinned)) return e;
    jobs.assign(max_jobs, Job{});
    drained = std::vector<std::atomic<int>>(T);
    for (auto& a : drained) a.store(0);
    for (int t = 0; t < T; ++t)
      if (cudaError_t e = cudaStreamWaitEvent(pool->out[t].stream, after, 0)) return e;
    if (pinned) return cudaSuccess;  // direct DMA from the calling thread, no workers
    for (int t = 0; t < T; ++t) try {
      threads.emplace_back([=]() {
        Pool::Lane& L = pool->out[t];
        cudaError_t e = cudaSetDevice(pool->device);
        const size_t chunk = chunk_bytes();
        const bool stream = (stream_copies() & 2) != 0;
        for (int k = 0;; ++k) {
          while (njobs.load(std::memory_order_acquire) <= k && !closed.load(std::memory_order_acquire)) nap();
          if (njobs.load(std::memory_order_acquire) <= k) break;
          const Job& j = jobs[k];
          if (e == cudaSuccess) e = cudaStreamWaitEvent(L.stream, j.ready, 0);
          for (int part = 0; part < 2 && e == cudaSuccess; ++part) {
            // this lane's byte range of the part, in chunks; DMA of chunk i overlaps the host copy of chunk i - 1
            const size_t a = j.bytes[part] * t / T / 16 * 16, b = t == T - 1 ? j.bytes[part] : j.bytes[part] * (t + 1) / T / 16 * 16;
            size_t pend_off = 0, pend_n = 0;
            int pend_buf = -1, tog = 0;
            for (size_t off = a; off < b && e == cudaSuccess; off += chunk) {
              const size_t n = std::min(chunk, b - off);
              e = cudaMemcpyAsync(L.buf[tog], j.src[part] + off, n, cudaMemcpyDeviceToHost, L.stream);
              if (e == cudaSuccess) e = cudaEventRecord(L.buf_ev[tog], L.stream);
              if (pend_buf >= 0 && e == cudaSuccess) {
                e = cudaEventSynchronize(L.buf_ev[pend_buf]);
                if (e == cudaSuccess) copy_staged(j.dst[part] + pend_off, L.buf[pend_buf], pend_n, stream);
              }
              pend_buf = tog, pend_off = off, pend_n = n, tog ^= 1;
            }
            if (pend_buf >= 0 && e == cudaSuccess) {
              e = cudaEventSynchronize(L.buf_ev[pend_buf]);
              if (e == cudaSuccess) copy_staged(j.dst[part] + pend_off, L.buf[pend_buf], pend_n, stream);
            }
          }
          copy_fence();
          if (e != cudaSuccess) err.store((int)e);
          drained[t].store(k + 1, std::memory_order_release);
        }
      });
    } catch (...) {  // thread creation failed: stop the lanes that did start
      err.store((int)cudaErrorUnknown);
      closed.store(true, std::memory_order_release);
      return cudaErrorUnknown;
    }
    return cudaSuccess;
  }
  // enqueue job k (jobs are pushed in order, k = 0, 1, ...)
  cudaError_t push(int k, const Job& j) {
    jobs[k] = j;
    if (pinned) {
      Pool::Lane& L = pool->out[0];
      if (cudaError_t e = cudaStreamWaitEvent(L.stream, j.ready, 0)) return e;
      for (int part = 0; part < 2; ++part)
        if (j.bytes[part])
          if (cudaError_t e = cudaMemcpyAsync(j.dst[part], j.src[part], j.bytes[part], cudaMemcpyDeviceToHost, L.stream)) return e;
      if (cudaError_t e = cudaEventRecord(L.slab_ev[k], L.stream)) return e;
    }
    njobs.store(k + 1, std::memory_order_release);
    return cudaSuccess;
  }
  // blocks the host until job k has left its device staging (so the staging may be overwritten)
  cudaError_t wait_job(int k) {
    if (pinned) return cudaEventSynchronize(pool->out[0].slab_ev[k]);
    for (int t = 0; t < T; ++t)
      while (drained[t].load(std::memory_order_acquire) <= k) nap();
    return (cudaError_t)err.load();
  }
  // all pushed jobs are in host memory on return
  cudaError_t finish() {
    closed.store(true, std::memory_order_release);
    for (auto& th : threads)
      if (th.joinable()) th.join();
    threads.clear();
    if (!pool || pool->out.empty()) return cudaSuccess;  // never started
    if (pinned) return cudaStreamSynchronize(pool->out[0].stream);
    return (cudaError_t)err.load();
  }
  ~Downloader() { finish(); }
};

}  // namespace hostpipe
