// Development tool: what the host can copy per second between pageable and pinned memory with T threads (memcpy vs
// non-temporal stores), i.e. the ceiling of the bounce stage of the host pipeline (cabi.cu: CopyPool).  Not product code.
//   nvcc -O2 -o scripts/_build/host_copy_probe scripts/host_copy_probe.cu -Xcompiler -pthread,-mavx2
#include <cuda_runtime.h>
#include <immintrin.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void ntCopy(uint8_t *dst, const uint8_t *src, size_t bytes) {
  size_t i = 0;
  for (; i + 128 <= bytes; i += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64)), d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
    _mm256_stream_si256((__m256i *)(dst + i), a); _mm256_stream_si256((__m256i *)(dst + i + 32), b);
    _mm256_stream_si256((__m256i *)(dst + i + 64), c); _mm256_stream_si256((__m256i *)(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < bytes) memcpy(dst + i, src + i, bytes - i);
}

static double run(uint8_t *dst, const uint8_t *src, size_t bytes, int threads, bool nt, size_t piece) {
  // the pool's shape: pieces handed out round-robin
  const size_t pieces = (bytes + piece - 1) / piece;
  double best = 1e9;
  for (int rep = 0; rep < 3; rep++) {
    const double t0 = now();
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
      th.emplace_back([=] {
        for (size_t p = t; p < pieces; p += threads) {
          const size_t off = p * piece, len = std::min(piece, bytes - off);
          if (nt) ntCopy(dst + off, src + off, len); else memcpy(dst + off, src + off, len);
        }
      });
    for (auto &x : th) x.join();
    best = std::min(best, now() - t0);
  }
  return bytes / best / 1e9;
}

int main(int argc, char **argv) {
  const size_t bytes = (size_t)(argc > 1 ? atof(argv[1]) : 1e9);
  uint8_t *pageable = (uint8_t *)aligned_alloc(4096, bytes), *pageable2 = (uint8_t *)aligned_alloc(4096, bytes), *pinned = nullptr;
  memset(pageable, 1, bytes);
  memset(pageable2, 2, bytes);
  if (cudaHostAlloc(&pinned, bytes, cudaHostAllocDefault) != cudaSuccess) { printf("no pinned memory\n"); return 1; }
  memset(pinned, 3, bytes);
  for (int threads : {1, 2, 4, 8, 12, 16})
    for (int nt = 0; nt < 2; nt++) {
      printf("{\"threads\": %d, \"stores\": \"%s\", \"pageable_to_pinned_gbs\": %.1f, \"pinned_to_pageable_gbs\": %.1f, \"pageable_to_pageable_gbs\": %.1f, \"piece_1MiB_to_pinned_gbs\": %.1f}\n",
             threads, nt ? "non-temporal" : "memcpy", run(pinned, pageable, bytes, threads, nt, 4 << 20), run(pageable, pinned, bytes, threads, nt, 4 << 20),
             run(pageable2, pageable, bytes, threads, nt, 4 << 20), run(pinned, pageable, bytes, threads, nt, 1 << 20));
      fflush(stdout);
    }
  // the driver's own staging of a pageable copy, and the bare pinned copy, for scale
  void *d = nullptr;
  cudaMalloc(&d, bytes);
  for (int k = 0; k < 2; k++) {
    double t0 = now();
    cudaMemcpy(d, pageable, bytes, cudaMemcpyHostToDevice);
    double t1 = now();
    cudaMemcpy(d, pinned, bytes, cudaMemcpyHostToDevice);
    double t2 = now();
    printf("{\"cudaMemcpy_h2d_pageable_gbs\": %.1f, \"cudaMemcpy_h2d_pinned_gbs\": %.1f}\n", bytes / (t1 - t0) / 1e9, bytes / (t2 - t1) / 1e9);
  }
  return 0;
}
