// Development tool: two shapes of the pageable <-> device bounce stage, measured bare (no kernels).
//   ranges : what cabi.cu's runPipeline does -- T threads copy a whole 60 MB range into one of 3 pinned buffers, then one
//            cudaMemcpyAsync per range; the copy of range c overlaps the DMA of ranges c-1, c-2.
//   workers: T independent mini-pipelines -- each worker owns a stream and S pinned slots of P bytes, copies its pieces
//            (k = t, t + T, ...) into a free slot and issues that piece's cudaMemcpyAsync itself; the staging footprint
//            (T * S * P) can stay inside the L3, so the DMA engine can read what the cores just wrote.
//   nvcc -O2 -o scripts/_build/bounce_probe scripts/bounce_probe.cu -Xcompiler -pthread,-mavx2
#include <cuda_runtime.h>
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
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
static void copyBytes(uint8_t *dst, const uint8_t *src, size_t n, bool nt) { if (nt) ntCopy(dst, src, n); else memcpy(dst, src, n); }

// ---- ranges -------------------------------------------------------------------------------------------------
static double rangesH2D(uint8_t *dev, const uint8_t *src, size_t bytes, int T, bool nt, size_t range) {
  uint8_t *pin[3];
  cudaStream_t st[3];
  cudaEvent_t ev[3];
  for (int s = 0; s < 3; s++) { CK(cudaHostAlloc(&pin[s], range, 0)); memset(pin[s], 0, range); CK(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking)); CK(cudaEventCreate(&ev[s])); }
  const size_t ranges = (bytes + range - 1) / range, piece = 4 << 20;
  const double t0 = now();
  for (size_t c = 0; c < ranges; c++) {
    const int s = c % 3;
    if (c >= 3) CK(cudaEventSynchronize(ev[s]));
    const size_t off = c * range, len = std::min(range, bytes - off), pieces = (len + piece - 1) / piece;
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
      th.emplace_back([=] { for (size_t p = t; p < pieces; p += T) copyBytes(pin[s] + p * piece, src + off + p * piece, std::min(piece, len - p * piece), nt); });
    for (auto &x : th) x.join();
    CK(cudaMemcpyAsync(dev + off, pin[s], len, cudaMemcpyHostToDevice, st[s]));
    CK(cudaEventRecord(ev[s], st[s]));
  }
  for (int s = 0; s < 3; s++) CK(cudaStreamSynchronize(st[s]));
  const double dt = now() - t0;
  for (int s = 0; s < 3; s++) { cudaFreeHost(pin[s]); cudaStreamDestroy(st[s]); cudaEventDestroy(ev[s]); }
  return bytes / dt / 1e9;
}

// ---- workers ------------------------------------------------------------------------------------------------
struct Worker { cudaStream_t st; std::vector<uint8_t *> slot; std::vector<cudaEvent_t> ev; };

static double workers(uint8_t *dev, uint8_t *host, size_t bytes, int T, int S, size_t P, bool nt, bool toDevice, std::vector<Worker> &w) {
  const size_t pieces = (bytes + P - 1) / P;
  const double t0 = now();
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++)
    th.emplace_back([&, t] {
      Worker &me = w[t];
      size_t round = 0;
      if (toDevice) {
        for (size_t k = t; k < pieces; k += T, round++) {
          const int s = round % S;
          if (round >= (size_t)S) CK(cudaEventSynchronize(me.ev[s]));
          const size_t off = k * P, len = std::min(P, bytes - off);
          copyBytes(me.slot[s], host + off, len, nt);
          CK(cudaMemcpyAsync(dev + off, me.slot[s], len, cudaMemcpyHostToDevice, me.st));
          CK(cudaEventRecord(me.ev[s], me.st));
        }
        CK(cudaStreamSynchronize(me.st));
      } else {
        // D2H: keep S pieces in flight, drain the oldest into the pageable destination
        std::vector<size_t> mine;
        for (size_t k = t; k < pieces; k += T) mine.push_back(k);
        for (size_t i = 0; i < mine.size() + S; i++) {
          if (i >= (size_t)S) {
            const size_t j = i - S, off = mine[j] * P, len = std::min(P, bytes - off);
            CK(cudaEventSynchronize(me.ev[j % S]));
            copyBytes(host + off, me.slot[j % S], len, nt);
          }
          if (i < mine.size()) {
            const size_t off = mine[i] * P, len = std::min(P, bytes - off);
            CK(cudaMemcpyAsync(me.slot[i % S], dev + off, len, cudaMemcpyDeviceToHost, me.st));
            CK(cudaEventRecord(me.ev[i % S], me.st));
          }
        }
      }
    });
  for (auto &x : th) x.join();
  return bytes / (now() - t0) / 1e9;
}

int main(int argc, char **argv) {
  const size_t bytes = (size_t)(argc > 1 ? atof(argv[1]) : 2.36e9);
  uint8_t *src = (uint8_t *)aligned_alloc(4096, bytes), *dev = nullptr;
  for (size_t i = 0; i < bytes; i += 4096) src[i] = (uint8_t)(i >> 12);
  memset(src, 5, bytes);
  CK(cudaMalloc(&dev, bytes));
  for (int T : {8, 12})
    for (int nt = 0; nt < 2; nt++)
      printf("{\"shape\": \"ranges\", \"dir\": \"h2d\", \"threads\": %d, \"stores\": \"%s\", \"range_mb\": 60, \"gbs\": %.1f}\n", T, nt ? "nt" : "memcpy",
             rangesH2D(dev, src, bytes, T, nt, (size_t)60 << 20)), fflush(stdout);
  for (int T : {4, 8, 12, 16})
    for (size_t P : {(size_t)512 << 10, (size_t)1 << 20, (size_t)2 << 20, (size_t)4 << 20})
      for (int S : {2, 3}) {
        std::vector<Worker> w(T);
        for (auto &x : w) {
          CK(cudaStreamCreateWithFlags(&x.st, cudaStreamNonBlocking));
          for (int s = 0; s < S; s++) { uint8_t *p; CK(cudaHostAlloc(&p, P, 0)); memset(p, 0, P); x.slot.push_back(p); cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); x.ev.push_back(e); }
        }
        for (int nt = 0; nt < 2; nt++) {
          const double up = workers(dev, src, bytes, T, S, P, nt, true, w);
          const double down = workers(dev, src, bytes, T, S, P, nt, false, w);
          printf("{\"shape\": \"workers\", \"threads\": %d, \"slots\": %d, \"piece_kb\": %zu, \"stores\": \"%s\", \"footprint_mb\": %.1f, \"h2d_gbs\": %.1f, \"d2h_gbs\": %.1f}\n",
                 T, S, P >> 10, nt ? "nt" : "memcpy", T * S * P / 1048576.0, up, down);
          fflush(stdout);
        }
        for (auto &x : w) { for (auto p : x.slot) cudaFreeHost(p); for (auto e : x.ev) cudaEventDestroy(e); cudaStreamDestroy(x.st); }
      }
  return 0;
}
