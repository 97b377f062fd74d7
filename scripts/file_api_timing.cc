// Development tool: wall time of the reference's file-level API -- saveSpz(cloud) -> bytes, loadSpz(bytes) -> cloud --
// and of packGaussians / unpackGaussians, on a synthetic SH3 cloud in std::vector planes.  The same source is built
// against this library (scripts/_build/file_api_timing) and against the reference's sources compiled in place
// (scripts/_build/file_api_timing_ref; made where /root/reference exists:
//   g++ -std=c++17 -O2 -ffp-contract=off -pthread -w -I/root/reference/src/cc scripts/file_api_timing.cc /root/reference/src/cc/{load-spz,splat-types,splat-c-types}.cc -o scripts/_build/file_api_timing_ref -lz
// ), run by `scripts/gpu.sh fileapi`.  Not part of the product or the tests.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "load-spz.h"

namespace spz {
PackedGaussians packGaussians(const GaussianCloud &g, const PackOptions &o);
GaussianCloud unpackGaussians(const PackedGaussians &p, const UnpackOptions &o);
}  // namespace spz

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void fill(std::vector<float> &v, size_t n, float lo, float hi, unsigned seed) {
  v.resize(n);
  const int T = 8;
  std::vector<std::thread> th;
  for (int k = 0; k < T; k++)
    th.emplace_back([&, k] {
      std::minstd_rand r(seed * 97u + (unsigned)k);
      const size_t a = n / T * k, b = k == T - 1 ? n : n / T * (k + 1);
      for (size_t i = a; i < b; i++) v[i] = lo + (hi - lo) * (float)(r() & 0xffffff) / 16777216.0f;
    });
  for (auto &t : th) t.join();
}

int main(int argc, char **argv) {
  const size_t n = argc > 1 ? (size_t)atof(argv[1]) : 1000000;
  const int reps = argc > 2 ? atoi(argv[2]) : 2;
  const char *impl = argc > 3 ? argv[3] : "?";
  spz::GaussianCloud g;
  g.numPoints = (int32_t)n;
  g.shDegree = 3;
  fill(g.positions, n * 3, -10, 10, 1);
  fill(g.scales, n * 3, -7, 1, 2);
  fill(g.rotations, n * 4, -1, 1, 3);
  fill(g.alphas, n, -6, 6, 4);
  fill(g.colors, n * 3, -2, 2, 5);
  fill(g.sh, n * 45, -0.5f, 0.5f, 6);
  for (int r = 0; r < reps; r++) {
    const double t0 = now();
    spz::PackedGaussians p = spz::packGaussians(g, {spz::CoordinateSystem::RUB});
    const double t1 = now();
    spz::GaussianCloud back = spz::unpackGaussians(p, {spz::CoordinateSystem::RUB});
    const double t2 = now();
    std::vector<uint8_t> file;
    const bool saved = spz::saveSpz(g, {spz::CoordinateSystem::RUB}, &file);
    const double t3 = now();
    spz::GaussianCloud loaded = spz::loadSpz(file, {spz::CoordinateSystem::RUB});
    const double t4 = now();
    if (!saved || p.numPoints != (int32_t)n || back.numPoints != (int32_t)n || loaded.numPoints != (int32_t)n) { printf("FAILED\n"); return 1; }
    double check = 0;
    for (size_t i = 0; i < back.sh.size(); i += 4099) check += (double)back.sh[i] - (double)loaded.sh[i];
    printf("{\"impl\": \"%s\", \"points\": %zu, \"rep\": %d, \"pack_ms\": %.1f, \"unpack_ms\": %.1f, \"saveSpz_ms\": %.1f, \"loadSpz_ms\": %.1f, \"file_mb\": %.1f, \"unpack_minus_load\": %.1f}\n",
           impl, n, r, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, file.size() / 1e6, check);
    fflush(stdout);
  }
  return 0;
}
