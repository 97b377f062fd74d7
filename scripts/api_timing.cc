// Development tool: wall time of the drop-in C++ API (spz::packGaussians / spz::unpackGaussians on
// std::vector planes) for a synthetic cloud.  Built by scripts/gpu_api.sh against include/spz and
// libspz_b200.so; not part of the product or the tests.
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
  const size_t n = argc > 1 ? (size_t)atof(argv[1]) : 10000000;
  const int reps = argc > 2 ? atoi(argv[2]) : 4;
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
    if (p.numPoints != (int32_t)n || back.numPoints != (int32_t)n) { printf("FAILED\n"); return 1; }
    printf("{\"api\": \"spz::packGaussians/unpackGaussians\", \"points\": %zu, \"rep\": %d, \"pack_ms\": %.1f, \"unpack_ms\": %.1f, \"check\": %.6f}\n", n, r,
           (t1 - t0) * 1e3, (t2 - t1) * 1e3, (double)back.sh[n * 45 - 1] + back.positions[0]);
    fflush(stdout);
  }
  // per-gaussian access: a loop over packed.unpack(i, c) against one unpackGaussiansAt call
  {
    const spz::PackedGaussians p = spz::packGaussians(g, {spz::CoordinateSystem::RUB});
    const spz::CoordinateConverter c = spz::coordinateConverter(spz::CoordinateSystem::RUB, spz::CoordinateSystem::RDF);
    const int loopN = 2000;
    double acc = 0;
    p.unpack(0, c);
    const double t0 = now();
    for (int i = 0; i < loopN; i++) acc += p.unpack((int32_t)((size_t)i * 7919u % n), c).position[0];
    const double t1 = now();
    std::vector<int32_t> idx(std::min<size_t>(n, 1000000));
    for (size_t i = 0; i < idx.size(); i++) idx[i] = (int32_t)(i * 7919u % n);
    spz::unpackGaussiansAt(p, idx, c);
    const double t2 = now();
    const std::vector<spz::UnpackedGaussian> many = spz::unpackGaussiansAt(p, idx, c);
    const double t3 = now();
    for (int i = 0; i < loopN; i++) acc -= many[(size_t)i].position[0];
    // the walk the accessor exists for: i = 0, 1, 2, ... (read-ahead windows; SPZ_B200_UNPACK_READAHEAD=0 turns them off)
    const size_t walkN = std::min<size_t>(n, std::getenv("SPZ_B200_UNPACK_READAHEAD") ? 20000 : 1000000);
    double walkAcc = 0;
    const double t4 = now();
    for (size_t i = 0; i < walkN; i++) walkAcc += p.unpack((int32_t)i, c).position[0];
    const double t5 = now();
    printf("{\"api\": \"PackedGaussians::unpack(i, c), i in order\", \"points\": %zu, \"walked\": %zu, \"walk_us_per_gaussian\": %.3f, \"walk_mgaussians_s\": %.2f, \"check\": %.4f}\n",
           n, walkN, (t5 - t4) * 1e6 / walkN, walkN / (t5 - t4) / 1e6, walkAcc);
    printf("{\"api\": \"PackedGaussians::unpack(i, c)\", \"points\": %zu, \"loop_us_per_gaussian\": %.2f, \"batched_gaussians\": %zu, \"batched_ms\": %.2f, "
           "\"batched_mgaussians_s\": %.2f, \"loop_minus_batched\": %.6f}\n",
           n, (t1 - t0) * 1e6 / loopN, many.size(), (t3 - t2) * 1e3, many.size() / (t3 - t2) / 1e6, acc);
  }
  return 0;
}
