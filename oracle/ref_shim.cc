// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product.
//
// extern "C" shim around the UNMODIFIED reference C++ (lanxinger/spz, src/cc/*.cc), compiled
// from the sources where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libspz_ref.so.  No reference source is copied into this repository; this file only
// *declares* the two hot-path functions the reference defines with external linkage but does not
// declare in any header (src/cc/load-spz.cc:257 packGaussians, :467 unpackGaussians) and marshals
// flat C arrays in and out of the reference's std::vector-based structs.
//
// Used for: (1) pinning oracle/spz_oracle.c (the C restatement) bit-for-bit, (2) generating the
// golden fixtures under tests/golden/, (3) the "reference" arm / cpu_baseline of bench.py.

#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "load-spz.h"  // resolved via -I/root/reference/src/cc

namespace spz {
PackedGaussians packGaussians(const GaussianCloud &g, const PackOptions &o);      // load-spz.cc:257
GaussianCloud unpackGaussians(const PackedGaussians &packed, const UnpackOptions &o);  // load-spz.cc:467
}  // namespace spz

namespace {

int dimForDegree(int d) { return d == 0 ? 0 : d == 1 ? 3 : d == 2 ? 8 : 15; }

double nowSeconds() {
  using clk = std::chrono::steady_clock;
  return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}

void fillCloud(spz::GaussianCloud &g, int32_t n, int32_t shDegree, int32_t antialiased,
               const float *positions, const float *scales, const float *rotations,
               const float *alphas, const float *colors, const float *sh) {
  const size_t N = static_cast<size_t>(n);
  const size_t D = static_cast<size_t>(dimForDegree(shDegree));
  g.numPoints = n;
  g.shDegree = shDegree;
  g.antialiased = antialiased != 0;
  g.positions.assign(positions, positions + N * 3);
  g.scales.assign(scales, scales + N * 3);
  g.rotations.assign(rotations, rotations + N * 4);
  g.alphas.assign(alphas, alphas + N);
  g.colors.assign(colors, colors + N * 3);
  g.sh.assign(sh, sh + N * D * 3);
}

template <class T>
void copyOut(T *dst, const std::vector<T> &src) {
  if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(T));
}

}  // namespace

extern "C" {

// Returns 0 on success, 1 if the reference rejected the input (it then returns an empty struct).
// *seconds receives the wall time of the packGaussians call alone (marshalling excluded).
int ref_pack(int32_t n, int32_t shDegree, int32_t from, const float *positions,
             const float *scales, const float *rotations, const float *alphas,
             const float *colors, const float *sh, uint8_t *oPositions, uint8_t *oScales,
             uint8_t *oRotations, uint8_t *oAlphas, uint8_t *oColors, uint8_t *oSh,
             double *seconds) {
  spz::GaussianCloud g;
  fillCloud(g, n, shDegree, 0, positions, scales, rotations, alphas, colors, sh);
  spz::PackOptions o;
  o.from = static_cast<spz::CoordinateSystem>(from);
  const double t0 = nowSeconds();
  spz::PackedGaussians p = spz::packGaussians(g, o);
  const double t1 = nowSeconds();
  if (seconds) *seconds = t1 - t0;
  if (p.numPoints != n || p.positions.size() != static_cast<size_t>(n) * 9) return 1;
  copyOut(oPositions, p.positions);
  copyOut(oScales, p.scales);
  copyOut(oRotations, p.rotations);
  copyOut(oAlphas, p.alphas);
  copyOut(oColors, p.colors);
  copyOut(oSh, p.sh);
  return 0;
}

// version: 1 (float16 positions, first-three quaternion), 2 (24-bit, first-three), 3 (24-bit,
// smallest-three) -- the mapping deserializePackedGaussians applies (load-spz.cc:571-572); 4 =
// float16 positions with usesQuaternionSmallestThree = true (a hand-built struct, no file form).
int ref_unpack(int32_t n, int32_t shDegree, int32_t fractionalBits, int32_t version, int32_t to,
               const uint8_t *positions, const uint8_t *scales, const uint8_t *rotations,
               const uint8_t *alphas, const uint8_t *colors, const uint8_t *sh, float *oPositions,
               float *oScales, float *oRotations, float *oAlphas, float *oColors, float *oSh,
               double *seconds) {
  const size_t N = static_cast<size_t>(n);
  const size_t D = static_cast<size_t>(dimForDegree(shDegree));
  spz::PackedGaussians p;
  p.numPoints = n;
  p.shDegree = shDegree;
  p.fractionalBits = fractionalBits;
  p.antialiased = false;
  p.usesQuaternionSmallestThree = version >= 3;
  p.positions.assign(positions, positions + N * 3 * ((version == 1 || version == 4) ? 2 : 3));
  p.scales.assign(scales, scales + N * 3);
  p.rotations.assign(rotations, rotations + N * (version >= 3 ? 4 : 3));
  p.alphas.assign(alphas, alphas + N);
  p.colors.assign(colors, colors + N * 3);
  p.sh.assign(sh, sh + N * D * 3);
  spz::UnpackOptions o;
  o.to = static_cast<spz::CoordinateSystem>(to);
  const double t0 = nowSeconds();
  spz::GaussianCloud g = spz::unpackGaussians(p, o);
  const double t1 = nowSeconds();
  if (seconds) *seconds = t1 - t0;
  if (g.numPoints != n || g.positions.size() != N * 3) return 1;
  copyOut(oPositions, g.positions);
  copyOut(oScales, g.scales);
  copyOut(oRotations, g.rotations);
  copyOut(oAlphas, g.alphas);
  copyOut(oColors, g.colors);
  copyOut(oSh, g.sh);
  return 0;
}

// PackedGaussians::unpack(i, c) (load-spz.cc:461-463) for a list of indices; conv = 21 floats of a
// CoordinateConverter; 59 floats out per gaussian in UnpackedGaussian's member order.
int ref_unpack_at(int32_t n, int32_t shDegree, int32_t fractionalBits, int32_t version,
                  const uint8_t *positions, const uint8_t *scales, const uint8_t *rotations,
                  const uint8_t *alphas, const uint8_t *colors, const uint8_t *sh,
                  const int64_t *indices, int64_t count, const float *conv, float *out) {
  const size_t N = static_cast<size_t>(n);
  const size_t D = static_cast<size_t>(dimForDegree(shDegree));
  spz::PackedGaussians p;
  p.numPoints = n;
  p.shDegree = shDegree;
  p.fractionalBits = fractionalBits;
  p.usesQuaternionSmallestThree = version >= 3;
  p.positions.assign(positions, positions + N * 3 * ((version == 1 || version == 4) ? 2 : 3));
  p.scales.assign(scales, scales + N * 3);
  p.rotations.assign(rotations, rotations + N * (version >= 3 ? 4 : 3));
  p.alphas.assign(alphas, alphas + N);
  p.colors.assign(colors, colors + N * 3);
  p.sh.assign(sh, sh + N * D * 3);
  spz::CoordinateConverter c;
  std::copy(conv, conv + 3, c.flipP.begin());
  std::copy(conv + 3, conv + 6, c.flipQ.begin());
  std::copy(conv + 6, conv + 21, c.flipSh.begin());
  static_assert(sizeof(spz::UnpackedGaussian) == 59 * sizeof(float), "UnpackedGaussian is 59 packed floats");
  for (int64_t k = 0; k < count; k++) {
    if (indices[k] < 0 || indices[k] >= n) return 2;
    const spz::UnpackedGaussian u = p.unpack(static_cast<int32_t>(indices[k]), c);
    std::memcpy(out + 59 * k, &u, sizeof u);
  }
  return 0;
}

// saveSpz(cloud, opts, vector*) (load-spz.cc:598).  Returns a malloc'd gzip buffer (caller frees
// with ref_free) or NULL.
uint8_t *ref_save_spz(int32_t n, int32_t shDegree, int32_t antialiased, int32_t from,
                      const float *positions, const float *scales, const float *rotations,
                      const float *alphas, const float *colors, const float *sh,
                      uint64_t *outSize) {
  spz::GaussianCloud g;
  fillCloud(g, n, shDegree, antialiased, positions, scales, rotations, alphas, colors, sh);
  spz::PackOptions o;
  o.from = static_cast<spz::CoordinateSystem>(from);
  std::vector<uint8_t> out;
  if (!spz::saveSpz(g, o, &out)) return nullptr;
  uint8_t *buf = static_cast<uint8_t *>(std::malloc(out.size() ? out.size() : 1));
  std::memcpy(buf, out.data(), out.size());
  *outSize = out.size();
  return buf;
}

// The uncompressed container the reference would gzip: packGaussians + serializePackedGaussians
// (load-spz.cc:533).  Caller frees with ref_free.
uint8_t *ref_serialize(int32_t n, int32_t shDegree, int32_t antialiased, int32_t from,
                       const float *positions, const float *scales, const float *rotations,
                       const float *alphas, const float *colors, const float *sh,
                       uint64_t *outSize) {
  spz::GaussianCloud g;
  fillCloud(g, n, shDegree, antialiased, positions, scales, rotations, alphas, colors, sh);
  spz::PackOptions o;
  o.from = static_cast<spz::CoordinateSystem>(from);
  spz::PackedGaussians p = spz::packGaussians(g, o);
  std::stringstream ss;
  spz::serializePackedGaussians(p, &ss);
  const std::string s = ss.str();
  uint8_t *buf = static_cast<uint8_t *>(std::malloc(s.size() ? s.size() : 1));
  std::memcpy(buf, s.data(), s.size());
  *outSize = s.size();
  return buf;
}

struct RefCloud {
  spz::GaussianCloud g;
};

// loadSpz(ptr, size, opts) (load-spz.cc:638).  Always returns a handle; an empty cloud (numPoints
// 0) is what the reference hands back on any failure.
RefCloud *ref_load_spz(const uint8_t *data, int32_t size, int32_t to) {
  spz::UnpackOptions o;
  o.to = static_cast<spz::CoordinateSystem>(to);
  RefCloud *c = new RefCloud;
  c->g = spz::loadSpz(data, size, o);
  return c;
}

void ref_cloud_info(const RefCloud *c, int32_t *n, int32_t *shDegree, int32_t *antialiased) {
  *n = c->g.numPoints;
  *shDegree = c->g.shDegree;
  *antialiased = c->g.antialiased ? 1 : 0;
}

void ref_cloud_copy(const RefCloud *c, float *positions, float *scales, float *rotations,
                    float *alphas, float *colors, float *sh) {
  copyOut(positions, c->g.positions);
  copyOut(scales, c->g.scales);
  copyOut(rotations, c->g.rotations);
  copyOut(alphas, c->g.alphas);
  copyOut(colors, c->g.colors);
  copyOut(sh, c->g.sh);
}

void ref_cloud_free(RefCloud *c) { delete c; }

// compressGzipped (load-spz.cc:186), timed; returns compressed size or 0.
uint64_t ref_gzip_size(const uint8_t *data, uint64_t size, double *seconds) {
  std::vector<uint8_t> out;
  const double t0 = nowSeconds();
  const bool ok = spz::compressGzipped(data, size, &out);
  const double t1 = nowSeconds();
  if (seconds) *seconds = t1 - t0;
  return ok ? out.size() : 0;
}

void ref_free(void *p) { std::free(p); }

}  // extern "C"
