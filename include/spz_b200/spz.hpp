// Drop-in C++ API of the B200-native .spz codec: namespace spz with the same public types and
// functions a consumer of lanxinger/spz compiles against (src/cc/load-spz.h, splat-types.h,
// splat-c-types.h), so existing callers rebuild and relink unchanged.  The two hot functions
//
//     PackedGaussians packGaussians(const GaussianCloud &, const PackOptions &)      load-spz.cc:257
//     GaussianCloud unpackGaussians(const PackedGaussians &, const UnpackOptions &)  load-spz.cc:467
//
// (defined with external linkage but not declared by the reference's headers) run on the GPU
// through the C-ABI in include/spz_b200.h; there is no CPU implementation of them here, and both
// return an empty struct when no sm_100 device is usable, after logging why.  Everything else in
// this header is host-side glue the reference also keeps on the host: the 16-byte container
// header and plane order, zlib's gzip framing, file I/O, one-element accessors, small math.
//
// Struct layouts (member order and types) match the reference so objects can cross a library
// boundary built against either header set.  Sizes the reference computes in int32 are computed
// in 64 bits here, and the reader's 10,000,000-point cap (load-spz.cc:549) is lifted to the
// int32 range of numPoints; both are deliberate, documented divergences.
#pragma once

#include <algorithm>
#include <array>
#include <math.h>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iosfwd>
#include <string>
#include <vector>

// ---- C bridge structs (reference: splat-c-types.h:11-26) ----------------------------------------
#ifndef SPZ_SPLAT_C_TYPES_H_
#define SPZ_SPLAT_C_TYPES_H_
typedef struct {
  size_t count;
  float *data;
} SpzFloatBuffer;

typedef struct {
  int32_t numPoints;
  int32_t shDegree;
  bool antialiased;
  SpzFloatBuffer positions;
  SpzFloatBuffer scales;
  SpzFloatBuffer rotations;
  SpzFloatBuffer alphas;
  SpzFloatBuffer colors;
  SpzFloatBuffer sh;
} GaussianCloudData;
#endif  // SPZ_SPLAT_C_TYPES_H_

namespace spz {

// Deep copy of a float plane into a new[]-allocated buffer the caller frees (reference: splat-types.h:14-22).
inline SpzFloatBuffer copyFloatBuffer(const std::vector<float> &vector) {
  SpzFloatBuffer buffer = {0, nullptr};
  if (vector.empty()) return buffer;
  buffer.count = vector.size();
  buffer.data = new float[buffer.count];
  std::copy(vector.begin(), vector.end(), buffer.data);
  return buffer;
}

// ---- coordinate systems (reference: splat-types.h:24-81) ----------------------------------------

// Axis directions are encoded in the value: bits 0,1,2 of (value - 1) say Right, Up, Front.
enum class CoordinateSystem {
  UNSPECIFIED = 0,
  LDB = 1,
  RDB = 2,
  LUB = 3,
  RUB = 4,  // three.js; the frame .spz stores
  LDF = 5,
  RDF = 6,  // PLY
  LUF = 7,  // GLB
  RUF = 8,  // Unity
};

struct CoordinateConverter {
  std::array<float, 3> flipP = {1.0f, 1.0f, 1.0f};
  std::array<float, 3> flipQ = {1.0f, 1.0f, 1.0f};  // x, y, z of the quaternion; w never flips
  std::array<float, 15> flipSh = {1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f,
                                  1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f};
};

constexpr std::array<bool, 3> axesMatch(CoordinateSystem a, CoordinateSystem b) {
  const int ia = static_cast<int>(a) - 1, ib = static_cast<int>(b) - 1;
  if (ia < 0 || ib < 0) return {true, true, true};  // UNSPECIFIED converts to nothing
  const int diff = ia ^ ib;
  return {(diff & 1) == 0, (diff & 2) == 0, (diff & 4) == 0};
}

constexpr CoordinateConverter coordinateConverter(CoordinateSystem from, CoordinateSystem to) {
  const std::array<bool, 3> same = axesMatch(from, to);
  const float x = same[0] ? 1.0f : -1.0f, y = same[1] ? 1.0f : -1.0f, z = same[2] ? 1.0f : -1.0f;
  // SH basis function j is odd in the axes listed: y, z, x | xy, yz, -, xz, - | y, xyz, y, z, x, z, x
  return CoordinateConverter{{x, y, z},
                             {y * z, x * z, x * y},
                             {y, z, x, x * y, y * z, 1.0f, x * z, 1.0f, y, x * y * z, y, z, x, z, x}};
}

// ---- GaussianCloud (reference: splat-types.h:90-186) ---------------------------------------------

struct GaussianCloud {
  int32_t numPoints = 0;
  int32_t shDegree = 0;
  bool antialiased = false;

  std::vector<float> positions;  // xyz
  std::vector<float> scales;     // log scale, xyz
  std::vector<float> rotations;  // xyzw
  std::vector<float> alphas;     // pre-sigmoid
  std::vector<float> colors;     // SH DC, rgb
  std::vector<float> sh;         // 0 / 9 / 24 / 45 per point; coefficient-major, rgb innermost

  // Deep copy into new[]-allocated buffers the caller frees.
  GaussianCloudData data() const {
    GaussianCloudData d;
    d.numPoints = numPoints;
    d.shDegree = shDegree;
    d.antialiased = antialiased;
    d.positions = copyFloatBuffer(positions);
    d.scales = copyFloatBuffer(scales);
    d.rotations = copyFloatBuffer(rotations);
    d.alphas = copyFloatBuffer(alphas);
    d.colors = copyFloatBuffer(colors);
    d.sh = copyFloatBuffer(sh);
    return d;
  }

  // In-place change of frame.  (The bulk decoder does not call this: unpackGaussians folds the
  // same flips into its kernel.)
  void convertCoordinates(CoordinateSystem from, CoordinateSystem to) {
    if (numPoints == 0) return;
    const CoordinateConverter c = coordinateConverter(from, to);
    for (size_t i = 0; i + 2 < positions.size(); i += 3)
      for (size_t a = 0; a < 3; a++) positions[i + a] *= c.flipP[a];
    for (size_t i = 0; i + 3 < rotations.size(); i += 4)
      for (size_t a = 0; a < 3; a++) rotations[i + a] *= c.flipQ[a];
    const size_t perPoint = sh.size() / 3 / static_cast<size_t>(numPoints);  // coefficients per point
    if (perPoint == 0) return;
    for (size_t i = 0; i + 2 < sh.size(); i += 3) {
      const float f = c.flipSh[(i / 3) % perPoint];
      sh[i] *= f;
      sh[i + 1] *= f;
      sh[i + 2] *= f;
    }
  }

  void rotate180DegAboutX() { convertCoordinates(CoordinateSystem::RUB, CoordinateSystem::RDF); }

  // Volume of the median ellipsoid, 4/3 pi exp(median(sx + sy + sz)); 0.01 for an empty cloud.
  float medianVolume() const {
    if (numPoints == 0) return 0.01f;
    std::vector<float> sums;
    sums.reserve(scales.size() / 3);
    for (size_t i = 0; i + 2 < scales.size(); i += 3) sums.push_back(scales[i] + scales[i + 1] + scales[i + 2]);
    if (sums.empty()) return 0.01f;  // inconsistent cloud (points but no scales): the reference reads out of bounds
    std::sort(sums.begin(), sums.end());
    const float median = sums[sums.size() / 2];
    return (3.14159265358979323846 * 4 / 3) * exp(median);
  }
};

// ---- small math (reference: splat-types.h:189-264, splat-types.cc) --------------------------------

using Vec3f = std::array<float, 3>;   // x, y, z
using Quat4f = std::array<float, 4>;  // w, x, y, z
using Half = uint16_t;

float halfToFloat(Half h);
Half floatToHalf(float f);
Vec3f normalized(const Vec3f &v);
float norm(const Vec3f &a);
float norm(const Quat4f &q);
Quat4f normalized(const Quat4f &v);
Quat4f axisAngleQuat(const Vec3f &scaledAxis);

constexpr Vec3f vec3f(const float *p) { return {p[0], p[1], p[2]}; }
constexpr Quat4f quat4f(const float *p) { return {p[0], p[1], p[2], p[3]}; }
constexpr float dot(const Vec3f &a, const Vec3f &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
constexpr float squaredNorm(const Vec3f &v) { return dot(v, v); }
constexpr Vec3f plus(const Vec3f &a, const Vec3f &b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
constexpr Vec3f times(const Vec3f &v, float s) { return {v[0] * s, v[1] * s, v[2] * s}; }
constexpr Vec3f times(const Vec3f &a, const Vec3f &b) { return {a[0] * b[0], a[1] * b[1], a[2] * b[2]}; }
constexpr Quat4f plus(const Quat4f &a, const Quat4f &b) {
  return {a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]};
}
constexpr Quat4f times(const Quat4f &a, float s) { return {a[0] * s, a[1] * s, a[2] * s, a[3] * s}; }

// Rotate p by the unit quaternion q (w, x, y, z).
constexpr Vec3f times(const Quat4f &q, const Vec3f &p) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  const float x2 = x + x, y2 = y + y, z2 = z + z;
  const float wx = w * x2, wy = w * y2, wz = w * z2;
  const float xx = x * x2, xy = x * y2, xz = x * z2;
  const float yy = y * y2, yz = y * z2, zz = z * z2;
  return {p[0] * (1.0f - (yy + zz)) + p[1] * (xy - wz) + p[2] * (xz + wy),
          p[0] * (xy + wz) + p[1] * (1.0f - (xx + zz)) + p[2] * (yz - wx),
          p[0] * (xz - wy) + p[1] * (yz + wx) + p[2] * (1.0f - (xx + yy))};
}

// Hamilton product, renormalised.
inline Quat4f times(const Quat4f &a, const Quat4f &b) {
  return normalized(Quat4f{a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                           a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                           a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                           a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]});
}

// ---- packed representation (reference: load-spz.h:13-67) -------------------------------------------

struct UnpackedGaussian {
  std::array<float, 3> position;
  std::array<float, 4> rotation;  // x, y, z, w
  std::array<float, 3> scale;
  std::array<float, 3> color;
  float alpha;
  std::array<float, 15> shR;
  std::array<float, 15> shG;
  std::array<float, 15> shB;
};

struct PackedGaussian {
  std::array<uint8_t, 9> position{};
  std::array<uint8_t, 4> rotation{};
  std::array<uint8_t, 3> scale{};
  std::array<uint8_t, 3> color{};
  uint8_t alpha = 0;
  std::array<uint8_t, 15> shR{};
  std::array<uint8_t, 15> shG{};
  std::array<uint8_t, 15> shB{};

  UnpackedGaussian unpack(bool usesFloat16, bool usesQuaternionSmallestThree, int32_t fractionalBits,
                          const CoordinateConverter &c) const;
};

struct PackedGaussians {
  int32_t numPoints = 0;
  int32_t shDegree = 0;
  int32_t fractionalBits = 0;
  bool antialiased = false;
  bool usesQuaternionSmallestThree = true;

  std::vector<uint8_t> positions;
  std::vector<uint8_t> scales;
  std::vector<uint8_t> rotations;
  std::vector<uint8_t> alphas;
  std::vector<uint8_t> colors;
  std::vector<uint8_t> sh;

  bool usesFloat16() const;
  PackedGaussian at(int32_t i) const;
  UnpackedGaussian unpack(int32_t i, const CoordinateConverter &c) const;
};

struct PackOptions {
  CoordinateSystem from = CoordinateSystem::UNSPECIFIED;
};

struct UnpackOptions {
  CoordinateSystem to = CoordinateSystem::UNSPECIFIED;
};

// ---- the codec (GPU) ---------------------------------------------------------------------------------
PackedGaussians packGaussians(const GaussianCloud &g, const PackOptions &o);
GaussianCloud unpackGaussians(const PackedGaussians &packed, const UnpackOptions &o);

// ---- container + file API (host glue around the codec; reference: load-spz.h:69-100) ----------------
bool saveSpz(const GaussianCloud &gaussians, const PackOptions &options, std::vector<uint8_t> *output);
bool saveSpz(const GaussianCloud &gaussians, const PackOptions &options, const std::string &filename);
GaussianCloud loadSpz(const std::vector<uint8_t> &data, const UnpackOptions &options);
GaussianCloud loadSpz(const std::string &filename, const UnpackOptions &o);
GaussianCloud loadSpz(const uint8_t *data, int32_t size, const UnpackOptions &options);
PackedGaussians loadSpzPacked(const std::string &filename);
PackedGaussians loadSpzPacked(const uint8_t *data, int32_t size);
PackedGaussians loadSpzPacked(const std::vector<uint8_t> &data);
bool saveSplatToPly(const spz::GaussianCloud &gaussians, const PackOptions &options, const std::string &filename);
GaussianCloud loadSplatFromPly(const std::string &filename, const UnpackOptions &options);
void serializePackedGaussians(const PackedGaussians &packed, std::ostream *out);
bool compressGzipped(const uint8_t *data, size_t size, std::vector<uint8_t> *out);

// ---- extensions (not in the reference's headers) --------------------------------------------------
// The version-2 form of the stream (rotations as the first three components, 3 bytes; the form
// unpackQuaternionFirstThree reads) and a container whose header says so.  PARITY UNPINNED: the reference
// has no encoder for it; every other plane equals packGaussians'.
PackedGaussians packGaussiansV2(const GaussianCloud &g, const PackOptions &o);
bool saveSpzV2(const GaussianCloud &gaussians, const PackOptions &options, std::vector<uint8_t> *output);
// packed.unpack(i, c) for every i of `indices` in ONE kernel launch (a loop over unpack(i, c) pays a
// launch per gaussian).  Empty result, after a logged line, when an index is out of range or no device
// is usable.
std::vector<UnpackedGaussian> unpackGaussiansAt(const PackedGaussians &packed, const std::vector<int32_t> &indices,
                                                const CoordinateConverter &c);
// The inverse of compressGzipped (the reference keeps it file-local, load-spz.cc:169-182).
bool decompressGzipped(const uint8_t *data, size_t size, std::vector<uint8_t> *out);
// zlib on several host threads: one standard gzip member made of independently deflated blocks,
// readable by any inflater; the inflated bytes equal compressGzipped's input, the compressed bytes
// do not equal its output.  decompressGzippedParallel inflates such members block-parallel and
// anything else serially.  saveSpz / loadSpz* use them when SPZ_B200_GZIP_THREADS > 1.
// .ply file -> .spz bytes with the loader's column shuffle and the encoder fused in one GPU kernel
// (no planar GaussianCloud in between).  options.from = RDF: the bytes of
// saveSpz(loadSplatFromPly(f, {to}), {from = to}); UNSPECIFIED: what the reference's ply_to_spz writes.
bool plyToSpz(const std::string &plyFilename, const PackOptions &options, std::vector<uint8_t> *output);
// The mirror: .spz bytes -> .ply file, decoder and the writer's record layout fused in one kernel.
// options.to = RDF: the file of saveSplatToPly(loadSpz(bytes, {to}), {from = to}, f); UNSPECIFIED: spz_to_ply's.
bool spzToPly(const std::vector<uint8_t> &spzBytes, const UnpackOptions &options, const std::string &plyFilename);
bool compressGzippedParallel(const uint8_t *data, size_t size, int threads, std::vector<uint8_t> *out);
bool decompressGzippedParallel(const uint8_t *data, size_t size, int threads, std::vector<uint8_t> *out);

}  // namespace spz
