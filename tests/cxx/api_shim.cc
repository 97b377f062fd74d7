// TEST INFRASTRUCTURE.  A consumer of the spz C++ API, written once and compiled twice:
//   * against this repo's drop-in headers + libspz_b200.so          (-DSHIM_PREFIX=b200_)
//   * against the reference's own headers and sources, in place      (-DSHIM_PREFIX=ref_)
// The two builds expose the same flat C functions so tests can feed both the same inputs and
// compare outputs byte for byte.  That the file compiles unchanged against either header set is
// itself the drop-in check.
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "load-spz.h"  // -I decides whose

namespace spz {
PackedGaussians packGaussians(const GaussianCloud &g, const PackOptions &o);
GaussianCloud unpackGaussians(const PackedGaussians &packed, const UnpackOptions &o);
}  // namespace spz

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
#define SHIM(name) CAT(SHIM_PREFIX, name)

namespace {
int dimFor(int d) { return d == 0 ? 0 : d == 1 ? 3 : d == 2 ? 8 : 15; }

spz::GaussianCloud makeCloud(int32_t n, int32_t deg, int32_t aa, const float *const planes[6]) {
  const size_t N = (size_t)n, D = (size_t)dimFor(deg);
  spz::GaussianCloud g;
  g.numPoints = n; g.shDegree = deg; g.antialiased = aa != 0;
  g.positions.assign(planes[0], planes[0] + N * 3);
  g.scales.assign(planes[1], planes[1] + N * 3);
  g.rotations.assign(planes[2], planes[2] + N * 4);
  g.alphas.assign(planes[3], planes[3] + N);
  g.colors.assign(planes[4], planes[4] + N * 3);
  g.sh.assign(planes[5], planes[5] + N * D * 3);
  return g;
}

// version: 1..3 container flavours, 4 = float16 positions + smallest-three
spz::PackedGaussians makePacked(int32_t n, int32_t deg, int32_t fb, int32_t version, int32_t aa,
                                const uint8_t *const planes[6]) {
  const size_t N = (size_t)n, D = (size_t)dimFor(deg);
  spz::PackedGaussians p;
  p.numPoints = n; p.shDegree = deg; p.fractionalBits = fb; p.antialiased = aa != 0;
  p.usesQuaternionSmallestThree = version >= 3;
  p.positions.assign(planes[0], planes[0] + N * ((version == 1 || version == 4) ? 6 : 9));
  p.scales.assign(planes[1], planes[1] + N * 3);
  p.rotations.assign(planes[2], planes[2] + N * (version >= 3 ? 4 : 3));
  p.alphas.assign(planes[3], planes[3] + N);
  p.colors.assign(planes[4], planes[4] + N * 3);
  p.sh.assign(planes[5], planes[5] + N * D * 3);
  return p;
}

template <class T>
void put(T *dst, const std::vector<T> &v) {
  if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(T));
}

uint8_t *dupBytes(const void *p, size_t n, uint64_t *outSize) {
  uint8_t *b = (uint8_t *)std::malloc(n ? n : 1);
  if (n) std::memcpy(b, p, n);
  *outSize = n;
  return b;
}
}  // namespace

extern "C" {

void SHIM(free)(void *p) { std::free(p); }

// ---- the hot path through the C++ API ------------------------------------------------------------
int SHIM(pack)(int32_t n, int32_t deg, int32_t from, const float *const planes[6], uint8_t *const out[6],
               int32_t *meta /* numPoints, shDegree, fractionalBits, antialiased, smallestThree */) {
  const spz::GaussianCloud g = makeCloud(n, deg, 1, planes);
  spz::PackOptions o;
  o.from = (spz::CoordinateSystem)from;
  const spz::PackedGaussians p = spz::packGaussians(g, o);
  meta[0] = p.numPoints; meta[1] = p.shDegree; meta[2] = p.fractionalBits; meta[3] = p.antialiased; meta[4] = p.usesQuaternionSmallestThree;
  if (p.numPoints != n || p.positions.size() != (size_t)n * 9) return 1;
  put(out[0], p.positions); put(out[1], p.scales); put(out[2], p.rotations);
  put(out[3], p.alphas); put(out[4], p.colors); put(out[5], p.sh);
  return 0;
}

int SHIM(unpack)(int32_t n, int32_t deg, int32_t fb, int32_t version, int32_t to, const uint8_t *const planes[6],
                 float *const out[6], int32_t *meta /* numPoints, shDegree, antialiased */) {
  const spz::PackedGaussians p = makePacked(n, deg, fb, version, 1, planes);
  spz::UnpackOptions o;
  o.to = (spz::CoordinateSystem)to;
  const spz::GaussianCloud g = spz::unpackGaussians(p, o);
  meta[0] = g.numPoints; meta[1] = g.shDegree; meta[2] = g.antialiased;
  if (g.numPoints != n || g.positions.size() != (size_t)n * 3) return 1;
  put(out[0], g.positions); put(out[1], g.scales); put(out[2], g.rotations);
  put(out[3], g.alphas); put(out[4], g.colors); put(out[5], g.sh);
  return 0;
}

// saveSpz(cloud, opts, vector*) -> malloc'd gzip bytes, or NULL when it returned false
uint8_t *SHIM(save_spz)(int32_t n, int32_t deg, int32_t aa, int32_t from, const float *const planes[6], uint64_t *outSize) {
  const spz::GaussianCloud g = makeCloud(n, deg, aa, planes);
  spz::PackOptions o;
  o.from = (spz::CoordinateSystem)from;
  std::vector<uint8_t> bytes;
  if (!spz::saveSpz(g, o, &bytes)) return nullptr;
  return dupBytes(bytes.data(), bytes.size(), outSize);
}

#ifdef SHIM_HAS_EXTENSIONS
// saveSpzV2 (this repo's extension: first-three rotations under a version-2 header) -> malloc'd gzip bytes
uint8_t *SHIM(save_spz_v2)(int32_t n, int32_t deg, int32_t aa, int32_t from, const float *const planes[6], uint64_t *outSize) {
  const spz::GaussianCloud g = makeCloud(n, deg, aa, planes);
  spz::PackOptions o;
  o.from = (spz::CoordinateSystem)from;
  std::vector<uint8_t> bytes;
  if (!spz::saveSpzV2(g, o, &bytes)) return nullptr;
  return dupBytes(bytes.data(), bytes.size(), outSize);
}
#endif

int SHIM(save_spz_file)(int32_t n, int32_t deg, int32_t aa, int32_t from, const float *const planes[6], const char *path) {
  const spz::GaussianCloud g = makeCloud(n, deg, aa, planes);
  spz::PackOptions o;
  o.from = (spz::CoordinateSystem)from;
  return spz::saveSpz(g, o, std::string(path)) ? 1 : 0;
}

struct CloudHandle { spz::GaussianCloud g; };
struct PackedHandle { spz::PackedGaussians p; };

CloudHandle *SHIM(load_spz)(const uint8_t *data, int32_t size, int32_t to, int32_t viaVector) {
  spz::UnpackOptions o;
  o.to = (spz::CoordinateSystem)to;
  auto *h = new CloudHandle;
  if (viaVector) h->g = spz::loadSpz(std::vector<uint8_t>(data, data + size), o);
  else h->g = spz::loadSpz(data, size, o);
  return h;
}

CloudHandle *SHIM(load_spz_file)(const char *path, int32_t to) {
  spz::UnpackOptions o;
  o.to = (spz::CoordinateSystem)to;
  auto *h = new CloudHandle;
  h->g = spz::loadSpz(std::string(path), o);
  return h;
}

void SHIM(cloud_info)(const CloudHandle *h, int64_t *sizes /* n, deg, aa, 6 plane sizes */) {
  sizes[0] = h->g.numPoints; sizes[1] = h->g.shDegree; sizes[2] = h->g.antialiased;
  sizes[3] = h->g.positions.size(); sizes[4] = h->g.scales.size(); sizes[5] = h->g.rotations.size();
  sizes[6] = h->g.alphas.size(); sizes[7] = h->g.colors.size(); sizes[8] = h->g.sh.size();
}
void SHIM(cloud_copy)(const CloudHandle *h, float *const out[6]) {
  put(out[0], h->g.positions); put(out[1], h->g.scales); put(out[2], h->g.rotations);
  put(out[3], h->g.alphas); put(out[4], h->g.colors); put(out[5], h->g.sh);
}
void SHIM(cloud_free)(CloudHandle *h) { delete h; }

// `threads` host threads pack and unpack the same cloud concurrently (the API is re-entrant: the
// reference has no shared state; this repo leases GPU contexts from a process-wide pool).  Returns the
// number of threads whose result differs from thread 0's, or -1 if any call failed.
int SHIM(concurrent_roundtrip)(int32_t n, int32_t deg, int32_t from, int32_t to, const float *const planes[6], int32_t threads) {
  const spz::GaussianCloud g = makeCloud(n, deg, 0, planes);
  spz::PackOptions po; po.from = (spz::CoordinateSystem)from;
  spz::UnpackOptions uo; uo.to = (spz::CoordinateSystem)to;
  std::vector<spz::PackedGaussians> packed((size_t)threads);
  std::vector<spz::GaussianCloud> back((size_t)threads);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&, t] {
      for (int rep = 0; rep < 3; rep++) {
        packed[t] = spz::packGaussians(g, po);
        back[t] = spz::unpackGaussians(packed[t], uo);
      }
    });
  for (auto &t : pool) t.join();
  int different = 0;
  for (int t = 0; t < threads; t++) {
    if (packed[t].numPoints != n || back[t].numPoints != n) return -1;
    const bool same = packed[t].positions == packed[0].positions && packed[t].rotations == packed[0].rotations &&
                      packed[t].sh == packed[0].sh && packed[t].alphas == packed[0].alphas &&
                      std::memcmp(back[t].sh.data(), back[0].sh.data(), back[0].sh.size() * 4) == 0 &&
                      std::memcmp(back[t].rotations.data(), back[0].rotations.data(), back[0].rotations.size() * 4) == 0;
    different += same ? 0 : 1;
  }
  return different;
}

int SHIM(save_ply)(int32_t n, int32_t deg, int32_t from, const float *const planes[6], const char *path) {
  const spz::GaussianCloud g = makeCloud(n, deg, 0, planes);
  spz::PackOptions o;
  o.from = (spz::CoordinateSystem)from;
  return spz::saveSplatToPly(g, o, std::string(path)) ? 1 : 0;
}

CloudHandle *SHIM(load_ply)(const char *path, int32_t to) {
  spz::UnpackOptions o;
  o.to = (spz::CoordinateSystem)to;
  auto *h = new CloudHandle;
  h->g = spz::loadSplatFromPly(std::string(path), o);
  return h;
}

// ---- host glue, no GPU involved ------------------------------------------------------------------
// which: 0 = bytes, 1 = vector, 2 = file (data = path)
PackedHandle *SHIM(load_packed)(const uint8_t *data, int32_t size, int32_t which) {
  auto *h = new PackedHandle;
  if (which == 0) h->p = spz::loadSpzPacked(data, size);
  else if (which == 1) h->p = spz::loadSpzPacked(std::vector<uint8_t>(data, data + size));
  else h->p = spz::loadSpzPacked(std::string(reinterpret_cast<const char *>(data), (size_t)size));
  return h;
}
void SHIM(packed_info)(const PackedHandle *h, int64_t *m /* n, deg, fb, aa, s3, half, 6 sizes */) {
  const spz::PackedGaussians &p = h->p;
  m[0] = p.numPoints; m[1] = p.shDegree; m[2] = p.fractionalBits; m[3] = p.antialiased;
  m[4] = p.usesQuaternionSmallestThree; m[5] = p.usesFloat16();
  m[6] = p.positions.size(); m[7] = p.scales.size(); m[8] = p.rotations.size();
  m[9] = p.alphas.size(); m[10] = p.colors.size(); m[11] = p.sh.size();
}
void SHIM(packed_copy)(const PackedHandle *h, uint8_t *const out[6]) {
  put(out[0], h->p.positions); put(out[1], h->p.scales); put(out[2], h->p.rotations);
  put(out[3], h->p.alphas); put(out[4], h->p.colors); put(out[5], h->p.sh);
}
void SHIM(packed_free)(PackedHandle *h) { delete h; }

uint8_t *SHIM(serialize)(int32_t n, int32_t deg, int32_t fb, int32_t version, int32_t aa,
                         const uint8_t *const planes[6], uint64_t *outSize) {
  const spz::PackedGaussians p = makePacked(n, deg, fb, version, aa, planes);
  std::stringstream ss;
  spz::serializePackedGaussians(p, &ss);
  const std::string s = ss.str();
  return dupBytes(s.data(), s.size(), outSize);
}

uint8_t *SHIM(gzip)(const uint8_t *data, uint64_t size, uint64_t *outSize) {
  std::vector<uint8_t> out;
  if (!spz::compressGzipped(data, size, &out)) return nullptr;
  return dupBytes(out.data(), out.size(), outSize);
}

#ifdef SHIM_HAS_EXTENSIONS  // this repo only: fused .ply -> .spz, block-parallel zlib (spz_gzip.cc)
uint8_t *SHIM(ply_to_spz)(const char *path, int32_t from, uint64_t *outSize) {
  spz::PackOptions o;
  o.from = (spz::CoordinateSystem)from;
  std::vector<uint8_t> out;
  if (!spz::plyToSpz(std::string(path), o, &out)) return nullptr;
  return dupBytes(out.data(), out.size(), outSize);
}
int SHIM(spz_to_ply)(const uint8_t *data, uint64_t size, int32_t to, const char *path) {
  spz::UnpackOptions o;
  o.to = (spz::CoordinateSystem)to;
  return spz::spzToPly(std::vector<uint8_t>(data, data + size), o, std::string(path)) ? 1 : 0;
}
uint8_t *SHIM(gzip_parallel)(const uint8_t *data, uint64_t size, int32_t threads, uint64_t *outSize) {
  std::vector<uint8_t> out;
  if (!spz::compressGzippedParallel(data, size, threads, &out)) return nullptr;
  return dupBytes(out.data(), out.size(), outSize);
}
uint8_t *SHIM(gunzip)(const uint8_t *data, uint64_t size, int32_t threads, uint64_t *outSize) {
  std::vector<uint8_t> out;
  const bool ok = threads > 0 ? spz::decompressGzippedParallel(data, size, threads, &out) : spz::decompressGzipped(data, size, &out);
  if (!ok) return nullptr;
  return dupBytes(out.data(), out.size(), outSize);
}
#endif

// PackedGaussians::at(i): 65 bytes out in member order position9 rotation4 scale3 color3 alpha1 shR15 shG15 shB15
void SHIM(packed_at)(int32_t n, int32_t deg, int32_t fb, int32_t version, const uint8_t *const planes[6], int32_t i, uint8_t *out65) {
  const spz::PackedGaussians p = makePacked(n, deg, fb, version, 0, planes);
  const spz::PackedGaussian g = p.at(i);
  uint8_t *o = out65;
  std::memcpy(o, g.position.data(), 9); o += 9;
  std::memcpy(o, g.rotation.data(), 4); o += 4;
  std::memcpy(o, g.scale.data(), 3); o += 3;
  std::memcpy(o, g.color.data(), 3); o += 3;
  *o++ = g.alpha;
  std::memcpy(o, g.shR.data(), 15); o += 15;
  std::memcpy(o, g.shG.data(), 15); o += 15;
  std::memcpy(o, g.shB.data(), 15);
}

// PackedGaussians::unpack(i, coordinateConverter(from, to)): 59 floats out (GPU in this repo)
void SHIM(packed_unpack_one)(int32_t n, int32_t deg, int32_t fb, int32_t version, const uint8_t *const planes[6],
                             int32_t i, int32_t from, int32_t to, float *out59) {
  const spz::PackedGaussians p = makePacked(n, deg, fb, version, 0, planes);
  const spz::CoordinateConverter c = spz::coordinateConverter((spz::CoordinateSystem)from, (spz::CoordinateSystem)to);
  const spz::UnpackedGaussian u = p.unpack(i, c);
  float *o = out59;
  std::memcpy(o, u.position.data(), 12); o += 3;
  std::memcpy(o, u.rotation.data(), 16); o += 4;
  std::memcpy(o, u.scale.data(), 12); o += 3;
  std::memcpy(o, u.color.data(), 12); o += 3;
  *o++ = u.alpha;
  std::memcpy(o, u.shR.data(), 60); o += 15;
  std::memcpy(o, u.shG.data(), 60); o += 15;
  std::memcpy(o, u.shB.data(), 60);
}

// PackedGaussians::unpack(idx[k], c) for a list of indices and ANY converter (21 floats: flipP, flipQ,
// flipSh), 59 floats out each.  use_batch != 0 takes this repo's one-launch extension
// (spz::unpackGaussiansAt) where the headers have it; otherwise, and always for the reference, a
// loop over unpack(i, c).  Returns the number of gaussians written.
int32_t SHIM(packed_unpack_many)(int32_t n, int32_t deg, int32_t fb, int32_t version, const uint8_t *const planes[6],
                                 const int32_t *idx, int32_t count, const float *conv21, int32_t use_batch, float *out) {
  const spz::PackedGaussians p = makePacked(n, deg, fb, version, 0, planes);
  spz::CoordinateConverter c;
  std::memcpy(c.flipP.data(), conv21, 12);
  std::memcpy(c.flipQ.data(), conv21 + 3, 12);
  std::memcpy(c.flipSh.data(), conv21 + 6, 60);
  static_assert(sizeof(spz::UnpackedGaussian) == 59 * sizeof(float), "UnpackedGaussian is 59 packed floats");
#ifdef SHIM_HAS_EXTENSIONS
  if (use_batch) {
    const std::vector<spz::UnpackedGaussian> got = spz::unpackGaussiansAt(p, std::vector<int32_t>(idx, idx + count), c);
    if (!got.empty()) std::memcpy(out, got.data(), got.size() * sizeof(spz::UnpackedGaussian));
    return (int32_t)got.size();
  }
#endif
  (void)use_batch;
  for (int32_t k = 0; k < count; k++) {
    const spz::UnpackedGaussian u = p.unpack(idx[k], c);
    std::memcpy(out + (size_t)k * 59, &u, sizeof u);
  }
  return count;
}

// A consumer's walk over a cloud: packed.unpack(idx[k], c) for k = 0..count-1 on ONE PackedGaussians that the caller
// keeps editing while it walks -- before step k (k % every == every - 1) the bytes of the three gaussians after idx[k]
// are XORed with 0x5a in every plane, and at k == count / 2 fractionalBits and the converter's sign of x change.  Any
// read-ahead inside unpack() has to notice all of that; the reference has none, so the two must agree bit for bit.
int32_t SHIM(packed_unpack_walk)(int32_t n, int32_t deg, int32_t fb, int32_t version, const uint8_t *const planes[6],
                                 const int32_t *idx, int32_t count, const float *conv21, int32_t every, float *out) {
  spz::PackedGaussians p = makePacked(n, deg, fb, version, 0, planes);
  spz::CoordinateConverter c;
  std::memcpy(c.flipP.data(), conv21, 12);
  std::memcpy(c.flipQ.data(), conv21 + 3, 12);
  std::memcpy(c.flipSh.data(), conv21 + 6, 60);
  std::vector<uint8_t> *all[6] = {&p.positions, &p.scales, &p.rotations, &p.alphas, &p.colors, &p.sh};
  for (int32_t k = 0; k < count; k++) {
    if (every > 0 && k % every == every - 1) {
      for (std::vector<uint8_t> *v : all) {
        const size_t per = n ? v->size() / (size_t)n : 0;
        for (int32_t g = idx[k] + 1; g <= idx[k] + 3 && g < n; g++)
          for (size_t b = 0; b < per; b++) (*v)[(size_t)g * per + b] ^= 0x5a;
      }
    }
    if (k == count / 2) {
      p.fractionalBits = fb + 3;
      c.flipP[0] = -c.flipP[0];
    }
    const spz::UnpackedGaussian u = p.unpack(idx[k], c);
    std::memcpy(out + (size_t)k * 59, &u, sizeof u);
  }
  return count;
}

// `threads` short-lived host threads, each packing the cloud ONCE and exiting (a request-per-thread
// server).  run_concurrently = 0 starts them one after the other, 2 = one thread does all the packs.  Returns the wall time in
// milliseconds, or -1 if any result differs from the first thread's.
double SHIM(short_lived_threads)(int32_t n, int32_t deg, int32_t from, const float *const planes[6], int32_t threads,
                                 int32_t run_concurrently) {
  const spz::GaussianCloud g = makeCloud(n, deg, 0, planes);
  spz::PackOptions po; po.from = (spz::CoordinateSystem)from;
  std::vector<spz::PackedGaussians> packed((size_t)threads);
  const auto t0 = std::chrono::steady_clock::now();
  if (run_concurrently == 2) {  // the yardstick: the same number of packs from ONE thread
    std::thread th([&] {
      for (int t = 0; t < threads; t++) packed[t] = spz::packGaussians(g, po);
    });
    th.join();
  } else if (run_concurrently) {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back([&, t] { packed[t] = spz::packGaussians(g, po); });
    for (auto &t : pool) t.join();
  } else {
    for (int t = 0; t < threads; t++) {
      std::thread th([&, t] { packed[t] = spz::packGaussians(g, po); });
      th.join();
    }
  }
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  for (int t = 0; t < threads; t++)
    if (packed[t].numPoints != n || packed[t].sh != packed[0].sh || packed[t].rotations != packed[0].rotations ||
        packed[t].positions != packed[0].positions)
      return -1.0;
  return ms;
}

void SHIM(converter)(int32_t from, int32_t to, float *out21) {
  const spz::CoordinateConverter c = spz::coordinateConverter((spz::CoordinateSystem)from, (spz::CoordinateSystem)to);
  std::memcpy(out21, c.flipP.data(), 12);
  std::memcpy(out21 + 3, c.flipQ.data(), 12);
  std::memcpy(out21 + 6, c.flipSh.data(), 60);
}

// in-place convertCoordinates + medianVolume + data() round trip on caller's planes
float SHIM(cloud_ops)(int32_t n, int32_t deg, int32_t from, int32_t to, float *const planes[6]) {
  spz::GaussianCloud g = makeCloud(n, deg, 0, planes);
  g.convertCoordinates((spz::CoordinateSystem)from, (spz::CoordinateSystem)to);
  const GaussianCloudData d = g.data();
  const SpzFloatBuffer bufs[6] = {d.positions, d.scales, d.rotations, d.alphas, d.colors, d.sh};
  for (int k = 0; k < 6; k++) {
    if (bufs[k].count) std::memcpy(planes[k], bufs[k].data, bufs[k].count * sizeof(float));
    delete[] bufs[k].data;
  }
  return g.medianVolume();
}

void SHIM(half_tables)(float *toFloat65536, const float *samples, int32_t nSamples, uint16_t *toHalf) {
  for (uint32_t h = 0; h < 65536; h++) toFloat65536[h] = spz::halfToFloat((spz::Half)h);
  for (int32_t i = 0; i < nSamples; i++) toHalf[i] = spz::floatToHalf(samples[i]);
}

void SHIM(math)(const float *axis3, const float *quatA, const float *quatB, const float *vec3, float *out /* 4+4+3+3+1+1 */) {
  const spz::Quat4f q = spz::axisAngleQuat(spz::vec3f(axis3));
  std::memcpy(out, q.data(), 16);
  const spz::Quat4f prod = spz::times(spz::quat4f(quatA), spz::quat4f(quatB));
  std::memcpy(out + 4, prod.data(), 16);
  const spz::Vec3f rot = spz::times(spz::normalized(spz::quat4f(quatA)), spz::vec3f(vec3));
  std::memcpy(out + 8, rot.data(), 12);
  const spz::Vec3f nv = spz::normalized(spz::vec3f(vec3));
  std::memcpy(out + 11, nv.data(), 12);
  out[14] = spz::norm(spz::vec3f(vec3));
  out[15] = spz::norm(spz::quat4f(quatA));
}

}  // extern "C"
