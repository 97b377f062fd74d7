// PLY import / export of the drop-in API (host I/O; reference behaviour: load-spz.cc:670-934).
// Gaussian-splat PLY files are binary little-endian, all-float vertex records:
//   x y z  nx ny nz  f_dc_0..2  f_rest_0..(3*S-1)  opacity  scale_0..2  rot_0..3
// with f_rest stored channel-major ([C][S]) and rot_0 = w, in the RDF frame.  GaussianCloud wants
// coefficient-major SH ([S][C]) and xyzw quaternions, so both directions are a per-point shuffle.
// This is file parsing around the codec, not codec arithmetic; it stays on the host.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <fstream>
#include <algorithm>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "spz_internal.hpp"

namespace spz {
namespace {

void say(const char *fmt, const char *a, const char *b = "") {
  printf(fmt, a, b);
  printf("\n");
  fflush(stdout);
}

// Next header line that is neither blank nor a comment, with leading whitespace removed.
bool nextHeaderLine(std::istream &in, std::string *line) {
  std::string raw;
  while (std::getline(in, raw)) {
    const size_t first = raw.find_first_not_of(" \t\n\r\f\v");
    if (first == std::string::npos) continue;
    if (raw.compare(first, 7, "comment") == 0) continue;
    *line = raw.substr(first);
    return true;
  }
  return false;
}

// fn(a, b) over [0, n) cut into contiguous ranges, one per host thread (small n: the calling thread alone).  The
// per-point shuffles below are memory-bound loops over GBs for a large cloud; one thread moves ~2 GB/s of them.
template <class Fn>
void overPointRanges(size_t n, Fn &&fn) {
  const size_t threads = n < 65536 ? 1 : std::min<size_t>({(size_t)16, std::max<size_t>(1, std::thread::hardware_concurrency()), n / 32768});
  if (threads <= 1) {
    fn((size_t)0, n);
    return;
  }
  std::vector<std::thread> pool;
  size_t started = 1;
  try {
    for (; started < threads; started++) pool.emplace_back([&fn, n, started, threads] { fn(n * started / threads, n * (started + 1) / threads); });
  } catch (const std::exception &) {
    // no more threads to be had: the calling thread takes the ranges that did not get one
  }
  fn((size_t)0, n / threads);
  for (size_t t = started; t < threads; t++) fn(n * t / threads, n * (t + 1) / threads);
  for (auto &th : pool) th.join();
}

bool startsWith(const std::string &s, const char *prefix) { return s.compare(0, std::strlen(prefix), prefix) == 0; }

int degreeForShDim(int dim) { return dim < 3 ? 0 : dim < 8 ? 1 : dim < 15 ? 2 : 3; }

// The header saveSplatToPly writes (load-spz.cc:892-918): x y z, zero normals, f_dc, f_rest
// channel-major, opacity, scales, rot wxyz -- 17 + 3 * shDim float properties.
std::string plyHeader(long long numPoints, size_t shDim) {
  std::string header = "ply\nformat binary_little_endian 1.0\nelement vertex " + std::to_string(numPoints) + "\n";
  for (const char *f : {"x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"}) header += std::string("property float ") + f + "\n";
  for (size_t i = 0; i < shDim * 3; i++) header += "property float f_rest_" + std::to_string(i) + "\n";
  for (const char *f : {"opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"}) header += std::string("property float ") + f + "\n";
  header += "end_header\n";
  return header;
}

}  // namespace

namespace detail {

bool readPlyRows(const std::string &filename, PlyLayout *layout, std::vector<float> *rows) {
  const char *name = filename.c_str();
  say("[SPZ] Loading: %s", name);
  std::ifstream in(filename, std::ios::binary);
  if (!in.good()) {
    say("[SPZ ERROR] Unable to open: %s", name);
    return false;
  }
  std::string line;
  std::getline(in, line);
  if (line != "ply") {
    say("[SPZ ERROR] %s: not a .ply file", name);
    return false;
  }
  if (!nextHeaderLine(in, &line) || line != "format binary_little_endian 1.0") {
    say("[SPZ ERROR] %s: unsupported .ply format", name);
    return false;
  }
  static const char kVertex[] = "element vertex ";
  if (!nextHeaderLine(in, &line) || !startsWith(line, kVertex)) {
    say("[SPZ ERROR] %s: missing vertex count", name);
    return false;
  }
  char *end = nullptr;
  const long long count = std::strtoll(line.c_str() + sizeof(kVertex) - 1, &end, 10);
  if (end == line.c_str() + sizeof(kVertex) - 1 || count <= 0 || count > 10LL * 1024 * 1024) {
    printf("[SPZ ERROR] %s: invalid vertex count: %lld\n", name, count);
    fflush(stdout);
    return false;
  }
  printf("[SPZ] Loading %lld points\n", count);
  fflush(stdout);

  // property name -> column
  std::unordered_map<std::string, int> column;
  static const char kProp[] = "property float ";
  int properties = 0;
  for (int i = 0;; i++) {
    if (!nextHeaderLine(in, &line)) {
      say("[SPZ ERROR] %s: unexpected EOF while reading header properties.", name);
      return false;
    }
    if (line == "end_header") break;
    if (!startsWith(line, kProp)) {
      say("[SPZ ERROR] %s: unsupported property data type: %s", name, line.c_str());
      return false;
    }
    column[line.substr(sizeof(kProp) - 1)] = i;
    properties = i + 1;
  }
  bool missing = false;
  auto col = [&](const char *field) {
    const auto it = column.find(field);
    if (it == column.end()) {
      say("[SPZ ERROR] Missing field: %s", field);
      missing = true;
      return -1;
    }
    return it->second;
  };
  PlyLayout &L = *layout;
  L.pos[0] = col("x"); L.pos[1] = col("y"); L.pos[2] = col("z");
  L.scale[0] = col("scale_0"); L.scale[1] = col("scale_1"); L.scale[2] = col("scale_2");
  L.rot[0] = col("rot_1"); L.rot[1] = col("rot_2"); L.rot[2] = col("rot_3"); L.rot[3] = col("rot_0");  // file is wxyz
  L.alpha = col("opacity");
  L.color[0] = col("f_dc_0"); L.color[1] = col("f_dc_1"); L.color[2] = col("f_dc_2");
  if (missing) return false;
  L.rest.clear();
  for (int i = 0; i < 45; i++) {
    const auto it = column.find("f_rest_" + std::to_string(i));
    if (it == column.end()) break;
    L.rest.push_back(it->second);
  }
  L.shDim = (int)(L.rest.size() / 3);
  // floats per vertex record = property lines, not distinct names: the reference sizes records by its map
  // (load-spz.cc:738), so a header that repeats a name makes it index past the end of its buffer; here every
  // column index stays inside the record
  L.width = properties;
  L.numPoints = count;

  rows->clear();
  detail::resizeUninitialized(*rows, (size_t)count * (size_t)L.width);  // filled by the read below
  in.read(reinterpret_cast<char *>(rows->data()), (std::streamsize)(rows->size() * sizeof(float)));
  if (!in.good()) {
    say("[SPZ ERROR] Unable to load data from: %s", name);
    return false;
  }
  return true;
}

}  // namespace detail

GaussianCloud loadSplatFromPly(const std::string &filename, const UnpackOptions &o) {
  detail::PlyLayout L;
  std::vector<float> rows;
  if (!detail::readPlyRows(filename, &L, &rows)) return {};
  const size_t numPoints = (size_t)L.numPoints, shDim = (size_t)L.shDim, width = (size_t)L.width;

  GaussianCloud g;
  g.numPoints = (int32_t)numPoints;
  g.shDegree = degreeForShDim((int)shDim);
  // every element is assigned in the loop below
  detail::resizeUninitialized(g.positions, numPoints * 3);
  detail::resizeUninitialized(g.scales, numPoints * 3);
  detail::resizeUninitialized(g.rotations, numPoints * 4);
  detail::resizeUninitialized(g.alphas, numPoints);
  detail::resizeUninitialized(g.colors, numPoints * 3);
  detail::resizeUninitialized(g.sh, numPoints * shDim * 3);
  // The reference shuffles, then runs convertCoordinates(RDF, to) over the finished cloud (load-spz.cc:818-848): every
  // position, quaternion xyz and SH value multiplied by its +-1 factor.  Same multiplications here, applied to each
  // point as it is shuffled, on several threads.
  const CoordinateConverter cv = coordinateConverter(CoordinateSystem::RDF, o.to);
  overPointRanges(numPoints, [&](size_t first, size_t last) {
    for (size_t p = first; p < last; p++) {
      const float *row = rows.data() + p * width;
      for (int a = 0; a < 3; a++) {
        g.positions[p * 3 + a] = row[L.pos[a]] * cv.flipP[a];
        g.scales[p * 3 + a] = row[L.scale[a]];
        g.colors[p * 3 + a] = row[L.color[a]];
      }
      for (int a = 0; a < 3; a++) g.rotations[p * 4 + a] = row[L.rot[a]] * cv.flipQ[a];
      g.rotations[p * 4 + 3] = row[L.rot[3]];
      g.alphas[p] = row[L.alpha];
      float *sh = g.sh.data() + p * shDim * 3;
      for (size_t s = 0; s < shDim; s++)
        for (size_t c = 0; c < 3; c++) sh[s * 3 + c] = row[L.rest[c * shDim + s]] * cv.flipSh[s];  // [C][S] -> [S][C]
    }
  });
  return g;
}

// Extension: .ply file -> .spz bytes without materialising the planar GaussianCloud.  The vertex
// records go to the GPU as they lie in the file and the fused kernel (ply_kernels.cu) shuffles and
// quantises them: same bytes as saveSpz(loadSplatFromPly(file, {to}), {from = to}) for any `to`
// when options.from == RDF, and as the reference's ply_to_spz tool (no conversion either way) when
// options.from == UNSPECIFIED.
bool plyToSpz(const std::string &plyFilename, const PackOptions &options, std::vector<uint8_t> *output) {
  detail::PlyLayout L;
  std::vector<float> rows;
  if (!detail::readPlyRows(plyFilename, &L, &rows)) return false;
  const int degree = degreeForShDim(L.shDim);
  const int usedDim = degree == 0 ? 0 : degree == 1 ? 3 : degree == 2 ? 8 : 15;
  if (usedDim != L.shDim) {
    // a file with e.g. 4 coefficients per channel: the loader would keep them all while declaring
    // degree 1, which packGaussians then rejects (load-spz.cc:115); same verdict here
    detail::logLine("[SPZ: ERROR] Check failed: %s:%d: sh coefficient count %d does not match a degree", __FILE__, __LINE__, L.shDim);
    return detail::finishSpz(PackedGaussians{}, output);
  }
  const size_t n = (size_t)L.numPoints;
  PackedGaussians packed;
  packed.numPoints = (int32_t)n;
  packed.shDegree = degree;
  packed.fractionalBits = 12;
  packed.antialiased = false;
  packed.usesQuaternionSmallestThree = true;
  detail::resizeUninitialized(packed.positions, n * 9);
  detail::resizeUninitialized(packed.scales, n * 3);
  detail::resizeUninitialized(packed.rotations, n * 4);
  detail::resizeUninitialized(packed.alphas, n);
  detail::resizeUninitialized(packed.colors, n * 3);
  detail::resizeUninitialized(packed.sh, n * (size_t)usedDim * 3);

  SpzB200PlyRows in;
  std::memset(&in, 0, sizeof in);
  in.num_points = (int64_t)n;
  in.width = L.width;
  in.sh_degree = degree;
  in.rows = rows.data();
  for (int a = 0; a < 3; a++) { in.col_pos[a] = L.pos[a]; in.col_scale[a] = L.scale[a]; in.col_color[a] = L.color[a]; }
  for (int a = 0; a < 4; a++) in.col_rot[a] = L.rot[a];
  in.col_alpha = L.alpha;
  // the loader reads channel c, coefficient s from rest[c * shDim + s] with shDim = coefficients in the FILE
  for (int c = 0; c < 3; c++)
    for (int s = 0; s < usedDim; s++) in.col_rest[c * usedDim + s] = L.rest[(size_t)(c * L.shDim + s)];
  SpzB200Packed out;
  std::memset(&out, 0, sizeof out);
  out.num_points = (int64_t)n;
  out.sh_degree = degree;
  out.positions = packed.positions.data(); out.scales = packed.scales.data(); out.rotations = packed.rotations.data();
  out.alphas = packed.alphas.data(); out.colors = packed.colors.data(); out.sh = packed.sh.data();
  detail::ContextLease lease(detail::configuredDevices()[0]);
  if (!lease.get()) return false;
  if (spzb200_encode_ply_host(lease.get(), &in, (int32_t)options.from, &out, nullptr) != SPZB200_OK) {
    detail::logLine("[SPZ ERROR] spz_b200: %s", spzb200_last_error());
    return false;
  }
  return detail::finishSpz(packed, output);
}

// Extension: .spz bytes -> .ply file without materialising the planar GaussianCloud: the packed
// planes go to the GPU and come back as finished vertex records (ply_kernels.cu).  Same file as
// saveSplatToPly(loadSpz(bytes, {to = X}), {from = X}, f) when options.to == RDF, and as the
// reference's spz_to_ply tool (no conversion either way) when options.to == UNSPECIFIED.
bool spzToPly(const std::vector<uint8_t> &spzBytes, const UnpackOptions &options, const std::string &plyFilename) {
  const PackedGaussians packed = loadSpzPacked(spzBytes);
  const size_t n = packed.numPoints < 0 ? 0 : (size_t)packed.numPoints;
  const int degree = packed.shDegree;
  const size_t shDim = degree == 0 ? 0 : degree == 1 ? 3 : degree == 2 ? 8 : 15;
  const size_t width = 17 + 3 * shDim;
  std::vector<float> rows;
  detail::resizeUninitialized(rows, n * width);  // the decoder writes every float, unmapped columns as 0
  if (n > 0) {
    SpzB200Packed in;
    std::memset(&in, 0, sizeof in);
    in.num_points = (int64_t)n;
    in.sh_degree = degree;
    in.fractional_bits = packed.fractionalBits;
    const bool half = packed.usesFloat16();
    in.version = half ? (packed.usesQuaternionSmallestThree ? SPZB200_STREAM_HALF_POSITIONS_SMALLEST_THREE : SPZB200_STREAM_V1)
                      : (packed.usesQuaternionSmallestThree ? SPZB200_STREAM_V3 : SPZB200_STREAM_V2);
    in.positions = const_cast<uint8_t *>(packed.positions.data()); in.scales = const_cast<uint8_t *>(packed.scales.data());
    in.rotations = const_cast<uint8_t *>(packed.rotations.data()); in.alphas = const_cast<uint8_t *>(packed.alphas.data());
    in.colors = const_cast<uint8_t *>(packed.colors.data()); in.sh = const_cast<uint8_t *>(packed.sh.data());
    SpzB200PlyRows out;
    std::memset(&out, 0, sizeof out);
    out.num_points = (int64_t)n;
    out.width = (int32_t)width;
    out.sh_degree = degree;
    out.rows = rows.data();
    // the writer's column order: x y z nx ny nz f_dc[3] f_rest[3 * shDim] opacity scale[3] rot_0(w) rot_1..3(xyz)
    for (int a = 0; a < 3; a++) { out.col_pos[a] = a; out.col_color[a] = 6 + a; out.col_scale[a] = (int32_t)(10 + 3 * shDim) + a; }
    for (size_t i = 0; i < 3 * shDim; i++) out.col_rest[i] = (int32_t)(9 + i);
    out.col_alpha = (int32_t)(9 + 3 * shDim);
    const int32_t rot0 = (int32_t)(13 + 3 * shDim);
    out.col_rot[3] = rot0;  // w
    for (int a = 0; a < 3; a++) out.col_rot[a] = rot0 + 1 + a;
    detail::ContextLease lease(detail::configuredDevices()[0]);
    if (!lease.get()) return false;
    if (spzb200_decode_ply_host(lease.get(), &in, (int32_t)options.to, &out, nullptr) != SPZB200_OK) {
      detail::logLine("[SPZ ERROR] spz_b200: %s", spzb200_last_error());
      return false;
    }
  }
  std::ofstream file(plyFilename, std::ios::binary);
  if (!file.good()) {
    say("[SPZ ERROR] Unable to open for writing: %s", plyFilename.c_str());
    return false;
  }
  const std::string header = plyHeader((long long)n, shDim);
  file.write(header.data(), (std::streamsize)header.size());
  file.write(reinterpret_cast<const char *>(rows.data()), (std::streamsize)(rows.size() * sizeof(float)));
  file.close();
  if (!file.good()) {
    say("[SPZ ERROR] Failed to write to: %s", plyFilename.c_str());
    return false;
  }
  return true;
}

bool saveSplatToPly(const GaussianCloud &g, const PackOptions &o, const std::string &filename) {
  const size_t n = g.numPoints < 0 ? 0 : (size_t)g.numPoints;
  auto sized = [&](const std::vector<float> &v, size_t per, const char *what) {
    if (v.size() == n * per) return true;
    printf("[SPZ: ERROR] Check failed: %s:%d: %s.size() == N * %zu\n", __FILE__, __LINE__, what, per);
    fflush(stdout);
    return false;
  };
  if (!sized(g.positions, 3, "positions") || !sized(g.scales, 3, "scales") || !sized(g.rotations, 4, "rotations") ||
      !sized(g.alphas, 1, "alphas") || !sized(g.colors, 3, "colors"))
    return false;
  // (an empty cloud divides by zero in the reference, load-spz.cc:853; here it writes a 0-vertex file)
  const size_t shDim = n ? g.sh.size() / n / 3 : 0;
  const size_t width = 17 + shDim * 3;
  const CoordinateConverter c = coordinateConverter(o.from, CoordinateSystem::RDF);

  std::ofstream out(filename, std::ios::binary);
  if (!out.good()) {
    say("[SPZ ERROR] Unable to open for writing: %s", filename.c_str());
    return false;
  }
  const std::string header = plyHeader(g.numPoints, shDim);
  out.write(header.data(), (std::streamsize)header.size());

  const size_t kBatch = n < 65536 ? 4096 : 262144;  // points per write; large batches are filled by several threads
  std::vector<float> rows(std::min(kBatch, std::max<size_t>(n, 1)) * width);
  for (size_t base = 0; base < n; base += kBatch) {
    const size_t m = std::min(kBatch, n - base);
    overPointRanges(m, [&](size_t first, size_t last) {
    for (size_t k = first; k < last; k++) {
      const size_t p = base + k;
      float *row = rows.data() + k * width;
      for (int a = 0; a < 3; a++) {
        row[a] = c.flipP[a] * g.positions[p * 3 + a];
        row[3 + a] = 0.0f;  // normals: always zero, some viewers want the columns
        row[6 + a] = g.colors[p * 3 + a];
      }
      const float *sh = g.sh.data() + p * shDim * 3;
      float *rest = row + 9;
      for (size_t ch = 0; ch < 3; ch++)
        for (size_t s = 0; s < shDim; s++) rest[ch * shDim + s] = (s < 15 ? c.flipSh[s] : 1.0f) * sh[s * 3 + ch];  // [S][C] -> [C][S]
      float *tail = rest + shDim * 3;
      tail[0] = g.alphas[p];
      for (int a = 0; a < 3; a++) tail[1 + a] = g.scales[p * 3 + a];
      tail[4] = g.rotations[p * 4 + 3];  // w first
      for (int a = 0; a < 3; a++) tail[5 + a] = c.flipQ[a] * g.rotations[p * 4 + a];
    }
    });
    out.write(reinterpret_cast<const char *>(rows.data()), (std::streamsize)(m * width * sizeof(float)));
  }
  out.close();
  if (!out.good()) {
    say("[SPZ ERROR] Failed to write to: %s", filename.c_str());
    return false;
  }
  return true;
}

}  // namespace spz
