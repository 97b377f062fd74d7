// PLY import / export of the drop-in API (host I/O; reference behaviour: load-spz.cc:670-934).
// Gaussian-splat PLY files are binary little-endian, all-float vertex records:
//   x y z  nx ny nz  f_dc_0..2  f_rest_0..(3*S-1)  opacity  scale_0..2  rot_0..3
// with f_rest stored channel-major ([C][S]) and rot_0 = w, in the RDF frame.  GaussianCloud wants
// coefficient-major SH ([S][C]) and xyzw quaternions, so both directions are a per-point shuffle.
// This is file parsing around the codec, not codec arithmetic; it stays on the host.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/spz_b200/spz.hpp"

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

bool startsWith(const std::string &s, const char *prefix) { return s.compare(0, std::strlen(prefix), prefix) == 0; }

int degreeForShDim(int dim) { return dim < 3 ? 0 : dim < 8 ? 1 : dim < 15 ? 2 : 3; }

}  // namespace

GaussianCloud loadSplatFromPly(const std::string &filename, const UnpackOptions &o) {
  const char *name = filename.c_str();
  say("[SPZ] Loading: %s", name);
  std::ifstream in(filename, std::ios::binary);
  if (!in.good()) {
    say("[SPZ ERROR] Unable to open: %s", name);
    return {};
  }
  std::string line;
  std::getline(in, line);
  if (line != "ply") {
    say("[SPZ ERROR] %s: not a .ply file", name);
    return {};
  }
  if (!nextHeaderLine(in, &line) || line != "format binary_little_endian 1.0") {
    say("[SPZ ERROR] %s: unsupported .ply format", name);
    return {};
  }
  static const char kVertex[] = "element vertex ";
  if (!nextHeaderLine(in, &line) || !startsWith(line, kVertex)) {
    say("[SPZ ERROR] %s: missing vertex count", name);
    return {};
  }
  char *end = nullptr;
  const long long count = std::strtoll(line.c_str() + sizeof(kVertex) - 1, &end, 10);
  if (end == line.c_str() + sizeof(kVertex) - 1 || count <= 0 || count > 10LL * 1024 * 1024) {
    printf("[SPZ ERROR] %s: invalid vertex count: %lld\n", name, count);
    fflush(stdout);
    return {};
  }
  const size_t numPoints = (size_t)count;
  printf("[SPZ] Loading %lld points\n", count);
  fflush(stdout);

  // property name -> column
  std::unordered_map<std::string, int> column;
  static const char kProp[] = "property float ";
  for (int i = 0;; i++) {
    if (!nextHeaderLine(in, &line)) {
      say("[SPZ ERROR] %s: unexpected EOF while reading header properties.", name);
      return {};
    }
    if (line == "end_header") break;
    if (!startsWith(line, kProp)) {
      say("[SPZ ERROR] %s: unsupported property data type: %s", name, line.c_str());
      return {};
    }
    column[line.substr(sizeof(kProp) - 1)] = i;
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
  const int pos[3] = {col("x"), col("y"), col("z")};
  const int scl[3] = {col("scale_0"), col("scale_1"), col("scale_2")};
  const int rot[4] = {col("rot_1"), col("rot_2"), col("rot_3"), col("rot_0")};  // file is wxyz
  const int alp = col("opacity");
  const int dc[3] = {col("f_dc_0"), col("f_dc_1"), col("f_dc_2")};
  if (missing) return {};
  std::vector<int> rest;
  for (int i = 0; i < 45; i++) {
    const auto it = column.find("f_rest_" + std::to_string(i));
    if (it == column.end()) break;
    rest.push_back(it->second);
  }
  const size_t shDim = rest.size() / 3;
  const size_t width = column.size();

  std::vector<float> rows(numPoints * width);
  in.read(reinterpret_cast<char *>(rows.data()), (std::streamsize)(rows.size() * sizeof(float)));
  if (!in.good()) {
    say("[SPZ ERROR] Unable to load data from: %s", name);
    return {};
  }

  GaussianCloud g;
  g.numPoints = (int32_t)numPoints;
  g.shDegree = degreeForShDim((int)shDim);
  g.positions.resize(numPoints * 3);
  g.scales.resize(numPoints * 3);
  g.rotations.resize(numPoints * 4);
  g.alphas.resize(numPoints);
  g.colors.resize(numPoints * 3);
  g.sh.resize(numPoints * shDim * 3);
  for (size_t p = 0; p < numPoints; p++) {
    const float *row = rows.data() + p * width;
    for (int a = 0; a < 3; a++) {
      g.positions[p * 3 + a] = row[pos[a]];
      g.scales[p * 3 + a] = row[scl[a]];
      g.colors[p * 3 + a] = row[dc[a]];
    }
    for (int a = 0; a < 4; a++) g.rotations[p * 4 + a] = row[rot[a]];
    g.alphas[p] = row[alp];
    float *sh = g.sh.data() + p * shDim * 3;
    for (size_t s = 0; s < shDim; s++)
      for (size_t c = 0; c < 3; c++) sh[s * 3 + c] = row[rest[c * shDim + s]];  // [C][S] -> [S][C]
  }
  g.convertCoordinates(CoordinateSystem::RDF, o.to);
  return g;
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
  std::string header = "ply\nformat binary_little_endian 1.0\nelement vertex " + std::to_string(g.numPoints) + "\n";
  for (const char *f : {"x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"}) header += std::string("property float ") + f + "\n";
  for (size_t i = 0; i < shDim * 3; i++) header += "property float f_rest_" + std::to_string(i) + "\n";
  for (const char *f : {"opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"}) header += std::string("property float ") + f + "\n";
  header += "end_header\n";
  out.write(header.data(), (std::streamsize)header.size());

  constexpr size_t kBatch = 4096;  // points per write
  std::vector<float> rows(kBatch * width);
  for (size_t base = 0; base < n; base += kBatch) {
    const size_t m = std::min(kBatch, n - base);
    for (size_t k = 0; k < m; k++) {
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
