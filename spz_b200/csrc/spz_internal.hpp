// Internals shared by the host-side sources of the drop-in API (spz_api.cc, spz_ply.cc).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/spz_b200.h"
#include "../../include/spz_b200/spz.hpp"

namespace spz {
namespace detail {

// Column layout of a gaussian-splat .ply body: every property is one float.
struct PlyLayout {
  int64_t numPoints = 0;
  int width = 0;   // floats per vertex record
  int shDim = 0;   // SH coefficients per channel present in the file (f_rest count / 3)
  int pos[3], scale[3], rot[4] /* x, y, z, w */, alpha, color[3];
  std::vector<int> rest;  // columns of f_rest_0 .. in file order ([channel][coefficient])
};

// Parses the header and reads the vertex records (load-spz.cc:691-806).  Logs and returns false the
// way loadSplatFromPly does on any malformed input.
bool readPlyRows(const std::string &filename, PlyLayout *layout, std::vector<float> *rows);

SpzB200Context *contextFor(int32_t device);           // this thread's context (spz_api.cc)
std::vector<int32_t> configuredDevices();             // SPZ_B200_DEVICE(S)
bool finishSpz(const PackedGaussians &packed, std::vector<uint8_t> *out);  // container + gzip (serial or SPZ_B200_GZIP_THREADS)
void logLine(const char *fmt, ...);

}  // namespace detail
}  // namespace spz
