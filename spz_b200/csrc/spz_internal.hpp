// Internals shared by the host-side sources of the drop-in API (spz_api.cc, spz_ply.cc).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <type_traits>
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

// v.resize(n) without the value-initialisation pass, for vectors the codec is about to overwrite
// completely.  The reference's value-returning signatures force std::vector outputs, and at 10M
// SH-degree-3 points resize() spends 1.7 s zero-filling 2.36 GB on one thread -- twenty times the
// whole H2D | kernel | D2H pipeline.  reserve() only maps address space;
// the pages are then first touched by the pipeline's copy threads, in parallel.  libstdc++ and libc++
// both lay a vector out as {begin, end, end-of-storage}; the layout is probed once and anything else
// falls back to resize().  `v` must be empty on entry.  Large planes are also marked for transparent
// huge pages, which the image's kernels grant on request (THP mode "madvise").
bool vectorLayoutIsThreePointers();
void adviseHugePages(void *p, size_t bytes);  // MADV_HUGEPAGE on the 2 MiB-aligned interior (Linux; no-op elsewhere)
template <class T>
void resizeUninitialized(std::vector<T> &v, size_t n) {
  static_assert(std::is_trivial<T>::value, "only for trivially constructible and destructible elements");
  static const bool hack = vectorLayoutIsThreePointers();
  if (!hack || !v.empty() || n == 0) {
    v.resize(n);
    return;
  }
  v.reserve(n);
  T *raw[3];
  static_assert(sizeof(std::vector<T>) == sizeof raw, "std::vector is not three pointers wide");
  std::memcpy(raw, &v, sizeof raw);
  raw[1] = raw[0] + n;
  std::memcpy(&v, raw, sizeof raw);
  // first touch then faults 2 MiB at a time instead of 4 KiB (3x faster fill of a fresh 2.4 GB plane)
  if (n * sizeof(T) >= ((size_t)32 << 20)) adviseHugePages(raw[0], n * sizeof(T));
}

SpzB200Context *contextFor(int32_t device);           // this thread's context (spz_api.cc)
std::vector<int32_t> configuredDevices();             // SPZ_B200_DEVICE(S)
bool finishSpz(const PackedGaussians &packed, std::vector<uint8_t> *out);  // container + gzip (serial or SPZ_B200_GZIP_THREADS)
void logLine(const char *fmt, ...);

}  // namespace detail
}  // namespace spz
