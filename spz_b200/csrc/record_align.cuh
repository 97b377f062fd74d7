// Realignment of per-gaussian byte records inside the packed planes (used by pergaussian_kernels.cu).
//
// Gaussian g's B bytes of a packed plane (B = 9 positions, 3 scales / colours / v2 rotations, 6 half
// positions, 3*D spherical harmonics) start at byte B*g, i.e. at offset (B*g) mod 4 inside a 32-bit
// word.  loadRecord brings them to byte 0 of a register array with funnel shifts; emitRecordWords
// does the reverse, completing the word a gaussian shares with its predecessor from the
// predecessor's last four bytes (`prevTail`, which the kernel passes down with one warp shuffle;
// B*32 is a multiple of 4, so lane 0 of a warp always starts a word).  All memory traffic is whole
// words.  Like codec_math.cuh the header also compiles for the host (tests/host_emul) so the index
// arithmetic is checked without a GPU; nothing in the product runs it on the CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SPZ_REC_HD __host__ __device__ __forceinline__
#else
#define SPZ_REC_HD inline
#endif

namespace spzb200 {

#if defined(__CUDA_ARCH__)
SPZ_REC_HD uint32_t funnelR(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
SPZ_REC_HD uint32_t funnelL(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }
#else
// (hi:lo) >> s and the high word of (hi:lo) << s, shift taken mod 32 as the SHF instruction does
SPZ_REC_HD uint32_t funnelR(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (s & 31u)); }
SPZ_REC_HD uint32_t funnelL(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)((((((uint64_t)hi) << 32) | lo) << (s & 31u)) >> 32); }
#endif

template <int B>
struct Rec {
  static constexpr int kMaxShift = (B % 4 == 0) ? 0 : (B % 2 == 0) ? 2 : 3;
  static constexpr int NL = (B + kMaxShift + 3) / 4;  // words a lane reads to cover its B bytes at any alignment
  static constexpr int NV = (B + 3) / 4;              // words holding one record, byte 0 first
};

// v <- the B bytes of gaussian g of a packed plane (record byte k = byte k & 3 of v[k >> 2]).
// Reads words (B*g) >> 2 .. + NL - 1: for B = 3 the last gaussian of a plane reads one word past it.
template <int B>
SPZ_REC_HD void loadRecord(const uint32_t *plane, int g, uint32_t (&v)[Rec<B>::NL]) {
  constexpr int NL = Rec<B>::NL;
  const uint32_t *p = plane + ((B * g) >> 2);
  uint32_t w[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) w[i] = p[i];
  if constexpr (B % 4 == 0) {
#pragma unroll
    for (int i = 0; i < NL; i++) v[i] = w[i];
  } else {
    const uint32_t sh = ((uint32_t)(B * g) & 3u) * 8u;
#pragma unroll
    for (int i = 0; i + 1 < NL; i++) v[i] = funnelR(w[i], w[i + 1], sh);
    v[NL - 1] = w[NL - 1] >> sh;
  }
}

// record bytes B-4 .. B-1 (bytes before the record's start, for B < 4, are don't-care)
template <int B>
SPZ_REC_HD uint32_t recordTail(const uint32_t (&v)[Rec<B>::NV]) {
  if constexpr (B >= 4) {
    constexpr int NV = Rec<B>::NV;
    constexpr int j = (B - 4) >> 2;
    constexpr uint32_t s = ((B - 4) & 3) * 8;
    return funnelR(v[j], v[j + 1 < NV ? j + 1 : j], s);
  } else {
    return v[0] << (8 * (4 - B));
  }
}

// the B bytes in v -> gaussian g's place in a packed plane.  Gaussian g stores the words that END
// inside its byte range [B*g, B*(g+1)); bytes at or above B in v's last word are never stored.
// prevTail = recordTail of gaussian g - 1 (unused when B*g is a multiple of 4).
template <int B>
SPZ_REC_HD void emitRecordWords(uint32_t *plane, int g, const uint32_t (&v)[Rec<B>::NV], uint32_t prevTail) {
  constexpr int NV = Rec<B>::NV;
  uint32_t *p = plane + ((B * g) >> 2);
  if constexpr (B % 4 == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) p[i] = v[i];
  } else {
    const uint32_t sh = ((uint32_t)(B * g) & 3u) * 8u;
    uint32_t w[NV];
    w[0] = funnelL(prevTail, v[0], sh);
#pragma unroll
    for (int i = 1; i < NV; i++) w[i] = funnelL(v[i - 1], v[i], sh);
#pragma unroll
    for (int i = 0; i + 1 < NV; i++) p[i] = w[i];
    if (((B * (g + 1)) >> 2) - ((B * g) >> 2) == NV) p[NV - 1] = w[NV - 1];
  }
}

}  // namespace spzb200
