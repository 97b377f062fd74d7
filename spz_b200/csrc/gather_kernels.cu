// Batched per-gaussian access: PackedGaussians::at(i) + PackedGaussian::unpack
// (load-spz.cc:383-463) for a LIST of gaussians in one launch.
//
// The unit is the reference's own pair of structs, so callers hand over arrays of them unchanged:
//   in   PackedGaussian   (load-spz.h:28-37)  65 bytes, alignment 1:
//          position[9] rotation[4] scale[3] color[3] alpha shR[15] shG[15] shB[15]
//   out  UnpackedGaussian (load-spz.h:13-24)  59 floats:
//          position[3] rotation[4] (x, y, z, w) scale[3] color[3] alpha shR[15] shG[15] shB[15]
//
// Two sources of the 65-byte record:
//   RecordSource  -- an array of PackedGaussian (what at() returns; the host entry point gathers them
//                    into a pinned buffer with plain byte copies and ships that)
//   PlaneSource   -- device-resident packed planes + an index list: thread g performs at(idx[g]) itself
//                    (SH de-interleaved per channel, padded with 128 = the code of 0.0 up to 15 coefficients)
//
// Unlike unpackGaussians, this path takes the caller's CoordinateConverter as 21 FLOATS and multiplies
// by them in the reference's order -- c.flipP[i] * (float)fixed * scale (load-spz.cc:401), the
// quaternion's xyz *= flipQ after the square root (:377-379) resp. before it for the first-three form
// (:335-343), c.flipSh[j] * unquantizeSH(b) (:426-428) -- because a hand-built converter may hold any
// value, not just +-1.  For +-1 the results equal the bulk decoder's bit for bit.
//
// One thread per gaussian, 128 per CTA.  The tile's records are staged in shared memory (coalesced
// word loads for RecordSource), each thread expands its record into a 59-float row of a shared-memory
// tile (stride 59 words: conflict free), and the tile leaves as contiguous 128-bit stores.
#include "codec_kernels.cuh"

#include "codec_math.cuh"
#include "kernel_utils.cuh"

namespace spzb200 {
namespace {

constexpr int kG = 128;             // gaussians per CTA
constexpr int kRec = 65;            // sizeof(PackedGaussian)
constexpr int kOut = 59;            // floats of UnpackedGaussian
constexpr int kRecBytes = kG * kRec;  // 8320, a multiple of 16

// byte offsets inside PackedGaussian / float offsets inside UnpackedGaussian
constexpr int rPos = 0, rRot = 9, rScale = 13, rColor = 16, rAlpha = 19, rShR = 20, rShG = 35, rShB = 50;
constexpr int uPos = 0, uRot = 3, uScale = 7, uColor = 10, uAlpha = 13, uShR = 14, uShG = 29, uShB = 44;

struct RecordSource {
  static __device__ __forceinline__ void stage(const GatherArgs &a, long long g0, int count, unsigned char *rec, int t) {
    const unsigned char *src = a.records + g0 * kRec;
    const int bytes = count * kRec;
    if ((reinterpret_cast<uintptr_t>(src) & 3u) == 0) {
      const int words = bytes >> 2;
      for (int i = t; i < words; i += kG) reinterpret_cast<uint32_t *>(rec)[i] = __ldg(reinterpret_cast<const uint32_t *>(src) + i);
      for (int i = (words << 2) + t; i < bytes; i += kG) rec[i] = __ldg(src + i);
    } else {
      for (int i = t; i < bytes; i += kG) rec[i] = __ldg(src + i);
    }
  }
};

struct PlaneSource {
  // at(i), load-spz.cc:431-459: a pure byte gather
  static __device__ __forceinline__ void stage(const GatherArgs &a, long long g0, int count, unsigned char *rec, int t) {
    if (t >= count) return;
    const long long i = a.indices ? a.indices[g0 + t] : g0 + t;
    unsigned char *r = rec + t * kRec;
    const bool half = a.version == 1 || a.version == 4;
    const int posBytes = half ? 6 : 9, rotBytes = a.version >= 3 ? 4 : 3;
    for (int k = 0; k < 9; k++) r[rPos + k] = k < posBytes ? a.positions[i * posBytes + k] : 0;
    for (int k = 0; k < 4; k++) r[rRot + k] = k < rotBytes ? a.rotations[i * rotBytes + k] : 0;
    for (int k = 0; k < 3; k++) {
      r[rScale + k] = a.scales[i * 3 + k];
      r[rColor + k] = a.colors[i * 3 + k];
    }
    r[rAlpha] = a.alphas[i];
    const unsigned char *s = a.sh + i * (3 * a.shDim);
    for (int j = 0; j < 15; j++) {
      const bool have = j < a.shDim;
      r[rShR + j] = have ? s[3 * j] : 128;
      r[rShG + j] = have ? s[3 * j + 1] : 128;
      r[rShB + j] = have ? s[3 * j + 2] : 128;
    }
  }
};

template <class Src>
__global__ void __launch_bounds__(kG)
unpackRecordsKernel(const GatherArgs a) {
  __shared__ __align__(16) unsigned char rec[kRecBytes];
  __shared__ __align__(16) float tile[kG * kOut];
  const int t = threadIdx.x;
  const long long g0 = (long long)blockIdx.x * kG;
  const int count = (int)(a.n - g0 < kG ? a.n - g0 : kG);
  Src::stage(a, g0, count, rec, t);
  __syncthreads();
  if (t < count) {
    const unsigned char *r = rec + t * kRec;
    float *u = tile + t * kOut;
    // positions, load-spz.cc:392-404
    if (a.version == 1 || a.version == 4) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const uint32_t h = (uint32_t)r[rPos + 2 * i] | ((uint32_t)r[rPos + 2 * i + 1] << 8);
        u[uPos + i] = m::mul(a.flipP[i], m::half_bits_to_float(h));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const uint32_t lo24 = (uint32_t)r[rPos + 3 * i] | ((uint32_t)r[rPos + 3 * i + 1] << 8) | ((uint32_t)r[rPos + 3 * i + 2] << 16);
        const int32_t fixed = (int32_t)(lo24 << 8) >> 8;
        u[uPos + i] = m::mul(m::mul(a.flipP[i], m::i2f(fixed)), a.positionScale);
      }
    }
    // scales :406-408, colours :421-423, alpha :419
#pragma unroll
    for (int i = 0; i < 3; i++) {
      u[uScale + i] = m::dequant_scale(r[rScale + i]);
      u[uColor + i] = __ldg(a.tables + 256 + r[rColor + i]);
    }
    u[uAlpha] = __ldg(a.tables + r[rAlpha]);
    // rotations
    float q[4];
    if (a.version >= 3) {  // :347-381
      uint32_t comp = (uint32_t)r[rRot] | ((uint32_t)r[rRot + 1] << 8) | ((uint32_t)r[rRot + 2] << 16) | ((uint32_t)r[rRot + 3] << 24);
      const uint32_t big = comp >> 30;
      float sum = 0.0f;
#pragma unroll
      for (int i = 3; i >= 0; --i) {
        if ((uint32_t)i != big) {
          const float v = m::bitsf(m::fbits(__ldg(a.tables + 512 + (comp & 511u))) | ((comp & 512u) << 22));
          comp >>= 10;
          q[i] = v;
          sum = m::add(sum, m::mul(v, v));
        } else {
          q[i] = 0.0f;
        }
      }
      const float largest = m::sqrt_rn(m::sub(1.0f, sum));
#pragma unroll
      for (int i = 0; i < 4; i++)
        if ((uint32_t)i == big) q[i] = largest;
#pragma unroll
      for (int i = 0; i < 3; i++) q[i] = m::mul(q[i], a.flipQ[i]);
    } else {  // :333-345: the flip comes before the squared norm
      const float k = 1.0f / 127.5f;
#pragma unroll
      for (int i = 0; i < 3; i++) q[i] = m::mul(m::add(m::mul(m::i2f((int32_t)r[rRot + i]), k), -1.0f), a.flipQ[i]);
      const float rem = m::sub(1.0f, m::add(m::add(m::mul(q[0], q[0]), m::mul(q[1], q[1])), m::mul(q[2], q[2])));
      q[3] = m::sqrt_rn(0.0f < rem ? rem : 0.0f);  // std::max(0.0f, rem): NaN -> 0
    }
#pragma unroll
    for (int i = 0; i < 4; i++) u[uRot + i] = q[i];
    // spherical harmonics :425-429: c.flipSh[j] * ((b - 128) / 128)
#pragma unroll
    for (int j = 0; j < 15; j++) {
      u[uShR + j] = m::mul(a.flipSh[j], m::dequant_sh(r[rShR + j], 0.0078125f));
      u[uShG + j] = m::mul(a.flipSh[j], m::dequant_sh(r[rShG + j], 0.0078125f));
      u[uShB + j] = m::mul(a.flipSh[j], m::dequant_sh(r[rShB + j], 0.0078125f));
    }
  }
  __syncthreads();
  float *dst = a.out + g0 * kOut;
  const int floats = count * kOut;
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    const int quads = floats >> 2;
    for (int i = t; i < quads; i += kG) reinterpret_cast<float4 *>(dst)[i] = reinterpret_cast<const float4 *>(tile)[i];
    for (int i = (quads << 2) + t; i < floats; i += kG) dst[i] = tile[i];
  } else {
    for (int i = t; i < floats; i += kG) dst[i] = tile[i];
  }
}

}  // namespace

cudaError_t launchUnpackRecords(const GatherArgs &a, cudaStream_t stream, int *launches) {
  if (launches) *launches = 0;
  if (a.n <= 0) return cudaSuccess;
  const long long blocks = (a.n + kG - 1) / kG;
  if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
  if (a.records) unpackRecordsKernel<RecordSource><<<(unsigned)blocks, kG, 0, stream>>>(a);
  else unpackRecordsKernel<PlaneSource><<<(unsigned)blocks, kG, 0, stream>>>(a);
  if (launches) *launches = 1;
  return cudaGetLastError();
}

}  // namespace spzb200
