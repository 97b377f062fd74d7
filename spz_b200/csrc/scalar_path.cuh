// The scalar one-thread-per-gaussian form of the codec: whole clouds whose pointers are under-aligned
// for the vector kernels, clouds smaller than a tile, and the sub-tile remainder of every other call.
// The remainder does not get a launch of its own: the vector kernels reserve their first CTA(s) for it
// (they start first, so the slow byte-wise path overlaps the tiles instead of trailing them), which
// makes every aligned encode or decode ONE kernel launch.  Called, not inlined: the branch is cold
// and must not cost the tile path registers under its 48-register bound.  The kernels declare their
// argument struct __grid_constant__ so that its address can be handed to these functions as it is;
// without that the compiler copies the whole struct into every thread's local memory at kernel entry
// (ten STL.128 per thread -- measured: the per-gaussian decoder fell from 6.5 to 4.7 TB/s).
#pragma once
#include "codec_kernels.cuh"
#include "codec_math.cuh"
#include "kernel_utils.cuh"

#define SPZ_SCALAR_FN static __device__ __noinline__

namespace spzb200 {

// packGaussians for gaussian g (load-spz.cc:281-328)
SPZ_SCALAR_FN void encodeOneGaussian(const EncodeArgs &a, const long long g) {
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    const uint32_t n = m::quant_position24(a.positions[g * 3 + ax],
                                           signedConst(4096.0f, (a.flipP >> ax) & 1u));
    uint8_t *o = a.oPositions + (g * 3 + ax) * 3;
    o[0] = (uint8_t)n; o[1] = (uint8_t)(n >> 8); o[2] = (uint8_t)(n >> 16);
    a.oScales[g * 3 + ax] = (uint8_t)m::quant_scale(a.scales[g * 3 + ax]);
    a.oColors[g * 3 + ax] = (uint8_t)m::quant_color(a.colors[g * 3 + ax]);
  }
  a.oAlphas[g] = (uint8_t)m::quant_alpha(a.alphas[g], a.alphaThresholds);
  const float *r = a.rotations + g * 4;
  if (a.version == 2) {
    const uint32_t b = m::quant_rotation_first3(r[0], r[1], r[2], r[3], a.flipQ);
    uint8_t *ro = a.oRotations + g * 3;
    ro[0] = (uint8_t)b; ro[1] = (uint8_t)(b >> 8); ro[2] = (uint8_t)(b >> 16);
  } else {
    const uint32_t comp = m::quant_rotation_smallest3(r[0], r[1], r[2], r[3], a.flipQ);
    uint8_t *ro = a.oRotations + g * 4;
    ro[0] = (uint8_t)comp; ro[1] = (uint8_t)(comp >> 8); ro[2] = (uint8_t)(comp >> 16); ro[3] = (uint8_t)(comp >> 24);
  }
  const int per = a.shDim * 3;
  const float *s = a.sh + g * per;
  uint8_t *so = a.oSh + g * per;
  for (int j = 0; j < per; j++) {
    const uint32_t bucket = j < 9 ? 8u : 16u;
    so[j] = (uint8_t)m::quant_sh(s[j], signedConst(128.0f, (a.flipSh >> (j / 3)) & 1u),
                                 128u + bucket / 2u, ~(bucket - 1u));
  }
}

// unpackGaussians + convertCoordinates for gaussian g (load-spz.cc:493-529)
SPZ_SCALAR_FN void decodeOneGaussian(const DecodeArgs &a, const long long g) {
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    const uint32_t flip = ((a.flipP >> ax) & 1u) << 31;
    float p;
    if (a.version == 1 || a.version == 4) {
      const uint8_t *h = a.positions + (g * 3 + ax) * 2;
      p = __uint_as_float(__float_as_uint(m::half_bits_to_float((uint32_t)h[0] | ((uint32_t)h[1] << 8))) ^ flip);
    } else {
      const uint8_t *b = a.positions + (g * 3 + ax) * 3;
      const uint32_t lo24 = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16);
      p = m::dequant_position24(lo24, __uint_as_float(__float_as_uint(a.positionScale) ^ flip));
    }
    a.oPositions[g * 3 + ax] = p;
    a.oScales[g * 3 + ax] = m::dequant_scale(a.scales[g * 3 + ax]);
    a.oColors[g * 3 + ax] = m::dequant_color(a.colors[g * 3 + ax]);
  }
  a.oAlphas[g] = a.tables[a.alphas[g]];
  float r[4];
  if (a.version >= 3) {
    const uint8_t *b = a.rotations + g * 4;
    uint32_t comp = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
    // no table here: compute the three magnitudes directly (same expression as the table fill)
    const uint32_t big = comp >> 30;
    float sum = 0.0f;
    for (int i = 3; i >= 0; --i) {
      if ((uint32_t)i == big) continue;
      const float v = __uint_as_float(__float_as_uint(m::dequant_s3_magnitude(comp & 511u)) | ((comp & 512u) << 22));
      comp >>= 10;
      r[i] = v;
      sum = m::add(sum, m::mul(v, v));
    }
    r[big] = m::sqrt_rn(m::sub(1.0f, sum));
    for (int i = 0; i < 3; i++) r[i] = __uint_as_float(__float_as_uint(r[i]) ^ (((a.flipQ >> i) & 1u) << 31));
  } else {
    const uint8_t *b = a.rotations + g * 3;
    m::dequant_rotation_first3(b[0], b[1], b[2], a.flipQ, r);
  }
  for (int i = 0; i < 4; i++) a.oRotations[g * 4 + i] = r[i];
  const int per = a.shDim * 3;
  const uint8_t *s = a.sh + g * per;
  float *so = a.oSh + g * per;
  for (int j = 0; j < per; j++)
    so[j] = m::dequant_sh(s[j], signedConst(0.0078125f, (a.flipSh >> (j / 3)) & 1u));
}

}  // namespace spzb200
