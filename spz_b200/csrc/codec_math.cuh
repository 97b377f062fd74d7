// Per-value quantizers and dequantizers of the .spz codec, written for sm_100a.
//
// Every function states the reference expression it reproduces (file:line under
// /root/reference/src/cc) and why the cheaper device formulation is bit-identical.  Parity rules
// that shape the code:
//   * no FMA contraction anywhere a product feeds a sum (the reference is built without FMA);
//     all float arithmetic goes through the explicit *_rn intrinsics, which nvcc never fuses;
//   * IEEE division and square root (__fdiv_rn / __fsqrt_rn), denormals kept (no -ftz);
//   * std::round is round-half-AWAY-from-zero, not the hardware's round-half-even;
//   * the reference's float->int casts are undefined for NaN / out-of-range operands; the x86-64
//     build yields the "integer indefinite" value (0x80000000, or its low bits).  Where matching
//     that costs one select we match it, so parity holds for NaN/Inf/huge inputs as well.
//
// The same header compiles for the host (tests/host_emul) with the intrinsics emulated, so the
// integer reformulations can be checked exhaustively against the oracle without a GPU.  The host
// build is test-only: nothing in the product calls these functions on the CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SPZ_HD __host__ __device__ __forceinline__
// rarely-taken general forms.  Inlined by default (the tile encoder is 6 % faster at SH degree 0 with
// them inline: a call under its 48-register cap spills); a translation unit that defines
// SPZ_COLD_NOINLINE before including this header gets one called copy instead.
#if defined(SPZ_COLD_NOINLINE)
#define SPZ_HD_COLD static __host__ __device__ __noinline__
#else
#define SPZ_HD_COLD __host__ __device__ __forceinline__
#endif
#else
#define SPZ_HD inline
#define SPZ_HD_COLD inline
#endif

#if defined(__CUDA_ARCH__)
#define SPZ_DEVICE_PASS 1
#else
#define SPZ_DEVICE_PASS 0
#include <cfenv>
#include <cmath>
#include <cstring>
#endif

namespace spzb200 {
namespace m {

// ---- thin wrappers: device intrinsics, or their host emulation (tests only) -----------------
#if SPZ_DEVICE_PASS
SPZ_HD float mul(float a, float b) { return __fmul_rn(a, b); }
SPZ_HD float add(float a, float b) { return __fadd_rn(a, b); }
SPZ_HD float sub(float a, float b) { return __fsub_rn(a, b); }
SPZ_HD float div(float a, float b) { return __fdiv_rn(a, b); }
SPZ_HD float sqrt_rn(float a) { return __fsqrt_rn(a); }
SPZ_HD float rcp_rn(float a) { return __frcp_rn(a); }
SPZ_HD float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }  // one rounding: used only where the reference has none to match
SPZ_HD float add_rz(float a, float b) { return __fadd_rz(a, b); }
SPZ_HD int32_t f2i_rz(float a) { return __float2int_rz(a); }        // saturating, NaN -> 0
SPZ_HD uint32_t f2u_rz(float a) { return __float2uint_rz(a); }      // saturating, NaN -> 0
SPZ_HD long long f2ll_rz(float a) { return __float2ll_rz(a); }
SPZ_HD float i2f(int32_t a) { return __int2float_rn(a); }
SPZ_HD uint32_t fbits(float a) { return __float_as_uint(a); }
SPZ_HD float bitsf(uint32_t a) { return __uint_as_float(a); }
SPZ_HD float minnum(float a, float b) { return fminf(a, b); }       // IEEE minNum: NaN loses
#else
inline float mul(float a, float b) { volatile float r = a * b; return r; }
inline float add(float a, float b) { volatile float r = a + b; return r; }
inline float sub(float a, float b) { volatile float r = a - b; return r; }
inline float div(float a, float b) { volatile float r = a / b; return r; }
inline float sqrt_rn(float a) { return std::sqrt(a); }
inline float rcp_rn(float a) { volatile float r = 1.0f / a; return r; }
inline float fma_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float add_rz(float a, float b) {
  const int old = std::fegetround();
  std::fesetround(FE_TOWARDZERO);
  volatile float x = a, y = b;
  volatile float r = x + y;
  std::fesetround(old);
  return r;
}
inline int32_t f2i_rz(float a) {
  if (a != a) return 0;
  if (a >= 2147483648.0f) return INT32_MAX;
  if (a < -2147483648.0f) return INT32_MIN;
  return (int32_t)a;
}
inline uint32_t f2u_rz(float a) {
  if (a != a || a <= 0.0f) return 0u;
  if (a >= 4294967296.0f) return 0xffffffffu;
  return (uint32_t)a;
}
inline long long f2ll_rz(float a) {
  if (a != a) return 0;
  if (a >= 9223372036854775808.0f) return INT64_MAX;
  if (a < -9223372036854775808.0f) return INT64_MIN;
  return (long long)a;
}
inline float i2f(int32_t a) { return (float)a; }
inline uint32_t fbits(float a) { uint32_t u; std::memcpy(&u, &a, 4); return u; }
inline float bitsf(uint32_t a) { float f; std::memcpy(&f, &a, 4); return f; }
inline float minnum(float a, float b) { return std::fmin(a, b); }
#endif

// ---- rounding --------------------------------------------------------------------------------

// (int32)std::round(t) for |t| < 2^31, saturating outside, 0 for NaN.
//   round-half-away(t) = sgn(t) * floor(|t| + 1/2).  Adding copysign(1/2, t) with round-TOWARD-
//   ZERO can only lose magnitude, never reach the next integer, so truncating the sum gives exactly
//   floor(|t| + 1/2) with the sign of t -- including the case t = 0.49999997 where a
//   round-to-nearest add would produce 1.0.  Three instructions: LOP3, FADD.RZ, F2I.TRUNC.
SPZ_HD int32_t round_half_away_i32(float t) {
  const float half = bitsf((fbits(t) & 0x80000000u) | 0x3f000000u);
  return f2i_rz(add_rz(t, half));
}

SPZ_HD int32_t clamp_u8(int32_t v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// toUint8 (load-spz.cc:74): (uint8)clamp(round(x), 0, 255).  NaN -> clamp passes NaN -> the x86
// cast gives 0; the saturating device conversion gives 0 for NaN too, +-Inf clamp as on x86.
SPZ_HD uint32_t to_u8(float x) { return (uint32_t)clamp_u8(round_half_away_i32(x)); }

// ---- encode-side quantizers -----------------------------------------------------------------

// The *_raw forms return the rounded integer BEFORE the [0,255] clamp, for callers that saturate
// and pack four values at once (cvt.pack.sat.u8.s32); the plain forms clamp.

// scales, load-spz.cc:291: toUint8((s + 10.0f) * 16.0f)
SPZ_HD int32_t quant_scale_raw(float s) { return round_half_away_i32(mul(add(s, 10.0f), 16.0f)); }
SPZ_HD uint32_t quant_scale(float s) { return (uint32_t)clamp_u8(quant_scale_raw(s)); }

// colours, load-spz.cc:306: toUint8(c * (0.15f*255.0f) + (0.5f*255.0f)); the two constants fold
// to 38.25f and 127.5f in float arithmetic.  Product and sum are rounded separately.
SPZ_HD int32_t quant_color_raw(float c) {
  return round_half_away_i32(add(mul(c, 0.15f * 255.0f), 0.5f * 255.0f));
}
SPZ_HD uint32_t quant_color(float c) { return (uint32_t)clamp_u8(quant_color_raw(c)); }

// positions, load-spz.cc:281-288: (int32)round((flip * p) * 4096.0f), low 24 bits kept.
// signedScale = +-4096.0f: multiplying by a power of two commutes exactly with the sign flip.
// The x86 cast yields 0x80000000 (low 24 bits zero) for NaN and for |t| >= 2^31; the saturating
// device cast differs only for t >= 2^31 (0x7fffffff), fixed by one select.
SPZ_HD uint32_t quant_position24(float p, float signedScale) {
  const float t = mul(p, signedScale);
  int32_t n = round_half_away_i32(t);
  if (t >= 2147483648.0f) n = 0;
  return (uint32_t)n & 0x00ffffffu;
}

// SH coefficients, load-spz.cc:77-81 + :312-326:
//     q = (int)(round((flip * x) * 128.0f) + 128.0f);  q = (q + b/2) / b * b;  clamp(q, 0, 255)
// with b = 8 for the first 9 values of a point and 16 for the rest.
//   signed128 = +-128.0f (flip folded in), addk = 128 + b/2, mask = ~(b-1).
//   * round(t) + 128.0f is exact whenever the result is below 2^31, so q = n + 128 in integers.
//   * C's truncating division only differs from a floor (the AND) when q + b/2 < 0, and then
//     both give a value <= 0 that the clamp turns into 0.
//   * NaN, and anything whose q would be >= 2^31, make the x86 cast return INT_MIN -> byte 0.
//     minnum() turns NaN into a huge finite value; the saturated INT_MAX plus addk wraps
//     negative in unsigned arithmetic -> byte 0, the same answer (checked on the boundary floats
//     2147483392.0f -> 255 and 2147483520.0f -> 0).
SPZ_HD int32_t quant_sh_raw(float x, float signed128, uint32_t addk, uint32_t mask) {
  const float t = minnum(mul(x, signed128), 4.0e9f);
  return (int32_t)(((uint32_t)round_half_away_i32(t) + addk) & mask);
}
SPZ_HD uint32_t quant_sh(float x, float signed128, uint32_t addk, uint32_t mask) {
  return (uint32_t)clamp_u8(quant_sh_raw(x, signed128, addk, mask));
}

// alphas, load-spz.cc:85 + :301: toUint8(sigmoid(a) * 255.0f) with libm expf inside.  Over all
// floats this is a monotone step function with 255 steps, so the byte is the
// number of thresholds <= a.  thr[0..254] ascending is built once per context on the host from
// the reference expression and the host's own libm (spzb200 context init); thr[255] = +Inf pads
// the table to a power of two.  NaN compares false everywhere -> 0, as on x86.
SPZ_HD uint32_t quant_alpha(float a, const float *thr) {
  uint32_t lo = 0;  // invariant: thr[lo-1] <= a (or lo == 0)
#pragma unroll
  for (uint32_t step = 128; step >= 1; step >>= 1) {
    if (a >= thr[lo + step - 1]) lo += step;
  }
  return lo;
}

// the cast (uint32_t)x as gcc emits it on x86-64: cvttss2si to 64 bits, keep the low half.
SPZ_HD uint32_t cast_u32_like_x86(float v) {
  if (v < 4294967296.0f && v > -1.0f) return f2u_rz(v);  // the only range valid input reaches
  if (!(v >= -9223372036854775808.0f && v < 9223372036854775808.0f)) return 0u;
  return (uint32_t)(unsigned long long)f2ll_rz(v);
}

// rotations, load-spz.cc:216-255 (packQuaternionSmallestThree) + splat-types.cc:71-74
// (normalized).  rot = x,y,z,w; flipBits bit i set = negate component i (i < 3) after the
// normalisation.  Same operations in the same order as the reference; valid for every input
// (NaN, Inf, zero and denormal norms included).  The kernels reach it only through the guard of
// quant_rotation_smallest3 below.
SPZ_HD_COLD uint32_t quant_rotation_smallest3_general(float x, float y, float z, float w, uint32_t flipBits) {
  const float kInvSqrt2 = 0.70710678118654752440f;
  const float n2 = add(add(add(mul(x, x), mul(y, y)), mul(z, z)), mul(w, w));
  const float nrm = sqrt_rn(n2);
  float q[4];
  q[0] = bitsf(fbits(div(x, nrm)) ^ ((flipBits & 1u) << 31));
  q[1] = bitsf(fbits(div(y, nrm)) ^ ((flipBits & 2u) << 30));
  q[2] = bitsf(fbits(div(z, nrm)) ^ ((flipBits & 4u) << 29));
  q[3] = div(w, nrm);
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = bitsf(fbits(q[i]) & 0x7fffffffu);
  uint32_t big = 0;
  float best = a[0];
#pragma unroll
  for (uint32_t i = 1; i < 4; i++) {
    if (a[i] > best) { best = a[i]; big = i; }  // strict >, NaN never wins (ref :227)
  }
  const float qbig = big == 0 ? q[0] : big == 1 ? q[1] : big == 2 ? q[2] : q[3];
  const uint32_t negate = qbig < 0.0f ? 1u : 0u;
  uint32_t comp = big;
#pragma unroll
  for (uint32_t i = 0; i < 4; i++) {
    if (i != big) {
      const uint32_t negbit = (q[i] < 0.0f ? 1u : 0u) ^ negate;
      const uint32_t mag = cast_u32_like_x86(add(mul(511.0f, div(a[i], kInvSqrt2)), 0.5f));
      comp = (comp << 10) | (negbit << 9) | mag;
    }
  }
  return comp;
}

// The IEEE quotient x / b from rb = RN(1 / b), the correctly rounded reciprocal (Markstein 1990; Muller
// et al., Handbook of Floating-Point Arithmetic, "division with an FMA"): q0 = RN(x * rb) is within 2 ulp
// of x / b; one residual correction leaves an error of 2^-46 relative before rounding, i.e. a
// faithful q1; for a faithful q1 the residual x - b * q1 is exact and RN(q1 + residual * rb) is the
// correctly rounded quotient.  Preconditions, established by the caller: x, b, the quotient and the
// residuals are in the normal range (or x == 0), and b's significand is not all ones.  Five
// branch-free instructions instead of the ~14 of a guarded division, and the reciprocal is shared by
// the four components.  tests/test_device_math_host.py checks it against the plain division on random
// operand pairs from that domain (2^33 pairs once, profiles/r1_tuning_notes.txt; 2^27 per test run).
SPZ_HD float div_by_rcp(float x, float b, float rb) {
  float q = mul(x, rb);
  q = fma_rn(fma_rn(-b, q, x), rb, q);
  return fma_rn(fma_rn(-b, q, x), rb, q);
}

// a / sqrt1_2 for a in {0} U [2^-81, 1.01]: one correction is enough for this divisor -- checked
// against the plain division for every float of that range (tests/test_device_math_host.py).
SPZ_HD float div_by_sqrt1_2(float a) {
  const float c = 0.70710678118654752440f;            // 0x3f3504f3, the reference's sqrt1_2 (load-spz.cc:218)
  const float rc = 1.41421353816986083984375f;        // RN(1 / c) = 0x3fb504f3
  const float q = mul(a, rc);
  return fma_rn(fma_rn(-c, q, a), rc, q);
}

// packQuaternionSmallestThree for the inputs real clouds hold -- squared norm in [2^-40, 2^40], every
// component zero or at least 2^-60 in magnitude -- with the seven divisions done by div_by_rcp;
// anything else (NaN, Inf, vanishing or huge norms, denormal components) takes the general form
// above.  Inside the guard every normalised component is within [2^-81, 1 + 2^-21] or zero, so the
// magnitude 511 * (a / sqrt1_2) + 0.5 stays below 724 and the plain truncating conversion is what the
// x86 cast computes; no NaN can occur, so the strict-> argmax and the < 0 tests are plain compares.
SPZ_HD uint32_t quant_rotation_smallest3(float x, float y, float z, float w, uint32_t flipBits) {
  const float n2 = add(add(add(mul(x, x), mul(y, y)), mul(z, z)), mul(w, w));
  const float nrm = sqrt_rn(n2);
  // (u - 1) >= 2^-60 - 1 as unsigned: true for u == 0 (wraps) and for u >= bits(2^-60)
  const uint32_t lowest = 0x21800000u - 1u;
  const bool componentsOk = ((fbits(x) & 0x7fffffffu) - 1u) >= lowest && ((fbits(y) & 0x7fffffffu) - 1u) >= lowest &&
                            ((fbits(z) & 0x7fffffffu) - 1u) >= lowest && ((fbits(w) & 0x7fffffffu) - 1u) >= lowest;
  const bool normOk = n2 >= 9.094947017729282e-13f && n2 <= 1099511627776.0f && (fbits(nrm) & 0x007fffffu) != 0x007fffffu;
  if (!(componentsOk && normOk)) return quant_rotation_smallest3_general(x, y, z, w, flipBits);

  const float rn = rcp_rn(nrm);
  float q[4];
  q[0] = bitsf(fbits(div_by_rcp(x, nrm, rn)) ^ ((flipBits & 1u) << 31));
  q[1] = bitsf(fbits(div_by_rcp(y, nrm, rn)) ^ ((flipBits & 2u) << 30));
  q[2] = bitsf(fbits(div_by_rcp(z, nrm, rn)) ^ ((flipBits & 4u) << 29));
  q[3] = div_by_rcp(w, nrm, rn);
  float a[4];
  uint32_t sgn[4], field[4];  // q < 0 as bit 9, and that bit over the 9-bit magnitude
#pragma unroll
  for (int i = 0; i < 4; i++) {
    a[i] = bitsf(fbits(q[i]) & 0x7fffffffu);
    sgn[i] = q[i] < 0.0f ? 512u : 0u;
    // <= 511 for every component but the largest (a <= sqrt1_2 * (1 + 2^-21)), whose field is not stored
    field[i] = f2u_rz(add(mul(511.0f, div_by_sqrt1_2(a[i])), 0.5f)) | sgn[i];
  }
  uint32_t big = 0;
  float best = a[0];
#pragma unroll
  for (uint32_t i = 1; i < 4; i++) {
    if (a[i] > best) { best = a[i]; big = i; }  // strict >: first index of the maximum (ref :227)
  }
  const uint32_t negate = big == 0 ? sgn[0] : big == 1 ? sgn[1] : big == 2 ? sgn[2] : sgn[3];  // q[big] < 0: kept signs are relative to it
  const uint32_t first = big == 0 ? field[1] : field[0];
  const uint32_t second = big <= 1 ? field[2] : field[1];
  const uint32_t third = big == 3 ? field[2] : field[3];
  return (big << 30) | ((first ^ negate) << 20) | ((second ^ negate) << 10) | (third ^ negate);
}

// rotations, version-2 streams (first three components as bytes).  PARITY UNPINNED: the reference tree
// has only the decoder for this form (load-spz.cc:333-345); this is the encoder upstream nianticlabs/spz
// shipped before the smallest-three change -- normalise, flip, make w non-negative, store
// toUint8(q.xyz * 127.5 + 127.5) -- written with the reference's own primitives (normalized,
// splat-types.cc:71-74; toUint8, load-spz.cc:74).  Returns the three bytes in bits 0..23.
SPZ_HD_COLD uint32_t quant_rotation_first3(float x, float y, float z, float w, uint32_t flipBits) {
  const float n2 = add(add(add(mul(x, x), mul(y, y)), mul(z, z)), mul(w, w));
  const float nrm = sqrt_rn(n2);
  const float qx = bitsf(fbits(div(x, nrm)) ^ ((flipBits & 1u) << 31));
  const float qy = bitsf(fbits(div(y, nrm)) ^ ((flipBits & 2u) << 30));
  const float qz = bitsf(fbits(div(z, nrm)) ^ ((flipBits & 4u) << 29));
  const float s = div(w, nrm) < 0.0f ? -127.5f : 127.5f;
  return to_u8(add(mul(qx, s), 127.5f)) | (to_u8(add(mul(qy, s), 127.5f)) << 8) | (to_u8(add(mul(qz, s), 127.5f)) << 16);
}

// ---- decode-side dequantizers ---------------------------------------------------------------

// scales, load-spz.cc:506: s / 16.0f - 10.0f.  s/16 is exact and the difference is a multiple of
// 1/16 below 16 in magnitude, so nothing rounds.
SPZ_HD float dequant_scale(uint32_t s) { return sub(mul(i2f((int32_t)s), 0.0625f), 10.0f); }

// colours, load-spz.cc:522: ((c / 255.0f) - 0.5f) / 0.15f  (used to fill the 256-entry table).
SPZ_HD float dequant_color(uint32_t c) { return div(sub(div(i2f((int32_t)c), 255.0f), 0.5f), 0.15f); }

// SH, load-spz.cc:83 then the flip of splat-types.h:158-161: ((x - 128) / 128) * (+-1).
// signedInv128 = +-0.0078125f; x == 128 gives +0 * -1/128 = -0, as the reference's 0 * -1.
SPZ_HD float dequant_sh(uint32_t x, float signedInv128) {
  return mul(i2f((int32_t)x - 128), signedInv128);
}

// positions, load-spz.cc:495-502 then splat-types.h:140-144: sign-extended 24-bit integer times
// (float)(1.0 / (1 << fractionalBits)), times +-1.  signedScale carries both factors.
SPZ_HD float dequant_position24(uint32_t lo24, float signedScale) {
  const int32_t fixed = (int32_t)(lo24 << 8) >> 8;
  return mul(i2f(fixed), signedScale);
}

// one magnitude of a smallest-three component, load-spz.cc:367: sqrt1_2 * (float)mag / 511.0f
SPZ_HD float dequant_s3_magnitude(uint32_t mag) {
  return div(mul(0.70710678118654752440f, i2f((int32_t)mag)), 511.0f);
}

// rotations v3, load-spz.cc:347-381.  magLut[m] = dequant_s3_magnitude(m), m < 512.  Fields are
// popped from the low end into components 3,2,1,0 skipping the largest; the sum of squares is
// accumulated in that order with separate multiply and add; the largest is sqrt(1 - sum) with no
// clamp (the reference's double-overload sqrt then narrowing equals the float sqrt bit for bit,
// SURVEY.md section 7).  flipBits as in the encoder; applied last (splat-types.h:145-150).
SPZ_HD void dequant_rotation_smallest3(uint32_t comp, const float *magLut, uint32_t flipBits,
                                       float out[4]) {
  const uint32_t big = comp >> 30;
  float sum = 0.0f;
  float r[4];
#pragma unroll
  for (int i = 3; i >= 0; --i) {
    if ((uint32_t)i != big) {
      const float v = bitsf(fbits(magLut[comp & 511u]) | ((comp & 512u) << 22));
      comp >>= 10;
      r[i] = v;
      sum = add(sum, mul(v, v));
    } else {
      r[i] = 0.0f;
    }
  }
  const float largest = sqrt_rn(sub(1.0f, sum));
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if ((uint32_t)i == big) r[i] = largest;
  }
  out[0] = bitsf(fbits(r[0]) ^ ((flipBits & 1u) << 31));
  out[1] = bitsf(fbits(r[1]) ^ ((flipBits & 2u) << 30));
  out[2] = bitsf(fbits(r[2]) ^ ((flipBits & 4u) << 29));
  out[3] = r[3];
}

// rotations v1/v2, load-spz.cc:333-345: xyz = byte * (1/127.5f) + (-1); w = sqrt(max(0, 1 - |xyz|^2)).
SPZ_HD void dequant_rotation_first3(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t flipBits,
                                    float out[4]) {
  const float k = 1.0f / 127.5f;
  const float x = add(mul(i2f((int32_t)b0), k), -1.0f);
  const float y = add(mul(i2f((int32_t)b1), k), -1.0f);
  const float z = add(mul(i2f((int32_t)b2), k), -1.0f);
  const float rem = sub(1.0f, add(add(mul(x, x), mul(y, y)), mul(z, z)));
  out[3] = sqrt_rn(0.0f < rem ? rem : 0.0f);
  out[0] = bitsf(fbits(x) ^ ((flipBits & 1u) << 31));
  out[1] = bitsf(fbits(y) ^ ((flipBits & 2u) << 30));
  out[2] = bitsf(fbits(z) ^ ((flipBits & 4u) << 29));
}

// v1 positions, splat-types.cc:8-28 halfToFloat: every branch is an exact product of powers of
// two and a 11-bit integer, i.e. the IEEE binary16 value; NaNs become the canonical quiet NaN.
SPZ_HD float half_bits_to_float(uint32_t h) {
  const uint32_t sgn = (h >> 15) & 1u, e = (h >> 10) & 0x1fu, mant = h & 0x3ffu;
  uint32_t bits;
  if (e == 0) {
    // subnormal half: mant * 2^-24, exact in float
    const float v = mul(i2f((int32_t)mant), 5.9604644775390625e-8f);
    bits = fbits(v) | (sgn << 31);
  } else if (e == 31) {
    bits = mant ? 0x7fc00000u : ((sgn << 31) | 0x7f800000u);
  } else {
    bits = (sgn << 31) | ((e + 112u) << 23) | (mant << 13);
  }
  return bitsf(bits);
}

// ---- coordinate flips (splat-types.h:43-81) as sign bit sets ---------------------------------
// bit i of .p = negate position axis i; .q = negate quaternion component i (x,y,z); .sh = negate
// SH coefficient j (all three colour channels), j < 15.
struct FlipBits {
  uint32_t p, q, sh;
};

SPZ_HD FlipBits make_flip_bits(int32_t from, int32_t to) {
  FlipBits f = {0u, 0u, 0u};
  const int32_t a = from - 1, b = to - 1;
  if (a < 0 || b < 0) return f;  // UNSPECIFIED on either side: no flips (splat-types.h:46-48)
  const uint32_t d = (uint32_t)(a ^ b) & 7u;
  const uint32_t x = d & 1u, y = (d >> 1) & 1u, z = (d >> 2) & 1u;
  f.p = x | (y << 1) | (z << 2);
  f.q = (y ^ z) | ((x ^ z) << 1) | ((x ^ y) << 2);
  // flipSh = {y, z, x, xy, yz, 1, xz, 1, y, xyz, y, z, x, z, x}
  const uint32_t s[15] = {y, z, x, x ^ y, y ^ z, 0u, x ^ z, 0u, y, x ^ y ^ z, y, z, x, z, x};
  for (int j = 0; j < 15; j++) f.sh |= s[j] << j;
  return f;
}

}  // namespace m
}  // namespace spzb200
