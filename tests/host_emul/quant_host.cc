// TEST-ONLY host build of spz_b200/csrc/codec_math.cuh: the device intrinsics are emulated
// (see the #else branch at the top of that header) so the integer reformulations the kernels
// rely on can be swept against the oracle without a GPU.  Never part of the product library.
//   g++ -O2 -std=c++17 -ffp-contract=off -frounding-math -shared -fPIC quant_host.cc
#include <cstdint>
#include <cstring>

#include "../../spz_b200/csrc/codec_math.cuh"

using namespace spzb200::m;

static float fromBits(uint32_t b) { float f; std::memcpy(&f, &b, 4); return f; }

extern "C" {

// which: 0 alpha (needs thr[256]), 1 scale, 2 colour, 3 sh bucket 8, 4 sh bucket 16,
//        5 sh bucket 8 flipped, 6 sh bucket 16 flipped
void emul_sweep_u8(int which, uint32_t first, uint32_t stride, int64_t count, const float *thr, uint8_t *out) {
  for (int64_t i = 0; i < count; i++) {
    const float x = fromBits(first + (uint32_t)i * stride);
    uint32_t r;
    switch (which) {
      case 0: r = quant_alpha(x, thr); break;
      case 1: r = quant_scale(x); break;
      case 2: r = quant_color(x); break;
      case 3: r = quant_sh(x, 128.0f, 132u, ~7u); break;
      case 4: r = quant_sh(x, 128.0f, 136u, ~15u); break;
      case 5: r = quant_sh(x, -128.0f, 132u, ~7u); break;
      default: r = quant_sh(x, -128.0f, 136u, ~15u); break;
    }
    out[i] = (uint8_t)r;
  }
}

// 24-bit position codes for a bit-pattern sweep; flip selects -4096
void emul_sweep_position(uint32_t first, uint32_t stride, int64_t count, int flip, uint32_t *out) {
  for (int64_t i = 0; i < count; i++)
    out[i] = quant_position24(fromBits(first + (uint32_t)i * stride), flip ? -4096.0f : 4096.0f);
}

void emul_rotations(int64_t n, const float *rot, uint32_t flipBits, uint32_t *out) {
  for (int64_t i = 0; i < n; i++)
    out[i] = quant_rotation_smallest3(rot[4 * i], rot[4 * i + 1], rot[4 * i + 2], rot[4 * i + 3], flipBits);
}

void emul_unrotations_s3(int64_t n, const uint32_t *comp, uint32_t flipBits, float *out) {
  float lut[512];
  for (uint32_t m = 0; m < 512; m++) lut[m] = dequant_s3_magnitude(m);
  for (int64_t i = 0; i < n; i++) dequant_rotation_smallest3(comp[i], lut, flipBits, out + 4 * i);
}

void emul_unrotations_f3(int64_t n, const uint8_t *b, uint32_t flipBits, float *out) {
  for (int64_t i = 0; i < n; i++) dequant_rotation_first3(b[3 * i], b[3 * i + 1], b[3 * i + 2], flipBits, out + 4 * i);
}

void emul_dequant_tables(float *scale256, float *color256, float *sh256, float *shFlipped256, float *half65536) {
  for (uint32_t i = 0; i < 256; i++) {
    scale256[i] = dequant_scale(i);
    color256[i] = dequant_color(i);
    sh256[i] = dequant_sh(i, 0.0078125f);
    shFlipped256[i] = dequant_sh(i, -0.0078125f);
  }
  for (uint32_t h = 0; h < 65536; h++) half65536[h] = half_bits_to_float(h);
}

void emul_positions_decode(int64_t count, const uint32_t *lo24, float signedScale, float *out) {
  for (int64_t i = 0; i < count; i++) out[i] = dequant_position24(lo24[i], signedScale);
}

void emul_flip_bits(int32_t from, int32_t to, uint32_t *p, uint32_t *q, uint32_t *sh) {
  const FlipBits f = make_flip_bits(from, to);
  *p = f.p; *q = f.q; *sh = f.sh;
}

}  // extern "C"

// ---- record realignment (spz_b200/csrc/record_align.cuh), one "lane" after the other --------------
#include "../../spz_b200/csrc/record_align.cuh"

namespace {
template <int B>
void emitAll(int64_t n, const uint8_t *records, uint32_t *plane, uint8_t garbage) {
  using R = spzb200::Rec<B>;
  uint32_t prevTail = 0x5a5a5a5au;  // what lane 0's shuffle hands back is its own tail: must not matter
  for (int64_t g = 0; g < n; g++) {
    uint8_t bytes[4 * R::NV];
    for (int i = 0; i < 4 * R::NV; i++) bytes[i] = i < B ? records[g * B + i] : garbage;
    uint32_t v[R::NV];
    std::memcpy(v, bytes, sizeof v);
    const uint32_t mine = spzb200::recordTail<B>(v);
    spzb200::emitRecordWords<B>(plane, (int)g, v, (g % 32) ? prevTail : mine);
    prevTail = mine;
  }
}
template <int B>
void loadAll(int64_t n, const uint32_t *plane, uint8_t *records) {
  using R = spzb200::Rec<B>;
  for (int64_t g = 0; g < n; g++) {
    uint32_t v[R::NL];
    spzb200::loadRecord<B>(plane, (int)g, v);
    uint8_t bytes[4 * R::NL];
    std::memcpy(bytes, v, sizeof v);
    std::memcpy(records + g * B, bytes, B);
  }
}
}  // namespace

extern "C" {
// records: n x B bytes -> plane words (caller sizes plane to n * B / 4 + slack and pre-fills it)
int emul_emit_records(int B, int64_t n, const uint8_t *records, uint32_t *plane, uint8_t garbage) {
  switch (B) {
    case 3: emitAll<3>(n, records, plane, garbage); return 0;
    case 9: emitAll<9>(n, records, plane, garbage); return 0;
    case 24: emitAll<24>(n, records, plane, garbage); return 0;
    case 45: emitAll<45>(n, records, plane, garbage); return 0;
    default: return -1;
  }
}
int emul_load_records(int B, int64_t n, const uint32_t *plane, uint8_t *records) {
  switch (B) {
    case 3: loadAll<3>(n, plane, records); return 0;
    case 6: loadAll<6>(n, plane, records); return 0;
    case 9: loadAll<9>(n, plane, records); return 0;
    case 24: loadAll<24>(n, plane, records); return 0;
    case 45: loadAll<45>(n, plane, records); return 0;
    default: return -1;
  }
}
}

// ---- div_by_rcp (codec_math.cuh) against the plain IEEE division ---------------------------------------
namespace {
inline uint64_t splitmix(uint64_t &s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
}  // namespace

extern "C" {
// Pseudo-random operand pairs from the domain quant_rotation_smallest3's guard admits: divisor b (the
// norm) in [2^-20, 2^20] with a significand that is not all ones, dividend x = 0 or 2^-60 <= |x| <=
// b * (1 + 2^-21).  Returns the number of quotients that differ from x / b; the first offender's bits
// go to bad[0..1].
int64_t emul_check_div_by_rcp_random(uint64_t seed, int64_t count, uint32_t *bad) {
  int64_t wrong = 0;
  uint64_t s = seed;
  for (int64_t i = 0; i < count; i++) {
    const uint64_t r = splitmix(s), r2 = splitmix(s);
    uint32_t mant = (uint32_t)r & 0x7fffffu;
    if (mant == 0x7fffffu) mant = 0x7ffffeu;
    const uint32_t be = 107u + (uint32_t)((r >> 23) % 41u);  // exponent of 2^-20 .. 2^20
    const float b = bitsf((be << 23) | mant);
    // dividend: random significand, exponent anywhere from 2^-60 up to b's; some exactly at the top
    uint32_t xe = 67u + (uint32_t)((r2 >> 23) % (be - 67u + 1u));
    if (((r2 >> 40) & 3u) == 0) xe = be;  // a quarter of the pairs near |x| ~ b, where the quotient is near 1
    float x = bitsf((xe << 23) | ((uint32_t)r2 & 0x7fffffu) | ((uint32_t)(r2 >> 63) << 31));
    if (std::fabs(x) > b * 1.0000005f) x = bitsf((fbits(x) & 0x80000000u) | fbits(b));
    if (((r2 >> 44) & 1023u) == 0) x = bitsf((uint32_t)(r2 >> 63) << 31);  // +-0
    volatile float want = x / b;
    const float got = div_by_rcp(x, b, rcp_rn(b));
    const float w = want;
    const bool same = fbits(got) == fbits(w) || (got == 0.0f && w == 0.0f);  // the sign of a zero quotient is not used
    if (!same) {
      if (wrong == 0 && bad) { bad[0] = fbits(x); bad[1] = fbits(b); }
      wrong++;
    }
  }
  return wrong;
}

// Every float a with bits in [lo, hi] divided by the constant sqrt1_2 through its precomputed
// reciprocal, against a / sqrt1_2.
int64_t emul_check_div_by_sqrt1_2(uint32_t lo, uint32_t hi, uint32_t *bad) {
  const float c = 0.70710678118654752440f;
  int64_t wrong = 0;
  for (uint64_t u = lo; u <= hi; u++) {
    const float a = bitsf((uint32_t)u);
    volatile float want = a / c;
    const float got = div_by_sqrt1_2(a);
    const float w = want;
    if (fbits(got) != fbits(w)) {
      if (wrong == 0 && bad) bad[0] = (uint32_t)u;
      wrong++;
    }
  }
  return wrong;
}
}
