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
