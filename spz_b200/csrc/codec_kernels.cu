// sm_100a kernels of the .spz per-gaussian codec: packGaussians (load-spz.cc:257-331) and
// unpackGaussians (load-spz.cc:467-531, with convertCoordinates splat-types.h:134-164 fused in).
//
// Layout insight that shapes the kernels: every plane of GaussianCloud / PackedGaussians is a
// flat array in which four consecutive floats correspond to a whole number of packed 32-bit
// words (4 floats <-> 4 bytes for scales / colours / alphas / SH, 4 floats <-> 12 bytes for
// 24-bit positions, 1 quaternion <-> 4 bytes).  So the codec is an ELEMENTWISE stream transform
// on 16-byte granules: consecutive lanes touch consecutive float4s (512 B per warp instruction)
// and consecutive packed words (128 B per warp instruction).  No transposition of planes is
// needed; the only position-dependent parameters are the coordinate-flip sign and the SH bucket
// size, which depend on (element index mod 3*shDim) resp. (element index mod 3).
//
// Tiling: a CTA has 320 threads (10 full warps); a "row" is 320 consecutive float4s (5120 B) of
// a float plane and the 320 (or 960) packed words that belong to them, so every warp-level access
// is 128-byte aligned and covers whole 32-byte sectors -- the first version of this kernel used
// rows of 315 and paid 11% (encode) / 24% (decode) extra DRAM reads for the misalignment (L2
// fills of partially written sectors, re-fetched straddled sectors; profiles/r1_*).  A tile is
// 4*320*M gaussians (M = 5 for SH degree 1, else 1) = 3*shDim*M rows of the SH plane.  The phase
// of a thread's four elements inside the 3*shDim-float SH record advances by a constant
// (4*320 mod 3*shDim) per row and therefore repeats with a short cycle (9 rows for degree 3 and
// 1, 3 rows for degree 2): rows c, c+CYC, c+2*CYC, ... share their flip signs and bucket sizes, so
// they are loaded together (5 or 8 float4 in flight per thread) under one set of constants.
// The 24-bit position words (3 per float4) are exchanged through a per-warp shared-memory
// stage so that their global loads and stores are contiguous 128-byte lines as well.
// One CTA per tile (gridDim = tiles): the hardware block scheduler hands tiles out in order, which
// keeps the set of pages being touched compact; persistent grid-stride CTAs drift apart over a
// 100M-point cloud and lost 4-12 % (kept as the SPZB200_GRID=persistent knob).  The encoder works
// from registers.  The tile decoder stages the SH plane -- three quarters of its bytes -- through
// shared memory with bulk async copies (decodeTilesBulkKernel below).  At SH degree 1 - 3 (and
// 16-byte aligned planes) launchDecode hands the cloud to the one-thread-per-gaussian bulk-copy
// decoder of pergaussian_kernels.cu instead: as fast at 100M points, faster below (64 / 128
// gaussians per CTA shorten a launch's ramp-up and tail); the tile decoders keep SH-less clouds
// and planes that are only 4-byte aligned (SPZB200_DECODE=bulk / direct force them).  The remainder (< one tile) and any call with
// under-aligned pointers goes to a scalar one-thread-per-gaussian kernel.
//
// HBM traffic is exactly the algorithmic 301 B per gaussian at SH degree 3 (236 B floats + 65 B
// packed); nothing is read twice.
#include "codec_kernels.cuh"

#include "codec_math.cuh"
#include "kernel_utils.cuh"
#include "scalar_path.cuh"

namespace spzb200 {
namespace {

constexpr int kThreads = 320;
constexpr int kWarps = kThreads / 32;
#ifndef SPZ_CTAS_PER_SM
#define SPZ_CTAS_PER_SM 4
#endif
constexpr int kCtasPerSm = SPZ_CTAS_PER_SM;  // resident CTAs per SM the register budget is set for
#ifndef SPZ_SMALL_CTAS_PER_SM
#define SPZ_SMALL_CTAS_PER_SM 10  // 128-thread CTAs: the same 48 registers per thread
#endif

constexpr int gcdc(int a, int b) { return b == 0 ? a : gcdc(b, a % b); }

// S = threads per CTA.  320 is the geometry everything was tuned for at 100M points; 128 (round 2) gives tiles 2.5x
// smaller for launches whose tail is a visible part of their duration (SH degree 0, 1, 2; at degree 3 the phase
// cycle would be 45 rows long with S = 128, and small degree-3 clouds use the per-gaussian kernels anyway).
#ifndef SPZ_SMALL_THREADS
#define SPZ_SMALL_THREADS 128  // must be = 2 mod 3 (Geo): 128, 224, 320, 416 ...; 224 x 6 CTAs/SM and 416 x 3 were tried for the SH-less decoder (notes section 25)
#endif
constexpr int kSmallThreads = SPZ_SMALL_THREADS;
#ifndef SPZ_SH1_SMALL_M
#define SPZ_SH1_SMALL_M 3  // SH degree 1, 128-thread geometry: sub-tiles per tile = SH rows in flight per phase class.  Measured (encode GB/s at
                           // 10M / 20M / 100M gaussians, profiles/r2_tuning_notes.txt section 16): M = 5 6290 / 6490 / 6750, M = 2 6375 / 6585 / 6636, M = 3 6478 / 6760 / 6866
#endif
template <int D, int S = kThreads>
struct Geo {
  static_assert((4 * S) % 3 == 2, "the xyz phase of element e in row i is taken as (t + e + 2 i) mod 3");
  static constexpr int M = (D == 3) ? (S == kThreads ? 5 : SPZ_SH1_SMALL_M) : 1;  // sub-tiles (4*S gaussians each) per tile
  static constexpr int TG = 4 * S * M;                // gaussians per tile
  static constexpr int MOD = 3 * D;                   // floats per SH record
  static constexpr int STEP = D ? (4 * S) % (D ? MOD : 1) : 0;  // phase advance per row
  static constexpr int CYC = D ? MOD / gcdc(STEP, MOD) : 1;            // rows until the phase repeats
  static constexpr int ROWS = 3 * D * M;              // SH rows per tile
  static constexpr int U = D ? ROWS / CYC : 1;        // rows sharing one set of constants
  static_assert(D == 0 || ROWS % CYC == 0, "tile must hold whole phase cycles");
  static_assert(D == 0 || (CYC * STEP) % MOD == 0, "");
};

// ---- memory helpers ---------------------------------------------------------------------------
// Every byte is touched once, so loads and stores carry the streaming (evict-first) hint.  The
// SPZ_LD_MODE / SPZ_ST_MODE macros exist so scripts/ can build and time the alternatives
// (profiles/ records the comparison); the defaults are what ships.
#ifndef SPZ_LD_MODE
#define SPZ_LD_MODE 1  // 0: ld.global.cs   1: ld.global.nc   2: ld.global   3: ld.global.lu
#endif
#ifndef SPZ_DEC_HOIST
#define SPZ_DEC_HOIST false  // true: request all small-plane words of a sub-tile up front (measured: no gain)
#endif
#ifndef SPZ_DEC_PREFETCH
#define SPZ_DEC_PREFETCH 0
#endif
#ifndef SPZ_ST_MODE
#define SPZ_ST_MODE 0  // 0: st.global.cs   1: st.global      2: st.global.wt   3: st.global.cg
#endif
template <class T>
__device__ __forceinline__ T ldStream(const T *p) {
#if SPZ_LD_MODE == 0
  return __ldcs(p);
#elif SPZ_LD_MODE == 1
  return __ldg(p);
#elif SPZ_LD_MODE == 2
  return *p;
#else
  return __ldlu(p);
#endif
}
template <class T>
__device__ __forceinline__ void stStream(T *p, T v) {
#if SPZ_ST_MODE == 0
  __stcs(p, v);
#elif SPZ_ST_MODE == 1
  *p = v;
#elif SPZ_ST_MODE == 2
  __stwt(p, v);
#else
  __stcg(p, v);
#endif
}

// Phase bookkeeping of the SH plane: pos[e] = index of the thread's element e inside its
// 3*D-float SH record for the current row class; advance() moves to the next class.
template <int D, int S = kThreads>
struct ShPhase {
  int pos[4];
  __device__ __forceinline__ void init(int t) {
#pragma unroll
    for (int e = 0; e < 4; e++) pos[e] = (4 * t + e) % Geo<D, S>::MOD;
  }
  __device__ __forceinline__ void advance() {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      pos[e] += Geo<D, S>::STEP;
      if (pos[e] >= Geo<D, S>::MOD) pos[e] -= Geo<D, S>::MOD;
    }
  }
  __device__ __forceinline__ uint32_t flip(int e, uint32_t flipSh) const {
    return (flipSh >> ((uint32_t)pos[e] / 3u)) & 1u;
  }
};

// =================================================================================================
// encode, vector path
// =================================================================================================
template <int D, int MODE, bool V2 = false, int S = kThreads>
__global__ void __launch_bounds__(S, S == kThreads ? kCtasPerSm : SPZ_SMALL_CTAS_PER_SM)
encodeTilesKernel(const __grid_constant__ EncodeArgs a, const long long numTiles, const int restCtas) {
  constexpr int M = Geo<D, S>::M;
  __shared__ float sThr[256];
  __shared__ uint32_t sStage[S / 32][3 * 96];  // per warp: 3 rows x 96 position words
  pdlTrigger();
  if ((int)blockIdx.x < restCtas) {  // the sub-tile remainder rides in the first CTA(s): scalar_path.cuh
    const long long g = numTiles * Geo<D, S>::TG + (long long)blockIdx.x * S + threadIdx.x;
    pdlWait();
    if (g < a.n) encodeOneGaussian(a, g);
    return;
  }
  for (int i = threadIdx.x; i < 256; i += S) sThr[i] = a.alphaThresholds[i];
  __syncthreads();
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  uint32_t *stage = sStage[warp];

  // +-4096 for the xyz phase (t + k) mod 3; element e of row i has phase (t + e + 2i) mod 3
  // because a row is 4*320 = 1280 = 2 (mod 3) floats long
  float posScale3[3];
#pragma unroll
  for (int k = 0; k < 3; k++) posScale3[k] = signedConst(4096.0f, (a.flipP >> ((t + k) % 3)) & 1u);
  pdlWait();  // tables and constants are in place; the planes may only be touched from here on

  for (long long tile = (int)blockIdx.x - restCtas; tile < numTiles; tile += (int)gridDim.x - restCtas) {
#pragma unroll 1
    for (int mm = 0; mm < M; mm++) {
      const long long q = tile * M + mm;  // sub-tile: gaussians [q*1280, (q+1)*1280)
      // ---- positions: float4 -> three words (4 x 24 bit), staged so the stores are contiguous --
      {
        const float4 *in = reinterpret_cast<const float4 *>(a.positions) + q * (3 * S) + t;
        uint32_t *out = reinterpret_cast<uint32_t *>(a.oPositions) + q * (9 * S) + warp * 96 + lane;
        float4 v[3];
#pragma unroll
        for (int i = 0; i < 3; i++) v[i] = ldStream(in + i * S);
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const uint32_t n0 = m::quant_position24(v[i].x, posScale3[(0 + 2 * i) % 3]);
          const uint32_t n1 = m::quant_position24(v[i].y, posScale3[(1 + 2 * i) % 3]);
          const uint32_t n2 = m::quant_position24(v[i].z, posScale3[(2 + 2 * i) % 3]);
          const uint32_t n3 = m::quant_position24(v[i].w, posScale3[(3 + 2 * i) % 3]);
          uint32_t *sp = stage + i * 96 + 3 * lane;  // stride 3 words: conflict free
          sp[0] = prmt(n0, n1, 0x4210u);
          sp[1] = prmt(n1, n2, 0x5421u);
          sp[2] = prmt(n2, n3, 0x6542u);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
          for (int k = 0; k < 3; k++) stStream(out + i * (3 * S) + k * 32, stage[i * 96 + k * 32 + lane]);
        }
        __syncwarp();
      }
      // ---- scales and colours: float4 -> word ------------------------------------------------
      {
        const float4 *inS = reinterpret_cast<const float4 *>(a.scales) + q * (3 * S) + t;
        const float4 *inC = reinterpret_cast<const float4 *>(a.colors) + q * (3 * S) + t;
        uint32_t *outS = reinterpret_cast<uint32_t *>(a.oScales) + q * (3 * S) + t;
        uint32_t *outC = reinterpret_cast<uint32_t *>(a.oColors) + q * (3 * S) + t;
        float4 vs[3], vc[3];
#pragma unroll
        for (int i = 0; i < 3; i++) { vs[i] = ldStream(inS + i * S); vc[i] = ldStream(inC + i * S); }
#pragma unroll
        for (int i = 0; i < 3; i++) {
          stStream(outS + i * S, packSat4<MODE>(m::quant_scale_raw(vs[i].x), m::quant_scale_raw(vs[i].y),
                                                m::quant_scale_raw(vs[i].z), m::quant_scale_raw(vs[i].w)));
          stStream(outC + i * S, packSat4<MODE>(m::quant_color_raw(vc[i].x), m::quant_color_raw(vc[i].y),
                                                m::quant_color_raw(vc[i].z), m::quant_color_raw(vc[i].w)));
        }
      }
      // ---- alphas (float4 -> word) and rotations (quaternion -> word) -------------------------
      {
        const float4 va = ldStream(reinterpret_cast<const float4 *>(a.alphas) + q * S + t);
        // version 2 (3 bytes per quaternion): a warp owns 128 consecutive quaternions = 96 packed words; lane L takes
        // quaternions L, L+32, L+64, L+96 (four contiguous 512-byte warp loads), drops its 3 bytes each into the warp's
        // stage, and the 96 words leave as three contiguous lines -- the mirror image of the v2 decoder
        const float4 *inR = reinterpret_cast<const float4 *>(a.rotations) + q * (4 * S) + (V2 ? warp * 128 + lane : t);
        float4 vr[4];
#pragma unroll
        for (int i = 0; i < 4; i++) vr[i] = ldStream(inR + i * (V2 ? 32 : S));
        const uint32_t a0 = m::quant_alpha(va.x, sThr), a1 = m::quant_alpha(va.y, sThr);
        const uint32_t a2 = m::quant_alpha(va.z, sThr), a3 = m::quant_alpha(va.w, sThr);
        stStream(reinterpret_cast<uint32_t *>(a.oAlphas) + q * S + t,
                 a0 | (a1 << 8) | (a2 << 16) | (a3 << 24));
        if constexpr (V2) {
          uint8_t *sb = reinterpret_cast<uint8_t *>(stage);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const uint32_t b = m::quant_rotation_first3(vr[i].x, vr[i].y, vr[i].z, vr[i].w, a.flipQ);
            const int qi = 3 * (lane + 32 * i);
            sb[qi] = (uint8_t)b; sb[qi + 1] = (uint8_t)(b >> 8); sb[qi + 2] = (uint8_t)(b >> 16);
          }
          __syncwarp();
          uint32_t *outR = reinterpret_cast<uint32_t *>(a.oRotations) + q * (3 * S) + warp * 96 + lane;
#pragma unroll
          for (int k = 0; k < 3; k++) stStream(outR + k * 32, stage[k * 32 + lane]);
          __syncwarp();
        } else {
          uint32_t *outR = reinterpret_cast<uint32_t *>(a.oRotations) + q * (4 * S) + t;
#pragma unroll
          for (int i = 0; i < 4; i++)
            stStream(outR + i * S,
                     m::quant_rotation_smallest3(vr[i].x, vr[i].y, vr[i].z, vr[i].w, a.flipQ));
        }
      }
    }
    // ---- spherical harmonics: float4 -> word; rows c, c+CYC, ... share their constants ---------
    if (D > 0) {
      constexpr int U = Geo<D, S>::U, CYC = Geo<D, S>::CYC, ROWS = Geo<D, S>::ROWS;
      const float4 *in = reinterpret_cast<const float4 *>(a.sh) + tile * ((long long)ROWS * S) + t;
      uint32_t *out = reinterpret_cast<uint32_t *>(a.oSh) + tile * ((long long)ROWS * S) + t;
      ShPhase<D, S> ph;
      ph.init(t);
#pragma unroll 1
      for (int c = 0; c < CYC; c++) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ldStream(in + (c + u * CYC) * S);
        float shMul[4];
        uint32_t shAdd[4], shMask[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          shMul[e] = signedConst(128.0f, ph.flip(e, a.flipSh));
          const bool fine = ph.pos[e] < 9;  // 5-bit band: first 9 values, load-spz.cc:312-326
          shAdd[e] = fine ? 132u : 136u;    // 128 + bucket/2
          shMask[e] = fine ? ~7u : ~15u;    // ~(bucket-1)
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          stStream(out + (c + u * CYC) * S,
                   packSat4<MODE>(m::quant_sh_raw(v[u].x, shMul[0], shAdd[0], shMask[0]),
                                  m::quant_sh_raw(v[u].y, shMul[1], shAdd[1], shMask[1]),
                                  m::quant_sh_raw(v[u].z, shMul[2], shAdd[2], shMask[2]),
                                  m::quant_sh_raw(v[u].w, shMul[3], shAdd[3], shMask[3])));
        }
        ph.advance();
      }
    }
  }
}

// =================================================================================================
// encode, scalar path: remainders, tiny clouds, under-aligned pointers.  One thread per gaussian.
// =================================================================================================
__global__ void __launch_bounds__(128)
encodeGenericKernel(const __grid_constant__ EncodeArgs a, const long long first) {
  const long long g = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  pdlTrigger();
  pdlWait();
  if (g < a.n) encodeOneGaussian(a, g);
}

// =================================================================================================
// decode, vector path.  VER: 1 = half positions + first-three quaternion, 2 = 24-bit positions +
// first-three, 3 = 24-bit + smallest-three (load-spz.cc:571-572); 4 = half positions +
// smallest-three, which no file produces but a hand-built PackedGaussians can (load-spz.cc:465,509).
// =================================================================================================

// Per-thread constants of the xyz planes: sign bit / signed scale for the phase (t + k) mod 3.
struct DecodePosConsts {
  uint32_t posFlip3[3];
  float posScale3[3];
  __device__ __forceinline__ void init(const DecodeArgs &a, int t) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      posFlip3[k] = ((a.flipP >> ((t + k) % 3)) & 1u) << 31;
      posScale3[k] = __uint_as_float(__float_as_uint(a.positionScale) ^ posFlip3[k]);
    }
  }
};

// positions, scales, colours, alphas and rotations of sub-tile q (gaussians [q*1280, (q+1)*1280)):
// direct 128-byte-aligned loads and stores, shared by both decode kernels.  `stage` is this warp's
// 288-word scratch.
// Where the expanded float4s of a sub-tile go: one float4 pointer per plane, already advanced to the
// sub-tile; `idx` is the float4 index inside that plane's part of the sub-tile.  (Round 2 also tried a
// shared-memory sink feeding bulk async stores for SH-less clouds -- commit 251cd9c, slower at every
// size, profiles/r2_tuning_notes.txt section 6 -- which is why the sink is a policy.)
struct GlobalSink {  // straight to the output planes with streaming stores
  float4 *pos, *scale, *rot, *alpha, *color;
  template <int S>
  static __device__ __forceinline__ GlobalSink at(const DecodeArgs &a, long long q) {
    return {reinterpret_cast<float4 *>(a.oPositions) + q * (3 * S), reinterpret_cast<float4 *>(a.oScales) + q * (3 * S),
            reinterpret_cast<float4 *>(a.oRotations) + q * (4 * S), reinterpret_cast<float4 *>(a.oAlphas) + q * S,
            reinterpret_cast<float4 *>(a.oColors) + q * (3 * S)};
  }
  static __device__ __forceinline__ void put(float4 *plane, int idx, float4 v) { stStream(plane + idx, v); }
};
template <int VER, bool HOIST, int S = kThreads, class Sink = GlobalSink>
__device__ __forceinline__ void decodeSmallPlanes(const DecodeArgs &a, const long long q, const int t, uint32_t *stage,
                                                  const float *sAlpha, const float *sColor, const float *sMag,
                                                  const DecodePosConsts &pc, const Sink &sink) {
  static_assert((4 * S) % 3 == 2, "the xyz phase of element e in row i is taken as (t + e + 2 i) mod 3");
  constexpr bool kHalf = (VER == 1 || VER == 4);  // float16 positions
  constexpr bool kS3 = (VER >= 3);                // smallest-three rotations
  const int lane = t & 31, warp = t >> 5;
  const uint32_t *posFlip3 = pc.posFlip3;
  const float *posScale3 = pc.posScale3;
  // The packed words of the five planes: 20 registers.  HOIST requests all of them before the
  // first value is expanded (the bulk kernel runs 20 warps per SM and has the registers to spare);
  // otherwise each plane's words are requested right before use.
  uint2 w[3];
  uint32_t g[3][3], ws[3], wc[3], wa, wr[4];
  auto loadPositions = [&] {
    if (kHalf) {
      const uint2 *in = reinterpret_cast<const uint2 *>(a.positions) + q * (3 * S) + t;
#pragma unroll
      for (int i = 0; i < 3; i++) w[i] = ldStream(in + i * S);
    } else {
      // contiguous 128-byte loads of the 96 words a warp needs per row (re-dealt through the stage below)
      const uint32_t *in = reinterpret_cast<const uint32_t *>(a.positions) + q * (9 * S) + warp * 96 + lane;
#pragma unroll
      for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int k = 0; k < 3; k++) g[i][k] = ldStream(in + i * (3 * S) + k * 32);
      }
    }
  };
  auto loadScalesColors = [&] {
    const uint32_t *inS = reinterpret_cast<const uint32_t *>(a.scales) + q * (3 * S) + t;
    const uint32_t *inC = reinterpret_cast<const uint32_t *>(a.colors) + q * (3 * S) + t;
#pragma unroll
    for (int i = 0; i < 3; i++) { ws[i] = ldStream(inS + i * S); wc[i] = ldStream(inC + i * S); }
  };
  auto loadAlphas = [&] { wa = ldStream(reinterpret_cast<const uint32_t *>(a.alphas) + q * S + t); };
  auto loadRotations = [&] {
    if (kS3) {
      const uint32_t *in = reinterpret_cast<const uint32_t *>(a.rotations) + q * (4 * S) + t;
#pragma unroll
      for (int i = 0; i < 4; i++) wr[i] = ldStream(in + i * S);
    }
  };
  if (HOIST) {
    loadPositions();
    loadScalesColors();
    loadAlphas();
    loadRotations();
  }
  // ---- positions ------------------------------------------------------------------------
  {
    if (!HOIST) loadPositions();
    if (kHalf) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        float4 o;
        o.x = __uint_as_float(__float_as_uint(m::half_bits_to_float(w[i].x & 0xffffu)) ^ posFlip3[(0 + 2 * i) % 3]);
        o.y = __uint_as_float(__float_as_uint(m::half_bits_to_float(w[i].x >> 16)) ^ posFlip3[(1 + 2 * i) % 3]);
        o.z = __uint_as_float(__float_as_uint(m::half_bits_to_float(w[i].y & 0xffffu)) ^ posFlip3[(2 + 2 * i) % 3]);
        o.w = __uint_as_float(__float_as_uint(m::half_bits_to_float(w[i].y >> 16)) ^ posFlip3[(3 + 2 * i) % 3]);
        Sink::put(sink.pos, i * S + t, o);
      }
    } else {
      // re-deal through the per-warp stage so each lane gets the three words of its four 24-bit values
#pragma unroll
      for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int k = 0; k < 3; k++) stage[i * 96 + k * 32 + lane] = g[i][k];
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const uint32_t *sp = stage + i * 96 + 3 * lane;
        const uint32_t w0 = sp[0], w1 = sp[1], w2 = sp[2];
        // PRMT with the sign-replicate bit (selector nibble 8|idx) sign-extends 24 -> 32 bits
        const int32_t f0 = (int32_t)prmt(w0, w0, 0xA210u);
        const int32_t f1 = (int32_t)prmt(w0, w1, 0xD543u);
        const int32_t f2 = (int32_t)prmt(w1, w2, 0xC432u);
        const int32_t f3 = (int32_t)prmt(w2, w2, 0xB321u);
        float4 o;
        o.x = m::mul(m::i2f(f0), posScale3[(0 + 2 * i) % 3]);
        o.y = m::mul(m::i2f(f1), posScale3[(1 + 2 * i) % 3]);
        o.z = m::mul(m::i2f(f2), posScale3[(2 + 2 * i) % 3]);
        o.w = m::mul(m::i2f(f3), posScale3[(3 + 2 * i) % 3]);
        Sink::put(sink.pos, i * S + t, o);
      }
      __syncwarp();
    }
  }
  // ---- scales (exact FMA on the magic float) and colours (table) ---------------------------
  {
    if (!HOIST) loadScalesColors();
#pragma unroll
    for (int i = 0; i < 3; i++) {
      // (2^23 + s) / 16 - (2^19 + 10) = s/16 - 10: both steps exact, so the fused form equals
      // the reference's s / 16.0f - 10.0f (load-spz.cc:506) bit for bit, +0 at s = 160.
      float4 o;
      o.x = __fmaf_rn(byteAsMagicFloat<0>(ws[i]), 0.0625f, -524298.0f);
      o.y = __fmaf_rn(byteAsMagicFloat<1>(ws[i]), 0.0625f, -524298.0f);
      o.z = __fmaf_rn(byteAsMagicFloat<2>(ws[i]), 0.0625f, -524298.0f);
      o.w = __fmaf_rn(byteAsMagicFloat<3>(ws[i]), 0.0625f, -524298.0f);
      Sink::put(sink.scale, i * S + t, o);
      float4 c;
      c.x = sColor[wc[i] & 0xffu];
      c.y = sColor[(wc[i] >> 8) & 0xffu];
      c.z = sColor[(wc[i] >> 16) & 0xffu];
      c.w = sColor[wc[i] >> 24];
      Sink::put(sink.color, i * S + t, c);
    }
  }
  // ---- alphas (table) ---------------------------------------------------------------------
  {
    if (!HOIST) loadAlphas();
    float4 o;
    o.x = sAlpha[wa & 0xffu];
    o.y = sAlpha[(wa >> 8) & 0xffu];
    o.z = sAlpha[(wa >> 16) & 0xffu];
    o.w = sAlpha[wa >> 24];
    Sink::put(sink.alpha, t, o);
  }
  // ---- rotations ----------------------------------------------------------------------------
  if (kS3) {
    if (!HOIST) loadRotations();
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float r[4];
      m::dequant_rotation_smallest3(wr[i], sMag, a.flipQ, r);
      Sink::put(sink.rot, i * S + t, make_float4(r[0], r[1], r[2], r[3]));
    }
  } else {
    // 3 bytes per quaternion.  A warp owns 128 consecutive quaternions = 96 words, loaded as
    // three contiguous lines into the stage; lane L then decodes quaternions L, L+32, L+64,
    // L+96 so that each of its four float4 stores is a contiguous 512-byte warp store.
    const uint32_t *in = reinterpret_cast<const uint32_t *>(a.rotations) + q * (3 * S) + warp * 96 + lane;
#pragma unroll
    for (int k = 0; k < 3; k++) stage[k * 32 + lane] = ldStream(in + k * 32);
    __syncwarp();
    const uint8_t *sb = reinterpret_cast<const uint8_t *>(stage);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int qi = 3 * (lane + 32 * i);
      float r[4];
      m::dequant_rotation_first3(sb[qi], sb[qi + 1], sb[qi + 2], a.flipQ, r);
      Sink::put(sink.rot, warp * 128 + lane + 32 * i, make_float4(r[0], r[1], r[2], r[3]));
    }
    __syncwarp();
  }
}

// Development knobs of the register-path tile decoder (scripts/ build and time the alternatives):
//   SPZ_DEC_TAB_L1       1: the three decode tables are read through L1 instead of being copied into every CTA's shared memory
//   SPZ_DEC_TILES_HOIST  true: all 20 small-plane words of a sub-tile are requested before the first is expanded
//   SPZ_DEC0_CTAS_PER_SM resident CTAs per SM the SH-less instantiation is register-bounded for (4: 48 registers, 5: 40, 6: 32)
#ifndef SPZ_DEC_TAB_L1
#define SPZ_DEC_TAB_L1 0
#endif
#ifndef SPZ_DEC_TILES_HOIST
#define SPZ_DEC_TILES_HOIST false
#endif
#ifndef SPZ_DEC0_CTAS_PER_SM
#define SPZ_DEC0_CTAS_PER_SM SPZ_CTAS_PER_SM
#endif
template <int D, int VER, int S = kThreads>
__global__ void __launch_bounds__(S, S != kThreads ? SPZ_SMALL_CTAS_PER_SM : D == 0 ? SPZ_DEC0_CTAS_PER_SM : kCtasPerSm)
decodeTilesKernel(const __grid_constant__ DecodeArgs a, const long long numTiles, const int restCtas) {
  constexpr int M = Geo<D, S>::M;
  constexpr bool kS3 = (VER >= 3);
  __shared__ float sTab[SPZ_DEC_TAB_L1 ? 1 : (kS3 ? kDecodeTableFloats : 512)];
  __shared__ uint32_t sStage[S / 32][3 * 96];  // per warp: 3 rows x 96 position words
  pdlTrigger();
  if ((int)blockIdx.x < restCtas) {  // the sub-tile remainder rides in the first CTA(s): scalar_path.cuh
    const long long g = numTiles * Geo<D, S>::TG + (long long)blockIdx.x * S + threadIdx.x;
    pdlWait();
    if (g < a.n) decodeOneGaussian(a, g);
    return;
  }
#if SPZ_DEC_TAB_L1
  const float *sAlpha = a.tables, *sColor = a.tables + 256, *sMag = a.tables + 512;
#else
  for (int i = threadIdx.x; i < (kS3 ? kDecodeTableFloats : 512); i += S) sTab[i] = __ldg(a.tables + i);
  const float *sAlpha = sTab, *sColor = sTab + 256, *sMag = sTab + 512;
  __syncthreads();
#endif
  const int t = threadIdx.x;
  uint32_t *stage = sStage[t >> 5];
  DecodePosConsts pc;
  pc.init(a, t);
  pdlWait();  // tables and constants are in place; the planes may only be touched from here on

  for (long long tile = (int)blockIdx.x - restCtas; tile < numTiles; tile += (int)gridDim.x - restCtas) {
#pragma unroll 1
    for (int mm = 0; mm < M; mm++)
      decodeSmallPlanes<VER, SPZ_DEC_TILES_HOIST, S>(a, tile * M + mm, t, stage, sAlpha, sColor, sMag, pc, GlobalSink::at<S>(a, tile * M + mm));
    // ---- spherical harmonics: word -> float4 -------------------------------------------------
    if (D > 0) {
      constexpr int U = Geo<D, S>::U, CYC = Geo<D, S>::CYC, ROWS = Geo<D, S>::ROWS;
      const uint32_t *in = reinterpret_cast<const uint32_t *>(a.sh) + tile * ((long long)ROWS * S) + t;
      float4 *out = reinterpret_cast<float4 *>(a.oSh) + tile * ((long long)ROWS * S) + t;
      ShPhase<D, S> ph;
      ph.init(t);
#if SPZ_DEC_PREFETCH
      // software pipeline: the words of class c+1 are requested before class c is expanded
      uint32_t w[U], wn[U];
#pragma unroll
      for (int u = 0; u < U; u++) w[u] = ldStream(in + (u * CYC) * S);
#pragma unroll 1
      for (int c = 0; c < CYC; c++) {
        if (c + 1 < CYC) {
#pragma unroll
          for (int u = 0; u < U; u++) wn[u] = ldStream(in + (c + 1 + u * CYC) * S);
        }
#else
#pragma unroll 1
      for (int c = 0; c < CYC; c++) {
        uint32_t w[U];
#pragma unroll
        for (int u = 0; u < U; u++) w[u] = ldStream(in + (c + u * CYC) * S);
#endif
        float shMul[4];
#pragma unroll
        for (int e = 0; e < 4; e++) shMul[e] = signedConst(0.0078125f, ph.flip(e, a.flipSh));
#pragma unroll
        for (int u = 0; u < U; u++) {
          // (2^23 + x) - (2^23 + 128) = x - 128 exactly (+0 at x = 128), then * +-1/128: the
          // reference's ((float)x - 128.0f) / 128.0f followed by the flip (load-spz.cc:83).
          float4 o;
          o.x = m::mul(m::add(byteAsMagicFloat<0>(w[u]), -8388736.0f), shMul[0]);
          o.y = m::mul(m::add(byteAsMagicFloat<1>(w[u]), -8388736.0f), shMul[1]);
          o.z = m::mul(m::add(byteAsMagicFloat<2>(w[u]), -8388736.0f), shMul[2]);
          o.w = m::mul(m::add(byteAsMagicFloat<3>(w[u]), -8388736.0f), shMul[3]);
          stStream(out + (c + u * CYC) * S, o);
        }
        ph.advance();
#if SPZ_DEC_PREFETCH
#pragma unroll
        for (int u = 0; u < U; u++) w[u] = wn[u];
#endif
      }
    }
    }
}

// ---- bulk-copy (TMA) staging of the SH plane ---------------------------------------------------------
// scripts/membench.cu (profiles/r1_membench_patterns.txt): for decode's write-heavy mix, moving the
// packed words in with ONE bulk async copy per tile (UBLKCP.S.G, completion on an mbarrier) and
// the expanded rows out with bulk async stores from shared memory (UBLKCP.G.S) sustains 6.76 TB/s
// where LDG.32 / STG.128 from registers sustains 6.09; for encode's read-heavy mix the two are
// equal (7.12 vs 7.06), so the encoder keeps its register path.
template <int D>
struct BulkGeo {
  static constexpr int SB = (D == 8) ? 4 : 5;  // rows per store batch (a class has U = 5 or 8 rows)
  static_assert(Geo<D>::U % SB == 0, "a class must split into whole store batches");
  static constexpr int kTableBytes = kDecodeTableFloats * 4;
  static constexpr int kWordBytes = Geo<D>::ROWS * kThreads * 4;  // the tile's packed SH words
  static constexpr int kOutBytes = 2 * SB * kThreads * 16;        // two batches of expanded rows
  static constexpr int kSmemBytes = kTableBytes + kWordBytes + kOutBytes;
  static_assert(kOutBytes >= kWarps * 3 * 96 * 4, "the out buffers double as the per-warp word stage");
};

template <int D, int VER>
__global__ void __launch_bounds__(kThreads, 2)
decodeTilesBulkKernel(const __grid_constant__ DecodeArgs a, const long long numTiles, const int restCtas) {
  static_assert(D > 0, "SH-less clouds use decodeTilesKernel");
  pdlTrigger();
  if ((int)blockIdx.x < restCtas) {  // the sub-tile remainder rides in the first CTA(s): scalar_path.cuh
    const long long g = numTiles * Geo<D>::TG + (long long)blockIdx.x * kThreads + threadIdx.x;
    pdlWait();
    if (g < a.n) decodeOneGaussian(a, g);
    return;
  }
  constexpr int S = kThreads;
  constexpr int M = Geo<D>::M;
  constexpr int U = Geo<D>::U, CYC = Geo<D>::CYC, ROWS = Geo<D>::ROWS, SB = BulkGeo<D>::SB;
  extern __shared__ __align__(128) unsigned char dynSmem[];
  __shared__ __align__(8) unsigned long long bar;
  float *sTab = reinterpret_cast<float *>(dynSmem);
  uint32_t *win = reinterpret_cast<uint32_t *>(dynSmem + BulkGeo<D>::kTableBytes);
  float4 *obuf = reinterpret_cast<float4 *>(dynSmem + BulkGeo<D>::kTableBytes + BulkGeo<D>::kWordBytes);
  const int t = threadIdx.x;
  for (int i = t; i < kDecodeTableFloats; i += kThreads) sTab[i] = __ldg(a.tables + i);
  if (t == 0) mbarInit(&bar);
  const float *sAlpha = sTab, *sColor = sTab + 256, *sMag = sTab + 512;
  __syncthreads();
  uint32_t *stage = reinterpret_cast<uint32_t *>(obuf) + (t >> 5) * (3 * 96);
  DecodePosConsts pc;
  pc.init(a, t);
  pdlWait();  // tables and constants are in place; the planes may only be touched from here on

  uint32_t parity = 0;
  for (long long tile = (int)blockIdx.x - restCtas; tile < numTiles; tile += (int)gridDim.x - restCtas, parity ^= 1u) {
    // the SH words of the whole tile: one bulk copy, in flight while the small planes are decoded
    if (t == 0)
      bulkLoad(win, reinterpret_cast<const uint32_t *>(a.sh) + tile * ((long long)ROWS * S), BulkGeo<D>::kWordBytes, &bar);
#pragma unroll 1
    for (int mm = 0; mm < M; mm++)
      decodeSmallPlanes<VER, SPZ_DEC_HOIST>(a, tile * M + mm, t, stage, sAlpha, sColor, sMag, pc, GlobalSink::at<kThreads>(a, tile * M + mm));
    mbarWait(&bar, parity);
    __syncthreads();  // every warp is done with its word stage, which the out buffers overlay

    float4 *out = reinterpret_cast<float4 *>(a.oSh) + tile * ((long long)ROWS * S);
    ShPhase<D> ph;
    ph.init(t);
    int batch = 0;
#pragma unroll 1
    for (int c = 0; c < CYC; c++) {
      float shMul[4];
#pragma unroll
      for (int e = 0; e < 4; e++) shMul[e] = signedConst(0.0078125f, ph.flip(e, a.flipSh));
#pragma unroll 1
      for (int h = 0; h < U / SB; h++, batch++) {
        float4 *b = obuf + (batch & 1) * (SB * S);
        if (batch >= 2) {  // the bulk stores that read this buffer two batches ago must be done with it
          if (t == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < SB; k++) {
          const uint32_t w = win[(c + (h * SB + k) * CYC) * S + t];
          // (2^23 + x) - (2^23 + 128) = x - 128 exactly (+0 at x = 128), then * +-1/128: the
          // reference's ((float)x - 128.0f) / 128.0f followed by the flip (load-spz.cc:83).
          float4 o;
          o.x = m::mul(m::add(byteAsMagicFloat<0>(w), -8388736.0f), shMul[0]);
          o.y = m::mul(m::add(byteAsMagicFloat<1>(w), -8388736.0f), shMul[1]);
          o.z = m::mul(m::add(byteAsMagicFloat<2>(w), -8388736.0f), shMul[2]);
          o.w = m::mul(m::add(byteAsMagicFloat<3>(w), -8388736.0f), shMul[3]);
          b[k * S + t] = o;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // make the STS visible to the copy engine
        __syncthreads();
        if (t == 0) {
#pragma unroll
          for (int k = 0; k < SB; k++) bulkStore(out + (long long)(c + (h * SB + k) * CYC) * S, b + k * S, S * 16);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      ph.advance();
    }
    // before the buffers are reused (next tile's word stage / words): stores have read them, and
    // every thread has read its last words
    if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
  }
  if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// =================================================================================================
// decode, scalar path
// =================================================================================================
__global__ void __launch_bounds__(128)
decodeGenericKernel(const __grid_constant__ DecodeArgs a, const long long first) {
  const long long g = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  pdlTrigger();
  pdlWait();
  if (g < a.n) decodeOneGaussian(a, g);
}

__global__ void buildDecodeTablesKernel(float *tables) {
  const int i = threadIdx.x;  // 512 threads
  if (i < 256) tables[256 + i] = m::dequant_color((uint32_t)i);
  tables[512 + i] = m::dequant_s3_magnitude((uint32_t)i);
}

__global__ void probePackKernel(int *ok) {
  const int32_t v[3][4] = {{-5, 300, 17, 255}, {0, 1, 2, 3}, {1000, -1000, 128, 64}};
  int good = 1;
  for (int i = 0; i < 3; i++) {
    const uint32_t x = packSat4<kPackAlu>(v[i][0], v[i][1], v[i][2], v[i][3]);
    const uint32_t y = packSat4<kPackCvt>(v[i][0], v[i][1], v[i][2], v[i][3]);
    if (x != y) good = 0;
  }
  *ok = good;
}

bool aligned(const void *p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

#ifndef SPZ_SMALL_TILES_MAX_POINTS
#define SPZ_SMALL_TILES_MAX_POINTS 16000000
#endif
constexpr long long kSmallTilesMaxPoints = SPZ_SMALL_TILES_MAX_POINTS;
// mode: 0 never, 1 up to kSmallTilesMaxPoints gaussians per launch, 2 always (LaunchPlan::smallTilesEncode / smallTilesDecode)
bool smallTilesWanted(const LaunchPlan &plan, int mode, int shDim, long long n) {
  if (!plan.flatGrid || shDim == 15 || mode == 0) return false;
  return mode == 2 || n <= kSmallTilesMaxPoints;
}

template <int D, int MODE, int S = kThreads>
cudaError_t launchEncodeTiles(const EncodeArgs &a, long long tiles, int grid, int restCtas, bool pdl, cudaStream_t s) {
  if (a.version == 2) return launchKernel(encodeTilesKernel<D, MODE, true, S>, grid + restCtas, S, 0, s, pdl, a, tiles, restCtas);
  return launchKernel(encodeTilesKernel<D, MODE, false, S>, grid + restCtas, S, 0, s, pdl, a, tiles, restCtas);
}

template <int D, int VER, int S = kThreads>
cudaError_t launchDecodeTilesVer(const DecodeArgs &a, long long tiles, int grid, int restCtas, bool bulk, bool pdl, cudaStream_t s) {
  if constexpr (D > 0 && S == kThreads) {
    if (bulk) {
      // per launch, not once: the attribute belongs to the current device, and one process may drive several
      const cudaError_t attr = cudaFuncSetAttribute(decodeTilesBulkKernel<D, VER>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    BulkGeo<D>::kSmemBytes);
      if (attr != cudaSuccess) return attr;
      return launchKernel(decodeTilesBulkKernel<D, VER>, grid + restCtas, kThreads, BulkGeo<D>::kSmemBytes, s, pdl, a, tiles, restCtas);
    }
  }
  return launchKernel(decodeTilesKernel<D, VER, S>, grid + restCtas, S, 0, s, pdl, a, tiles, restCtas);
}

template <int D, int S = kThreads>
cudaError_t launchDecodeTiles(const DecodeArgs &a, long long tiles, int grid, int restCtas, bool bulk, bool pdl, cudaStream_t s) {
  switch (a.version) {
    case 1: return launchDecodeTilesVer<D, 1, S>(a, tiles, grid, restCtas, bulk, pdl, s);
    case 2: return launchDecodeTilesVer<D, 2, S>(a, tiles, grid, restCtas, bulk, pdl, s);
    case 4: return launchDecodeTilesVer<D, 4, S>(a, tiles, grid, restCtas, bulk, pdl, s);
    default: return launchDecodeTilesVer<D, 3, S>(a, tiles, grid, restCtas, bulk, pdl, s);
  }
}

}  // namespace

int tileGaussians(int shDim) { return shDim == 3 ? Geo<3>::TG : Geo<15>::TG; }

cudaError_t launchEncode(const EncodeArgs &a, const LaunchPlan &plan, cudaStream_t stream,
                         int *launches) {
  int count = 0;
  if (launches) *launches = 0;
  if (a.n <= 0) return cudaSuccess;
  const bool vec = !plan.forceGeneric && aligned(a.positions, 16) && aligned(a.scales, 16) &&
                   aligned(a.rotations, 16) && aligned(a.alphas, 16) && aligned(a.colors, 16) &&
                   (a.shDim == 0 || aligned(a.sh, 16)) && aligned(a.oPositions, 4) &&
                   aligned(a.oScales, 4) && aligned(a.oRotations, 4) && aligned(a.oAlphas, 4) &&
                   aligned(a.oColors, 4) && (a.shDim == 0 || aligned(a.oSh, 4));
  // SH degree 3 and at most 24M gaussians: one thread per gaussian, planes in and out by bulk async copies
  // (pergaussian_kernels.cu: 7045 vs 6414 GB/s at 10M points); larger launches and the other degrees use the
  // register-path tiles below, which lead from ~40M points up.  SPZB200_ENCODE=tiles / =bulk force either.
  long long bulkDone = 0;
  if (cudaError_t e = launchEncodePerGaussianPlanar(a, plan, stream, &bulkDone); e != cudaSuccess) return e;
  if (bulkDone > 0) count++;
  // The 128-thread geometry (tiles 2.5x smaller, 10 CTAs per SM) for SH degree 0, 1, 2: measured ahead of the 320-thread one for the
  // encoder from 2.5M gaussians up (SH0 +2 % at 10M, +7 % at 40M; SH2 +1.5 %; SH1 +1-2 %: profiles/r2_tuning_notes.txt section 12) and
  // 1-2 % behind it for the SH-less decoder, which keeps 320.  SPZB200_TILE=128 / 320 force either for both directions.
  if (vec && bulkDone == 0 && smallTilesWanted(plan, plan.smallTilesEncode, a.shDim, a.n)) {
    const long long tgS = a.shDim == 3 ? Geo<3, kSmallThreads>::TG : Geo<0, kSmallThreads>::TG;
    const long long tilesS = a.n / tgS;
    if (tilesS > 0 && tilesS < 0x7fffffffLL) {
      const int restS = plan.foldRest ? (int)((a.n - tilesS * tgS + kSmallThreads - 1) / kSmallThreads) : 0;
      const bool cvt = plan.packMode == kPackCvt;
      cudaError_t e;
      switch (a.shDim) {
        case 0: e = cvt ? launchEncodeTiles<0, kPackCvt, kSmallThreads>(a, tilesS, (int)tilesS, restS, plan.pdl, stream) : launchEncodeTiles<0, kPackAlu, kSmallThreads>(a, tilesS, (int)tilesS, restS, plan.pdl, stream); break;
        case 3: e = cvt ? launchEncodeTiles<3, kPackCvt, kSmallThreads>(a, tilesS, (int)tilesS, restS, plan.pdl, stream) : launchEncodeTiles<3, kPackAlu, kSmallThreads>(a, tilesS, (int)tilesS, restS, plan.pdl, stream); break;
        default: e = cvt ? launchEncodeTiles<8, kPackCvt, kSmallThreads>(a, tilesS, (int)tilesS, restS, plan.pdl, stream) : launchEncodeTiles<8, kPackAlu, kSmallThreads>(a, tilesS, (int)tilesS, restS, plan.pdl, stream); break;
      }
      if (e != cudaSuccess) return e;
      count++;
      bulkDone = restS > 0 ? a.n : tilesS * tgS;
    }
  }
  const long long tg = tileGaussians(a.shDim);
  const long long tiles = vec && bulkDone == 0 ? a.n / tg : 0;
  int restCtas = 0;
  if (tiles > 0) {
    const int per = plan.ctasPerSm >= 1 && plan.ctasPerSm <= kCtasPerSm ? plan.ctasPerSm : kCtasPerSm;
    const long long cap = plan.flatGrid ? 0x7fffffffLL : (long long)plan.smCount * per;
    const int grid = (int)(tiles < cap ? tiles : cap);
    // one CTA per tile: the remainder (< one tile) is folded into the same launch as leading CTAs
    restCtas = plan.flatGrid && plan.foldRest ? (int)((a.n - tiles * tg + kThreads - 1) / kThreads) : 0;
    cudaError_t e;
    const bool cvt = plan.packMode == kPackCvt;
    switch (a.shDim) {
      case 0: e = cvt ? launchEncodeTiles<0, kPackCvt>(a, tiles, grid, restCtas, plan.pdl, stream) : launchEncodeTiles<0, kPackAlu>(a, tiles, grid, restCtas, plan.pdl, stream); break;
      case 3: e = cvt ? launchEncodeTiles<3, kPackCvt>(a, tiles, grid, restCtas, plan.pdl, stream) : launchEncodeTiles<3, kPackAlu>(a, tiles, grid, restCtas, plan.pdl, stream); break;
      case 8: e = cvt ? launchEncodeTiles<8, kPackCvt>(a, tiles, grid, restCtas, plan.pdl, stream) : launchEncodeTiles<8, kPackAlu>(a, tiles, grid, restCtas, plan.pdl, stream); break;
      case 15: e = cvt ? launchEncodeTiles<15, kPackCvt>(a, tiles, grid, restCtas, plan.pdl, stream) : launchEncodeTiles<15, kPackAlu>(a, tiles, grid, restCtas, plan.pdl, stream); break;
      default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    count++;
  }
  const long long first = restCtas > 0 ? a.n : bulkDone + tiles * tg;
  if (first < a.n) {
    const long long rest = a.n - first;
    const long long blocks = (rest + 127) / 128;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    cudaError_t e = launchKernel(encodeGenericKernel, (unsigned)blocks, 128, 0, stream, plan.pdl, a, first);
    if (e != cudaSuccess) return e;
    count++;
  }
  if (launches) *launches = count;
  return cudaSuccess;
}

cudaError_t launchDecode(const DecodeArgs &a, const LaunchPlan &plan, cudaStream_t stream,
                         int *launches) {
  int count = 0;
  if (launches) *launches = 0;
  if (a.n <= 0) return cudaSuccess;
  const bool vec = !plan.forceGeneric && aligned(a.positions, (a.version == 1 || a.version == 4) ? 8 : 4) &&
                   aligned(a.scales, 4) && aligned(a.rotations, 4) && aligned(a.alphas, 4) &&
                   aligned(a.colors, 4) && (a.shDim == 0 || aligned(a.sh, 4)) &&
                   aligned(a.oPositions, 16) && aligned(a.oScales, 16) &&
                   aligned(a.oRotations, 16) && aligned(a.oAlphas, 16) && aligned(a.oColors, 16) &&
                   (a.shDim == 0 || aligned(a.oSh, 16));
  // SH degree 1 - 3, 16-byte aligned planes: one thread per gaussian, 64 / 128 gaussians per CTA, planes in and out
  // by bulk async copies (pergaussian_kernels.cu).  Equal to the tile kernels below at 100M points (6697 vs 6684 GB/s
  // at degree 3) and ahead of them on smaller clouds, whose launches are dominated by ramp-up and tail (1.25M points
  // 5168 vs 4760; degree 1, 10M points 6373 vs 5463; degree 2, 10M 6447 vs 6268): profiles/r1_tuning_notes.txt.
  // Anything else falls through.
  long long pgDone = 0;
  if (cudaError_t e = launchDecodePerGaussianPlanar(a, plan, stream, &pgDone); e != cudaSuccess) return e;
  if (pgDone > 0) count++;
  if (vec && pgDone == 0 && smallTilesWanted(plan, plan.smallTilesDecode, a.shDim, a.n)) {  // as in launchEncode; off by default (measured 1-2 % slower)
    const long long tgS = a.shDim == 3 ? Geo<3, kSmallThreads>::TG : Geo<0, kSmallThreads>::TG;
    const long long tilesS = a.n / tgS;
    if (tilesS > 0 && tilesS < 0x7fffffffLL) {
      const int restS = plan.foldRest ? (int)((a.n - tilesS * tgS + kSmallThreads - 1) / kSmallThreads) : 0;
      cudaError_t e;
      switch (a.shDim) {
        case 0: e = launchDecodeTiles<0, kSmallThreads>(a, tilesS, (int)tilesS, restS, false, plan.pdl, stream); break;
        case 3: e = launchDecodeTiles<3, kSmallThreads>(a, tilesS, (int)tilesS, restS, false, plan.pdl, stream); break;
        default: e = launchDecodeTiles<8, kSmallThreads>(a, tilesS, (int)tilesS, restS, false, plan.pdl, stream); break;
      }
      if (e != cudaSuccess) return e;
      count++;
      pgDone = restS > 0 ? a.n : tilesS * tgS;
    }
  }
  const long long tg = tileGaussians(a.shDim);
  const long long tiles = vec && pgDone == 0 ? a.n / tg : 0;
  int restCtas = 0;
  if (tiles > 0) {
    const int per = plan.ctasPerSm >= 1 && plan.ctasPerSm <= kCtasPerSm ? plan.ctasPerSm : kCtasPerSm;
    const long long cap = plan.flatGrid ? 0x7fffffffLL : (long long)plan.smCount * per;
    const int grid = (int)(tiles < cap ? tiles : cap);
    // the bulk-copy kernel needs the packed SH plane 16-byte aligned (cp.async.bulk); 4-byte aligned
    // planes still take the register-path tile kernel
    const bool bulk = plan.decodeBulk && aligned(a.sh, 16);
    restCtas = plan.flatGrid && plan.foldRest ? (int)((a.n - tiles * tg + kThreads - 1) / kThreads) : 0;
    cudaError_t e;
    switch (a.shDim) {
      case 0: e = launchDecodeTiles<0>(a, tiles, grid, restCtas, false, plan.pdl, stream); break;
      case 3: e = launchDecodeTiles<3>(a, tiles, grid, restCtas, bulk, plan.pdl, stream); break;
      case 8: e = launchDecodeTiles<8>(a, tiles, grid, restCtas, bulk, plan.pdl, stream); break;
      case 15: e = launchDecodeTiles<15>(a, tiles, grid, restCtas, bulk, plan.pdl, stream); break;
      default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    count++;
  }
  const long long first = restCtas > 0 ? a.n : pgDone + tiles * tg;
  if (first < a.n) {
    const long long rest = a.n - first;
    const long long blocks = (rest + 127) / 128;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    cudaError_t e = launchKernel(decodeGenericKernel, (unsigned)blocks, 128, 0, stream, plan.pdl, a, first);
    if (e != cudaSuccess) return e;
    count++;
  }
  if (launches) *launches = count;
  return cudaSuccess;
}

cudaError_t buildDecodeTables(float *tables, cudaStream_t stream) {
  buildDecodeTablesKernel<<<1, 512, 0, stream>>>(tables);
  return cudaGetLastError();
}

// Self-check of the two division identities the rotation quantizer rests on (codec_math.cuh), on the
// device's own FMA / reciprocal / division units.  part 0: a / sqrt1_2 for every float a in
// {0} U [2^-81, 1.01] (thread i takes bit patterns i, i + threads, ...); part 1: x / b for
// `count` pseudo-random pairs of the guard's domain per thread.  Counts quotients that differ in any bit.
namespace {
__device__ __forceinline__ uint64_t splitmixStep(uint64_t &s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void divisionSelfCheckKernel(int part, uint64_t count, uint64_t seed, unsigned long long *wrong) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, threads = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  if (part == 0) {
    const uint32_t lo = 0x17000000u /* 2^-81 */, hi = 0x3f8147aeu /* 1.01f */;
    for (uint64_t u = (uint64_t)lo + tid; u <= hi; u += threads) {
      const float a = __uint_as_float((uint32_t)u);
      bad += __float_as_uint(m::div_by_sqrt1_2(a)) != __float_as_uint(__fdiv_rn(a, 0.70710678118654752440f));
    }
    if (tid == 0) bad += __float_as_uint(m::div_by_sqrt1_2(0.0f)) != 0u;
  } else {
    uint64_t s = seed ^ (tid * 0xd1342543de82ef95ull);
    for (uint64_t i = 0; i < count; i++) {
      const uint64_t r = splitmixStep(s), r2 = splitmixStep(s);
      uint32_t mant = (uint32_t)r & 0x7fffffu;
      if (mant == 0x7fffffu) mant = 0x7ffffeu;                  // the guard sends all-ones significands to the general form
      const uint32_t be = 107u + (uint32_t)((r >> 23) % 41u);   // b in [2^-20, 2^21)
      const float b = __uint_as_float((be << 23) | mant);
      uint32_t xe = 67u + (uint32_t)((r2 >> 23) % (be - 67u + 1u));  // |x| from 2^-60 up to b's binade
      if (((r2 >> 40) & 3u) == 0) xe = be;
      float x = __uint_as_float((xe << 23) | ((uint32_t)r2 & 0x7fffffu) | ((uint32_t)(r2 >> 63) << 31));
      if (fabsf(x) > __fmul_rn(b, 1.0000005f)) x = __uint_as_float((__float_as_uint(x) & 0x80000000u) | __float_as_uint(b));
      const float got = m::div_by_rcp(x, b, m::rcp_rn(b)), want = __fdiv_rn(x, b);
      bad += __float_as_uint(got) != __float_as_uint(want);
    }
  }
  if (bad) atomicAdd(wrong, bad);
}
}  // namespace

cudaError_t divisionSelfCheck(int part, unsigned long long pairsPerThread, unsigned long long seed, int smCount, cudaStream_t stream,
                              unsigned long long *wrong, unsigned long long *checked) {
  unsigned long long *d = nullptr;
  cudaError_t e = cudaMalloc(&d, sizeof *d);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(d, 0, sizeof *d, stream);
  const int blocks = smCount * 8, threads = 256;
  if (e == cudaSuccess) {
    divisionSelfCheckKernel<<<blocks, threads, 0, stream>>>(part, pairsPerThread, seed, d);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(wrong, d, sizeof *d, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(d);
  if (checked) *checked = part == 0 ? (unsigned long long)(0x3f8147aeu - 0x17000000u) + 2ull : pairsPerThread * (unsigned long long)blocks * threads;
  return e;
}

cudaError_t probePackCvt(cudaStream_t stream, int *scratch, int *ok) {
  probePackKernel<<<1, 1, 0, stream>>>(scratch);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(ok, scratch, sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  return e;
}

}  // namespace spzb200
