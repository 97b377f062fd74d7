// Fused PLY-rows encoder: the vertex records of a gaussian-splat .ply (row-major, `width` floats
// per vertex, the layout load-spz.cc:691-844 parses) straight to PackedGaussians planes, skipping
// the planar float GaussianCloud in between.  It computes exactly
//
//     packGaussians(loadSplatFromPly(rows, to = X), from = X)          
//
// i.e. the [N,C,S] -> [N,S,C] SH shuffle and wxyz -> xyzw quaternion reorder of the PLY loader
// (load-spz.cc:808-838), its RDF -> X flips and the encoder's X -> RUB flips (both are sign
// products, so they compose into one set of sign bits), then the quantizers of codec_math.cuh.
//
// This IS a transposition: a vertex record is 62 consecutive floats (248 B at SH degree 3) while
// every output plane wants consecutive gaussians.  A CTA therefore brings 128 records (31.7 KB,
// contiguous in memory) into shared memory with one bulk async copy and gathers from there: thread
// t produces output word t, t + 256, ... of each plane, picking its four source floats by
// (gaussian, column) -- record stride 62 words is 2-way bank conflicted at worst -- so all global
// stores are contiguous words and nothing is read from HBM twice: 313 B per gaussian at degree 3.
#include "codec_kernels.cuh"

#include "codec_math.cuh"

namespace spzb200 {
namespace {

constexpr int kPlyThreads = 256;
#ifndef SPZ_PLY_CTAS
#define SPZ_PLY_CTAS 6  // 40 registers: 4 -> 6 resident CTAs was +8 % at SH degree 3
#endif
constexpr int kPlyTileBase = 128;  // tile sizes are multiples of this
// gaussians per CTA: about 32 KB of records at the standard widths (62 / 41 / 26 / 17 floats)
template <int D>
constexpr int kPlyTileFor = D == 15 ? 128 : D == 8 ? 256 : D == 3 ? 256 : 512;

__device__ __forceinline__ uint32_t smemAddrPly(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float signedConstPly(float magnitude, uint32_t negate) {
  return __uint_as_float(__float_as_uint(magnitude) | (negate << 31));
}

// Column maps live in shared memory: indexing the kernel-parameter arrays with a lane-dependent
// index turns into divergent constant-bank loads, which replay once per distinct address (the first
// version did that and ran at 0.9 TB/s).
//   shMap[k], k = 3 * coefficient + channel: bits 0..15 source column, bit 30 = 5-bit band, bit 31 = flip
//   xyzMap[plane][axis]: source column (| flip << 31 for positions)
template <int D>
__global__ void __launch_bounds__(kPlyThreads, SPZ_PLY_CTAS)
encodePlyTilesKernel(const PlyEncodeArgs a, const long long numTiles) {
  extern __shared__ __align__(128) unsigned char dynSmem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ float sThr[256];
  __shared__ uint32_t shMap[D > 0 ? 3 * D : 1];
  __shared__ uint32_t xyzMap[3][3];
  float *rows = reinterpret_cast<float *>(dynSmem);
  const int t = threadIdx.x;
  for (int i = t; i < 256; i += kPlyThreads) sThr[i] = a.alphaThresholds[i];
  if (D > 0 && t < 3 * D) {
    const int sCoef = t / 3, ch = t - 3 * sCoef;
    shMap[t] = (uint32_t)a.colRest[ch * D + sCoef] | (t < 9 ? 1u << 30 : 0u) | (((a.flipSh >> sCoef) & 1u) << 31);
  }
  if (t < 3) {
    xyzMap[0][t] = (uint32_t)a.colPos[t] | (((a.flipP >> t) & 1u) << 31);
    xyzMap[1][t] = (uint32_t)a.colScale[t];
    xyzMap[2][t] = (uint32_t)a.colColor[t];
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddrPly(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int width = a.width;
  constexpr int G = kPlyTileFor<D>;
  const uint32_t tileBytes = (uint32_t)(G * width * 4);
  const int colAlpha = a.colAlpha, colRx = a.colRot[0], colRy = a.colRot[1], colRz = a.colRot[2], colRw = a.colRot[3];

  uint32_t parity = 0;
  for (long long tile = blockIdx.x; tile < numTiles; tile += gridDim.x, parity ^= 1u) {
    if (t == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddrPly(&bar)), "r"(tileBytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddrPly(rows)),
                   "l"(a.rows + tile * (long long)G * width), "r"(tileBytes), "r"(smemAddrPly(&bar))
                   : "memory");
    }
    {
      uint32_t done = 0;
      for (uint32_t spin = 0; !done; spin++) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smemAddrPly(&bar)), "r"(parity)
                     : "memory");
        if (spin > (1u << 28)) __trap();
      }
    }
    const long long g0 = tile * G;
    // ---- positions: 3G values -> 3G/4 groups of three words --------------------------------------
    for (int j = t; j < 3 * G / 4; j += kPlyThreads) {
      uint32_t n[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int i = 4 * j + e, g = i / 3, ax = i - 3 * g;
        const uint32_t mp = xyzMap[0][ax];
        n[e] = m::quant_position24(rows[g * width + (int)(mp & 0xffffu)], __uint_as_float(0x45800000u | (mp & 0x80000000u)));  // +-4096
      }
      uint32_t *o = reinterpret_cast<uint32_t *>(a.oPositions + g0 * 9) + 3 * j;
      o[0] = n[0] | (n[1] << 24);
      o[1] = (n[1] >> 8) | (n[2] << 16);
      o[2] = (n[2] >> 16) | (n[3] << 8);
    }
    // ---- scales, colours: 3G values -> 3G/4 words each -------------------------------------------
    for (int j = t; j < 3 * G / 4; j += kPlyThreads) {
      uint32_t ws = 0, wc = 0;
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int i = 4 * j + e, g = i / 3, ax = i - 3 * g;
        ws |= m::quant_scale(rows[g * width + (int)xyzMap[1][ax]]) << (8 * e);
        wc |= m::quant_color(rows[g * width + (int)xyzMap[2][ax]]) << (8 * e);
      }
      reinterpret_cast<uint32_t *>(a.oScales + g0 * 3)[j] = ws;
      reinterpret_cast<uint32_t *>(a.oColors + g0 * 3)[j] = wc;
    }
    // ---- alphas: G values -> G/4 words;  rotations: one word per gaussian ---------------------------
    for (int j = t; j < G / 4; j += kPlyThreads) {
      uint32_t w = 0;
#pragma unroll
      for (int e = 0; e < 4; e++) w |= m::quant_alpha(rows[(4 * j + e) * width + colAlpha], sThr) << (8 * e);
      reinterpret_cast<uint32_t *>(a.oAlphas + g0)[j] = w;
    }
    for (int g = t; g < G; g += kPlyThreads) {
      const float *row = rows + g * width;
      reinterpret_cast<uint32_t *>(a.oRotations + g0 * 4)[g] =
          m::quant_rotation_smallest3(row[colRx], row[colRy], row[colRz], row[colRw], a.flipQ);
    }
    // ---- spherical harmonics: 3*D*G values -> 3*D*G/4 words ---------------------------------------------
    if constexpr (D > 0) {
      uint32_t *o = reinterpret_cast<uint32_t *>(a.oSh + g0 * (3 * D));
      for (int j = t; j < 3 * D * G / 4; j += kPlyThreads) {
        const int i0 = 4 * j;
        int g = i0 / (3 * D), k = i0 - g * (3 * D);  // element 0 of the word; the others follow, wrapping into the next gaussian
        uint32_t word = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const uint32_t mp = shMap[k];
          const float x = rows[g * width + (int)(mp & 0xffffu)];
          const uint32_t fine = (mp >> 30) & 1u;
          word |= m::quant_sh(x, __uint_as_float(0x43000000u | (mp & 0x80000000u)), 136u - (fine << 2), ~15u | (fine << 3)) << (8 * e);
          if (++k == 3 * D) { k = 0; g++; }
        }
        o[j] = word;
      }
    }
    __syncthreads();  // everyone is done with the records before the next tile's copy lands
  }
}

// remainder / under-aligned rows: one thread per gaussian straight from global memory
__global__ void __launch_bounds__(128)
encodePlyGenericKernel(const PlyEncodeArgs a, const long long first) {
  const long long g = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.n) return;
  const float *row = a.rows + g * a.width;
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    const uint32_t n = m::quant_position24(row[a.colPos[ax]], signedConstPly(4096.0f, (a.flipP >> ax) & 1u));
    uint8_t *o = a.oPositions + (g * 3 + ax) * 3;
    o[0] = (uint8_t)n; o[1] = (uint8_t)(n >> 8); o[2] = (uint8_t)(n >> 16);
    a.oScales[g * 3 + ax] = (uint8_t)m::quant_scale(row[a.colScale[ax]]);
    a.oColors[g * 3 + ax] = (uint8_t)m::quant_color(row[a.colColor[ax]]);
  }
  a.oAlphas[g] = (uint8_t)m::quant_alpha(row[a.colAlpha], a.alphaThresholds);
  const uint32_t comp = m::quant_rotation_smallest3(row[a.colRot[0]], row[a.colRot[1]], row[a.colRot[2]], row[a.colRot[3]], a.flipQ);
  uint8_t *ro = a.oRotations + g * 4;
  ro[0] = (uint8_t)comp; ro[1] = (uint8_t)(comp >> 8); ro[2] = (uint8_t)(comp >> 16); ro[3] = (uint8_t)(comp >> 24);
  const int D = a.shDim;
  uint8_t *so = a.oSh + g * (3 * D);
  for (int k = 0; k < 3 * D; k++) {
    const int sCoef = k / 3, ch = k - 3 * sCoef;
    const uint32_t bucket = k < 9 ? 8u : 16u;
    so[k] = (uint8_t)m::quant_sh(row[a.colRest[ch * D + sCoef]], signedConstPly(128.0f, (a.flipSh >> sCoef) & 1u),
                                 128u + bucket / 2u, ~(bucket - 1u));
  }
}

// =================================================================================================
// decode to PLY rows: the mirror image.  The records of a tile are assembled in shared memory
// (scatter by column, 2-way bank conflicts at worst) and leave with one bulk async store.
// =================================================================================================
__device__ __forceinline__ float unpackS3Component(uint32_t field, const float *magLut) {
  return __uint_as_float(__float_as_uint(magLut[field & 511u]) | ((field & 512u) << 22));
}

// rotation of gaussian g (global index) -> out[4] = x, y, z, w with flips applied
__device__ __forceinline__ void decodeRotationAt(const PlyDecodeArgs &a, long long g, const float *sMag, float out[4]) {
  if (a.version >= 3) {
    const uint32_t comp = __ldg(reinterpret_cast<const uint32_t *>(a.rotations) + g);
    m::dequant_rotation_smallest3(comp, sMag, a.flipQ, out);
  } else {
    const uint8_t *b = a.rotations + g * 3;
    m::dequant_rotation_first3(b[0], b[1], b[2], a.flipQ, out);
  }
}

template <int D>
__global__ void __launch_bounds__(kPlyThreads)
decodePlyTilesKernel(const PlyDecodeArgs a, const long long numTiles) {
  extern __shared__ __align__(128) unsigned char dynSmem[];
  __shared__ float sTab[kDecodeTableFloats];
  __shared__ uint32_t shMap[D > 0 ? 3 * D : 1];  // k = 3 * coefficient + channel: column | flip << 31
  __shared__ uint32_t xyzMap[3][3];
  float *rows = reinterpret_cast<float *>(dynSmem);
  const int t = threadIdx.x;
  for (int i = t; i < kDecodeTableFloats; i += kPlyThreads) sTab[i] = __ldg(a.tables + i);
  if (D > 0 && t < 3 * D) {
    const int sCoef = t / 3, ch = t - 3 * sCoef;
    shMap[t] = (uint32_t)a.colRest[ch * D + sCoef] | (((a.flipSh >> sCoef) & 1u) << 31);
  }
  if (t < 3) {
    xyzMap[0][t] = (uint32_t)a.colPos[t] | (((a.flipP >> t) & 1u) << 31);
    xyzMap[1][t] = (uint32_t)a.colScale[t];
    xyzMap[2][t] = (uint32_t)a.colColor[t];
  }
  // columns no plane writes (the three normals, anything extra) must read 0; the host lists them so
  // only they are cleared per tile, not the whole 32 KB of records
  __shared__ uint16_t unmapped[kMaxUnmappedColumns];
  const int numUnmapped = a.numUnmapped;  // -1: too many for the list, clear everything
  if (t < numUnmapped) unmapped[t] = a.unmapped[t];
  __syncthreads();
  const float *sAlpha = sTab, *sColor = sTab + 256, *sMag = sTab + 512;
  const int width = a.width;
  constexpr int G = kPlyTileFor<D>;
  const bool half = a.version == 1 || a.version == 4;

  for (long long tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
    const long long g0 = tile * G;
    if (numUnmapped < 0) {
      for (int i = t; i < G * width; i += kPlyThreads) rows[i] = 0.0f;
      __syncthreads();  // the scatter below overwrites some of these
    } else {
      for (int i = t; i < G * numUnmapped; i += kPlyThreads) {
        const int g = i / numUnmapped;
        rows[g * width + unmapped[i - g * numUnmapped]] = 0.0f;
      }
    }
    // ---- positions --------------------------------------------------------------------------------------
    if (!half) {
      const uint32_t *in = reinterpret_cast<const uint32_t *>(a.positions + g0 * 9);
      for (int j = t; j < 3 * G / 4; j += kPlyThreads) {
        const uint32_t w0 = __ldg(in + 3 * j), w1 = __ldg(in + 3 * j + 1), w2 = __ldg(in + 3 * j + 2);
        const uint32_t lo[4] = {w0 & 0xffffffu, (w0 >> 24) | ((w1 & 0xffffu) << 8), (w1 >> 16) | ((w2 & 0xffu) << 16), w2 >> 8};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int i = 4 * j + e, g = i / 3, ax = i - 3 * g;
          const uint32_t mp = xyzMap[0][ax];
          rows[g * width + (int)(mp & 0xffffu)] = m::dequant_position24(lo[e], __uint_as_float(__float_as_uint(a.positionScale) ^ (mp & 0x80000000u)));
        }
      }
    } else {
      const uint16_t *in = reinterpret_cast<const uint16_t *>(a.positions + g0 * 6);
      for (int i = t; i < 3 * G; i += kPlyThreads) {
        const int g = i / 3, ax = i - 3 * g;
        const uint32_t mp = xyzMap[0][ax];
        rows[g * width + (int)(mp & 0xffffu)] = __uint_as_float(__float_as_uint(m::half_bits_to_float(in[i])) ^ (mp & 0x80000000u));
      }
    }
    // ---- scales, colours -----------------------------------------------------------------------------------
    for (int j = t; j < 3 * G / 4; j += kPlyThreads) {
      const uint32_t ws = __ldg(reinterpret_cast<const uint32_t *>(a.scales + g0 * 3) + j);
      const uint32_t wc = __ldg(reinterpret_cast<const uint32_t *>(a.colors + g0 * 3) + j);
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int i = 4 * j + e, g = i / 3, ax = i - 3 * g;
        rows[g * width + (int)xyzMap[1][ax]] = m::dequant_scale((ws >> (8 * e)) & 0xffu);
        rows[g * width + (int)xyzMap[2][ax]] = sColor[(wc >> (8 * e)) & 0xffu];
      }
    }
    // ---- alphas, rotations -------------------------------------------------------------------------------
    for (int j = t; j < G / 4; j += kPlyThreads) {
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(a.alphas + g0) + j);
#pragma unroll
      for (int e = 0; e < 4; e++) rows[(4 * j + e) * width + a.colAlpha] = sAlpha[(w >> (8 * e)) & 0xffu];
    }
    for (int g = t; g < G; g += kPlyThreads) {
      float r[4];
      decodeRotationAt(a, g0 + g, sMag, r);
      float *row = rows + g * width;
      row[a.colRot[0]] = r[0]; row[a.colRot[1]] = r[1]; row[a.colRot[2]] = r[2]; row[a.colRot[3]] = r[3];
    }
    // ---- spherical harmonics --------------------------------------------------------------------------------
    if constexpr (D > 0) {
      const uint32_t *in = reinterpret_cast<const uint32_t *>(a.sh + g0 * (3 * D));
      for (int j = t; j < 3 * D * G / 4; j += kPlyThreads) {
        const uint32_t w = __ldg(in + j);
        const int i0 = 4 * j;
        int g = i0 / (3 * D), k = i0 - g * (3 * D);
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const uint32_t mp = shMap[k];
          rows[g * width + (int)(mp & 0xffffu)] =
              m::dequant_sh((w >> (8 * e)) & 0xffu, __uint_as_float(0x3c000000u | (mp & 0x80000000u)));  // +-1/128
          if (++k == 3 * D) { k = 0; g++; }
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (t == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.rows + g0 * width), "r"(smemAddrPly(rows)),
                   "r"((uint32_t)(G * width * 4))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the records are reused by the next tile
    }
    __syncthreads();
  }
  if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// remainder / under-aligned rows: one thread per gaussian
__global__ void __launch_bounds__(128)
decodePlyGenericKernel(const PlyDecodeArgs a, const long long first) {
  const long long g = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.n) return;
  float *row = a.rows + g * a.width;
  for (int i = 0; i < a.width; i++) row[i] = 0.0f;
  const bool half = a.version == 1 || a.version == 4;
#pragma unroll
  for (int ax = 0; ax < 3; ax++) {
    const uint32_t flip = ((a.flipP >> ax) & 1u) << 31;
    float p;
    if (half) {
      const uint8_t *h = a.positions + (g * 3 + ax) * 2;
      p = __uint_as_float(__float_as_uint(m::half_bits_to_float((uint32_t)h[0] | ((uint32_t)h[1] << 8))) ^ flip);
    } else {
      const uint8_t *b = a.positions + (g * 3 + ax) * 3;
      p = m::dequant_position24((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16),
                                __uint_as_float(__float_as_uint(a.positionScale) ^ flip));
    }
    row[a.colPos[ax]] = p;
    row[a.colScale[ax]] = m::dequant_scale(a.scales[g * 3 + ax]);
    row[a.colColor[ax]] = a.tables[256 + a.colors[g * 3 + ax]];
  }
  row[a.colAlpha] = a.tables[a.alphas[g]];
  float r[4];
  if (a.version >= 3) {
    const uint8_t *b = a.rotations + g * 4;
    m::dequant_rotation_smallest3((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24),
                                  a.tables + 512, a.flipQ, r);
  } else {
    const uint8_t *b = a.rotations + g * 3;
    m::dequant_rotation_first3(b[0], b[1], b[2], a.flipQ, r);
  }
  for (int i = 0; i < 4; i++) row[a.colRot[i]] = r[i];
  const int D = a.shDim;
  const uint8_t *s = a.sh + g * (3 * D);
  for (int k = 0; k < 3 * D; k++) {
    const int sCoef = k / 3, ch = k - 3 * sCoef;
    row[a.colRest[ch * D + sCoef]] = m::dequant_sh(s[k], signedConstPly(0.0078125f, (a.flipSh >> sCoef) & 1u));
  }
}

bool alignedTo(const void *p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

template <int D>
cudaError_t launchPlyTiles(const PlyEncodeArgs &a, long long tiles, cudaStream_t s) {
  const int smem = kPlyTileFor<D> * a.width * 4;
  cudaError_t e = cudaFuncSetAttribute(encodePlyTilesKernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  encodePlyTilesKernel<D><<<(unsigned)(tiles < 0x7fffffffLL ? tiles : 0x7fffffffLL), kPlyThreads, smem, s>>>(a, tiles);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launchEncodePly(const PlyEncodeArgs &a, const LaunchPlan &plan, cudaStream_t stream, int *launches) {
  int count = 0;
  if (launches) *launches = 0;
  if (a.n <= 0) return cudaSuccess;
  // the bulk copy wants 16-byte aligned records; a tile of `width`-float records always is a
  // multiple of 16 bytes, so only the base pointer matters
  const bool vec = !plan.forceGeneric && alignedTo(a.rows, 16) && alignedTo(a.oPositions, 4) && alignedTo(a.oScales, 4) &&
                   alignedTo(a.oRotations, 4) && alignedTo(a.oAlphas, 4) && alignedTo(a.oColors, 4) &&
                   (a.shDim == 0 || alignedTo(a.oSh, 4));
  const int tileG = a.shDim == 15 ? kPlyTileFor<15> : a.shDim == 8 ? kPlyTileFor<8> : a.shDim == 3 ? kPlyTileFor<3> : kPlyTileFor<0>;
  // 200 KB of shared memory bound the record width the staged kernel accepts (wider: scalar kernel)
  // the property order the reference writes: compile-time columns (pergaussian_kernels.cu)
  long long canon = 0;
  if (cudaError_t e = launchEncodePlyCanonical(a, plan, stream, &canon); e != cudaSuccess) return e;
  if (canon > 0) count++;
  const long long tiles = canon == 0 && vec && (long long)tileG * a.width * 4 <= 200 * 1024 ? a.n / tileG : 0;
  if (tiles > 0) {
    cudaError_t e;
    switch (a.shDim) {
      case 0: e = launchPlyTiles<0>(a, tiles, stream); break;
      case 3: e = launchPlyTiles<3>(a, tiles, stream); break;
      case 8: e = launchPlyTiles<8>(a, tiles, stream); break;
      case 15: e = launchPlyTiles<15>(a, tiles, stream); break;
      default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    count++;
  }
  const long long first = canon + tiles * tileG;
  if (first < a.n) {
    const long long blocks = (a.n - first + 127) / 128;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    encodePlyGenericKernel<<<(unsigned)blocks, 128, 0, stream>>>(a, first);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    count++;
  }
  if (launches) *launches = count;
  return cudaSuccess;
}

template <int D>
cudaError_t launchPlyDecodeTiles(const PlyDecodeArgs &a, long long tiles, cudaStream_t s) {
  const int smem = kPlyTileFor<D> * a.width * 4;
  cudaError_t e = cudaFuncSetAttribute(decodePlyTilesKernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  decodePlyTilesKernel<D><<<(unsigned)(tiles < 0x7fffffffLL ? tiles : 0x7fffffffLL), kPlyThreads, smem, s>>>(a, tiles);
  return cudaGetLastError();
}

cudaError_t launchDecodePly(const PlyDecodeArgs &a, const LaunchPlan &plan, cudaStream_t stream, int *launches) {
  int count = 0;
  if (launches) *launches = 0;
  if (a.n <= 0) return cudaSuccess;
  const bool half = a.version == 1 || a.version == 4;
  const bool vec = !plan.forceGeneric && alignedTo(a.rows, 16) && alignedTo(a.positions, half ? 2 : 4) && alignedTo(a.scales, 4) &&
                   alignedTo(a.alphas, 4) && alignedTo(a.colors, 4) && (a.version < 3 || alignedTo(a.rotations, 4)) &&
                   (a.shDim == 0 || alignedTo(a.sh, 4));
  const int tileG = a.shDim == 15 ? kPlyTileFor<15> : a.shDim == 8 ? kPlyTileFor<8> : a.shDim == 3 ? kPlyTileFor<3> : kPlyTileFor<0>;
  long long canon = 0;
  if (cudaError_t e = launchDecodePlyCanonical(a, plan, stream, &canon); e != cudaSuccess) return e;
  if (canon > 0) count++;
  const long long tiles = canon == 0 && vec && (long long)tileG * a.width * 4 <= 200 * 1024 ? a.n / tileG : 0;
  if (tiles > 0) {
    cudaError_t e;
    switch (a.shDim) {
      case 0: e = launchPlyDecodeTiles<0>(a, tiles, stream); break;
      case 3: e = launchPlyDecodeTiles<3>(a, tiles, stream); break;
      case 8: e = launchPlyDecodeTiles<8>(a, tiles, stream); break;
      case 15: e = launchPlyDecodeTiles<15>(a, tiles, stream); break;
      default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    count++;
  }
  const long long first = canon + tiles * tileG;
  if (first < a.n) {
    const long long blocks = (a.n - first + 127) / 128;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    decodePlyGenericKernel<<<(unsigned)blocks, 128, 0, stream>>>(a, first);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    count++;
  }
  if (launches) *launches = count;
  return cudaSuccess;
}

int plyTileGaussians() { return kPlyTileFor<0>; }  // the largest tile: a granule every tile size divides

}  // namespace spzb200
