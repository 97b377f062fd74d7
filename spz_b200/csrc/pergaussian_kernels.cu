// One-thread-per-gaussian kernels: the fused PLY-rows encoder and decoder for the CANONICAL record
// layout, and (same kernel, other source) the planar encoder of packGaussians.  The canonical
// layout is the property order the reference's saveSplatToPly writes (load-spz.cc:892-916) and the
// 3DGS trainers emit:
//
//     x y z  nx ny nz  f_dc_0..2  f_rest_0..3D-1 (channel-major)  opacity  scale_0..2  rot_0..3 (w x y z)
//
// so every column index is a compile-time constant and the column maps of ply_kernels.cu (which
// stay for any other property order) disappear.  One thread per gaussian, 64 - 256 gaussians per CTA:
//
//   rows -> packed (encodePerGaussianKernel<RowsSource>): the tile's records (31.7 KB at SH degree 3) arrive with
//     one bulk async copy; thread g reads record g from shared memory (64-bit loads, conflict
//     free at the even widths, 32-bit at the odd ones), quantises its 59 values with immediate
//     constants and deposits its bytes into a shared-memory image of the six packed planes, which
//     leaves with six bulk async stores.
//   planes -> packed (encodePerGaussianKernel<PlanarSource>): as above, but the tile's six float
//     planes arrive with six bulk copies and thread g picks its 56 + 12D bytes out of them (odd
//     word strides, conflict free).  The default encoder for SH degree 3 clouds of up to 24M gaussians
//     (7045 vs 6414 GB/s at 10M points); above that the register-path tiles of codec_kernels.cu lead by
//     about 1 %, and the single 31.7 KB copy of the rows source stays ahead of both (7145 GB/s);
//     profiles/r1_tuning_notes.txt.
//   packed -> rows (decodePerGaussianKernel<RowsSink>): the mirror image; six bulk loads bring the tile's packed
//     planes in, thread g expands gaussian g into record g of the shared-memory tile, one bulk
//     async store writes the records.
//
// A gaussian's bytes in a packed plane start at byte B*g (B = 9, 3, 3 or 3D), i.e. at a lane-
// dependent offset (B*g) mod 4 inside a word.  loadRecord/emitRecord do the realignment with funnel
// shifts; on the way out the word a lane shares with its predecessor is completed from the
// predecessor's last four bytes, passed down with one warp shuffle (B*32 is a multiple of 4, so lane
// 0 of every warp starts a word).  All shared-memory traffic is therefore whole words.
#include "codec_kernels.cuh"

#define SPZ_COLD_NOINLINE  // one rotation per thread here: the general form as a called function measured 2 % faster
#include "codec_math.cuh"
#include "kernel_utils.cuh"
#include "record_align.cuh"
#include "scalar_path.cuh"

namespace spzb200 {
namespace {

template <int D, int G_>
struct Canon {
  static constexpr int W = 17 + 3 * D;  // floats per record
  static constexpr int G = G_;          // gaussians per tile = threads per CTA
  static constexpr int kColor = 6, kRest = 9, kAlpha = 9 + 3 * D, kScale = 10 + 3 * D, kRot = 13 + 3 * D;
  static constexpr int kRowBytes = G * W * 4;
  // the tile's packed planes in shared memory, in the container's order (load-spz.cc:533-546);
  // every offset and size is a multiple of 16 bytes, as the bulk copies require
  static constexpr int oPos = 0, oAlpha = 9 * G, oColor = 10 * G, oScale = 13 * G, oRot = 16 * G, oSh = 20 * G;
  static constexpr int kPackedBytes = (20 + 3 * D) * G + 16;  // + one granule: loadRecord<3> of the last lane reads a word ahead
  static constexpr int kSmemBytes = kRowBytes + kPackedBytes;
};

// Gaussians per tile (= threads per CTA) and the __launch_bounds__ CTA count, per direction and SH
// dimension, from device-timed sweeps at 40M points (profiles/r1_tuning_notes.txt).  The decoder
// wants small tiles -- each CTA holds its shared memory through load wait, expansion and store
// drain, and more, smaller CTAs per SM overlap those phases better (64 instead of 128 gaussians:
// 6166 -> 6675 GB/s at degree 3) -- while the SH-less encoder is issue-bound and wants every thread
// slot of the SM filled (256 x 8).
template <int D>
struct EncGeo {
  static constexpr int G = D == 0 ? 256 : 128;
  static constexpr int CTAS = D <= 3 ? 8 : 6;  // degree 2/3: 80 registers instead of 115 under a bound of 4: planar source +6 %, rows source unchanged
  using C = Canon<D, G>;
};
template <int D>
struct DecGeo {
  static constexpr int G = D <= 3 ? 128 : 64;
  static constexpr int CTAS = 8;
  using C = Canon<D, G>;
};
constexpr int encTileGaussians(int shDim) { return shDim == 0 ? EncGeo<0>::G : shDim == 3 ? EncGeo<3>::G : shDim == 8 ? EncGeo<8>::G : EncGeo<15>::G; }
constexpr int decTileGaussians(int shDim) { return shDim == 0 ? DecGeo<0>::G : shDim == 3 ? DecGeo<3>::G : shDim == 8 ? DecGeo<8>::G : DecGeo<15>::G; }

// Must be called by all 32 lanes of a warp, g consecutive per lane, lane 0's g a multiple of 32
// (record_align.cuh has the word arithmetic).
template <int B>
__device__ __forceinline__ void emitRecord(uint32_t *plane, int g, const uint32_t (&v)[Rec<B>::NV]) {
  uint32_t prev = 0;
  if constexpr (B % 4 != 0) prev = __shfl_up_sync(0xffffffffu, recordTail<B>(v), 1);
  emitRecordWords<B>(plane, g, v, prev);
}

__device__ __forceinline__ uint32_t byteOf(const uint32_t *v, int k) { return (v[k >> 2] >> (8 * (k & 3))) & 0xffu; }
// record byte k as the float 2^23 + byte (one PRMT; k is a constant after unrolling)
__device__ __forceinline__ float magicByte(const uint32_t *v, int k) {
  return __uint_as_float(prmt(v[k >> 2], 0x4b000000u, 0x7650u + (uint32_t)(k & 3)));
}

// =================================================================================================
// encode.  Two sources of the same per-gaussian record r[0..W) (canonical column order):
//   RowsSource   -- .ply vertex records: packGaussians(loadSplatFromPly(rows, to = X), from = X),
//                   load-spz.cc:808-838 + :257-331
//   PlanarSource -- the six GaussianCloud planes (splat-types.h:90-115): packGaussians itself
// Everything after the record is in registers is shared.
// =================================================================================================
template <int D>
struct RowsSource {
  using Args = PlyEncodeArgs;
  static constexpr bool kFoldRest = false;  // the remainder of a rows call goes to the column-map kernels
  static __device__ __forceinline__ void rest(const Args &, long long) {}
  using C = typename EncGeo<D>::C;
  static constexpr int kBytes = C::kRowBytes;
  static __device__ __forceinline__ void request(const Args &a, long long tile, unsigned char *buf, unsigned long long *bar) {
    bulkLoad(buf, a.rows + tile * (long long)(C::G * C::W), kBytes, bar);
  }
  static __device__ __forceinline__ void read(const unsigned char *buf, int t, float (&r)[C::W]) {
    constexpr int W = C::W;
    const float *row = reinterpret_cast<const float *>(buf) + t * W;
    if constexpr (W % 2 == 0) {
#pragma unroll
      for (int j = 0; j < W / 2; j++) {
        const float2 p = reinterpret_cast<const float2 *>(row)[j];
        r[2 * j] = p.x;
        r[2 * j + 1] = p.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < W; c++) r[c] = row[c];
    }
  }
};

template <int D>
struct PlanarSource {
  using Args = EncodeArgs;
  static constexpr bool kFoldRest = true;  // the sub-tile remainder rides in CTA 0 (scalar_path.cuh)
  static __device__ __forceinline__ void rest(const Args &a, long long g) { encodeOneGaussian(a, g); }
  using C = typename EncGeo<D>::C;
  static constexpr int G = C::G;
  // float planes of one tile in shared memory; sizes and offsets are multiples of 16 bytes
  static constexpr int oPos = 0, oScale = 12 * G, oRot = 24 * G, oAlpha = 40 * G, oColor = 44 * G, oSh = 56 * G;
  static constexpr int kBytes = (56 + 12 * D) * G;
  static __device__ __forceinline__ void request(const Args &a, long long tile, unsigned char *buf, unsigned long long *bar) {
    const long long g0 = tile * G;
    mbarExpect(bar, kBytes);
    bulkLoadOn(buf + oPos, a.positions + g0 * 3, 12 * G, bar);
    bulkLoadOn(buf + oScale, a.scales + g0 * 3, 12 * G, bar);
    bulkLoadOn(buf + oRot, a.rotations + g0 * 4, 16 * G, bar);
    bulkLoadOn(buf + oAlpha, a.alphas + g0, 4 * G, bar);
    bulkLoadOn(buf + oColor, a.colors + g0 * 3, 12 * G, bar);
    if constexpr (D > 0) bulkLoadOn(buf + oSh, a.sh + g0 * (3 * D), 12 * D * G, bar);
  }
  // lane strides of 3 and 3*D words are odd (D = 0, 3, 15), so the 32-bit loads are bank-conflict free; D = 8 below
  static __device__ __forceinline__ void read(const unsigned char *buf, int t, float (&r)[C::W]) {
    const float *f = reinterpret_cast<const float *>(buf);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      r[i] = f[oPos / 4 + 3 * t + i];
      r[C::kScale + i] = f[oScale / 4 + 3 * t + i];
      r[C::kColor + i] = f[oColor / 4 + 3 * t + i];
    }
    r[3] = r[4] = r[5] = 0.0f;
    const float4 q = reinterpret_cast<const float4 *>(buf + oRot)[t];  // x, y, z, w
    r[C::kRot] = q.w;
    r[C::kRot + 1] = q.x;
    r[C::kRot + 2] = q.y;
    r[C::kRot + 3] = q.z;
    r[C::kAlpha] = f[oAlpha / 4 + t];
    if constexpr (D == 8) {
      // 24-word lane stride: 32-bit loads would be 8-way bank conflicts; six 128-bit loads per lane are 2-way
      // (quarter-warp phases, stride 6 quad-words), which the shared-memory bandwidth has room for
      const float4 *s4 = reinterpret_cast<const float4 *>(buf + oSh) + 6 * t;
#pragma unroll
      for (int j = 0; j < 6; j++) {
        const float4 v = s4[j];
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int k = 4 * j + q;
          r[C::kRest + (k % 3) * D + k / 3] = e[q];
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3 * D; k++) r[C::kRest + (k % 3) * D + k / 3] = f[oSh / 4 + 3 * D * t + k];
    }
  }
};

template <int D, int MODE, class Src>
__global__ void __launch_bounds__(EncGeo<D>::G, EncGeo<D>::CTAS)
encodePerGaussianKernel(const __grid_constant__ typename Src::Args a, const long long numTiles, const int restCtas) {
  using C = typename EncGeo<D>::C;
  constexpr int W = C::W, kG = C::G;
  extern __shared__ __align__(128) unsigned char dynSmem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ float sThr[256];
  unsigned char *stage = dynSmem + Src::kBytes;
  const int t = threadIdx.x;
  pdlTrigger();
  if constexpr (Src::kFoldRest) {
    if ((int)blockIdx.x < restCtas) {  // fewer than kG gaussians past the last whole tile: one scalar CTA, started first
      const long long g = numTiles * kG + t;
      pdlWait();
      if (g < a.n) Src::rest(a, g);
      return;
    }
  }
  if (t == 0) mbarInit(&bar);
  __syncthreads();
  pdlWait();  // the planes may only be touched from here on
  uint32_t parity = 0;
  const long long firstTile = (int)blockIdx.x - restCtas;
  for (long long tile = firstTile; tile < numTiles; tile += (int)gridDim.x - restCtas, parity ^= 1u) {
    if (tile != firstTile) {  // multi-tile CTAs only: the previous tile's stores are done reading the stage
      if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncthreads();
    }
    if (t == 0) Src::request(a, tile, dynSmem, &bar);
    if (tile == firstTile) {  // the threshold table arrives while the tile is in flight (these CTAs live for one tile)
      for (int i = t; i < 256; i += kG) sThr[i] = __ldg(a.alphaThresholds + i);
      __syncthreads();
    }
    mbarWait(&bar, parity);

    float r[W];
    Src::read(dynSmem, t, r);
    // positions: 3 x 24 bit
    uint32_t pv[3];
    {
      const uint32_t n0 = m::quant_position24(r[0], signedConst(4096.0f, a.flipP & 1u));
      const uint32_t n1 = m::quant_position24(r[1], signedConst(4096.0f, (a.flipP >> 1) & 1u));
      const uint32_t n2 = m::quant_position24(r[2], signedConst(4096.0f, (a.flipP >> 2) & 1u));
      pv[0] = prmt(n0, n1, 0x4210u);
      pv[1] = prmt(n1, n2, 0x5421u);
      pv[2] = n2 >> 16;
    }
    const uint32_t cv[1] = {packSat4<MODE>(m::quant_color_raw(r[C::kColor]), m::quant_color_raw(r[C::kColor + 1]),
                                           m::quant_color_raw(r[C::kColor + 2]), 0)};
    const uint32_t sv[1] = {packSat4<MODE>(m::quant_scale_raw(r[C::kScale]), m::quant_scale_raw(r[C::kScale + 1]),
                                           m::quant_scale_raw(r[C::kScale + 2]), 0)};
    const uint32_t av = m::quant_alpha(r[C::kAlpha], sThr);
    // canonical order has w first (load-spz.cc:826-829)
    const uint32_t rv = m::quant_rotation_smallest3(r[C::kRot + 1], r[C::kRot + 2], r[C::kRot + 3], r[C::kRot], a.flipQ);
    uint32_t hv[D > 0 ? Rec<3 * D>::NV : 1];
    if constexpr (D > 0) {
#pragma unroll
      for (int k4 = 0; k4 < Rec<3 * D>::NV; k4++) {
        int32_t q[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int k = 4 * k4 + e;  // packed order: coefficient-major, channel innermost
          if (k < 3 * D) {
            const int coef = k / 3, ch = k - 3 * coef;
            q[e] = m::quant_sh_raw(r[C::kRest + ch * D + coef], signedConst(128.0f, (a.flipSh >> coef) & 1u), k < 9 ? 132u : 136u,
                                   k < 9 ? ~7u : ~15u);
          } else {
            q[e] = 0;
          }
        }
        hv[k4] = packSat4<MODE>(q[0], q[1], q[2], q[3]);
      }
    }

    emitRecord<9>(reinterpret_cast<uint32_t *>(stage + C::oPos), t, pv);
    stage[C::oAlpha + t] = (unsigned char)av;
    emitRecord<3>(reinterpret_cast<uint32_t *>(stage + C::oColor), t, cv);
    emitRecord<3>(reinterpret_cast<uint32_t *>(stage + C::oScale), t, sv);
    reinterpret_cast<uint32_t *>(stage + C::oRot)[t] = rv;
    if constexpr (D > 0) emitRecord<3 * D>(reinterpret_cast<uint32_t *>(stage + C::oSh), t, hv);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();  // also: everyone has read its record, the next tile's copy may land
    if (t == 0) {
      const long long g0 = tile * kG;
      bulkStore(a.oPositions + g0 * 9, stage + C::oPos, 9 * kG);
      bulkStore(a.oAlphas + g0, stage + C::oAlpha, kG);
      bulkStore(a.oColors + g0 * 3, stage + C::oColor, 3 * kG);
      bulkStore(a.oScales + g0 * 3, stage + C::oScale, 3 * kG);
      bulkStore(a.oRotations + g0 * 4, stage + C::oRot, 4 * kG);
      if constexpr (D > 0) bulkStore(a.oSh + g0 * (3 * D), stage + C::oSh, 3 * D * kG);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory must outlive the stores' reads
}

// =================================================================================================
// packed -> rows:  saveSplatToPly(unpackGaussians(in, to = X), from = X)'s vertex records, load-spz.cc:467-531 + :858-890
// =================================================================================================
// Two sinks for the same per-gaussian record rec[0..W) (canonical column order):
//   RowsSink   -- .ply vertex records, one bulk store per tile
//   PlanarSink -- the six GaussianCloud planes (unpackGaussians itself), six bulk stores per tile: the default
//                 decoder at SH degree 1, 2 and 3 (launchDecodePerGaussianPlanar)
template <int D>
struct RowsSink {
  using Args = PlyDecodeArgs;
  static constexpr bool kFoldRest = false;
  static __device__ __forceinline__ void rest(const Args &, long long) {}
  using C = typename DecGeo<D>::C;
  static constexpr int kBytes = C::kRowBytes;
  static __device__ __forceinline__ void write(unsigned char *buf, int t, const float (&rec)[C::W]) {
    constexpr int W = C::W;
    float *row = reinterpret_cast<float *>(buf) + t * W;
    if constexpr (W % 2 == 0) {
#pragma unroll
      for (int j = 0; j < W / 2; j++) reinterpret_cast<float2 *>(row)[j] = make_float2(rec[2 * j], rec[2 * j + 1]);
    } else {
#pragma unroll
      for (int c = 0; c < W; c++) row[c] = rec[c];
    }
  }
  static __device__ __forceinline__ void store(const Args &a, long long tile, const unsigned char *buf) {
    bulkStore(a.rows + tile * (long long)(C::G * C::W), buf, kBytes);
  }
};

template <int D>
struct PlanarSink {
  using Args = DecodeArgs;
  static constexpr bool kFoldRest = true;
  static __device__ __forceinline__ void rest(const Args &a, long long g) { decodeOneGaussian(a, g); }
  using C = typename DecGeo<D>::C;
  static constexpr int G = C::G;
  static constexpr int oPos = 0, oScale = 12 * G, oRot = 24 * G, oAlpha = 40 * G, oColor = 44 * G, oSh = 56 * G;
  static constexpr int kBytes = (56 + 12 * D) * G;
  static __device__ __forceinline__ void write(unsigned char *buf, int t, const float (&rec)[C::W]) {
    float *f = reinterpret_cast<float *>(buf);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      f[oPos / 4 + 3 * t + i] = rec[i];
      f[oScale / 4 + 3 * t + i] = rec[C::kScale + i];
      f[oColor / 4 + 3 * t + i] = rec[C::kColor + i];
    }
    reinterpret_cast<float4 *>(buf + oRot)[t] = make_float4(rec[C::kRot + 1], rec[C::kRot + 2], rec[C::kRot + 3], rec[C::kRot]);  // x, y, z, w
    f[oAlpha / 4 + t] = rec[C::kAlpha];
    if constexpr ((3 * D) % 4 == 0) {
      // degree 2: a lane stride of 24 words would be an 8-way bank conflict for 32-bit stores; 128-bit ones are 2-way
#pragma unroll
      for (int k = 0; k < 3 * D; k += 4) {
        reinterpret_cast<float4 *>(buf + oSh)[(3 * D * t + k) / 4] =
            make_float4(rec[C::kRest + (k % 3) * D + k / 3], rec[C::kRest + ((k + 1) % 3) * D + (k + 1) / 3],
                        rec[C::kRest + ((k + 2) % 3) * D + (k + 2) / 3], rec[C::kRest + ((k + 3) % 3) * D + (k + 3) / 3]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3 * D; k++) f[oSh / 4 + 3 * D * t + k] = rec[C::kRest + (k % 3) * D + k / 3];
    }
  }
  static __device__ __forceinline__ void store(const Args &a, long long tile, const unsigned char *buf) {
    const long long g0 = tile * G;
    bulkStore(a.oPositions + g0 * 3, buf + oPos, 12 * G);
    bulkStore(a.oScales + g0 * 3, buf + oScale, 12 * G);
    bulkStore(a.oRotations + g0 * 4, buf + oRot, 16 * G);
    bulkStore(a.oAlphas + g0, buf + oAlpha, 4 * G);
    bulkStore(a.oColors + g0 * 3, buf + oColor, 12 * G);
    if constexpr (D > 0) bulkStore(a.oSh + g0 * (3 * D), buf + oSh, 12 * D * G);
  }
};

template <int D, class Sink>
__global__ void __launch_bounds__(DecGeo<D>::G, DecGeo<D>::CTAS)
decodePerGaussianKernel(const __grid_constant__ typename Sink::Args a, const long long numTiles, const int restCtas) {
  using C = typename DecGeo<D>::C;
  constexpr int W = C::W, kG = C::G;
  extern __shared__ __align__(128) unsigned char dynSmem[];
  __shared__ __align__(8) unsigned long long bar;
  unsigned char *in = dynSmem + Sink::kBytes;
  const int t = threadIdx.x;
  pdlTrigger();
  if constexpr (Sink::kFoldRest) {
    if ((int)blockIdx.x < restCtas) {  // fewer than kG gaussians past the last whole tile: one scalar CTA, started first
      const long long g = numTiles * kG + t;
      pdlWait();
      if (g < a.n) Sink::rest(a, g);
      return;
    }
  }
  const bool half = a.version == 1 || a.version == 4;
  const bool s3 = a.version >= 3;
  const uint32_t posBytes = half ? 6 * kG : 9 * kG, rotBytes = s3 ? 4 * kG : 3 * kG;
  // 4 KB of tables read through L1 (the bulk copies do not pass through it); copying them to shared
  // memory per CTA measured 4 % slower
  const float *tab = a.tables;
  if (t == 0) mbarInit(&bar);
  __syncthreads();
  pdlWait();  // the planes may only be touched from here on
  uint32_t parity = 0;
  const long long firstTile = (int)blockIdx.x - restCtas;
  for (long long tile = firstTile; tile < numTiles; tile += (int)gridDim.x - restCtas, parity ^= 1u) {
    if (tile != firstTile) {  // multi-tile CTAs only: the previous tile's store is done reading the records
      if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncthreads();
    }
    if (t == 0) {
      const long long g0 = tile * kG;
      mbarExpect(&bar, posBytes + rotBytes + (uint32_t)((7 + 3 * D) * kG));
      bulkLoadOn(in + C::oPos, a.positions + g0 * (half ? 6 : 9), posBytes, &bar);
      bulkLoadOn(in + C::oAlpha, a.alphas + g0, kG, &bar);
      bulkLoadOn(in + C::oColor, a.colors + g0 * 3, 3 * kG, &bar);
      bulkLoadOn(in + C::oScale, a.scales + g0 * 3, 3 * kG, &bar);
      bulkLoadOn(in + C::oRot, a.rotations + g0 * (s3 ? 4 : 3), rotBytes, &bar);
      if constexpr (D > 0) bulkLoadOn(in + C::oSh, a.sh + g0 * (3 * D), 3 * D * kG, &bar);
    }
    mbarWait(&bar, parity);

    // every shared-memory read of the inputs comes before the first store to the records
    uint32_t pv[3], cv[2], sv[2], rv[2], hv[D > 0 ? Rec<3 * D>::NL : 1];
    if (half) {
      uint32_t h[2];
      loadRecord<6>(reinterpret_cast<const uint32_t *>(in + C::oPos), t, h);
      pv[0] = h[0]; pv[1] = h[1]; pv[2] = 0;
    } else {
      loadRecord<9>(reinterpret_cast<const uint32_t *>(in + C::oPos), t, pv);
    }
    loadRecord<3>(reinterpret_cast<const uint32_t *>(in + C::oColor), t, cv);
    loadRecord<3>(reinterpret_cast<const uint32_t *>(in + C::oScale), t, sv);
    const uint32_t av = in[C::oAlpha + t];
    if (s3) {
      rv[0] = reinterpret_cast<const uint32_t *>(in + C::oRot)[t];
    } else {
      loadRecord<3>(reinterpret_cast<const uint32_t *>(in + C::oRot), t, rv);
    }
    if constexpr (D > 0) loadRecord<3 * D>(reinterpret_cast<const uint32_t *>(in + C::oSh), t, hv);

    float rec[W];
    if (half) {
      rec[0] = __uint_as_float(__float_as_uint(m::half_bits_to_float(pv[0] & 0xffffu)) ^ ((a.flipP & 1u) << 31));
      rec[1] = __uint_as_float(__float_as_uint(m::half_bits_to_float(pv[0] >> 16)) ^ ((a.flipP & 2u) << 30));
      rec[2] = __uint_as_float(__float_as_uint(m::half_bits_to_float(pv[1] & 0xffffu)) ^ ((a.flipP & 4u) << 29));
    } else {
      // PRMT with the sign-replicate bit (selector nibble 8|idx) sign-extends 24 -> 32 bits
      const int32_t f0 = (int32_t)prmt(pv[0], pv[0], 0xA210u);
      const int32_t f1 = (int32_t)prmt(pv[0], pv[1], 0xD543u);
      const int32_t f2 = (int32_t)prmt(pv[1], pv[2], 0xC432u);
      const uint32_t ps = __float_as_uint(a.positionScale);
      rec[0] = m::mul(m::i2f(f0), __uint_as_float(ps ^ ((a.flipP & 1u) << 31)));
      rec[1] = m::mul(m::i2f(f1), __uint_as_float(ps ^ ((a.flipP & 2u) << 30)));
      rec[2] = m::mul(m::i2f(f2), __uint_as_float(ps ^ ((a.flipP & 4u) << 29)));
    }
    rec[3] = rec[4] = rec[5] = 0.0f;  // normals (load-spz.cc:866)
#pragma unroll
    for (int i = 0; i < 3; i++) {
      rec[C::kColor + i] = tab[256 + byteOf(cv, i)];
      // (2^23 + s) / 16 - (2^19 + 10) = s/16 - 10, both steps exact (load-spz.cc:506)
      rec[C::kScale + i] = __fmaf_rn(magicByte(sv, i), 0.0625f, -524298.0f);
    }
    rec[C::kAlpha] = tab[av];
    {
      float q[4];
      if (s3) m::dequant_rotation_smallest3(rv[0], tab + 512, a.flipQ, q);
      else m::dequant_rotation_first3(byteOf(rv, 0), byteOf(rv, 1), byteOf(rv, 2), a.flipQ, q);
      rec[C::kRot] = q[3];
      rec[C::kRot + 1] = q[0];
      rec[C::kRot + 2] = q[1];
      rec[C::kRot + 3] = q[2];
    }
    if constexpr (D > 0) {
#pragma unroll
      for (int k = 0; k < 3 * D; k++) {
        const int coef = k / 3, ch = k - 3 * coef;
        // (2^23 + x) - (2^23 + 128) = x - 128 exactly, then * +-1/128 (load-spz.cc:83, splat-types.h:158-161)
        rec[C::kRest + ch * D + coef] =
            m::mul(m::add(magicByte(hv, k), -8388736.0f), signedConst(0.0078125f, (a.flipSh >> coef) & 1u));
      }
    }
    Sink::write(dynSmem, t, rec);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();  // (one bulk store per warp instead, issued after __syncwarp, measured 2 % slower)
    if (t == 0) {
      Sink::store(a, tile, dynSmem);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <class Args>
bool canonicalColumns(const Args &a) {
  const int D = a.shDim;
  if (a.width != 17 + 3 * D) return false;
  bool ok = a.colAlpha == 9 + 3 * D && a.colRot[3] == 13 + 3 * D;  // colRot = x, y, z, w; the file has w first
  for (int i = 0; i < 3; i++) ok = ok && a.colPos[i] == i && a.colColor[i] == 6 + i && a.colScale[i] == 10 + 3 * D + i && a.colRot[i] == 14 + 3 * D + i;
  for (int i = 0; i < 3 * D; i++) ok = ok && a.colRest[i] == 9 + i;
  return ok;
}

unsigned gridFor(long long tiles, const LaunchPlan &plan, int ctasPerSm) {
  if (!plan.flatGrid) {
    const long long g = (long long)plan.smCount * ctasPerSm;
    return (unsigned)(tiles < g ? tiles : g);
  }
  return (unsigned)(tiles < 0x7fffffffLL ? tiles : 0x7fffffffLL);
}

template <int D, int MODE, class Src>
cudaError_t launchEncodePerGaussian(const typename Src::Args &a, long long tiles, int restCtas, const LaunchPlan &plan, cudaStream_t s) {
  constexpr int smem = Src::kBytes + EncGeo<D>::C::kPackedBytes;
  static_assert(smem <= 48 * 1024, "above 48 KB the kernel would need cudaFuncAttributeMaxDynamicSharedMemorySize on every device");
  return launchKernel(encodePerGaussianKernel<D, MODE, Src>, gridFor(tiles, plan, EncGeo<D>::CTAS) + restCtas, EncGeo<D>::G, smem, s, plan.pdl, a, tiles,
                      restCtas);
}

template <template <int> class Src, class Args>
cudaError_t dispatchEncodePerGaussian(const Args &a, long long tiles, int restCtas, const LaunchPlan &plan, cudaStream_t s) {
  const bool cvt = plan.packMode == kPackCvt;
  switch (a.shDim) {
    case 0: return cvt ? launchEncodePerGaussian<0, kPackCvt, Src<0>>(a, tiles, restCtas, plan, s) : launchEncodePerGaussian<0, kPackAlu, Src<0>>(a, tiles, restCtas, plan, s);
    case 3: return cvt ? launchEncodePerGaussian<3, kPackCvt, Src<3>>(a, tiles, restCtas, plan, s) : launchEncodePerGaussian<3, kPackAlu, Src<3>>(a, tiles, restCtas, plan, s);
    case 8: return cvt ? launchEncodePerGaussian<8, kPackCvt, Src<8>>(a, tiles, restCtas, plan, s) : launchEncodePerGaussian<8, kPackAlu, Src<8>>(a, tiles, restCtas, plan, s);
    case 15: return cvt ? launchEncodePerGaussian<15, kPackCvt, Src<15>>(a, tiles, restCtas, plan, s) : launchEncodePerGaussian<15, kPackAlu, Src<15>>(a, tiles, restCtas, plan, s);
    default: return cudaErrorInvalidValue;
  }
}

template <int D, class Sink>
cudaError_t launchDecodePerGaussian(const typename Sink::Args &a, long long tiles, int restCtas, const LaunchPlan &plan, cudaStream_t s) {
  constexpr int smem = Sink::kBytes + DecGeo<D>::C::kPackedBytes;
  static_assert(smem <= 48 * 1024, "above 48 KB the kernel would need cudaFuncAttributeMaxDynamicSharedMemorySize on every device");
  return launchKernel(decodePerGaussianKernel<D, Sink>, gridFor(tiles, plan, DecGeo<D>::CTAS) + restCtas, DecGeo<D>::G, smem, s, plan.pdl, a, tiles,
                      restCtas);
}

template <template <int> class Sink, class Args>
cudaError_t dispatchDecodePerGaussian(const Args &a, long long tiles, int restCtas, const LaunchPlan &plan, cudaStream_t s) {
  switch (a.shDim) {
    case 0: return launchDecodePerGaussian<0, Sink<0>>(a, tiles, restCtas, plan, s);
    case 3: return launchDecodePerGaussian<3, Sink<3>>(a, tiles, restCtas, plan, s);
    case 8: return launchDecodePerGaussian<8, Sink<8>>(a, tiles, restCtas, plan, s);
    case 15: return launchDecodePerGaussian<15, Sink<15>>(a, tiles, restCtas, plan, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

// Both return the number of leading gaussians handled (a multiple of 128; 0 when the layout is not
// canonical, a pointer is not 16-byte aligned or the cloud is smaller than a tile); the caller
// finishes the rest with the column-map / scalar kernels of ply_kernels.cu.
cudaError_t launchEncodePlyCanonical(const PlyEncodeArgs &a, const LaunchPlan &plan, cudaStream_t stream, long long *done) {
  *done = 0;
  if (plan.forceGeneric || plan.plyMapped || !canonicalColumns(a)) return cudaSuccess;
  if (!(aligned16(a.rows) && aligned16(a.oPositions) && aligned16(a.oScales) && aligned16(a.oRotations) && aligned16(a.oAlphas) &&
        aligned16(a.oColors) && (a.shDim == 0 || aligned16(a.oSh))))
    return cudaSuccess;
  const int G = encTileGaussians(a.shDim);
  const long long tiles = a.n / G;
  if (tiles == 0) return cudaSuccess;
  const cudaError_t e = dispatchEncodePerGaussian<RowsSource>(a, tiles, 0, plan, stream);
  if (e == cudaSuccess) *done = tiles * G;
  return e;
}

// packGaussians on planar float planes through the same kernel; *done as above.  By default only where it measured
// faster than the register-path tile encoder: SH degree 3 clouds (or shards, or pipeline ranges) of at most 24M
// gaussians -- 10M points 7045 vs 6414 GB/s, 2.5M 5616 vs 5215, 20M 6902 vs 6793; from 40M up the tile encoder
// leads by about 1 % (6761 vs 6671 at 100M), and at degree 1 by more.
constexpr long long kEncodePerGaussianMaxPoints = 24000000;
// Round 2, SH degree 1 (profiles/r2_ab_encode_pergaussian_vs_tiles.jsonl): the tile encoder's 6400-gaussian tiles quantise small launches
// badly (1.25M points = 195 CTAs on 148 SMs).  Per-gaussian vs tiles, same box, GB/s: 1.25M 5062 vs 2734, 2.5M 5583 vs 4780, 5M 5862 vs 5327,
// 10M 6056 vs 6267, 20M 6141 vs 6424 -> per-gaussian up to 6M points.
constexpr long long kEncodePerGaussianMaxPointsSh1 = 6000000;
// SH degree 2 (planar source with 128-bit shared-memory reads of the 24-word SH records; profiles/r2_ab_encode_pergaussian_sh2.jsonl), per-gaussian vs
// 128-thread tiles, GB/s: 600K 5204 vs 5410, 1.25M 6051 vs 5759, 2.5M 6558 vs 6237, 5M 6857 vs 6689, 10M 7005 vs 6865, 20M 7094 vs 6981, 40M 6950 vs 7006
// -> per-gaussian from 1M to 24M points.
constexpr long long kEncodePerGaussianMinPointsSh2 = 1000000, kEncodePerGaussianMaxPointsSh2 = 24000000;
cudaError_t launchEncodePerGaussianPlanar(const EncodeArgs &a, const LaunchPlan &plan, cudaStream_t stream, long long *done) {
  *done = 0;
  if (plan.forceGeneric || plan.encodeBulk == 0 || a.version != 3) return cudaSuccess;  // version-2 streams: tile encoder only
  if (plan.encodeBulk < 2 && !((a.shDim == 15 && a.n <= kEncodePerGaussianMaxPoints) || (a.shDim == 3 && a.n <= kEncodePerGaussianMaxPointsSh1) ||
                               (a.shDim == 8 && a.n >= kEncodePerGaussianMinPointsSh2 && a.n <= kEncodePerGaussianMaxPointsSh2)))
    return cudaSuccess;
  if (!(aligned16(a.positions) && aligned16(a.scales) && aligned16(a.rotations) && aligned16(a.alphas) && aligned16(a.colors) &&
        (a.shDim == 0 || aligned16(a.sh)) && aligned16(a.oPositions) && aligned16(a.oScales) && aligned16(a.oRotations) &&
        aligned16(a.oAlphas) && aligned16(a.oColors) && (a.shDim == 0 || aligned16(a.oSh))))
    return cudaSuccess;
  const int G = encTileGaussians(a.shDim);
  const long long tiles = a.n / G;
  if (tiles == 0) return cudaSuccess;
  const int restCtas = plan.flatGrid && plan.foldRest && a.n > tiles * G ? 1 : 0;
  const cudaError_t e = dispatchEncodePerGaussian<PlanarSource>(a, tiles, restCtas, plan, stream);
  if (e == cudaSuccess) *done = restCtas ? a.n : tiles * G;
  return e;
}

cudaError_t launchDecodePlyCanonical(const PlyDecodeArgs &a, const LaunchPlan &plan, cudaStream_t stream, long long *done) {
  *done = 0;
  if (plan.forceGeneric || plan.plyMapped || !canonicalColumns(a)) return cudaSuccess;
  if (!(aligned16(a.rows) && aligned16(a.positions) && aligned16(a.scales) && aligned16(a.rotations) && aligned16(a.alphas) &&
        aligned16(a.colors) && (a.shDim == 0 || aligned16(a.sh))))
    return cudaSuccess;
  const int G = decTileGaussians(a.shDim);
  const long long tiles = a.n / G;
  if (tiles == 0) return cudaSuccess;
  const cudaError_t e = dispatchDecodePerGaussian<RowsSink>(a, tiles, 0, plan, stream);
  if (e == cudaSuccess) *done = tiles * G;
  return e;
}

// unpackGaussians onto planar float planes through the same kernel; *done as above
cudaError_t launchDecodePerGaussianPlanar(const DecodeArgs &a, const LaunchPlan &plan, cudaStream_t stream, long long *done) {
  *done = 0;
  // SH-less clouds are faster through the register-path tiles (6.60 vs 5.90 TB/s), so they come here only when asked to (tests)
  if (plan.forceGeneric || plan.decodePerGaussian == 0 || (a.shDim == 0 && plan.decodePerGaussian < 2)) return cudaSuccess;
  if (!(aligned16(a.positions) && aligned16(a.scales) && aligned16(a.rotations) && aligned16(a.alphas) && aligned16(a.colors) &&
        (a.shDim == 0 || aligned16(a.sh)) && aligned16(a.oPositions) && aligned16(a.oScales) && aligned16(a.oRotations) &&
        aligned16(a.oAlphas) && aligned16(a.oColors) && (a.shDim == 0 || aligned16(a.oSh))))
    return cudaSuccess;
  const int G = decTileGaussians(a.shDim);
  const long long tiles = a.n / G;
  if (tiles == 0) return cudaSuccess;
  const int restCtas = plan.flatGrid && plan.foldRest && a.n > tiles * G ? 1 : 0;
  const cudaError_t e = dispatchDecodePerGaussian<PlanarSink>(a, tiles, restCtas, plan, stream);
  if (e == cudaSuccess) *done = restCtas ? a.n : tiles * G;
  return e;
}

}  // namespace spzb200
