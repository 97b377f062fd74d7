// Small device helpers shared by the codec kernels: byte permutes and packers, and the bulk async
// copy (TMA) / mbarrier wrappers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "codec_kernels.cuh"
#include "codec_math.cuh"

namespace spzb200 {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// Saturate four int32 to [0,255] and pack them little-endian (v0 lowest byte).
template <int MODE>
__device__ __forceinline__ uint32_t packSat4(int32_t v0, int32_t v1, int32_t v2, int32_t v3) {
  if (MODE == kPackCvt) {
    // cvt.pack.sat.u8.s32.b32 d, a, b, c:  d = sat(a) << 8 | sat(b) | c << 16   (I2IP in SASS)
    uint32_t hi, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(v1), "r"(v0), "r"(hi));
    return r;
  }
  return (uint32_t)m::clamp_u8(v0) | ((uint32_t)m::clamp_u8(v1) << 8) |
         ((uint32_t)m::clamp_u8(v2) << 16) | ((uint32_t)m::clamp_u8(v3) << 24);
}

// byte k of w as the float 2^23 + byte (exact), built with one PRMT
template <int K>
__device__ __forceinline__ float byteAsMagicFloat(uint32_t w) {
  return __uint_as_float(prmt(w, 0x4b000000u, 0x7650u + K));
}

__device__ __forceinline__ float signedConst(float magnitude, uint32_t negate) {
  return __uint_as_float(__float_as_uint(magnitude) | (negate << 31));
}

// Programmatic dependent launch (PDL).  Every vector / scalar codec kernel is launched with the
// programmatic-stream-serialization attribute (launchKernel below): its CTAs may become resident while
// the previous kernel of the stream is still draining, run their prologue (table staging, barrier
// init, constants), and then block in pdlWait() until that kernel has completed and its writes are
// visible.  So: NO global read of a plane and NO global write before pdlWait().  pdlTrigger() lets the
// NEXT kernel's CTAs in as soon as all of this grid's CTAs have started.  Without the attribute (or
// behind a kernel that never triggers) both are no-ops and the stream serialises as usual.
__device__ __forceinline__ void pdlTrigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launchKernel(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, bool pdl,
                                Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(unsigned long long *bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// SPZ_BULK_HINT (development knob): bit 0 = bulk stores carry an L2 evict-first policy, bit 1 = bulk
// loads do (every byte is touched once).  scripts/ build and time the alternatives.
#ifndef SPZ_BULK_HINT
#define SPZ_BULK_HINT 0
#endif
__device__ __forceinline__ uint64_t policyEvictFirst() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// the copy alone, counted on a barrier whose expected byte count the caller has set (mbarExpect)
__device__ __forceinline__ void bulkLoadOn(void *dstSmem, const void *srcGlobal, uint32_t bytes, unsigned long long *bar) {
#if SPZ_BULK_HINT & 2
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smemAddr(dstSmem)),
               "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)), "l"(policyEvictFirst())
               : "memory");
#else
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)),
               "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
               : "memory");
#endif
}
__device__ __forceinline__ void mbarExpect(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkLoad(void *dstSmem, const void *srcGlobal, uint32_t bytes, unsigned long long *bar) {
  mbarExpect(bar, bytes);
  bulkLoadOn(dstSmem, srcGlobal, bytes, bar);
}
__device__ __forceinline__ void mbarWait(unsigned long long *bar, uint32_t parity) {
  uint32_t done = 0;
  // bounded so a lost completion surfaces as a launch failure instead of a hung GPU
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(smemAddr(bar)), "r"(parity)
                 : "memory");
    if (spin > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void bulkStore(void *dstGlobal, const void *srcSmem, uint32_t bytes) {
#if SPZ_BULK_HINT & 1
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dstGlobal), "r"(smemAddr(srcSmem)),
               "r"(bytes), "l"(policyEvictFirst())
               : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstGlobal), "r"(smemAddr(srcSmem)), "r"(bytes)
               : "memory");
#endif
}

}  // namespace spzb200
