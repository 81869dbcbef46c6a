// Shared device helpers: mbarrier / bulk-copy (TMA engine) PTX wrappers, warp reductions, error plumbing.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace spatten {

// ---- error plumbing (host) -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* where);
int sm_count();

#define SPATTEN_CHECK_ARG(cond, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::spatten::set_error(__VA_ARGS__);  \
      return SPATTEN_ERR_INVALID;         \
    }                                     \
  } while (0)

// ---- PTX wrappers (device) -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
  // make mbarrier.init visible to the async proxy (bulk copies complete_tx on it)
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// L2 eviction policy for data that is streamed once (the KV cache during one decode step).
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// 1-D bulk async copy global -> shared through the TMA engine (SASS: UBLKCP), completion on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Mixed-precision FMA of sm_100 (SASS FHFMA): c + a * b with fp16 a, b (selected halves of half2 registers, no
// conversion instructions) and an fp32 accumulator.  The fp16 x fp16 product is exact in fp32, so the result is
// bit-identical to fmaf(__half2float(a), __half2float(b), c).
__device__ __forceinline__ float fhfma_lo(__half2 a, __half2 b, float c) {
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(__half_as_ushort(__low2half(a))), "h"(__half_as_ushort(__low2half(b))));
  return c;
}
__device__ __forceinline__ float fhfma_hi(__half2 a, __half2 b, float c) {
  asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(__half_as_ushort(__high2half(a))), "h"(__half_as_ushort(__high2half(b))));
  return c;
}

template <int BYTES>
struct VecT;
template <>
struct VecT<16> {
  using type = uint4;
};
template <>
struct VecT<8> {
  using type = uint2;
};

constexpr float kLog2e = 1.4426950408889634f;

template <int N>
struct HalfVec {
  __half2 h[N];
};

// x / c with the result of an IEEE fp32 division (what torch's fp16 `tensor / sqrt(D)` computes in its fp32
// opmath, modify_llama.py:111-113) for finite x: one Markstein correction of x * RN(1/c).  Falls back to the
// uncorrected product when the correction is not finite (x = +-inf).
__device__ __forceinline__ float div_by_const(float x, float c, float rc) {
  const float y = x * rc;
  const float r = fmaf(-y, c, x);
  const float y2 = fmaf(r, rc, y);
  return (fabsf(y2) <= 3.4028234e38f) ? y2 : y;
}

}  // namespace spatten
