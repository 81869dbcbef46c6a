// RTL-ordered stages beyond the reference's Python path: MultiplyValue over top-k survivors ("local V pruning",
// SpAttenController.scala:554,588-591,668-726,749) and head selection (README.md:21).  Parity unpinned: checked
// against oracle/spatten_oracle.py::local_value_attention / head_prune_mask.
#include <math_constants.h>

#include "../../include/spatten_b200.h"
#include "common.cuh"

namespace spatten {

constexpr int kPvThreads = 1024;
constexpr int kPvWarps = kPvThreads / 32;
constexpr float kLog2eS = 1.4426950408889634f;

struct SparsePvParams {
  const __half* logits;
  int64_t logits_stride;
  const int32_t* keep;
  int64_t keep_stride;
  const __half* v;
  int64_t stride_b, stride_h;
  __half* out;
  int H, n_rep, keep_n;
};

// One CTA per (batch, head) row.  Pass 1: softmax statistics over the kept logits.  Pass 2: every warp gathers a
// contiguous share of the kept V rows (16 lanes x 16 B per 256-byte row, 2 rows per load instruction, 8 loads in
// flight) and accumulates p * v in fp32; warps are merged through shared memory.
template <int D>
__global__ void __launch_bounds__(kPvThreads) sparse_pv_kernel(const SparsePvParams p) {
  constexpr int E = D / 16;                  // halves per lane per row (16 lanes per row, 2 rows per warp load)
  constexpr int NH2 = E / 2;
  using Vec = typename VecT<E * 2>::type;
  __shared__ float s_red[kPvWarps];
  __shared__ float s_acc[kPvWarps][D];
  const int row = blockIdx.x, b = row / p.H, h = row % p.H, kvh = h / p.n_rep;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 4, c = lane & 15;
  const __half* lg = p.logits + static_cast<int64_t>(row) * p.logits_stride;
  const int32_t* keep = p.keep + static_cast<int64_t>(row) * p.keep_stride;
  const __half* V = p.v + b * p.stride_b + kvh * p.stride_h;

  float mx = -CUDART_INF_F;
  for (int i = tid; i < p.keep_n; i += kPvThreads) mx = fmaxf(mx, __half2float(lg[keep[i]]));
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[lane];                          // kPvWarps == 32
  mx = warp_max(mx);
  __syncthreads();
  float sum = 0.f;
  for (int i = tid; i < p.keep_n; i += kPvThreads) sum += fast_exp2((__half2float(lg[keep[i]]) - mx) * kLog2eS);
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = warp_sum(s_red[lane]);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;

  float acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;
  const int per_warp = (p.keep_n + kPvWarps - 1) / kPvWarps;
  const int i0 = warp * per_warp, i1 = min(i0 + per_warp, p.keep_n);
  constexpr int kU = 8;                      // gathered rows in flight per half-warp
  for (int i = i0 + g; i < i1; i += 2 * kU) {
    Vec v[kU];
    float pr[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int ii = i + 2 * u;
      pr[u] = 0.f;
      v[u] = Vec{};
      if (ii < i1) {
        const int j = keep[ii];
        pr[u] = fast_exp2((__half2float(lg[j]) - mx) * kLog2eS) * inv;
        v[u] = __ldg(reinterpret_cast<const Vec*>(V + static_cast<int64_t>(j) * D + c * E));
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
      for (int e = 0; e < NH2; ++e) {
        const float2 f = __half22float2(h2[e]);
        acc[2 * e] = fmaf(pr[u], f.x, acc[2 * e]);
        acc[2 * e + 1] = fmaf(pr[u], f.y, acc[2 * e + 1]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
  }
  if (g == 0) {
#pragma unroll
    for (int e = 0; e < E; ++e) s_acc[warp][c * E + e] = acc[e];
  }
  __syncthreads();
  if (tid < D) {
    float o = 0.f;
#pragma unroll 8
    for (int w = 0; w < kPvWarps; ++w) o += s_acc[w][tid];
    p.out[static_cast<int64_t>(row) * D + tid] = __float2half_rn(o);
  }
}

__global__ void head_select_kernel(const float* imp, int H, int keep, uint8_t* mask) {
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const float v = imp[h];
    int rank = 0;
    for (int o = 0; o < H; ++o) {
      const float w = imp[o];
      rank += (w > v) || (w == v && o < h);
    }
    mask[h] = rank < keep ? 1 : 0;
  }
}

}  // namespace spatten

using namespace spatten;

extern "C" int spatten_sparse_pv(const SpattenSparsePvArgs* a, void* stream) {
  SPATTEN_CHECK_ARG(a != nullptr && a->struct_size == sizeof(SpattenSparsePvArgs), "spatten_sparse_pv: bad args/struct_size");
  SPATTEN_CHECK_ARG(a->batch > 0 && a->num_heads > 0 && a->num_kv_heads > 0 && a->num_heads % a->num_kv_heads == 0 &&
                        a->keep > 0, "spatten_sparse_pv: bad shape");
  SPATTEN_CHECK_ARG(a->logits && a->keep_idx && a->v_cache && a->out, "spatten_sparse_pv: NULL pointer");
  if (a->head_dim != 64 && a->head_dim != 128) {
    set_error("spatten_sparse_pv: head_dim %d unsupported (64 or 128)", a->head_dim);
    return SPATTEN_ERR_UNSUPPORTED;
  }
  SPATTEN_CHECK_ARG((reinterpret_cast<uintptr_t>(a->v_cache) & 15) == 0 && a->cache_stride_b % 8 == 0 &&
                        a->cache_stride_h % 8 == 0, "spatten_sparse_pv: v_cache must be 16-byte aligned");
  SparsePvParams p{static_cast<const __half*>(a->logits), a->logits_stride, a->keep_idx, a->keep_stride,
                   static_cast<const __half*>(a->v_cache), a->cache_stride_b, a->cache_stride_h,
                   static_cast<__half*>(a->out), a->num_heads, a->num_heads / a->num_kv_heads, a->keep};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rows = a->batch * a->num_heads;
  if (a->head_dim == 128) sparse_pv_kernel<128><<<rows, kPvThreads, 0, st>>>(p);
  else sparse_pv_kernel<64><<<rows, kPvThreads, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "sparse_pv_kernel launch");
}

extern "C" int spatten_head_select(const float* head_importance, int32_t num_heads, int32_t keep_heads, uint8_t* head_mask,
                                   void* stream) {
  SPATTEN_CHECK_ARG(head_importance && head_mask && num_heads > 0 && keep_heads >= 0 && keep_heads <= num_heads,
                    "spatten_head_select: bad arguments");
  head_select_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(head_importance, num_heads, keep_heads, head_mask);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "head_select_kernel launch");
}
