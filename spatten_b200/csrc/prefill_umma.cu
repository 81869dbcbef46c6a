// Prefill attention on the 5th-generation tensor cores -- spatten_prefill_attn(): the entry point and the THREE-LAUNCH path.
// The default path is the single fused kernel in prefill_fused.cu (taken whenever the caller's scratch is large enough for
// the operand images); the kernels below are the round-1 structure, kept as the fallback (SPATTEN_PREFILL_LEGACY, scratch_bytes
// == 0) and as the in-tree cross-check of the fused kernel (tests/test_prefill_umma_gpu.py).
//
// Replaces llama_pos_shift_attention_forward for q_len > 1 (reference spatten_llm/pos_shift/modify_llama.py:86-144)
// with the reference's own three-step structure, because the reference *exports* the full pre-mask logit matrix
// (`self.attn_scores`, :116-119) and masks additively (:127-132), so S = Q.K^T has to be materialised anyway:
//   1. umma_qk_kernel   S[q, kv] = rope(Q) . rope(K)^T          tcgen05.mma (UMMA) 128x128 tiles, fp32 in TMEM; RoPE is
//                                                               applied by the producer threads on the way into shared
//                                                               memory (fp16, op for op, :27); epilogue writes the fp16
//                                                               logit exactly as the reference forms it (:111-113)
//   2. softmax_rows_kernel  P = softmax(S + mask) in fp32 -> fp16 (:132-137)
//   3. umma_pv_kernel   O[q, D] = P . V                          tcgen05.mma with an MN-major B operand (V rows as stored)
// Operands are staged in the canonical no-swizzle UMMA layout (8x8-element core matrices of 128 contiguous bytes;
// K-major: LBO = 128 B between K chunks, SBO between 8-row groups) by plain 128-bit loads/stores: the tiles are
// rotated (Q, K) or used once (P, V), so there is no TMA tensor map.  One elected thread issues the MMAs and
// tcgen05.commit signals an mbarrier; four warps read the accumulator back with tcgen05.ld (32 lanes each).
#include <math_constants.h>

#include <algorithm>
#include <cmath>

#include "../../include/spatten_b200.h"
#include "common.cuh"
#include "umma.cuh"

namespace spatten {

constexpr int kUThreads = 128;     // 4 warps: producers, then epilogue (warp w owns TMEM lanes 32w..32w+31)
constexpr int kTileM = 128;
constexpr int kTileN = 128;

struct PrefillParams {
  const __half* q;          // [b, q_len, H, D] un-rotated
  const __half* k_cache;    // [b, H_kv, >=kv_len, D] un-rotated (new rows already stored)
  const __half* v_cache;
  int64_t stride_b, stride_h;
  const __half* rope;       // [rope_len, D]: cos | sin first halves
  const int64_t* q_pos;     // [b, q_len]
  const __half* mask;       // optional [b, q_len, kv_len]
  __half* logits;           // [b, H, q_len, kv_len] strides below
  int64_t ls_b, ls_h, ls_q;
  __half* probs;            // scratch, same shape/strides as logits
  __half* out;              // [b, q_len, H, D]
  int batch, H, n_rep, q_len, kv_len, rope_len;
  float sqrt_d, rsqrt_d;
  float* importance;        // optional [b, H, imp_stride]: += column sums of P
  int64_t imp_stride;
};

// ---- 1. S = rope(Q) . rope(K)^T -----------------------------------------------------------------------------------
// grid (ceil(kv/128), ceil(q/128), b*H); 128 threads.  Whole head_dim in one K block (D/16 MMAs of K = 16).
template <int D>
__global__ void __launch_bounds__(kUThreads) umma_qk_kernel(const PrefillParams p) {
  constexpr int kChunks = D / 8;                       // 16-byte K chunks per row
  constexpr int kTileBytes = kTileM * D * 2;           // 32 KiB (D = 128)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kTileBytes;
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.z, b = bh / p.H, h = bh % p.H, kvh = h / p.n_rep;
  const int m0 = blockIdx.y * kTileM, n0 = blockIdx.x * kTileN;

  if (warp == 0) tmem_alloc(&s_tmem, kTileN);
  if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }

  // producers: rows of Q (rotated at q_pos) and of K (rotated at the slot index), fp16 op for op (:21-28)
  const __half* Kc = p.k_cache + b * p.stride_b + kvh * p.stride_h;
  // work unit of a warp-iteration: 8 rows x 4 chunk pairs.  lane % 8 -> row, lane / 8 -> chunk: a quarter-warp writes
  // one whole 128-byte core matrix (bank-conflict free) and each row is read as 64 contiguous bytes
  static_assert(kChunks / 2 == 8, "producer mapping assumes head_dim 128");
  for (int unit = warp; unit < (kTileM / 8) * 2; unit += kUThreads / 32) {
    const int r = (unit >> 1) * 8 + (lane & 7), c = (unit & 1) * 4 + (lane >> 3);   // row, chunk of the FIRST half
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const int row = (which ? n0 : m0) + r;
      const bool ok = row < (which ? p.kv_len : p.q_len);
      uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
      if (ok) {
        const __half* src = which ? Kc + static_cast<int64_t>(row) * D
                                  : p.q + ((static_cast<int64_t>(b) * p.q_len + row) * p.H + h) * D;
        int64_t pos = which ? row : p.q_pos[static_cast<int64_t>(b) * p.q_len + row];
        pos = max(static_cast<int64_t>(0), min(pos, static_cast<int64_t>(p.rope_len - 1)));
        const uint4 xl = *reinterpret_cast<const uint4*>(src + c * 8);
        const uint4 xh = *reinterpret_cast<const uint4*>(src + D / 2 + c * 8);
        const uint4 cs = __ldg(reinterpret_cast<const uint4*>(p.rope + pos * D + c * 8));
        const uint4 sn = __ldg(reinterpret_cast<const uint4*>(p.rope + pos * D + D / 2 + c * 8));
        const __half2* a = reinterpret_cast<const __half2*>(&xl);
        const __half2* bb = reinterpret_cast<const __half2*>(&xh);
        const __half2* c2 = reinterpret_cast<const __half2*>(&cs);
        const __half2* s2 = reinterpret_cast<const __half2*>(&sn);
        __half2* ol = reinterpret_cast<__half2*>(&lo);
        __half2* oh = reinterpret_cast<__half2*>(&hi);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ol[i] = __hsub2_rn(__hmul2_rn(a[i], c2[i]), __hmul2_rn(bb[i], s2[i]));
          oh[i] = __hadd2_rn(__hmul2_rn(bb[i], c2[i]), __hmul2_rn(a[i], s2[i]));
        }
      }
      uint8_t* dst = which ? sB : sA;
      *reinterpret_cast<uint4*>(dst + kmajor_off(r, c, kChunks)) = lo;
      *reinterpret_cast<uint4*>(dst + kmajor_off(r, c + kChunks / 2, kChunks)) = hi;
    }
  }
  fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc(kTileM, kTileN, 0);
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      const uint64_t ad = umma_desc(smem_u32(sA) + ks * 256, 128, kChunks * 128);
      const uint64_t bd = umma_desc(smem_u32(sB) + ks * 256, 128, kChunks * 128);
      umma_f16(tmem, ad, bd, idesc, ks > 0);
    }
    umma_commit(&s_bar);
  }
  mbar_wait_bounded(&s_bar, 0);
  tc_fence_after();

  // epilogue: warp w owns rows 32w..32w+31 (TMEM lanes); fp16(acc) / sqrt(D) -> fp16 (:111-113).  The fp16 tile is
  // staged in the (now idle) operand shared memory with a padded pitch and written out as whole 256-byte row
  // segments, 16 lanes x 16 bytes per row.
  constexpr int kPitch = kTileN * 2 + 16;              // bytes; +16 keeps the 16-byte stores of 32 rows off one bank group
  uint8_t* sC = smem;                                   // 128 x 272 B = 34 KiB <= the two operand tiles
  {
    const int r = warp * 32 + lane;
#pragma unroll 1
    for (int cb = 0; cb < kTileN; cb += 32) {
      float v[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + cb, v);
      __align__(16) __half o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i)
        o[i] = __float2half_rn(pf_div_by_const(__half2float(__float2half_rn(v[i])), p.sqrt_d, p.rsqrt_d));
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(sC + r * kPitch + cb * 2 + i * 16) = reinterpret_cast<const uint4*>(o)[i];
    }
  }
  __syncthreads();
  {
    __half* base = p.logits + b * p.ls_b + h * p.ls_h + n0;
    const bool vec_ok = (p.ls_q % 8) == 0 && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
    const int half_w = lane >> 4, piece = lane & 15;    // 2 rows per warp instruction, 16 pieces of 8 logits per row
#pragma unroll 1
    for (int rr = warp * 2 + half_w; rr < kTileM; rr += 8) {
      const int row = m0 + rr;
      if (row >= p.q_len) continue;
      const int col = piece * 8;
      if (n0 + col >= p.kv_len) continue;
      const uint4 val = *reinterpret_cast<const uint4*>(sC + rr * kPitch + col * 2);
      __half* dst = base + static_cast<int64_t>(row) * p.ls_q + col;
      if (vec_ok && n0 + col + 8 <= p.kv_len) {
        *reinterpret_cast<uint4*>(dst) = val;
      } else {
        const __half* hv = reinterpret_cast<const __half*>(&val);
        for (int i = 0; i < 8; ++i)
          if (n0 + col + i < p.kv_len) dst[i] = hv[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTileN);
}

// ---- 2. P = softmax(S + mask), fp32 (:132-137) ---------------------------------------------------------------------
// one warp per (b, h, q) row.  Pass 1 reads the row once with 128-bit loads and keeps a running (max, sum) per lane
// (online softmax, so no row-sized register array and full occupancy); pass 2 re-reads it -- an L2 hit, the row
// was touched microseconds ago -- and writes the fp16 probabilities.
__device__ __forceinline__ void sm_load8(const __half* s, const __half* mk, int j, float (&f)[8]) {
  uint4 x = *reinterpret_cast<const uint4*>(s + j);
  __half2* h2 = reinterpret_cast<__half2*>(&x);
  if (mk) {
    const uint4 mm = __ldg(reinterpret_cast<const uint4*>(mk + j));
    const __half2* m2 = reinterpret_cast<const __half2*>(&mm);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __hadd2_rn(h2[e], m2[e]);      // fp16 add, as attn_weights + attention_mask (:132)
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __half22float2(h2[e]);
    f[2 * e] = t.x; f[2 * e + 1] = t.y;
  }
}

__global__ void __launch_bounds__(256) softmax_rows_kernel(const PrefillParams p) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t rows = static_cast<int64_t>(p.batch) * p.H * p.q_len;
  if (row >= rows) return;
  const int qi = row % p.q_len, h = (row / p.q_len) % p.H, b = row / (static_cast<int64_t>(p.q_len) * p.H);
  const __half* s = p.logits + b * p.ls_b + h * p.ls_h + static_cast<int64_t>(qi) * p.ls_q;
  __half* o = p.probs + b * p.ls_b + h * p.ls_h + static_cast<int64_t>(qi) * p.ls_q;
  const __half* mk = p.mask ? p.mask + (static_cast<int64_t>(b) * p.q_len + qi) * p.kv_len : nullptr;
  const bool vec = (p.kv_len % 8) == 0 &&
                   ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(mk)) & 15) == 0;
  float mx = -CUDART_INF_F, sum = 0.f;
  if (vec) {
    for (int j = lane * 8; j < p.kv_len; j += 256) {
      float f[8];
      sm_load8(s, mk, j, f);
      float m8 = f[0];
#pragma unroll
      for (int e = 1; e < 8; ++e) m8 = fmaxf(m8, f[e]);
      const float m_new = fmaxf(mx, m8);
      float part = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) part += __expf(f[e] - m_new);
      sum = sum * __expf(mx - m_new) + part;
      mx = m_new;
    }
  } else {
    for (int j = lane; j < p.kv_len; j += 32) {
      __half x = s[j];
      if (mk) x = __hadd_rn(x, mk[j]);
      const float v = __half2float(x), m_new = fmaxf(mx, v);
      sum = sum * __expf(mx - m_new) + __expf(v - m_new);
      mx = m_new;
    }
  }
  const float M = warp_max(mx);
  sum = warp_sum(sum * __expf(mx - M));
  const float inv = 1.f / sum;
  if (vec) {
    for (int j = lane * 8; j < p.kv_len; j += 256) {
      float f[8];
      sm_load8(s, mk, j, f);
      uint4 outv;
      __half2* h2 = reinterpret_cast<__half2*>(&outv);
#pragma unroll
      for (int e = 0; e < 4; ++e) h2[e] = __floats2half2_rn(__expf(f[2 * e] - M) * inv, __expf(f[2 * e + 1] - M) * inv);
      *reinterpret_cast<uint4*>(o + j) = outv;
    }
  } else {
    for (int j = lane; j < p.kv_len; j += 32) {
      __half x = s[j];
      if (mk) x = __hadd_rn(x, mk[j]);
      o[j] = __float2half_rn(__expf(__half2float(x) - M) * inv);
    }
  }
}

// ---- 3. O = P . V --------------------------------------------------------------------------------------------------
// grid (1, ceil(q/128), b*H); N = D; K loop over kv in blocks of 64.  A = P rows (K-major), B = V rows as stored
// (MN-major: 16 bytes along D, 8 tokens per core matrix).
template <int D>
__global__ void __launch_bounds__(kUThreads) umma_pv_kernel(const PrefillParams p) {
  constexpr int BK = 64;
  constexpr int kAChunks = BK / 8;                      // 8 K chunks per P row
  constexpr int kABytes = kTileM * BK * 2;              // 16 KiB
  constexpr int kBBytes = BK * D * 2;                   // 16 KiB (D = 128)
  constexpr int kNChunks = D / 8;                       // 16-byte chunks along D
  static_assert(D == 128 && kTileM * kAChunks == 32 * 32 && BK * kNChunks == 32 * 32, "producer mapping assumes these tiles");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kABytes;
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.z, b = bh / p.H, h = bh % p.H, kvh = h / p.n_rep;
  const int m0 = blockIdx.y * kTileM;
  if (warp == 0) tmem_alloc(&s_tmem, D < 32 ? 32 : D);
  if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const __half* P = p.probs + b * p.ls_b + h * p.ls_h;
  const __half* V = p.v_cache + b * p.stride_b + kvh * p.stride_h;
  const bool vec_ok = (p.ls_q % 8) == 0 && ((reinterpret_cast<uintptr_t>(P) & 15) == 0);
  const int nk = (p.kv_len + BK - 1) / BK;
  constexpr int kAPer = kTileM * kAChunks / kUThreads;      // 8 x 16-byte pieces of P per thread per block
  constexpr int kBPer = BK * kNChunks / kUThreads;          // 8 x 16-byte pieces of V per thread per block
  uint4 ra[kAPer], rb[kBPer];
  // global -> registers for K block kb (software pipelining: issued while the tensor core works on block kb-1)
  auto fetch = [&](int kb) {
    const int k0 = kb * BK;
#pragma unroll
    for (int i = 0; i < kAPer; ++i) {
      const int unit = warp + i * (kUThreads / 32);                    // 8 rows x 4 chunks per warp-iteration
      const int r = (unit >> 1) * 8 + (lane & 7), c = (unit & 1) * 4 + (lane >> 3);
      uint4 val = make_uint4(0, 0, 0, 0);
      const int row = m0 + r, col = k0 + c * 8;
      if (row < p.q_len && col < p.kv_len) {
        const __half* src = P + static_cast<int64_t>(row) * p.ls_q + col;
        if (vec_ok && col + 8 <= p.kv_len) {
          val = *reinterpret_cast<const uint4*>(src);
        } else {
          __align__(16) __half t[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) t[e] = (col + e < p.kv_len) ? src[e] : __float2half(0.f);
          val = *reinterpret_cast<const uint4*>(t);
        }
      }
      ra[i] = val;
    }
#pragma unroll
    for (int i = 0; i < kBPer; ++i) {
      const int unit = warp + i * (kUThreads / 32);                    // 8 tokens x 4 D-chunks per warp-iteration
      const int j = (unit >> 2) * 8 + (lane & 7), c = (unit & 3) * 4 + (lane >> 3);
      rb[i] = (k0 + j < p.kv_len) ? *reinterpret_cast<const uint4*>(V + static_cast<int64_t>(k0 + j) * D + c * 8)
                                  : make_uint4(0, 0, 0, 0);
    }
  };
  fetch(0);
  for (int kb = 0; kb < nk; ++kb) {
    if (kb > 0) {                       // the MMAs of block kb-1 have consumed shared memory
      mbar_wait_bounded(&s_bar, (kb - 1) & 1);
      tc_fence_after();
    }
    // A: P[m0 + r, k0 + 8c .. +8) K-major;  B: V[k0 + j, 8c .. +8) MN-major: core matrix (token group j/8, D chunk c)
    // at (j/8) * LBO + c * SBO, row j%8 inside
#pragma unroll
    for (int i = 0; i < kAPer; ++i) {
      const int unit = warp + i * (kUThreads / 32);
      *reinterpret_cast<uint4*>(sA + kmajor_off((unit >> 1) * 8 + (lane & 7), (unit & 1) * 4 + (lane >> 3), kAChunks)) = ra[i];
    }
#pragma unroll
    for (int i = 0; i < kBPer; ++i) {
      const int unit = warp + i * (kUThreads / 32);
      const int j = (unit >> 2) * 8 + (lane & 7), c = (unit & 3) * 4 + (lane >> 3);
      *reinterpret_cast<uint4*>(sB + (j >> 3) * (kNChunks * 128) + c * 128 + (j & 7) * 16) = rb[i];
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
      constexpr uint32_t idesc = umma_idesc(kTileM, D, 1);
#pragma unroll
      for (int ks = 0; ks < BK / 16; ++ks) {
        const uint64_t ad = umma_desc(smem_u32(sA) + ks * 256, 128, kAChunks * 128);
        // MN-major B: 2 token groups per MMA; LBO = stride between token groups, SBO = stride between D chunks
        const uint64_t bd = umma_desc(smem_u32(sB) + ks * 2 * (kNChunks * 128), kNChunks * 128, 128);
        umma_f16(tmem, ad, bd, idesc, kb > 0 || ks > 0);
      }
      umma_commit(&s_bar);
    }
    if (kb + 1 < nk) fetch(kb + 1);     // in flight while the tensor core runs block kb
  }
  mbar_wait_bounded(&s_bar, (nk - 1) & 1);
  tc_fence_after();

  const int row = m0 + warp * 32 + lane;
  __half* dst = p.out + ((static_cast<int64_t>(b) * p.q_len + row) * p.H + h) * D;
#pragma unroll 1
  for (int cb = 0; cb < D; cb += 32) {
    float v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + cb, v);
    if (row < p.q_len) {
      __align__(16) __half o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __float2half_rn(v[i]);
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(dst + cb)[i] = reinterpret_cast<const uint4*>(o)[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, D < 32 ? 32 : D);
}

// importance[b, h, j] += sum_q P[b, h, q, j]: the column sums of the probability matrix (cumulative token importance).
// A flash-style kernel never materialises P; this path keeps it as the fp16 scratch the P.V tiles read anyway, so the
// column sums are one more streaming pass over it: a thread owns 8 adjacent columns (one 16-byte load per row), a CTA
// 1024 columns x kColsumRows rows; one atomicAdd per (thread, column) at the end.
constexpr int kColsumThreads = 128;
constexpr int kColsumRows = 1024;
__global__ void __launch_bounds__(kColsumThreads) prob_colsum_kernel(const PrefillParams p) {
  const int bh = blockIdx.z, h = bh % p.H, b = bh / p.H;
  const int j0 = (blockIdx.x * kColsumThreads + threadIdx.x) * 8;
  if (j0 >= p.kv_len) return;
  const int r0 = blockIdx.y * kColsumRows, r1 = min(p.q_len, r0 + kColsumRows);
  const __half* P = p.probs + b * p.ls_b + h * p.ls_h;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const bool vec = j0 + 8 <= p.kv_len && (p.ls_q % 8) == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0;
  if (vec) {
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(P + static_cast<int64_t>(r) * p.ls_q + j0));
      const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        acc[2 * e] += f.x;
        acc[2 * e + 1] += f.y;
      }
    }
  } else {
    for (int r = r0; r < r1; ++r)
      for (int e = 0; e < 8 && j0 + e < p.kv_len; ++e) acc[e] += __half2float(P[static_cast<int64_t>(r) * p.ls_q + j0 + e]);
  }
  float* imp = p.importance + (static_cast<int64_t>(b) * p.H + h) * p.imp_stride + j0;
#pragma unroll
  for (int e = 0; e < 8; ++e)
    if (j0 + e < p.kv_len && acc[e] != 0.f) atomicAdd(imp + e, acc[e]);
}

}  // namespace spatten

using namespace spatten;

namespace spatten {
int64_t prefill_fused_scratch_bytes(int batch, int H, int H_kv, int q_len, int kv_len);   // prefill_fused.cu
int launch_prefill_fused(const SpattenPrefillArgs* a, cudaStream_t st);
}  // namespace spatten

extern "C" int64_t spatten_prefill_scratch_bytes(int32_t batch, int32_t num_heads, int32_t num_kv_heads, int32_t q_len,
                                                 int32_t kv_len) {
  return prefill_fused_scratch_bytes(batch, num_heads, num_kv_heads, q_len, kv_len);
}

extern "C" int spatten_prefill_attn(const SpattenPrefillArgs* a, void* stream) {
  SPATTEN_CHECK_ARG(a != nullptr && a->struct_size == sizeof(SpattenPrefillArgs), "spatten_prefill_attn: bad args/struct_size");
  SPATTEN_CHECK_ARG(a->batch > 0 && a->num_heads > 0 && a->num_kv_heads > 0 && a->num_heads % a->num_kv_heads == 0,
                    "spatten_prefill_attn: bad head counts");
  if (a->head_dim != 128) { set_error("spatten_prefill_attn: head_dim %d unsupported (128)", a->head_dim); return SPATTEN_ERR_UNSUPPORTED; }
  SPATTEN_CHECK_ARG(a->q_len > 0 && a->kv_len >= a->q_len && a->rope_len >= a->kv_len, "spatten_prefill_attn: bad q_len / kv_len / rope_len");
  SPATTEN_CHECK_ARG(a->q && a->k_cache && a->v_cache && a->rope && a->q_pos && a->logits && a->probs && a->out,
                    "spatten_prefill_attn: NULL pointer");
  SPATTEN_CHECK_ARG(a->logits_stride_q >= a->kv_len, "spatten_prefill_attn: logits_stride_q < kv_len");
  SPATTEN_CHECK_ARG(a->importance == nullptr || a->importance_stride >= a->kv_len,
                    "spatten_prefill_attn: importance_stride < kv_len");
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k_cache) |
                       reinterpret_cast<uintptr_t>(a->v_cache) | reinterpret_cast<uintptr_t>(a->rope) |
                       reinterpret_cast<uintptr_t>(a->out);
  SPATTEN_CHECK_ARG((al & 15) == 0 && a->cache_stride_b % 8 == 0 && a->cache_stride_h % 8 == 0,
                    "spatten_prefill_attn: q/k_cache/v_cache/rope/out must be 16-byte aligned");
  PrefillParams p{};
  p.q = static_cast<const __half*>(a->q);
  p.k_cache = static_cast<const __half*>(a->k_cache);
  p.v_cache = static_cast<const __half*>(a->v_cache);
  p.stride_b = a->cache_stride_b; p.stride_h = a->cache_stride_h;
  p.rope = static_cast<const __half*>(a->rope); p.q_pos = a->q_pos; p.mask = static_cast<const __half*>(a->mask);
  p.logits = static_cast<__half*>(a->logits); p.probs = static_cast<__half*>(a->probs); p.out = static_cast<__half*>(a->out);
  p.ls_b = a->logits_stride_b; p.ls_h = a->logits_stride_h; p.ls_q = a->logits_stride_q;
  p.batch = a->batch; p.H = a->num_heads; p.n_rep = a->num_heads / a->num_kv_heads; p.q_len = a->q_len; p.kv_len = a->kv_len; p.rope_len = a->rope_len;
  p.sqrt_d = static_cast<float>(sqrt(static_cast<double>(a->head_dim))); p.rsqrt_d = 1.0f / p.sqrt_d;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!(a->flags & SPATTEN_PREFILL_LEGACY) && a->scratch_bytes > 0) {
    SPATTEN_CHECK_ARG((reinterpret_cast<uintptr_t>(a->probs) & 15) == 0, "spatten_prefill_attn: scratch must be 16-byte aligned");
    if (a->scratch_bytes >= prefill_fused_scratch_bytes(a->batch, a->num_heads, a->num_kv_heads, a->q_len, a->kv_len))
      return launch_prefill_fused(a, st);
    SPATTEN_CHECK_ARG(a->scratch_bytes >= (static_cast<int64_t>(a->batch - 1) * a->logits_stride_b +
                                           static_cast<int64_t>(a->num_heads - 1) * a->logits_stride_h +
                                           static_cast<int64_t>(a->q_len - 1) * a->logits_stride_q + a->kv_len) * 2,
                      "spatten_prefill_attn: scratch_bytes too small for either path");
  }
  constexpr int D = 128;
  constexpr int smem_qk = 2 * kTileM * D * 2, smem_pv = kTileM * 64 * 2 + 64 * D * 2;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    cudaError_t e1 = cudaFuncSetAttribute(umma_qk_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_qk);
    cudaError_t e2 = cudaFuncSetAttribute(umma_pv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pv);
    if (e1 != cudaSuccess || e2 != cudaSuccess) return cuda_fail(e1 != cudaSuccess ? e1 : e2, "cudaFuncSetAttribute(umma kernels)");
    configured[dev] = true;
  }
  const int mt = (a->q_len + kTileM - 1) / kTileM, nt = (a->kv_len + kTileN - 1) / kTileN, bh = a->batch * a->num_heads;
  umma_qk_kernel<D><<<dim3(nt, mt, bh), kUThreads, smem_qk, st>>>(p);
  const int64_t rows = static_cast<int64_t>(a->batch) * a->num_heads * a->q_len;
  softmax_rows_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(p);
  umma_pv_kernel<D><<<dim3(1, mt, bh), kUThreads, smem_pv, st>>>(p);
  if (a->importance != nullptr) {
    p.importance = a->importance;
    p.imp_stride = a->importance_stride;
    prob_colsum_kernel<<<dim3((a->kv_len + kColsumThreads * 8 - 1) / (kColsumThreads * 8),
                              (a->q_len + kColsumRows - 1) / kColsumRows, bh), kColsumThreads, 0, st>>>(p);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "spatten_prefill_attn launch");
}
