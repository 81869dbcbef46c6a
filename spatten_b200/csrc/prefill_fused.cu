// Fused prefill attention: one warp-specialised tcgen05 kernel per (128-query tile, head) -- the single-pass successor of
// the three-launch path in prefill_umma.cu.  Same function: llama_pos_shift_attention_forward for q_len > 1
// (reference spatten_llm/pos_shift/modify_llama.py:86-144): fp16 RoPE on Q and K (:21-28, :104-108), the pre-mask fp16
// logit matrix `attn_scores` as an OUTPUT (:111-119), additive mask + fp32 softmax -> fp16 (:127-137), P.V (:138), plus
// the column sums of P (cumulative token importance, README.md:21).
//
//   0. prefill_pack_kernel   Q and K are rotated once (fp16, op for op) and Q / K / V are re-tiled into 32 KiB operand
//                            IMAGES -- the exact shared-memory byte order the tensor core reads (8x8 core matrices, no
//                            swizzle) -- so the main kernel feeds its rings with one contiguous 32 KiB bulk copy (TMA
//                            engine, SASS UBLKCP) per tile and no CUDA-core staging.
//   1. prefill_fused_kernel  576 threads: warps 0-15 = two softmax groups (group g owns TMEM buffer S[g] and P buffer g,
//                            two threads per query row), warp 16 = loaders (K lane, V lane), warp 17 = MMA issuer.
//        phase A  for every key tile j:      S[j&1] = Q.K_j^T (tcgen05)  ->  fp16 logits written ONCE (coalesced through
//                                            warp-private shared-memory staging); running row max / row sum (online) for the tiles
//                                            the mask leaves alive.  HBM-bound: 32 KiB of logits per tile.
//        phase B  for every live key tile j: S = Q.K_j^T again (cheaper than re-reading the logits from HBM),
//                                            p = exp(s - M) / L with the FINAL row statistics -> fp16 P tile in shared memory
//                                            -> O += P.V_j (tcgen05, accumulator stays in TMEM, no rescaling pass);
//                                            column sums of the same P tile -> importance (stats-then-accumulate,
//                                            SURVEY.md 7 "hard part 7").  Fully masked tiles are skipped.
//   The additive mask is classified on the device (mask_classify_kernel): the standard causal mask (0 / <= -60000) turns
//   into an in-register predicate and tile skipping; anything else is read from global memory per element.
#include <math_constants.h>

#include <algorithm>
#include <cmath>

#include "../../include/spatten_b200.h"
#include "common.cuh"
#include "umma.cuh"

namespace spatten {

constexpr int kFT = 128;                       // tile edge: queries x keys
constexpr int kFD = 128;                       // head_dim
constexpr int kImgBytes = kFT * kFD * 2;       // 32 KiB operand image
constexpr int kGroupThreads = 256;                // one softmax group: 8 warps, two threads per query row
constexpr int kLoadWarp = 16, kMmaWarp = 17;
constexpr int kFusedThreads = 18 * 32;
constexpr int kStagePitch = 64 * 2 + 16;       // bytes; a warp stages its 32 rows x 64 logits (16-byte stores of 8 rows hit 8 bank groups)
constexpr int kStageBytes = 32 * kStagePitch;  // 4608 per warp
constexpr int kOffQ = 0, kOffK = kImgBytes, kOffV = 3 * kImgBytes, kOffP = 5 * kImgBytes, kOffBars = 7 * kImgBytes;
constexpr int kOffStats = kOffBars + 256, kOffTmem = kOffStats + 2 * kFT * 8;
constexpr int kFusedSmem = kOffTmem + 16;
static_assert(16 * kStageBytes + 2048 <= 3 * kImgBytes, "phase-A staging + statistics exchange must fit in the V ring + P[0]");
static_assert(kFusedSmem <= 232448, "shared memory budget");

enum FusedBar { Q_FULL = 0, K_FULL = 1, K_EMPTY = 3, V_FULL = 5, V_EMPTY = 7, S_FULL = 9, S_EMPTY = 11, P_FULL = 13,
                P_EMPTY = 15, O_FULL = 17, A_DONE = 18, NUM_BARS = 19 };

// byte offset of the 16-byte piece (row r, 8-element chunk c) inside a 128 x 128 fp16 operand image: core matrix
// (r / 8, c) of 128 contiguous bytes.  As a K-major operand (Q, K, P: rows = M/N, chunks = K) LBO = 128, SBO = 2048;
// as the MN-major B operand (V: rows = tokens = K, chunks = head_dim = N) LBO = 2048, SBO = 128.
__device__ __forceinline__ uint32_t img_off(int r, int c) { return static_cast<uint32_t>((r >> 3) * 2048 + c * 128 + (r & 7) * 16); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void group_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

struct PackParams {
  const __half* q;          // [b, q_len, H, D]
  const __half* k_cache;    // [b, H_kv, cap, D]
  const __half* v_cache;
  int64_t stride_b, stride_h;
  const __half* rope;
  const int64_t* q_pos;
  uint8_t* q_img;           // [b, H, mt, 32 KiB]
  uint8_t* k_img;           // [b, H_kv, nt, 32 KiB]
  uint8_t* v_img;
  int* mode_flag;
  int batch, H, H_kv, q_len, kv_len, rope_len, mt, nt;
};

// grid (max(mt, nt), b*H + 2*b*H_kv), 128 threads: one operand image per CTA.  Rows past the end are zero-filled, so the
// main kernel never needs a bounds check on its loads.
__global__ void __launch_bounds__(128) prefill_pack_kernel(const PackParams p) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int z = blockIdx.y;
  const int tile = blockIdx.x;
  if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *p.mode_flag = 1;          // "mask is causal" until proven otherwise
  int kind = 0;                                                                // 0 = Q, 1 = K, 2 = V
  if (z >= p.batch * p.H) { z -= p.batch * p.H; kind = 1; if (z >= p.batch * p.H_kv) { z -= p.batch * p.H_kv; kind = 2; } }
  if (tile >= (kind == 0 ? p.mt : p.nt)) return;
  const int heads = kind == 0 ? p.H : p.H_kv;
  const int b = z / heads, h = z % heads;
  const int rows = kind == 0 ? p.q_len : p.kv_len;
  uint8_t* dst = (kind == 0 ? p.q_img : kind == 1 ? p.k_img : p.v_img) +
                 (static_cast<int64_t>(z) * (kind == 0 ? p.mt : p.nt) + tile) * kImgBytes;
  const __half* cache = (kind == 2 ? p.v_cache : p.k_cache) + b * p.stride_b + h * p.stride_h;
  // warp-iteration unit: 8 rows x 4 chunk pairs (lane % 8 -> row, lane / 8 -> chunk): a quarter-warp writes one whole
  // 128-byte core matrix and reads 64 contiguous bytes of each row half
  for (int unit = warp; unit < (kFT / 8) * 2; unit += 4) {
    const int r = (unit >> 1) * 8 + (lane & 7), c = (unit & 1) * 4 + (lane >> 3);
    const int row = tile * kFT + r;
    uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
    if (row < rows) {
      const __half* src = kind == 0 ? p.q + ((static_cast<int64_t>(b) * p.q_len + row) * p.H + h) * kFD
                                    : cache + static_cast<int64_t>(row) * kFD;
      const uint4 xl = *reinterpret_cast<const uint4*>(src + c * 8);
      const uint4 xh = *reinterpret_cast<const uint4*>(src + kFD / 2 + c * 8);
      if (kind == 2) {
        lo = xl; hi = xh;
      } else {
        int64_t pos = kind == 0 ? p.q_pos[static_cast<int64_t>(b) * p.q_len + row] : row;       // K rotates at its slot (:104-108)
        pos = max(static_cast<int64_t>(0), min(pos, static_cast<int64_t>(p.rope_len - 1)));
        const uint4 cs = __ldg(reinterpret_cast<const uint4*>(p.rope + pos * kFD + c * 8));
        const uint4 sn = __ldg(reinterpret_cast<const uint4*>(p.rope + pos * kFD + kFD / 2 + c * 8));
        const __half2* a = reinterpret_cast<const __half2*>(&xl);
        const __half2* bb = reinterpret_cast<const __half2*>(&xh);
        const __half2* c2 = reinterpret_cast<const __half2*>(&cs);
        const __half2* s2 = reinterpret_cast<const __half2*>(&sn);
        __half2* ol = reinterpret_cast<__half2*>(&lo);
        __half2* oh = reinterpret_cast<__half2*>(&hi);
#pragma unroll
        for (int i = 0; i < 4; ++i) {                                         // x*cos + rotate_half(x)*sin, fp16 op for op (:27)
          ol[i] = __hsub2_rn(__hmul2_rn(a[i], c2[i]), __hmul2_rn(bb[i], s2[i]));
          oh[i] = __hadd2_rn(__hmul2_rn(bb[i], c2[i]), __hmul2_rn(a[i], s2[i]));
        }
      }
    }
    *reinterpret_cast<uint4*>(dst + img_off(r, c)) = lo;
    *reinterpret_cast<uint4*>(dst + img_off(r, c + kFD / 16)) = hi;
  }
}

// *flag = 0 unless mask[b, i, j] is 0 for j <= i + (kv_len - q_len) and <= -60000 (or -inf) beyond: the standard additive
// causal mask (HF `_make_causal_mask` fills with finfo.min = -65504).  A thread checks 8 consecutive elements.
__global__ void __launch_bounds__(256) mask_classify_kernel(const __half* mask, int64_t total, int q_len, int kv_len, int* flag) {
  const int64_t e0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (e0 >= total) return;
  const int off = kv_len - q_len;
  bool ok = true;
  const bool vec = e0 + 8 <= total && (reinterpret_cast<uintptr_t>(mask + e0) & 15) == 0;
  __align__(16) __half v[8];
  if (vec) {
    *reinterpret_cast<uint4*>(v) = __ldg(reinterpret_cast<const uint4*>(mask + e0));
  } else {
    for (int i = 0; i < 8; ++i) v[i] = e0 + i < total ? mask[e0 + i] : __float2half(0.f);
  }
  const int64_t r = e0 / kv_len;
  int j = static_cast<int>(e0 - r * kv_len), qi = static_cast<int>(r % q_len);
  for (int i = 0; i < 8 && e0 + i < total; ++i) {
    const float m = __half2float(v[i]);
    ok = ok && (j <= qi + off ? m == 0.f : m <= -60000.f);
    if (++j == kv_len) { j = 0; if (++qi == q_len) qi = 0; }
  }
  if (!ok) *flag = 0;
}

struct FusedParams {
  const uint8_t* q_img;
  const uint8_t* k_img;
  const uint8_t* v_img;
  const int* mode_flag;
  const __half* mask;       // [b, q_len, kv_len] or null
  __half* logits;
  int64_t ls_b, ls_h, ls_q;
  __half* out;              // [b, q_len, H, D]
  float* importance;
  int64_t imp_stride;
  int batch, H, n_rep, q_len, kv_len, mt, nt;
  int mode;                 // 1 = causal, 2 = no mask, 0 = mask given: read *mode_flag (1 causal / 0 generic)
  int debug;                // timing experiments only (tools/microbench.py): 1 = no logit stores, 2 = skip phase B, 4 = no phase-A statistics
  float rs_up;              // nextafter(1/sqrt(D)): fp16(x * rs_up) == fp16(x / sqrt(D)) for every fp16 x (D = 128)
};

// fp16 logit exactly as the reference forms it: fp16(fp16(acc) / sqrt(D)) (:111-113), two at a time
__device__ __forceinline__ __half2 logit2(float a, float b, float rs_up) {
  const float2 f = __half22float2(__floats2half2_rn(a, b));
  return __floats2half2_rn(f.x * rs_up, f.y * rs_up);
}

// tcgen05.mma with the two shared-memory descriptors passed as (low, high) words: the high words are compile-time
// constants and the low words differ from a base by an immediate, so one MMA costs a handful of issue slots
__device__ __forceinline__ void umma_f16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}

// additive mask of columns 2i, 2i + 1 of a 32-column chunk; columns past `lim` (beyond kv_len) read as 0
__device__ __forceinline__ __half2 mask2(const __half* mchunk, int i, int lim) {
  const __half z = __float2half(0.f);
  return __halves2half2(2 * i <= lim ? __ldg(mchunk + 2 * i) : z, 2 * i + 1 <= lim ? __ldg(mchunk + 2 * i + 1) : z);
}

__global__ void __launch_bounds__(kFusedThreads, 1) prefill_fused_kernel(const FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sP = smem + kOffP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  float2* sStats = reinterpret_cast<float2*>(smem + kOffStats);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + kOffTmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mtile = p.mt - 1 - static_cast<int>(blockIdx.x);               // heavy (late) query tiles first
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H, kvh = h / p.n_rep;
  const int m0 = mtile * kFT;
  const int off = p.kv_len - p.q_len;

  if (warp == kLoadWarp) tmem_alloc(sTmem, 512);                           // S[0], S[1], O: 3 x 128 columns
  if (tid == kMmaWarp * 32) {
    mbar_init(&bars[Q_FULL], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[K_FULL + i], 1);  mbar_init(&bars[K_EMPTY + i], 1);
      mbar_init(&bars[V_FULL + i], 1);  mbar_init(&bars[V_EMPTY + i], 1);
      mbar_init(&bars[S_FULL + i], 1);  mbar_init(&bars[S_EMPTY + i], kGroupThreads);
      mbar_init(&bars[P_FULL + i], kGroupThreads); mbar_init(&bars[P_EMPTY + i], 1);
    }
    mbar_init(&bars[O_FULL], 1);
    mbar_init(&bars[A_DONE], 2 * kGroupThreads / 32);
    fence_mbar_init();
  }
  int mode = p.mode;
  if (mode == 0) mode = (*p.mode_flag != 0) ? 1 : 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sTmem;

  const int last_row = min(m0 + kFT - 1, p.q_len - 1);
  const int nt = p.nt;
  const int nb = (p.debug & 2) ? 0 : (mode == 1 ? min(nt, (last_row + off) / kFT + 1) : nt);      // key tiles the mask leaves alive
  const int nb_stats = (p.debug & 4) ? 0 : (mode == 1 ? min(nt, (last_row + off) / kFT + 1) : nt);

  if (warp == kLoadWarp) {
    // ===== loaders: contiguous 32 KiB images through the TMA engine; lane 0 feeds Q and the K ring, lane 1 the V ring =====
    const uint64_t keep = policy_evict_last();                             // K / V images are re-read by every query tile of the head
    const int64_t kv_img0 = (static_cast<int64_t>(b) * (p.H / p.n_rep) + kvh) * nt * kImgBytes;
    if (lane == 0) {
      const uint8_t* kimg = p.k_img + kv_img0;
      mbar_expect_tx(&bars[Q_FULL], kImgBytes);
      bulk_g2s(sQ, p.q_img + (static_cast<int64_t>(bh) * p.mt + mtile) * kImgBytes, kImgBytes, &bars[Q_FULL], policy_evict_first());
      for (int kuse = 0; kuse < nt + nb; ++kuse) {
        const int s = kuse & 1, j = kuse < nt ? kuse : kuse - nt;
        mbar_wait_bounded(&bars[K_EMPTY + s], ((kuse >> 1) & 1) ^ 1);
        mbar_expect_tx(&bars[K_FULL + s], kImgBytes);
        bulk_g2s(sK + s * kImgBytes, kimg + static_cast<int64_t>(j) * kImgBytes, kImgBytes, &bars[K_FULL + s], keep);
      }
    } else if (lane == 1) {
      const uint8_t* vimg = p.v_img + kv_img0;
      while (!mbar_try_wait(&bars[A_DONE], 0)) __nanosleep(500);           // the V ring doubles as phase A's logits staging
      for (int j = 0; j < nb; ++j) {
        const int vs = j & 1;
        mbar_wait_bounded(&bars[V_EMPTY + vs], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&bars[V_FULL + vs], kImgBytes);
        bulk_g2s(sV + vs * kImgBytes, vimg + static_cast<int64_t>(j) * kImgBytes, kImgBytes, &bars[V_FULL + vs], keep);
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer =====
    // One lane issues every tcgen05.mma.  The warp shares its scheduler with four softmax warps, so the issue sequence is
    // kept short: descriptor words are built once (only the 14-bit start-address field changes per MMA) and phase B
    // issues whichever of "next S tile" / "next P.V" is ready first instead of a fixed order.
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc(kFT, kFT, 0);
      constexpr uint32_t idesc_pv = umma_idesc(kFT, kFD, 1);
      constexpr uint32_t hi_kmajor = (2048u >> 4) | (1u << 14);            // SBO = 2048, descriptor version 1
      constexpr uint32_t hi_mnmajor = (128u >> 4) | (1u << 14);            // SBO = 128
      constexpr uint32_t lbo_kmajor = (128u >> 4) << 16, lbo_mnmajor = (2048u >> 4) << 16;
      const uint32_t dQ = (smem_u32(sQ) >> 4) | lbo_kmajor, dK = (smem_u32(sK) >> 4) | lbo_kmajor;
      const uint32_t dP = (smem_u32(sP) >> 4) | lbo_kmajor, dV = (smem_u32(sV) >> 4) | lbo_mnmajor;
      mbar_wait_bounded(&bars[Q_FULL], 0);
      uint32_t jt = 0;                                                     // S tiles issued so far: K slot = S slot = jt & 1
      auto qk_ready = [&]() {
        const int s = jt & 1;
        return mbar_try_wait(&bars[K_FULL + s], (jt >> 1) & 1) && mbar_try_wait(&bars[S_EMPTY + s], ((jt >> 1) & 1) ^ 1);
      };
      auto issue_qk = [&]() {
        const int s = jt & 1;
        tc_fence_after();
        const uint32_t d = tmem + s * kFT, bk = dK + s * (kImgBytes >> 4);
#pragma unroll
        for (int ks = 0; ks < kFD / 16; ++ks) umma_f16_split(d, dQ + ks * 16, hi_kmajor, bk + ks * 16, hi_kmajor, idesc_qk, ks > 0);
        umma_commit(&bars[K_EMPTY + s]);
        umma_commit(&bars[S_FULL + s]);
        ++jt;
      };
      auto pv_ready = [&](int j) {
        return mbar_try_wait(&bars[P_FULL + ((nt + j) & 1)], (j >> 1) & 1) && mbar_try_wait(&bars[V_FULL + (j & 1)], (j >> 1) & 1);
      };
      auto issue_pv = [&](int j) {
        const int g = (nt + j) & 1, vs = j & 1;
        tc_fence_after();
        const uint32_t d = tmem + 2 * kFT, ap = dP + g * (kImgBytes >> 4), bv = dV + vs * (kImgBytes >> 4);
        if (j == 0) {
#pragma unroll
          for (int ks = 0; ks < kFT / 16; ++ks) umma_f16_split(d, ap + ks * 16, hi_kmajor, bv + ks * 256, hi_mnmajor, idesc_pv, ks > 0);
        } else {
#pragma unroll
          for (int ks = 0; ks < kFT / 16; ++ks) umma_f16_split(d, ap + ks * 16, hi_kmajor, bv + ks * 256, hi_mnmajor, idesc_pv, true);
        }
        umma_commit(&bars[P_EMPTY + g]);
        umma_commit(&bars[V_EMPTY + vs]);
      };
      const long long t0 = clock64();
      for (int j = 0; j < nt; ++j) {
        while (!qk_ready()) { if (clock64() - t0 > 4000000000ll) __trap(); }
        issue_qk();
      }
      int qn = 0, pn = 0;                                                  // phase-B S tiles / P.V tiles issued
      while (pn < nb) {
        // S of tile qn may run at most two tiles ahead of P.V (its group must have drained S of tile qn - 2: S_EMPTY)
        if (qn < nb && qk_ready()) { issue_qk(); ++qn; continue; }
        if (pn < qn && pv_ready(pn)) { issue_pv(pn); ++pn; continue; }
        if (clock64() - t0 > 4000000000ll) __trap();
      }
      umma_commit(&bars[O_FULL]);
    }
    __syncwarp();
  } else {
    // ===== softmax groups: group g drains S[g]; two threads per query row (TMEM lane), 64 key columns each =====
    const int g = warp >> 3, wg = warp & 7, quarter = warp & 3, half = (warp >> 2) & 1, row = quarter * 32 + lane;
    const int col0 = half * 64;
    const int row_g = m0 + row;                                            // query index inside the chunk
    const bool row_ok = row_g < p.q_len;
    const uint32_t t_s = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + g * kFT + col0;
    const int row_lim = (mode == 1 ? min(row_g + off, p.kv_len - 1) : p.kv_len - 1);   // last live key of this row
    const __half* mrow = (mode == 0 && row_ok) ? p.mask + (static_cast<int64_t>(b) * p.q_len + row_g) * p.kv_len : nullptr;
    uint8_t* stage = sV + warp * kStageBytes;                              // warp-private: rows of this TMEM quarter, this column half
    __half* lbase = p.logits + b * p.ls_b + h * p.ls_h;
    const bool vec_ok = (p.ls_q % 8) == 0 && ((reinterpret_cast<uintptr_t>(lbase) & 15) == 0);
    uint32_t suse = 0;
    float m = -CUDART_INF_F, l = 0.f;
    __half2 lg[32];                                                        // this thread's 64 logits of the current tile

    // S[g] -> fp16 logits in registers, then the buffer goes straight back to the MMA issuer
    auto drain_s = [&]() {
      mbar_wait_bounded(&bars[S_FULL + g], suse & 1);
      ++suse;
      tc_fence_after();
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        float v[32];
        tmem_ld32(t_s + c2 * 32, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) lg[c2 * 16 + i] = logit2(v[2 * i], v[2 * i + 1], p.rs_up);
      }
      tc_fence_before();
      mbar_arrive(&bars[S_EMPTY + g]);
    };

    // ---- phase A: logits + row statistics ----
    for (int j = g; j < nt; j += 2) {
      const int n0 = j * kFT;
      drain_s();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<uint4*>(stage + lane * kStagePitch + i * 16) = *reinterpret_cast<const uint4*>(&lg[4 * i]);
      const int lim = row_lim - n0 - col0;                                 // columns 0..lim of this thread's 64 are live
      if (j < nb_stats && row_ok && lim >= 0) {
        if (lim >= 63 && mrow == nullptr) {
          __half2 hm = lg[0];
#pragma unroll
          for (int i = 1; i < 32; ++i) hm = __hmax2(hm, lg[i]);
          const float cm = fmaxf(__low2float(hm), __high2float(hm));
          if (cm > m) { l *= fast_exp2((m - cm) * kLog2e); m = cm; }      // m = -inf: l is 0 and stays 0
          const float ml2 = m * kLog2e;
          float part0 = 0.f, part1 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float2 f = __half22float2(lg[i]);
            part0 += fast_exp2(fmaf(f.x, kLog2e, -ml2));
            part1 += fast_exp2(fmaf(f.y, kLog2e, -ml2));
          }
          l += part0 + part1;
        } else {
#pragma unroll
          for (int cb = 0; cb < 64; cb += 16) {
            if (cb > lim) break;
            float x[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __half2 t = lg[(cb >> 1) + i];
              if (mrow) t = __hadd2_rn(t, mask2(mrow + n0 + col0 + cb, i, lim - cb));   // fp16 add, as attn_weights + attention_mask (:132)
              const float2 f = __half22float2(t);
              x[2 * i] = (cb + 2 * i <= lim) ? f.x : -CUDART_INF_F;
              x[2 * i + 1] = (cb + 2 * i + 1 <= lim) ? f.y : -CUDART_INF_F;
            }
            float cm = x[0];
#pragma unroll
            for (int i = 1; i < 16; ++i) cm = fmaxf(cm, x[i]);
            if (cm > m) { l *= fast_exp2((m - cm) * kLog2e); m = cm; }
            if (m > -CUDART_INF_F) {
              const float ml2 = m * kLog2e;
              float part = 0.f;
#pragma unroll
              for (int i = 0; i < 16; ++i) part += fast_exp2(fmaf(x[i], kLog2e, -ml2));
              l += part;
            }
          }
        }
      }
      __syncwarp();
      {
        // the warp writes its own 32 rows x 128 bytes: 4 rows per instruction, 8 lanes x 16 bytes per row (whole sectors)
        const int r4 = lane >> 3, piece = lane & 7, col = col0 + piece * 8;
        const uint8_t* src = stage + r4 * kStagePitch + piece * 16;
        const int grow0 = m0 + quarter * 32 + r4;
        __half* dst = lbase + static_cast<int64_t>(grow0) * p.ls_q + n0 + col;
        if (p.debug & 1) {
        } else if (vec_ok && m0 + kFT <= p.q_len && n0 + kFT <= p.kv_len) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            __stcs(reinterpret_cast<uint4*>(dst + static_cast<int64_t>(it) * 4 * p.ls_q),
                   *reinterpret_cast<const uint4*>(src + it * 4 * kStagePitch));
        } else {
#pragma unroll 1
          for (int it = 0; it < 8; ++it) {
            if (grow0 + it * 4 >= p.q_len || n0 + col >= p.kv_len) continue;
            const uint4 val = *reinterpret_cast<const uint4*>(src + it * 4 * kStagePitch);
            __half* d = dst + static_cast<int64_t>(it) * 4 * p.ls_q;
            if (vec_ok && n0 + col + 8 <= p.kv_len) {
              __stcs(reinterpret_cast<uint4*>(d), val);
            } else {
              const __half* hv = reinterpret_cast<const __half*>(&val);
              for (int i = 0; i < 8; ++i)
                if (n0 + col + i < p.kv_len) d[i] = hv[i];
            }
          }
        }
      }
      __syncwarp();
    }
    // merge the statistics: the two column halves of a row inside the group (through the dead staging tile), then the
    // two groups; the staging region is released to the V loader in between
    {
      float2* xch = reinterpret_cast<float2*>(sV + 16 * kStageBytes + g * 1024);   // past every warp's staging, inside P[0]
      if (half == 1) xch[row] = make_float2(m, l);
      group_sync(1 + g, kGroupThreads);
      if (half == 0) {
        const float2 o2 = xch[row];
        const float M = fmaxf(m, o2.x);
        if (M > -CUDART_INF_F) { l = l * fast_exp2((m - M) * kLog2e) + o2.y * fast_exp2((o2.x - M) * kLog2e); m = M; }
        sStats[g * kFT + row] = make_float2(m, l);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[A_DONE]);
      group_sync(3, 2 * kGroupThreads);
      const float2 s0 = sStats[row], s1 = sStats[kFT + row];
      const float M = fmaxf(s0.x, s1.x);
      if (M > -CUDART_INF_F) {
        l = s0.y * fast_exp2((s0.x - M) * kLog2e) + s1.y * fast_exp2((s1.x - M) * kLog2e);
        m = M;
      } else {
        m = 0.f; l = 1.f;                                                  // fully masked row (generic mask only)
      }
    }
    // p = exp2(x * log2e - sub): 1 / L folded into the exponent; padding rows get p = 0 (nothing for the column sums)
    const float sub_l2 = row_ok ? fmaf(m, kLog2e, __log2f(l)) : CUDART_INF_F;

    // ---- phase B: P = exp(S - M) / L -> shared memory -> O += P.V, column sums ----
    uint8_t* pimg = sP + g * kImgBytes;
    float* imp = p.importance ? p.importance + (static_cast<int64_t>(b) * p.H + h) * p.imp_stride : nullptr;
    uint32_t puse = 0;
    for (int j = (g + nt) & 1; j < nb; j += 2, ++puse) {
      const int n0 = j * kFT;
      drain_s();
      const int lim = row_lim - n0 - col0;
      if (lim >= 63 && mrow == nullptr) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float2 f = __half22float2(lg[i]);
          lg[i] = __floats2half2_rn(fast_exp2(fmaf(f.x, kLog2e, -sub_l2)), fast_exp2(fmaf(f.y, kLog2e, -sub_l2)));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          __half2 t = lg[i];
          if (mrow) t = __hadd2_rn(t, mask2(mrow + n0 + col0, i, lim));
          const float2 f = __half22float2(t);
          const float p0 = 2 * i <= lim ? fast_exp2(fmaf(f.x, kLog2e, -sub_l2)) : 0.f;
          const float p1 = 2 * i + 1 <= lim ? fast_exp2(fmaf(f.y, kLog2e, -sub_l2)) : 0.f;
          lg[i] = __floats2half2_rn(p0, p1);
        }
      }
      mbar_wait_bounded(&bars[P_EMPTY + g], (puse & 1) ^ 1);               // P.V of this group's previous tile has read the buffer
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<uint4*>(pimg + img_off(row, half * 8 + i)) = *reinterpret_cast<const uint4*>(&lg[4 * i]);
      fence_proxy_async();                                                 // generic-proxy writes -> visible to the tensor core
      mbar_arrive(&bars[P_FULL + g]);
      if (imp) {
        group_sync(1 + g, kGroupThreads);
        // column sums of the tile just written: thread (chunk c, part hp, sub) adds rows sub + 8i, i in [8hp, 8hp + 8), of
        // the 8 columns of chunk c; the 8 x 2 lanes that share a chunk then reduce-scatter
        const int tg = wg * 32 + lane, c = tg >> 4, hp = (tg >> 3) & 1, sub = tg & 7;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        const __half2 one = __floats2half2_rn(1.f, 1.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 x = *reinterpret_cast<const uint4*>(pimg + (hp * 8 + i) * 2048 + c * 128 + sub * 16);
          const __half2* h2 = reinterpret_cast<const __half2*>(&x);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[2 * e] = fhfma_lo(h2[e], one, acc[2 * e]);
            acc[2 * e + 1] = fhfma_hi(h2[e], one, acc[2 * e + 1]);
          }
        }
        float a4[4], a2[2];
        {
          const bool hi = sub & 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float send = hi ? acc[i] : acc[i + 4];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
            a4[i] = (hi ? acc[i + 4] : acc[i]) + recv;
          }
        }
        {
          const bool hi = sub & 2;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float send = hi ? a4[i] : a4[i + 2];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
            a2[i] = (hi ? a4[i + 2] : a4[i]) + recv;
          }
        }
        const bool hi1 = sub & 1;
        const float send = hi1 ? a2[0] : a2[1];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
        float total = (hi1 ? a2[1] : a2[0]) + recv;                        // column c*8 + sub over this lane's 64 rows
        total += __shfl_xor_sync(0xffffffffu, total, 8);                   // + the other 64 rows
        const int col = n0 + c * 8 + sub;
        if (hp == 0 && col < p.kv_len && total != 0.f) atomicAdd(imp + col, total);
        group_sync(1 + g, kGroupThreads);                                  // nobody overwrites P[g] while a neighbour still sums it
      }
    }

    // ---- O: TMEM -> fp16 -> out[b, q, h, :]; thread (group g, half) writes head_dim columns [32(2g + half), +32) ----
    mbar_wait_bounded(&bars[O_FULL], 0);
    tc_fence_after();
    const int oc = (g * 2 + half) * 32;
    float v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(quarter * 32) << 16) + 2 * kFT + oc, v);
    if (row_ok) {
      __half* dst = p.out + ((static_cast<int64_t>(b) * p.q_len + row_g) * p.H + h) * kFD + oc;
      __align__(16) __half2 o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(o)[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kLoadWarp) tmem_dealloc(tmem, 512);
}

int64_t prefill_fused_scratch_bytes(int batch, int H, int H_kv, int q_len, int kv_len) {
  const int64_t mt = (q_len + kFT - 1) / kFT, nt = (kv_len + kFT - 1) / kFT;
  return 1024 + (static_cast<int64_t>(batch) * H * mt + 2 * static_cast<int64_t>(batch) * H_kv * nt) * kImgBytes;
}

// Launches pack (+ mask classification) + the fused kernel.  `scratch` holds the operand images.
int launch_prefill_fused(const SpattenPrefillArgs* a, cudaStream_t st) {
  const int mt = (a->q_len + kFT - 1) / kFT, nt = (a->kv_len + kFT - 1) / kFT;
  uint8_t* ws = static_cast<uint8_t*>(a->probs);
  PackParams pk{};
  pk.q = static_cast<const __half*>(a->q);
  pk.k_cache = static_cast<const __half*>(a->k_cache);
  pk.v_cache = static_cast<const __half*>(a->v_cache);
  pk.stride_b = a->cache_stride_b; pk.stride_h = a->cache_stride_h;
  pk.rope = static_cast<const __half*>(a->rope); pk.q_pos = a->q_pos;
  pk.mode_flag = reinterpret_cast<int*>(ws);
  pk.q_img = ws + 1024;
  pk.k_img = pk.q_img + static_cast<int64_t>(a->batch) * a->num_heads * mt * kImgBytes;
  pk.v_img = pk.k_img + static_cast<int64_t>(a->batch) * a->num_kv_heads * nt * kImgBytes;
  pk.batch = a->batch; pk.H = a->num_heads; pk.H_kv = a->num_kv_heads; pk.q_len = a->q_len; pk.kv_len = a->kv_len;
  pk.rope_len = a->rope_len; pk.mt = mt; pk.nt = nt;

  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(prefill_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(prefill_fused_kernel)");
    configured[dev] = true;
  }
  prefill_pack_kernel<<<dim3(std::max(mt, nt), a->batch * (a->num_heads + 2 * a->num_kv_heads)), 128, 0, st>>>(pk);
  if (a->mask != nullptr) {
    const int64_t total = static_cast<int64_t>(a->batch) * a->q_len * a->kv_len;
    mask_classify_kernel<<<static_cast<unsigned>((total + 2047) / 2048), 256, 0, st>>>(static_cast<const __half*>(a->mask), total,
                                                                                      a->q_len, a->kv_len, pk.mode_flag);
  }
  FusedParams p{};
  p.q_img = pk.q_img; p.k_img = pk.k_img; p.v_img = pk.v_img; p.mode_flag = pk.mode_flag;
  p.mask = static_cast<const __half*>(a->mask);
  p.logits = static_cast<__half*>(a->logits);
  p.ls_b = a->logits_stride_b; p.ls_h = a->logits_stride_h; p.ls_q = a->logits_stride_q;
  p.out = static_cast<__half*>(a->out);
  p.importance = a->importance; p.imp_stride = a->importance_stride;
  p.batch = a->batch; p.H = a->num_heads; p.n_rep = a->num_heads / a->num_kv_heads; p.q_len = a->q_len; p.kv_len = a->kv_len;
  p.mt = mt; p.nt = nt;
  p.mode = a->mask != nullptr ? 0 : ((a->flags & SPATTEN_PREFILL_CAUSAL) ? 1 : 2);
  p.debug = static_cast<int>((a->flags >> 8) & 7);
  p.rs_up = nextafterf(1.0f / sqrtf(static_cast<float>(kFD)), 1.0f);
  prefill_fused_kernel<<<dim3(mt, a->batch * a->num_heads), kFusedThreads, kFusedSmem, st>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "spatten_prefill_attn (fused) launch");
}

}  // namespace spatten
