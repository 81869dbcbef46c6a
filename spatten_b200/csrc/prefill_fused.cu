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
// shared memory: [Q slot 0][Q slot 1 (phase A) = P buffer 0 (phase B)][K ring x2][V ring x2 | P buffer 1]; phase A's
// warp-private logit staging and the statistics exchange live in the (then idle) V ring + P buffer 1
constexpr int kOffQ = 0, kOffP0 = kImgBytes, kOffK = 2 * kImgBytes, kOffV = 4 * kImgBytes, kOffP1 = 6 * kImgBytes, kOffBars = 7 * kImgBytes;
constexpr int kOffStats = kOffBars + 256, kOffTmem = kOffStats + 2 * kFT * 8;
constexpr int kFusedSmem = kOffTmem + 16;
static_assert(16 * kStageBytes + 2048 <= 3 * kImgBytes, "phase-A staging + statistics exchange must fit in the V ring + P buffer 1");
static_assert(kFusedSmem <= 232448, "shared memory budget");

enum FusedBar { Q_FULL = 0, Q_EMPTY = 1, K_FULL = 2, K_EMPTY = 4, V_FULL = 6, V_EMPTY = 8, S_FULL = 10, S_EMPTY = 14, P_FULL = 18,
                P_EMPTY = 20, O_FULL = 22, O_EMPTY = 23, A_DONE = 24, NUM_BARS = 25 };

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
// causal mask (HF `_make_causal_mask` fills with finfo.min = -65504).  A thread checks 8 consecutive elements; chunks that
// lie wholly on one side of the diagonal are tested on the raw bit patterns (two packed compares per 4 bytes), so the
// kernel runs at memory speed.
__global__ void __launch_bounds__(256) mask_classify_kernel(const __half* mask, int64_t total, int q_len, int kv_len, int* flag) {
  const int64_t e0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (e0 >= total) return;
  const int off = kv_len - q_len;
  const int64_t r = e0 / kv_len;
  int j = static_cast<int>(e0 - r * kv_len), qi = static_cast<int>(r % q_len);
  bool ok = true;
  const bool vec = e0 + 8 <= total && (reinterpret_cast<uintptr_t>(mask + e0) & 15) == 0;
  if (vec && j + 8 <= kv_len && (j + 7 <= qi + off || j > qi + off)) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(mask + e0));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    if (j + 7 <= qi + off) {                                    // visible: +0 or -0
      ok = (((w[0] | w[1] | w[2] | w[3]) & 0x7fff7fffu) == 0);
    } else {                                                    // masked: fp16 bits in [0xFB53 (-60000), 0xFC00 (-inf)]
#pragma unroll
      for (int i = 0; i < 4; ++i)
        ok = ok && (__vcmpgeu2(w[i], 0xFB53FB53u) & __vcmpleu2(w[i], 0xFC00FC00u)) == 0xffffffffu;
    }
  } else {
    for (int i = 0; i < 8 && e0 + i < total; ++i) {
      const float m = __half2float(mask[e0 + i]);
      ok = ok && (j <= qi + off ? m == 0.f : m <= -60000.f);
      if (++j == kv_len) { j = 0; if (++qi == q_len) qi = 0; }
    }
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  float2* sStats = reinterpret_cast<float2*>(smem + kOffStats);
  uint32_t* sTmem = reinterpret_cast<uint32_t*>(smem + kOffTmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A CTA owns TWO query tiles of one head, the heaviest remaining one and the lightest: every K tile of phase A is
  // fetched once and multiplied with both (half the L2 -> shared-memory traffic of the logit pass), and under the causal
  // mask every CTA has the same phase-B work (nb[0] + nb[1] = nt + 1 tiles).
  const int qt0 = p.mt - 1 - static_cast<int>(blockIdx.x), qt1 = static_cast<int>(blockIdx.x);
  const bool two = qt1 < qt0;                                              // odd tile count: the middle tile is alone
  const int bh = blockIdx.y, b = bh / p.H, h = bh % p.H, kvh = h / p.n_rep;
  const int off = p.kv_len - p.q_len;

  if (warp == kLoadWarp) tmem_alloc(sTmem, 512);                           // phase A: S[2 tiles][2]; phase B: S[0], O, S[2]
  if (tid == kMmaWarp * 32) {
    mbar_init(&bars[Q_FULL], 1); mbar_init(&bars[Q_EMPTY], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[K_FULL + i], 1);  mbar_init(&bars[K_EMPTY + i], 1);
      mbar_init(&bars[V_FULL + i], 1);  mbar_init(&bars[V_EMPTY + i], 1);
      mbar_init(&bars[P_FULL + i], kGroupThreads); mbar_init(&bars[P_EMPTY + i], 1);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&bars[S_FULL + i], 1); mbar_init(&bars[S_EMPTY + i], kGroupThreads); }
    mbar_init(&bars[O_FULL], 1); mbar_init(&bars[O_EMPTY], 2 * kGroupThreads);
    mbar_init(&bars[A_DONE], 2 * kGroupThreads / 32);
    fence_mbar_init();
  }
  int mode = p.mode;
  if (mode == 0) mode = (*p.mode_flag != 0) ? 1 : 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *sTmem;

  const int nt = p.nt;
  // key tiles the mask leaves alive, per query tile
  auto live_tiles = [&](int qt) { return mode == 1 ? min(nt, (min(qt * kFT + kFT - 1, p.q_len - 1) + off) / kFT + 1) : nt; };
  const int nl0 = live_tiles(qt0), nl1 = two ? live_tiles(qt1) : 0;
  const int nb0 = (p.debug & 2) ? 0 : nl0, nb1 = (p.debug & 2) ? 0 : nl1;   // phase-B tiles
  const int nbt = nb0 + nb1;

  if (warp == kLoadWarp) {
    // ===== loaders: contiguous 32 KiB images through the TMA engine; lane 0 feeds Q and the K ring, lane 1 the V ring =====
    const uint64_t keep = policy_evict_last();                             // K / V images are re-read by every query tile of the head
    const int64_t kv_img0 = (static_cast<int64_t>(b) * (p.H / p.n_rep) + kvh) * nt * kImgBytes;
    if (lane == 0) {
      const uint8_t* kimg = p.k_img + kv_img0;
      const uint8_t* qimg = p.q_img + static_cast<int64_t>(bh) * p.mt * kImgBytes;
      const uint64_t once = policy_evict_first();
      mbar_expect_tx(&bars[Q_FULL], kImgBytes * (two ? 2 : 1));
      bulk_g2s(sQ, qimg + static_cast<int64_t>(qt0) * kImgBytes, kImgBytes, &bars[Q_FULL], once);
      if (two) bulk_g2s(sQ + kImgBytes, qimg + static_cast<int64_t>(qt1) * kImgBytes, kImgBytes, &bars[Q_FULL], once);
      for (int kuse = 0; kuse < nt + nbt; ++kuse) {
        const int s = kuse & 1;
        const int j = kuse < nt ? kuse : (kuse - nt < nb0 ? kuse - nt : kuse - nt - nb0);
        if (kuse == nt + nb0 && nb1 > 0) {                                 // second query tile of phase B: reload it into slot 0
          mbar_wait_bounded(&bars[Q_EMPTY], 0);
          mbar_expect_tx(&bars[Q_FULL], kImgBytes);
          bulk_g2s(sQ, qimg + static_cast<int64_t>(qt1) * kImgBytes, kImgBytes, &bars[Q_FULL], once);
        }
        mbar_wait_bounded(&bars[K_EMPTY + s], ((kuse >> 1) & 1) ^ 1);
        mbar_expect_tx(&bars[K_FULL + s], kImgBytes);
        bulk_g2s(sK + s * kImgBytes, kimg + static_cast<int64_t>(j) * kImgBytes, kImgBytes, &bars[K_FULL + s], keep);
      }
    } else if (lane == 1) {
      const uint8_t* vimg = p.v_img + kv_img0;
      while (!mbar_try_wait(&bars[A_DONE], 0)) __nanosleep(500);           // the V ring doubles as phase A's logits staging
      for (int t = 0; t < nbt; ++t) {
        const int vs = t & 1, j = t < nb0 ? t : t - nb0;
        mbar_wait_bounded(&bars[V_EMPTY + vs], ((t >> 1) & 1) ^ 1);
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
      const uint32_t dP0 = (smem_u32(smem + kOffP0) >> 4) | lbo_kmajor, dP1 = (smem_u32(smem + kOffP1) >> 4) | lbo_kmajor;
      const uint32_t dV = (smem_u32(sV) >> 4) | lbo_mnmajor;
      const long long t0 = clock64();
      auto spin_guard = [&]() { if (clock64() - t0 > 4000000000ll) __trap(); };
      mbar_wait_bounded(&bars[Q_FULL], 0);
      uint32_t kuse = 0;
      uint32_t sbits = 0;                                                   // bit sb = parity of the next use of S buffer sb
      auto qk = [&](int qslot, int sb) {                                    // S[sb] = Q[qslot] . K[kuse & 1]^T
        const int s = kuse & 1;
        const uint32_t d = tmem + sb * kFT, aq = dQ + qslot * (kImgBytes >> 4), bk = dK + s * (kImgBytes >> 4);
#pragma unroll
        for (int ks = 0; ks < kFD / 16; ++ks) umma_f16_split(d, aq + ks * 16, hi_kmajor, bk + ks * 16, hi_kmajor, idesc_qk, ks > 0);
        umma_commit(&bars[S_FULL + sb]);
      };
      // ---- phase A: every K tile against both query tiles ----
      for (int j = 0; j < nt; ++j, ++kuse) {
        const int s = kuse & 1, pp = j & 1;
        mbar_wait_bounded(&bars[K_FULL + s], (kuse >> 1) & 1);
#pragma unroll 1
        for (int g = 0; g < (two ? 2 : 1); ++g) {
          const int sb = g * 2 + pp;
          mbar_wait_bounded(&bars[S_EMPTY + sb], ((sbits >> sb) & 1) ^ 1);
          tc_fence_after();
          qk(g, sb);
          sbits ^= 1u << sb;
        }
        umma_commit(&bars[K_EMPTY + s]);
      }
      // ---- phase B: one query tile at a time (always in Q slot 0); tile t of the CTA goes to group t & 1 ----
      int tbase = 0;
#pragma unroll 1
      for (int qi = 0; qi < 2; ++qi) {
        const int nb = qi ? nb1 : nb0;
        if (nb == 0) continue;
        if (qi == 1) { mbar_wait_bounded(&bars[Q_FULL], 1); }
        int qn = 0, pn = 0;
        while (pn < nb) {
          if (qn < nb) {                                                    // next S tile: needs its K image and a drained S buffer
            const int t = tbase + qn, g = t & 1, sb = g * 2, s = kuse & 1;
            if (mbar_try_wait(&bars[K_FULL + s], (kuse >> 1) & 1) && mbar_try_wait(&bars[S_EMPTY + sb], ((sbits >> sb) & 1) ^ 1)) {
              tc_fence_after();
              qk(0, sb);
              umma_commit(&bars[K_EMPTY + s]);
              sbits ^= 1u << sb; ++kuse; ++qn;
              if (qn == nb && qi == 0 && nb1 > 0) umma_commit(&bars[Q_EMPTY]);   // slot 0 may take the second query tile
              continue;
            }
          }
          if (pn < qn) {                                                    // next P.V: needs the group's P tile and the V image
            const int t = tbase + pn, g = t & 1, vs = t & 1;
            if (mbar_try_wait(&bars[P_FULL + g], (t >> 1) & 1) && mbar_try_wait(&bars[V_FULL + vs], (t >> 1) & 1)) {
              if (pn == 0 && qi == 1 && nb0 > 0) mbar_wait_bounded(&bars[O_EMPTY], 0);   // O of the first tile has been read
              tc_fence_after();
              const uint32_t d = tmem + 1 * kFT, ap = g ? dP1 : dP0, bv = dV + vs * (kImgBytes >> 4);
              if (pn == 0) {
#pragma unroll
                for (int ks = 0; ks < kFT / 16; ++ks) umma_f16_split(d, ap + ks * 16, hi_kmajor, bv + ks * 256, hi_mnmajor, idesc_pv, ks > 0);
              } else {
#pragma unroll
                for (int ks = 0; ks < kFT / 16; ++ks) umma_f16_split(d, ap + ks * 16, hi_kmajor, bv + ks * 256, hi_mnmajor, idesc_pv, true);
              }
              umma_commit(&bars[P_EMPTY + g]);
              umma_commit(&bars[V_EMPTY + vs]);
              ++pn;
              continue;
            }
          }
          spin_guard();
        }
        umma_commit(&bars[O_FULL]);
        tbase += nb;
      }
    }
    __syncwarp();
  } else {
    // ===== softmax groups: two threads per query row (TMEM lane), 64 key columns each.  Phase A: group g owns query
    // tile g.  Phase B: the groups alternate key tiles of the current query tile. =====
    const int g = warp >> 3, quarter = warp & 3, half = (warp >> 2) & 1, row = quarter * 32 + lane;
    const int col0 = half * 64;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    uint8_t* stage = sV + warp * kStageBytes;                              // warp-private: rows of this TMEM quarter, this column half
    __half* lbase = p.logits + b * p.ls_b + h * p.ls_h;
    const bool vec_ok = (p.ls_q % 8) == 0 && ((reinterpret_cast<uintptr_t>(lbase) & 15) == 0);
    uint32_t sbits = 0;                                                    // bit pp = parity of the next use of this group's S buffer pp
    __half2 lg[32];                                                        // this thread's 64 logits of the current tile

    // S[sb] -> fp16 logits in registers, then the buffer goes straight back to the MMA issuer
    auto drain_s = [&](int pp) {
      const int sb = g * 2 + pp;
      mbar_wait_bounded(&bars[S_FULL + sb], (sbits >> pp) & 1);
      sbits ^= 1u << pp;
      tc_fence_after();
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        float v[32];
        tmem_ld32(t_lane + sb * kFT + col0 + c2 * 32, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) lg[c2 * 16 + i] = logit2(v[2 * i], v[2 * i + 1], p.rs_up);
      }
      tc_fence_before();
      mbar_arrive(&bars[S_EMPTY + sb]);
    };
    // per query tile: first row, this thread's row, last live key of the row, mask row
    auto tile_rows = [&](int qt, int& m0, int& row_g, bool& row_ok, int& row_lim, const __half*& mrow) {
      m0 = qt * kFT;
      row_g = m0 + row;
      row_ok = row_g < p.q_len;
      row_lim = (mode == 1 ? min(row_g + off, p.kv_len - 1) : p.kv_len - 1);
      mrow = (mode == 0 && row_ok) ? p.mask + (static_cast<int64_t>(b) * p.q_len + row_g) * p.kv_len : nullptr;
    };

    // ---- phase A: logits + row statistics of query tile g ----
    float m = -CUDART_INF_F, l = 0.f;
    if (g == 0 || two) {
      int m0, row_g, row_lim; bool row_ok; const __half* mrow;
      tile_rows(g ? qt1 : qt0, m0, row_g, row_ok, row_lim, mrow);
      const int nb_stats = (p.debug & 4) ? 0 : (g ? nl1 : nl0);
      for (int j = 0; j < nt; ++j) {
        const int n0 = j * kFT;
        drain_s(j & 1);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(stage + lane * kStagePitch + i * 16) = *reinterpret_cast<const uint4*>(&lg[4 * i]);
        const int lim = row_lim - n0 - col0;                               // columns 0..lim of this thread's 64 are live
        if (j < nb_stats && row_ok && lim >= 0) {
          if (lim >= 63 && mrow == nullptr) {
            __half2 hm = lg[0];
#pragma unroll
            for (int i = 1; i < 32; ++i) hm = __hmax2(hm, lg[i]);
            const float cm = fmaxf(__low2float(hm), __high2float(hm));
            if (cm > m) { l *= fast_exp2((m - cm) * kLog2e); m = cm; }    // m = -inf: l is 0 and stays 0
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
    }
    // merge the two column halves of every row (through shared memory past the staging tiles) and publish the tile's
    // statistics; the staging region is released to the V loader
    {
      float2* xch = reinterpret_cast<float2*>(sV + 16 * kStageBytes + g * 1024);
      if (half == 1) xch[row] = make_float2(m, l);
      group_sync(1 + g, kGroupThreads);
      if (half == 0) {
        const float2 o2 = xch[row];
        const float M = fmaxf(m, o2.x);
        if (M > -CUDART_INF_F) { l = l * fast_exp2((m - M) * kLog2e) + o2.y * fast_exp2((o2.x - M) * kLog2e); m = M; }
        else { m = 0.f; l = 1.f; }                                         // fully masked row (generic mask only)
        sStats[g * kFT + row] = make_float2(m, l);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[A_DONE]);
      group_sync(3, 2 * kGroupThreads);
    }

    // ---- phase B: P = exp(S - M) / L -> shared memory -> O += P.V, column sums; then O -> out ----
    uint8_t* pimg = smem + (g ? kOffP1 : kOffP0);
    float* imp = p.importance ? p.importance + (static_cast<int64_t>(b) * p.H + h) * p.imp_stride : nullptr;
    int tbase = 0;
#pragma unroll 1
    for (int qi = 0; qi < 2; ++qi) {
      const int nb = qi ? nb1 : nb0;
      if (nb == 0) continue;
      int m0, row_g, row_lim; bool row_ok; const __half* mrow;
      tile_rows(qi ? qt1 : qt0, m0, row_g, row_ok, row_lim, mrow);
      const float2 st = sStats[qi * kFT + row];
      // p = exp2(x * log2e - sub): 1 / L folded into the exponent; padding rows get p = 0 (nothing for the column sums)
      const float sub_l2 = row_ok ? fmaf(st.x, kLog2e, __log2f(st.y)) : CUDART_INF_F;
#pragma unroll 1
      for (int t = tbase + ((tbase ^ g) & 1); t < tbase + nb; t += 2) {     // this group's tiles: t & 1 == g
        const int n0 = (t - tbase) * kFT;
        drain_s(0);
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
            __half2 tt = lg[i];
            if (mrow) tt = __hadd2_rn(tt, mask2(mrow + n0 + col0, i, lim));
            const float2 f = __half22float2(tt);
            const float p0 = 2 * i <= lim ? fast_exp2(fmaf(f.x, kLog2e, -sub_l2)) : 0.f;
            const float p1 = 2 * i + 1 <= lim ? fast_exp2(fmaf(f.y, kLog2e, -sub_l2)) : 0.f;
            lg[i] = __floats2half2_rn(p0, p1);
          }
        }
        mbar_wait_bounded(&bars[P_EMPTY + g], ((t >> 1) & 1) ^ 1);         // P.V of this group's previous tile has read the buffer
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(pimg + img_off(row, half * 8 + i)) = *reinterpret_cast<const uint4*>(&lg[4 * i]);
        fence_proxy_async();                                               // generic-proxy writes -> visible to the tensor core
        mbar_arrive(&bars[P_FULL + g]);
        if (imp) {
          group_sync(1 + g, kGroupThreads);
          // column sums of the tile just written: thread (chunk c, part hp, sub) adds rows sub + 8i, i in [8hp, 8hp + 8), of
          // the 8 columns of chunk c; the 8 x 2 lanes that share a chunk then reduce-scatter
          const int tg = (warp & 7) * 32 + lane, c = tg >> 4, hp = (tg >> 3) & 1, sub = tg & 7;
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
          float total = (hi1 ? a2[1] : a2[0]) + recv;                      // column c*8 + sub over this lane's 64 rows
          total += __shfl_xor_sync(0xffffffffu, total, 8);                 // + the other 64 rows
          const int col = n0 + c * 8 + sub;
          if (hp == 0 && col < p.kv_len && total != 0.f) atomicAdd(imp + col, total);
          group_sync(1 + g, kGroupThreads);                                // nobody overwrites P[g] while a neighbour still sums it
        }
      }
      tbase += nb;

      // O: TMEM -> fp16 -> out[b, q, h, :]; thread (group g, half) writes head_dim columns [32(2g + half), +32)
      mbar_wait_bounded(&bars[O_FULL], (qi == 1 && nb0 > 0) ? 1 : 0);
      tc_fence_after();
      const int oc = (g * 2 + half) * 32;
      float v[32];
      tmem_ld32(t_lane + 1 * kFT + oc, v);
      tc_fence_before();
      mbar_arrive(&bars[O_EMPTY]);
      if (row_ok) {
        __half* dst = p.out + ((static_cast<int64_t>(b) * p.q_len + row_g) * p.H + h) * kFD + oc;
        __align__(16) __half2 o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
#pragma unroll
        for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(o)[i];
      }
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
  prefill_fused_kernel<<<dim3((mt + 1) / 2, a->batch * a->num_heads), kFusedThreads, kFusedSmem, st>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "spatten_prefill_attn (fused) launch");
}

}  // namespace spatten
