// Progressive-quantisation decode attention (BASELINE config 4) -- spatten_kv_quantize / spatten_decode_attn_quant.
//
// RTL stages named by north_star: MSB-first fetch of 4-bit K/V (MatrixFetcher.scala:48-51 profile (4,1),
// :344-345 MSB left-aligned), RequantDecision (RequantDecision.scala:47-72: need_requant = row max < threshold),
// LSB refetch + replay of the same request (SpAttenController.scala:140-175, 214-232, 296-303; :346-347 LSB
// right-aligned).  The reference Python never executes this (README.md:11), so parity is "unpinned": the
// arithmetic is the restatement in oracle/spatten_oracle.py::attn_core_quantized.
//
// HBM layout per (batch, kv head): K and V each as two 4-bit planes u8 [tokens, D/2] (MSB nibble, LSB nibble of an
// offset-128 int8) plus one fp16 scale per token.  Nibble order inside every 8-dim group is
// [d0,d2,d4,d6,d1,d3,d5,d7], so that `(w >> 4i) & 0x000f000f` yields the half2 pair (d_2i, d_2i+1) directly, and the
// sixteen 8-dim groups of a row are ordered [0,1,8,9, 2,3,10,11, ...]: one 16-byte chunk holds 16 dims and their RoPE
// partners (the fetch-friendly interleave plays the role of the RTL's bit-reversed BMR layout, BMR.scala:59-65).
// Pass 1 streams the MSB planes only (K: D/2 bytes per token instead of 2*D) and runs the whole attention on the
// tensor cores (mma.sync m16n8k16: the CUDA cores only unpack nibbles and rotate); rows whose largest softmax
// probability is below the threshold are replayed by pass 2 with MSB+LSB keys.
#include <math_constants.h>

#include <algorithm>
#include <cmath>

#include "../../include/spatten_b200.h"
#include "common.cuh"

namespace spatten {

constexpr int kQWarps = 16;     // 4 per scheduler: a 17th (producer) warp would cost every thread 32 registers
constexpr int kQThreads = kQWarps * 32;
constexpr int kQTile = 16;      // tokens per stage = one MMA tile (M = 16 tokens for Q.K^T, K = 16 tokens for P.V)
constexpr int kQStages = 3;
constexpr int kQMaxPartials = 8192;
constexpr int kRopePitch = 272; // bytes: a 256-byte RoPE row + 16, so the 16-byte reads of two adjacent tokens hit different banks
constexpr int kRopeStage = kQTile * kRopePitch;
constexpr int kRopeStages = 6;
constexpr float kLog2eQ = 1.4426950408889634f;
constexpr int kQMergeMaxSplits = 512;

struct QuantPlanes {
  const uint8_t* k_msb; const uint8_t* k_lsb; const __half* k_scale;
  const uint8_t* v_msb; const uint8_t* v_lsb; const __half* v_scale;
  int64_t plane_stride_b, plane_stride_h;   // bytes
  int64_t scale_stride_b, scale_stride_h;   // elements
};

struct QuantDecodeParams {
  QuantPlanes kv;
  const __half* q;
  const __half* rope;
  const int64_t* q_pos;
  __half* out;
  __half* logits;
  int64_t logits_stride;
  float* part;            // [rows][n_splits][D]
  float2* part_ml;        // [rows][n_splits]
  uint8_t* need;          // [rows] pass 1 writes (max prob < threshold); pass 2 reads
  int* counters;          // [rows] tickets, zero between launches
  int H, n_rep, kv_len, rope_len, v_bits;
  int rows, rpc, spc, n_rg, n_splits, tok_per_range;
  float rs_up, threshold;
};

// (a & m) | c in ONE LOP3: ptxas only folds one immediate into a LOP3, so mask and exponent pattern are kept in registers
// (DeqConst, made opaque to constant propagation) -- otherwise every unpacked pair costs two LOP3s.
struct DeqConst {
  uint32_t m_lo, m_hi, one_k;       // 0x000f000f, 0x00f000f0, 0x64006400 (fp16 1024 | 1024)
  __device__ __forceinline__ void init() {
    asm("mov.b32 %0, 0x000f000f;" : "=r"(m_lo));
    asm("mov.b32 %0, 0x00f000f0;" : "=r"(m_hi));
    asm("mov.b32 %0, 0x64006400;" : "=r"(one_k));
  }
};
__device__ __forceinline__ __half2 and_or(uint32_t a, uint32_t m, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(m), "r"(c));
  return *reinterpret_cast<__half2*>(&d);
}

// 8 values (one u32 of the MSB plane, optionally one of the LSB plane) -> 4 half2 holding the exact integers
// (16*m + l) - 128, or the MSB bucket midpoint (16*m + 8) - 128 when the LSB plane is not fetched.  half2 #i takes
// nibbles i and i + 4 of the word.  A nibble OR-ed into mantissa bits 4..7 of the fp16 pattern 0x6400 (= 1024) reads as
// 1024 + 16*m, so the MSB code 16*m needs no multiply: one shift (none for i = 1), one LOP3, one HADD2 per pair, all exact.
template <bool LSB>
__device__ __forceinline__ void deq8(uint32_t wm, uint32_t wl, __half2 (&x)[4], const DeqConst& k) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t sh = i == 0 ? (wm << 4) : (i == 1 ? wm : (wm >> (4 * i - 4)));
    const __half2 bm = and_or(sh, k.m_hi, k.one_k);                              // (1024 + 16 m_a, 1024 + 16 m_b)
    if (LSB) {
      const __half2 bl = and_or(wl >> (4 * i), k.m_lo, k.one_k);                 // (1024 + l_a, 1024 + l_b)
      x[i] = __hadd2(__hadd2(bm, __float2half2_rn(-1152.f)), __hadd2(bl, __float2half2_rn(-1024.f)));   // 16 m - 128 + l, exact
    } else {
      x[i] = __hadd2(bm, __float2half2_rn(-1144.f));                             // 16 m + 8 - 128
    }
  }
}

__device__ __forceinline__ uint64_t q_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// D(16x8, fp32) += A(16x16, fp16, row) . B(16x8, fp16, col) -- warp-level tensor-core MMA (SASS HMMA)
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mbar_arrive_q(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- pass kernel ---------------------------------------------------------------------------------------------------
// CTA = (group of `rpc` rows) x (one token range); the 16 compute warps are rpc rows x spc interleaved tile subsets of the
// SAME range, so the RoPE rows of a tile (16 tokens x 256 B, twice the bytes of the int4 K+V tile itself) are fetched
// once per CTA (by warp 0, on the side) into a shared ring.  Each compute warp streams its own K/V planes with bulk copies
// and does both matrix products on the tensor cores with the single query row broadcast over the N = 8 columns:
//   S (16 tokens x 8) = Krot (16 tokens x 128, A operand, dequantised + rotated in registers) . q (B operand)
//   O (16 dims x 8)  += V^T (16 dims x 16 tokens, A operand: token pairs merged with byte permutes) . p (B operand)
// Lane (r = lane / 4, kq = lane % 4): keys of tokens r, r + 8, dims [16 kq, 16 kq + 16) and their RoPE partners + 64;
// values of tokens 2 kq + {0, 1, 8, 9}, dims [16 r, 16 r + 16).  The CUDA cores only unpack nibbles and rotate.
template <bool LSB, bool V8>
__global__ void __launch_bounds__(kQThreads, 1) decode_attn_q4_kernel(const QuantDecodeParams p) {
  constexpr int D = 128;
  constexpr int kPlaneRow = D / 2;                      // 64 bytes per token per plane
  constexpr int kPlaneTile = kQTile * kPlaneRow;        // 1 KiB
  constexpr int kOffKl = kPlaneTile, kOffVm = kPlaneTile * (LSB ? 2 : 1), kOffVl = kOffVm + kPlaneTile;
  constexpr int kOffSc = kOffVm + kPlaneTile * (V8 ? 2 : 1);
  constexpr int kStageBytes = kOffSc + 64;              // + k scales (32 B) + v scales (32 B)
  constexpr int kWarpSmem = kQStages * kStageBytes;
  extern __shared__ __align__(128) uint8_t smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* rope_sm = smem + kQWarps * kWarpSmem;
  uint64_t* rope_full = reinterpret_cast<uint64_t*>(rope_sm + kRopeStages * kRopeStage);
  uint64_t* rope_empty = rope_full + kRopeStages;
  uint64_t* kv_bars = rope_empty + kRopeStages;

  const int rg = blockIdx.x % p.n_rg, range = blockIdx.x / p.n_rg;
  const int T0 = range * p.tok_per_range, T1 = min(T0 + p.tok_per_range, p.kv_len);
  const int ntiles = (T1 - T0 + kQTile - 1) / kQTile;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRopeStages; ++s) { mbar_init(&rope_full[s], 1); mbar_init(&rope_empty[s], kQWarps); }
    for (int s = 0; s < kQWarps * kQStages; ++s) mbar_init(&kv_bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const int r = lane >> 2, kq = lane & 3;
  uint8_t* wsm = smem + warp * kWarpSmem;
  uint64_t* bars = kv_bars + warp * kQStages;
  const int row = rg * p.rpc + (warp % p.rpc), sub = warp / p.rpc;
  const int split = range * p.spc + sub;
  const bool row_ok = row < p.rows;
  const int bi = row_ok ? row / p.H : 0, h = row_ok ? row % p.H : 0, kvh = h / p.n_rep;
  const bool live = row_ok && (!LSB || p.need[row] != 0);      // pass 2 replays only the rows flagged by the requant decision
  const int my_tiles = live ? (ntiles - sub + p.spc - 1) / p.spc : 0;       // tiles sub, sub + spc, ...
  if (!__syncthreads_or(live ? 1 : 0)) return;                 // pass 2: none of this CTA's rows is replayed

  const int64_t poff = bi * p.kv.plane_stride_b + kvh * p.kv.plane_stride_h;
  const int64_t soff = bi * p.kv.scale_stride_b + kvh * p.kv.scale_stride_h;
  const uint64_t pol = policy_evict_first();
  auto issue = [&](int k) {                                     // k-th tile of this warp
    const int s = k % kQStages;
    const int tok = T0 + (sub + k * p.spc) * kQTile;
    const int n = min(kQTile, T1 - tok);
    const uint32_t pb = static_cast<uint32_t>(n) * kPlaneRow;
    const uint32_t sb = static_cast<uint32_t>((n + 7) / 8 * 8) * 2;     // scales: 16-byte granules (buffers are padded)
    uint8_t* st = wsm + s * kStageBytes;
    mbar_expect_tx(&bars[s], pb * (2 + (LSB ? 1 : 0) + (V8 ? 1 : 0)) + 2 * sb);
    bulk_g2s(st, p.kv.k_msb + poff + static_cast<int64_t>(tok) * kPlaneRow, pb, &bars[s], pol);
    if (LSB) bulk_g2s(st + kOffKl, p.kv.k_lsb + poff + static_cast<int64_t>(tok) * kPlaneRow, pb, &bars[s], pol);
    bulk_g2s(st + kOffVm, p.kv.v_msb + poff + static_cast<int64_t>(tok) * kPlaneRow, pb, &bars[s], pol);
    if (V8) bulk_g2s(st + kOffVl, p.kv.v_lsb + poff + static_cast<int64_t>(tok) * kPlaneRow, pb, &bars[s], pol);
    bulk_g2s(st + kOffSc, p.kv.k_scale + soff + tok, sb, &bars[s], pol);
    bulk_g2s(st + kOffSc + 32, p.kv.v_scale + soff + tok, sb, &bars[s], pol);
  };
  if (lane == 0) {
#pragma unroll 1
    for (int k = 0; k < min(my_tiles, kQStages); ++k) issue(k);
  }

  // query B fragments: fp16 RoPE at its position; lane kq holds dims 16 kq + 2 s (+1) and their partners + 64, s = 0..7
  uint32_t qlo[8], qhi[8];
  {
    int64_t pos = p.q_pos ? p.q_pos[bi] : p.kv_len - 1;
    pos = max(static_cast<int64_t>(0), min(pos, static_cast<int64_t>(p.rope_len - 1)));
    const __half* qr = p.q + static_cast<int64_t>(row_ok ? row : 0) * D;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const uint4 a = *reinterpret_cast<const uint4*>(qr + kq * 16 + hf * 8), b = *reinterpret_cast<const uint4*>(qr + 64 + kq * 16 + hf * 8);
      const uint4 cs = __ldg(reinterpret_cast<const uint4*>(p.rope + pos * D + kq * 16 + hf * 8));
      const uint4 sn = __ldg(reinterpret_cast<const uint4*>(p.rope + pos * D + 64 + kq * 16 + hf * 8));
      const __half2* a2 = reinterpret_cast<const __half2*>(&a); const __half2* b2 = reinterpret_cast<const __half2*>(&b);
      const __half2* c2 = reinterpret_cast<const __half2*>(&cs); const __half2* s2 = reinterpret_cast<const __half2*>(&sn);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        qlo[hf * 4 + i] = h2u(__hsub2_rn(__hmul2_rn(a2[i], c2[i]), __hmul2_rn(b2[i], s2[i])));
        qhi[hf * 4 + i] = h2u(__hadd2_rn(__hmul2_rn(b2[i], c2[i]), __hmul2_rn(a2[i], s2[i])));
      }
    }
  }
  DeqConst dk;
  dk.init();
  float m = -CUDART_INF_F, l = 0.f;
  float acc[8][4];                      // P.V MMA mi: c0 = c1 = out[16 r + 2 mi], c2 = c3 = out[16 r + 2 mi + 1]
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  __half* lg = p.logits ? p.logits + static_cast<int64_t>(row_ok ? row : 0) * p.logits_stride : nullptr;
  // values: the lane reads 8 bytes (dims 16 r ..) of tokens 2 kq + {0, 1, 8, 9}; odd r starts with the odd token so that
  // the two half-rows of a 128-byte line are read together (the byte-permute selectors undo it for free)
  const int v_off = (r & 3) * 16 + (r >> 2) * 8;
  const int vt0 = 2 * kq + (r & 1), vt1 = 2 * kq + 1 - (r & 1);
  const uint32_t sel_lo = (r & 1) ? 0x1054u : 0x5410u, sel_hi = (r & 1) ? 0x3276u : 0x7632u;

  int k = 0, rope_issued = 0;
  const uint64_t rope_pol = q_policy_evict_last();
#pragma unroll 1
  for (int i = 0; i < ntiles; ++i) {
    const int rs = i % kRopeStages;
    const uint32_t rpar = (i / kRopeStages) & 1;
    if (warp == 0) {
      // RoPE-row ring, fed by warp 0 on the side: one 256-byte bulk copy per token (lane = token) into the padded stage.
      // Tile i itself is waited for; tiles i+1 .. i+stages-1 are issued only if their stage is already free.
      while (rope_issued < ntiles && rope_issued < i + kRopeStages) {
        const int ps = rope_issued % kRopeStages;
        const uint32_t ppar = ((rope_issued / kRopeStages) & 1) ^ 1;
        int ok = 1;
        if (rope_issued == i) mbar_wait(&rope_empty[ps], ppar);
        else ok = __shfl_sync(0xffffffffu, mbar_try_wait(&rope_empty[ps], ppar) ? 1 : 0, 0);
        if (!ok) break;
        const int tok = T0 + rope_issued * kQTile;
        const int n = min(kQTile, min(T1, p.rope_len) - tok);
        if (lane == 0) mbar_expect_tx(&rope_full[ps], static_cast<uint32_t>(max(n, 0)) * 256u);
        __syncwarp();
        if (lane < n)
          bulk_g2s(rope_sm + ps * kRopeStage + lane * kRopePitch, p.rope + static_cast<int64_t>(tok + lane) * D, 256, &rope_full[ps], rope_pol);
        ++rope_issued;
      }
    }
    if (!live || (i % p.spc) != sub) {                     // not this warp's tile: only keep the shared ring's protocol
      mbar_wait(&rope_full[rs], rpar);
      __syncwarp();
      if (lane == 0) mbar_arrive_q(&rope_empty[rs]);
      continue;
    }
    const int s = k % kQStages;
    const int tok0 = T0 + i * kQTile;
    const int ntok = min(kQTile, T1 - tok0);
    const uint8_t* st = wsm + s * kStageBytes;
    const uint8_t* rp = rope_sm + rs * kRopeStage;
    mbar_wait(&bars[s], (k / kQStages) & 1);
    mbar_wait(&rope_full[rs], rpar);

    // ---- S = Krot . q ----
    float sc4[2][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) sc4[0][e] = sc4[1][e] = 0.f;
    uint32_t rl[2][8], rh[2][8];
#pragma unroll
    for (int tt = 0; tt < 2; ++tt) {
      const int tl = r + tt * 8;
      const uint4 wm = *reinterpret_cast<const uint4*>(st + tl * kPlaneRow + kq * 16);       // lo0 | lo1 | hi0 | hi1
      uint4 wl = make_uint4(0, 0, 0, 0);
      if (LSB) wl = *reinterpret_cast<const uint4*>(st + kOffKl + tl * kPlaneRow + kq * 16);
      __half2 L[8], H[8];
      { __half2 t4[4]; deq8<LSB>(wm.x, wl.x, t4, dk); L[0] = t4[0]; L[1] = t4[1]; L[2] = t4[2]; L[3] = t4[3]; }
      { __half2 t4[4]; deq8<LSB>(wm.y, wl.y, t4, dk); L[4] = t4[0]; L[5] = t4[1]; L[6] = t4[2]; L[7] = t4[3]; }
      { __half2 t4[4]; deq8<LSB>(wm.z, wl.z, t4, dk); H[0] = t4[0]; H[1] = t4[1]; H[2] = t4[2]; H[3] = t4[3]; }
      { __half2 t4[4]; deq8<LSB>(wm.w, wl.w, t4, dk); H[4] = t4[0]; H[5] = t4[1]; H[6] = t4[2]; H[7] = t4[3]; }
      const uint8_t* rr = rp + tl * kRopePitch + kq * 32;
      const uint4 c0 = *reinterpret_cast<const uint4*>(rr), c1 = *reinterpret_cast<const uint4*>(rr + 16);
      const uint4 s0 = *reinterpret_cast<const uint4*>(rr + 128), s1 = *reinterpret_cast<const uint4*>(rr + 144);
      const __half2* cA = reinterpret_cast<const __half2*>(&c0); const __half2* cB = reinterpret_cast<const __half2*>(&c1);
      const __half2* sA = reinterpret_cast<const __half2*>(&s0); const __half2* sB = reinterpret_cast<const __half2*>(&s1);
#pragma unroll
      for (int e = 0; e < 8; ++e) {                         // x*cos + rotate_half(x)*sin on the integer codes, fp16 op for op
        const __half2 c = e < 4 ? cA[e] : cB[e - 4], sn = e < 4 ? sA[e] : sB[e - 4];
        rl[tt][e] = h2u(__hsub2_rn(__hmul2_rn(L[e], c), __hmul2_rn(H[e], sn)));
        rh[tt][e] = h2u(__hadd2_rn(__hmul2_rn(H[e], c), __hmul2_rn(L[e], sn)));
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) mma16816(sc4[e & 1], rl[0][e], rl[1][e], rh[0][e], rh[1][e], qlo[e], qhi[e]);   // two chains

    // ---- logits, online softmax (every lane holds the scores of tokens r and r + 8) ----
    float sc[2], vsc[2];
#pragma unroll
    for (int tt = 0; tt < 2; ++tt) {
      const int tl = r + tt * 8;
      const bool valid = tl < ntok;
      const float d = sc4[0][tt * 2] + sc4[1][tt * 2];
      const float ks = __half2float(*reinterpret_cast<const __half*>(st + kOffSc + tl * 2));
      vsc[tt] = valid ? __half2float(*reinterpret_cast<const __half*>(st + kOffSc + 32 + tl * 2)) : 0.f;
      const __half s16 = __float2half_rn(__half2float(__float2half_rn(d * ks)) * p.rs_up);   // fp16(fp16(d k) / sqrt(D))
      if (valid && kq == 0 && lg) lg[tok0 + tl] = s16;
      sc[tt] = valid ? __half2float(s16) : -CUDART_INF_F;
    }
    float mx = fmaxf(sc[0], sc[1]);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    const float m_new = fmaxf(m, mx);                       // warp-uniform; tile 0 always has a valid token
    if (m_new != m) {
      const float corr = fast_exp2((m - m_new) * kLog2eQ);
      l *= corr;
#pragma unroll
      for (int mi = 0; mi < 8; ++mi) { acc[mi][0] *= corr; acc[mi][1] *= corr; acc[mi][2] *= corr; acc[mi][3] *= corr; }
      m = m_new;
    }
    const float p0 = fast_exp2((sc[0] - m) * kLog2eQ), p1 = fast_exp2((sc[1] - m) * kLog2eQ);
    l += p0 + p1;
    // probability x value scale as the fp16 B operand: lanes 8 kq' and 8 kq' + 4 hold tokens 2 kq', 2 kq' + 1 (and + 8)
    const uint32_t z = h2u(__floats2half2_rn(p0 * vsc[0], p1 * vsc[1]));
    const uint32_t z0 = __shfl_sync(0xffffffffu, z, 8 * kq), z1 = __shfl_sync(0xffffffffu, z, 8 * kq + 4);
    const uint32_t pb0 = __byte_perm(z0, z1, 0x5410), pb1 = __byte_perm(z0, z1, 0x7632);   // tokens (2kq, 2kq+1), (2kq+8, 2kq+9)

    // ---- O += V^T . p ----
    {
      const uint8_t* vb = st + kOffVm + v_off;
      const uint2 a = *reinterpret_cast<const uint2*>(vb + vt0 * kPlaneRow), b = *reinterpret_cast<const uint2*>(vb + vt1 * kPlaneRow);
      const uint2 c = *reinterpret_cast<const uint2*>(vb + (vt0 + 8) * kPlaneRow), d = *reinterpret_cast<const uint2*>(vb + (vt1 + 8) * kPlaneRow);
      uint2 al = make_uint2(0, 0), bl = al, cl = al, dl = al;
      if (V8) {
        const uint8_t* vlb = st + kOffVl + v_off;
        al = *reinterpret_cast<const uint2*>(vlb + vt0 * kPlaneRow); bl = *reinterpret_cast<const uint2*>(vlb + vt1 * kPlaneRow);
        cl = *reinterpret_cast<const uint2*>(vlb + (vt0 + 8) * kPlaneRow); dl = *reinterpret_cast<const uint2*>(vlb + (vt1 + 8) * kPlaneRow);
      }
      // merged words: nibble i of the low half = token pair's dim (even dims from the low 16 bits, odd dims from the high)
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const uint32_t wa = w ? a.y : a.x, wb = w ? b.y : b.x, wc = w ? c.y : c.x, wd = w ? d.y : d.x;
        const uint32_t la = w ? al.y : al.x, lb = w ? bl.y : bl.x, lc = w ? cl.y : cl.x, ld = w ? dl.y : dl.x;
        __half2 e01[4], o01[4], e23[4], o23[4];             // even / odd dims of the word, token pairs (2kq, 2kq+1) / (+8)
        deq8<V8>(__byte_perm(wa, wb, sel_lo), __byte_perm(la, lb, sel_lo), e01, dk);
        deq8<V8>(__byte_perm(wa, wb, sel_hi), __byte_perm(la, lb, sel_hi), o01, dk);
        deq8<V8>(__byte_perm(wc, wd, sel_lo), __byte_perm(lc, ld, sel_lo), e23, dk);
        deq8<V8>(__byte_perm(wc, wd, sel_hi), __byte_perm(lc, ld, sel_hi), o23, dk);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma16816(acc[w * 4 + j], h2u(e01[j]), h2u(o01[j]), h2u(e23[j]), h2u(o23[j]), pb0, pb1);
      }
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive_q(&rope_empty[rs]);
      if (k + kQStages < my_tiles) issue(k + kQStages);
    }
    ++k;
  }

  l += __shfl_xor_sync(0xffffffffu, l, 4);
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
  if (!live) return;                                       // pass 2: rows that keep their MSB result take no ticket
  if (kq == 0) {
    float* rec = p.part + (static_cast<int64_t>(row) * p.n_splits + split) * D + r * 16;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      __stcg(reinterpret_cast<float4*>(rec + 4 * j), make_float4(acc[2 * j][0], acc[2 * j][2], acc[2 * j + 1][0], acc[2 * j + 1][2]));
    if (r == 0) __stcg(p.part_ml + static_cast<int64_t>(row) * p.n_splits + split, make_float2(m, l));
  }
  // ---- merge by the row's last split (ticket): no second launch.  Also the RequantDecision of pass 1. ----
  __threadfence();
  __syncwarp();
  int ticket = 0;
  if (lane == 0) ticket = atomicAdd(p.counters + row, 1);
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket != p.n_splits - 1) return;
  __threadfence();
  {
    const float2* ml = p.part_ml + static_cast<int64_t>(row) * p.n_splits;
    const float* rec = p.part + static_cast<int64_t>(row) * p.n_splits * D;
    float2 mine[kQMergeMaxSplits / 32];
    float M = -CUDART_INF_F;
#pragma unroll
    for (int kk = 0; kk < kQMergeMaxSplits / 32; ++kk) {
      const int sidx = lane + kk * 32;
      mine[kk] = sidx < p.n_splits ? __ldcg(ml + sidx) : make_float2(-CUDART_INF_F, 0.f);
      M = fmaxf(M, mine[kk].x);
    }
    M = warp_max(M);
    float L = 0.f;
#pragma unroll
    for (int kk = 0; kk < kQMergeMaxSplits / 32; ++kk) {
      const float w = (M > -CUDART_INF_F) ? fast_exp2((mine[kk].x - M) * kLog2eQ) : 0.f;
      L = fmaf(mine[kk].y, w, L);
      mine[kk].x = w;                                      // keep the weight
    }
    L = warp_sum(L);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);           // lane owns output dims 4 lane .. 4 lane + 3
#pragma unroll
    for (int kk = 0; kk < kQMergeMaxSplits / 32; ++kk) {
      if (kk * 32 >= p.n_splits) break;                    // warp-uniform
#pragma unroll 1
      for (int s0 = 0; s0 < 32; s0 += 8) {
        if (kk * 32 + s0 >= p.n_splits) break;
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          v[u] = __ldcg(reinterpret_cast<const float4*>(rec + static_cast<int64_t>(min(kk * 32 + s0 + u, p.n_splits - 1)) * D) + lane);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float w = __shfl_sync(0xffffffffu, mine[kk].x, s0 + u);
          if (kk * 32 + s0 + u >= p.n_splits) w = 0.f;
          o.x = fmaf(v[u].x, w, o.x); o.y = fmaf(v[u].y, w, o.y); o.z = fmaf(v[u].z, w, o.z); o.w = fmaf(v[u].w, w, o.w);
        }
      }
    }
    const float inv = L > 0.f ? 1.f / L : 0.f;
    __half2 h01 = __floats2half2_rn(o.x * inv, o.y * inv), h23 = __floats2half2_rn(o.z * inv, o.w * inv);
    *reinterpret_cast<uint2*>(p.out + static_cast<int64_t>(row) * D + lane * 4) = make_uint2(h2u(h01), h2u(h23));
    // RequantDecision (RequantDecision.scala:72): the row's largest probability is exp(M - M) / L = 1 / L
    if (!LSB && lane == 0) p.need[row] = (inv < p.threshold) ? 1 : 0;
    if (lane == 0) p.counters[row] = 0;                    // self-resetting: the workspace is zero-filled once by the caller
  }
}

// ---- fp16 rows -> planes ---------------------------------------------------------------------------------------
struct QuantizeParams {
  const __half* src[2];
  int64_t src_stride_b, src_stride_h, src_stride_t;     // elements
  uint8_t* msb[2]; uint8_t* lsb[2]; __half* scale[2];
  int64_t plane_stride_b, plane_stride_h, scale_stride_b, scale_stride_h;
  int heads, n_tokens, slot0;
};

// one warp per (tensor, batch, head, token) row of 128 dims: lane holds 4 consecutive dims
__global__ void kv_quantize_kernel(const QuantizeParams p) {
  constexpr int D = 128;
  const int wid = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wid >= p.n_tokens * p.heads) return;                 // warp-uniform
  const int t = wid % p.n_tokens, hh = wid / p.n_tokens, b = blockIdx.y, which = blockIdx.z;
  const __half* src = (which ? p.src[1] : p.src[0]) + b * p.src_stride_b + hh * p.src_stride_h + t * p.src_stride_t;
  const uint2 raw = *reinterpret_cast<const uint2*>(src + lane * 4);
  const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
  const float2 f0 = __half22float2(h2[0]), f1 = __half22float2(h2[1]);
  const float x[4] = {f0.x, f0.y, f1.x, f1.y};
  float amax = fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3])));
  amax = warp_max(amax);
  const __half sc16 = __float2half_rn(__fdiv_rn(amax, 127.f));
  const float sf = fmaxf(__half2float(sc16), 1e-8f);
  uint32_t wm = 0, wl = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int qv = __float2int_rn(__fdiv_rn(x[i], sf));
    qv = max(-127, min(127, qv));
    const uint32_t u = static_cast<uint32_t>(qv + 128);
    const int d = (lane & 1) * 4 + i;                       // dim inside the 8-dim group
    const int pos = (d & 1) ? 4 + (d >> 1) : (d >> 1);      // nibble order [d0,d2,d4,d6,d1,d3,d5,d7]
    wm |= (u >> 4) << (4 * pos);
    wl |= (u & 15u) << (4 * pos);
  }
  wm |= __shfl_xor_sync(0xffffffffu, wm, 1);
  wl |= __shfl_xor_sync(0xffffffffu, wl, 1);
  const int64_t poff = b * p.plane_stride_b + hh * p.plane_stride_h + static_cast<int64_t>(p.slot0 + t) * (D / 2);
  if ((lane & 1) == 0) {
    // word c (dims 8c .. 8c+7) lands at position 4*(c/2 % 4) + (c & 1) + 2*(c / 8): 16-byte chunk kq holds the words of
    // dims [16 kq, 16 kq + 16) followed by their RoPE partners [64 + 16 kq, 64 + 16 kq + 16)
    const int c = lane >> 1, pos = 4 * ((c >> 1) & 3) + (c & 1) + 2 * (c >> 3);
    *reinterpret_cast<uint32_t*>((which ? p.msb[1] : p.msb[0]) + poff + pos * 4) = wm;
    *reinterpret_cast<uint32_t*>((which ? p.lsb[1] : p.lsb[0]) + poff + pos * 4) = wl;
  }
  if (lane == 0) (which ? p.scale[1] : p.scale[0])[b * p.scale_stride_b + hh * p.scale_stride_h + p.slot0 + t] = sc16;
}

}  // namespace spatten

using namespace spatten;

static int check_planes(const SpattenQuantKV* kv, const char* who) {
  SPATTEN_CHECK_ARG(kv && kv->k_msb && kv->k_lsb && kv->k_scale && kv->v_msb && kv->v_lsb && kv->v_scale, "%s: NULL plane", who);
  const uintptr_t al = reinterpret_cast<uintptr_t>(kv->k_msb) | reinterpret_cast<uintptr_t>(kv->k_lsb) |
                       reinterpret_cast<uintptr_t>(kv->v_msb) | reinterpret_cast<uintptr_t>(kv->v_lsb) |
                       reinterpret_cast<uintptr_t>(kv->k_scale) | reinterpret_cast<uintptr_t>(kv->v_scale) |
                       static_cast<uintptr_t>(kv->plane_stride_b | kv->plane_stride_h);
  SPATTEN_CHECK_ARG((al & 15) == 0 && kv->scale_stride_b % 8 == 0 && kv->scale_stride_h % 8 == 0,
                    "%s: planes, scales and strides must be 16-byte aligned", who);
  return SPATTEN_OK;
}

extern "C" int spatten_kv_quantize(const void* k, const void* v, int64_t stride_b, int64_t stride_h, int64_t stride_t,
                                   int32_t batch, int32_t num_kv_heads, int32_t n_tokens, int32_t head_dim,
                                   const SpattenQuantKV* dst, int32_t slot0, void* stream) {
  SPATTEN_CHECK_ARG(k && v && batch > 0 && num_kv_heads > 0 && n_tokens > 0 && slot0 >= 0, "spatten_kv_quantize: bad arguments");
  if (head_dim != 128) { set_error("spatten_kv_quantize: head_dim %d unsupported (128)", head_dim); return SPATTEN_ERR_UNSUPPORTED; }
  if (int rc = check_planes(dst, "spatten_kv_quantize")) return rc;
  SPATTEN_CHECK_ARG(stride_t % 4 == 0 && stride_h % 4 == 0 && stride_b % 4 == 0, "spatten_kv_quantize: source rows must be 8-byte aligned");
  QuantizeParams p{};
  p.src[0] = static_cast<const __half*>(k); p.src[1] = static_cast<const __half*>(v);
  p.src_stride_b = stride_b; p.src_stride_h = stride_h; p.src_stride_t = stride_t;
  p.msb[0] = static_cast<uint8_t*>(dst->k_msb); p.msb[1] = static_cast<uint8_t*>(dst->v_msb);
  p.lsb[0] = static_cast<uint8_t*>(dst->k_lsb); p.lsb[1] = static_cast<uint8_t*>(dst->v_lsb);
  p.scale[0] = static_cast<__half*>(dst->k_scale); p.scale[1] = static_cast<__half*>(dst->v_scale);
  p.plane_stride_b = dst->plane_stride_b; p.plane_stride_h = dst->plane_stride_h;
  p.scale_stride_b = dst->scale_stride_b; p.scale_stride_h = dst->scale_stride_h;
  p.heads = num_kv_heads; p.n_tokens = n_tokens; p.slot0 = slot0;
  const int warps = n_tokens * num_kv_heads;
  dim3 grid((warps + 7) / 8, batch, 2);
  kv_quantize_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "kv_quantize_kernel launch");
}

extern "C" size_t spatten_decode_quant_workspace_bytes(int32_t batch, int32_t num_heads, int32_t head_dim) {
  if (batch <= 0 || num_heads <= 0 || head_dim <= 0) return 0;
  return static_cast<size_t>(kQMaxPartials) * (head_dim + 2) * sizeof(float) + static_cast<size_t>(kQMaxPartials) * sizeof(int);
}

extern "C" int spatten_decode_attn_quant(const SpattenQuantDecodeArgs* a, void* stream) {
  SPATTEN_CHECK_ARG(a != nullptr && a->struct_size == sizeof(SpattenQuantDecodeArgs), "spatten_decode_attn_quant: bad args/struct_size");
  SPATTEN_CHECK_ARG(a->batch > 0 && a->num_heads > 0 && a->num_kv_heads > 0 && a->num_heads % a->num_kv_heads == 0,
                    "spatten_decode_attn_quant: bad head counts");
  if (a->head_dim != 128) { set_error("spatten_decode_attn_quant: head_dim %d unsupported (128)", a->head_dim); return SPATTEN_ERR_UNSUPPORTED; }
  SPATTEN_CHECK_ARG(a->kv_len > 0 && a->rope_len >= a->kv_len, "spatten_decode_attn_quant: bad kv_len / rope_len");
  SPATTEN_CHECK_ARG(a->v_bits == 4 || a->v_bits == 8, "spatten_decode_attn_quant: v_bits must be 4 or 8");
  SPATTEN_CHECK_ARG(a->q && a->rope && a->out && a->need_requant && a->workspace, "spatten_decode_attn_quant: NULL pointer");
  SPATTEN_CHECK_ARG(!a->logits || a->logits_stride >= a->kv_len, "spatten_decode_attn_quant: logits_stride < kv_len");
  if (int rc = check_planes(&a->kv, "spatten_decode_attn_quant")) return rc;

  QuantDecodeParams p{};
  p.kv.k_msb = static_cast<const uint8_t*>(a->kv.k_msb); p.kv.k_lsb = static_cast<const uint8_t*>(a->kv.k_lsb);
  p.kv.k_scale = static_cast<const __half*>(a->kv.k_scale);
  p.kv.v_msb = static_cast<const uint8_t*>(a->kv.v_msb); p.kv.v_lsb = static_cast<const uint8_t*>(a->kv.v_lsb);
  p.kv.v_scale = static_cast<const __half*>(a->kv.v_scale);
  p.kv.plane_stride_b = a->kv.plane_stride_b; p.kv.plane_stride_h = a->kv.plane_stride_h;
  p.kv.scale_stride_b = a->kv.scale_stride_b; p.kv.scale_stride_h = a->kv.scale_stride_h;
  p.q = static_cast<const __half*>(a->q); p.rope = static_cast<const __half*>(a->rope); p.q_pos = a->q_pos;
  p.out = static_cast<__half*>(a->out); p.logits = static_cast<__half*>(a->logits); p.logits_stride = a->logits_stride;
  p.need = a->need_requant;
  p.H = a->num_heads; p.n_rep = a->num_heads / a->num_kv_heads; p.kv_len = a->kv_len; p.rope_len = a->rope_len;
  p.v_bits = a->v_bits; p.rows = a->batch * a->num_heads;
  p.rs_up = nextafterf(1.0f / sqrtf(static_cast<float>(a->head_dim)), 1.0f);   // fp16(x * rs_up) == fp16(x / sqrt(128)) for every fp16 x
  p.threshold = a->requant_threshold;

  const int sms = sm_count();
  if (sms <= 0) return SPATTEN_ERR_CUDA;
  p.rpc = 1;
  while (p.rpc < kQWarps && p.rpc * 2 <= p.rows) p.rpc *= 2;              // rows per CTA; the other warps interleave tiles
  p.spc = kQWarps / p.rpc;
  p.n_rg = (p.rows + p.rpc - 1) / p.rpc;
  // ranges: fill the SMs; a range is a whole number of tiles per interleaved subset
  const int tiles = (a->kv_len + kQTile - 1) / kQTile;
  int best = 1; long best_cost = -1;
  const int max_ranges = std::max(1, std::min({(tiles + p.spc - 1) / p.spc, kQMaxPartials / (p.n_rg * p.rpc * p.spc),
                                               kQMergeMaxSplits / p.spc}));
  for (int nr = 1; nr <= max_ranges; ++nr) {
    const long waves = (static_cast<long>(p.n_rg) * nr + sms - 1) / sms;
    const int tpw = ((tiles + nr - 1) / nr + p.spc - 1) / p.spc;          // tiles per warp
    const long cost = waves * (tpw + 3);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = nr; }
  }
  const int tiles_per_range = (tiles + best - 1) / best;
  const int n_ranges = (tiles + tiles_per_range - 1) / tiles_per_range;
  p.tok_per_range = tiles_per_range * kQTile;
  p.n_splits = n_ranges * p.spc;
  SPATTEN_CHECK_ARG(a->workspace_bytes >= spatten_decode_quant_workspace_bytes(a->batch, a->num_heads, a->head_dim),
                    "spatten_decode_attn_quant: workspace too small");
  p.part_ml = static_cast<float2*>(a->workspace);
  p.part = reinterpret_cast<float*>(static_cast<uint8_t*>(a->workspace) + static_cast<size_t>(kQMaxPartials) * sizeof(float2));
  p.counters = reinterpret_cast<int*>(p.part + static_cast<size_t>(kQMaxPartials) * 128);
  SPATTEN_CHECK_ARG(p.rows <= kQMaxPartials, "spatten_decode_attn_quant: too many rows");

  const bool v8 = a->v_bits == 8;
  auto smem_for = [](bool lsb, bool v8k) {
    return static_cast<size_t>(kQWarps) * kQStages * (kQTile * 64 * (2 + (lsb ? 1 : 0) + (v8k ? 1 : 0)) + 64) +
           static_cast<size_t>(kRopeStages) * kRopeStage + (2 * kRopeStages + kQWarps * kQStages) * 8;
  };
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    cudaError_t e1 = cudaFuncSetAttribute(decode_attn_q4_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_for(false, false)));
    cudaError_t e2 = cudaFuncSetAttribute(decode_attn_q4_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_for(true, false)));
    cudaError_t e3 = cudaFuncSetAttribute(decode_attn_q4_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_for(false, true)));
    cudaError_t e4 = cudaFuncSetAttribute(decode_attn_q4_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_for(true, true)));
    for (cudaError_t e : {e1, e2, e3, e4})
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(decode_attn_q4_kernel)");
    configured[dev] = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = p.n_rg * n_ranges;
  if (v8) decode_attn_q4_kernel<false, true><<<grid, kQThreads, smem_for(false, true), st>>>(p);     // pass 1: MSB keys
  else    decode_attn_q4_kernel<false, false><<<grid, kQThreads, smem_for(false, false), st>>>(p);
  if (a->requant_threshold > 0.f) {                       // pass 2: replay flagged rows with the LSB plane
    if (v8) decode_attn_q4_kernel<true, true><<<grid, kQThreads, smem_for(true, true), st>>>(p);
    else    decode_attn_q4_kernel<true, false><<<grid, kQThreads, smem_for(true, false), st>>>(p);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "decode_attn_q4 launch");
}
