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
// [d0,d2,d4,d6,d1,d3,d5,d7], so that `(w >> 4i) & 0x000f000f` yields the half2 pair (d_2i, d_2i+1) directly
// (the fetch-friendly interleave plays the role of the RTL's bit-reversed BMR layout, BMR.scala:59-65).
// Pass 1 streams the MSB planes only (K: D/2 bytes per token instead of 2*D) and runs the whole attention;
// rows whose largest softmax probability is below the threshold are replayed by pass 2 with MSB+LSB keys.
#include <math_constants.h>

#include <algorithm>
#include <cmath>

#include "../../include/spatten_b200.h"
#include "common.cuh"

namespace spatten {

constexpr int kQWarps = 16;     // the dequantise-rotate-dot loop is ALU heavy: 16 warps per SM hide its latencies
constexpr int kQThreads = kQWarps * 32;
constexpr int kQTile = 16;      // tokens per stage (per warp): 4 planes x 1 KiB + scales, 3 stages -> 200 KiB per CTA
constexpr int kQStages = 3;
constexpr int kQMaxPartials = 8192;
constexpr float kLog2eQ = 1.4426950408889634f;

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
  int H, n_rep, kv_len, rope_len, v_bits;
  int rows, rpc, spc, n_rg, n_splits, tok_per_split;
  float sqrt_d, rsqrt_d, threshold;
};

__device__ __forceinline__ float q_div_by_const(float x, float c, float rc) {
  const float y = x * rc;
  const float r = fmaf(-y, c, x);
  const float y2 = fmaf(r, rc, y);
  return (fabsf(y2) <= 3.4028234e38f) ? y2 : y;
}

// 8 dims (one u32 of the MSB plane, optionally one of the LSB plane) -> 4 half2 holding the exact integers
// (16*m + l) - 128, or the MSB bucket midpoint (16*m + 8) - 128 when the LSB plane is not fetched.
template <bool LSB>
__device__ __forceinline__ void deq8(uint32_t wm, uint32_t wl, __half2 (&x)[4]) {
  // A nibble OR-ed into mantissa bits 4..7 of the fp16 pattern 0x6400 (= 1024) reads as 1024 + 16*m, so the MSB code
  // 16*m needs no multiply: one shift (none for i = 1), one LOP3, one HADD2 per pair of dims, all exact.
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t sh = i == 0 ? (wm << 4) : (i == 1 ? wm : (wm >> (4 * i - 4)));
    const uint32_t bm = (sh & 0x00f000f0u) | 0x64006400u;                        // (1024 + 16 m_2i, 1024 + 16 m_2i+1)
    if (LSB) {
      const uint32_t bl = ((wl >> (4 * i)) & 0x000f000fu) | 0x64006400u;         // (1024 + l_2i, 1024 + l_2i+1)
      const __half2 hi = __hadd2(*reinterpret_cast<const __half2*>(&bm), __float2half2_rn(-1152.f));   // 16 m - 128
      x[i] = __hadd2(hi, __hadd2(*reinterpret_cast<const __half2*>(&bl), __float2half2_rn(-1024.f)));   // + l, exact
    } else {
      x[i] = __hadd2(*reinterpret_cast<const __half2*>(&bm), __float2half2_rn(-1144.f));             // 16 m + 8 - 128
    }
  }
}

// ---- pass kernel: stream the planes of one token range per warp, write (acc, m, l) partials ----------------
template <bool LSB>
__global__ void __launch_bounds__(kQThreads, 1) decode_attn_q4_kernel(const QuantDecodeParams p) {
  constexpr int D = 128;
  constexpr int kPlaneRow = D / 2;                      // 64 bytes per token per plane
  constexpr int kPlaneTile = kQTile * kPlaneRow;        // 2 KiB
  constexpr int kScaleTile = kQTile * 2;                // 64 B
  constexpr int kStageBytes = 4 * kPlaneTile + 2 * kScaleTile;   // Kmsb | Klsb | Vmsb | Vlsb | kscale | vscale
  constexpr int kWarpSmem = kQStages * kStageBytes;
  extern __shared__ __align__(128) uint8_t smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 3, c = lane & 7;
  uint8_t* wsm = smem + warp * kWarpSmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kQWarps * kWarpSmem) + warp * kQStages;

  const int rg = blockIdx.x % p.n_rg, range = blockIdx.x / p.n_rg;
  const int row = rg * p.rpc + (warp % p.rpc);
  const int split = range * p.spc + warp / p.rpc;
  const bool row_ok = row < p.rows;
  const int bi = row_ok ? row / p.H : 0, h = row_ok ? row % p.H : 0, kvh = h / p.n_rep;
  // pass 2 replays only the rows flagged by the requant decision
  const bool live = row_ok && (!LSB || p.need[row] != 0);
  int t0 = split * p.tok_per_split;
  int t1 = min(t0 + p.tok_per_split, p.kv_len);
  if (!live || t0 >= t1) { t0 = 0; t1 = 0; }
  const int nt = (t1 - t0 + kQTile - 1) / kQTile;

  const int64_t poff = bi * p.kv.plane_stride_b + kvh * p.kv.plane_stride_h;
  const int64_t soff = bi * p.kv.scale_stride_b + kvh * p.kv.scale_stride_h;
  const bool v8 = p.v_bits == 8;
  const uint64_t pol = policy_evict_first();
  auto issue = [&](int i) {
    const int s = i % kQStages;
    const int tok = t0 + i * kQTile;
    const int n = min(kQTile, t1 - tok);
    const uint32_t pb = static_cast<uint32_t>(n) * kPlaneRow;
    const uint32_t sb = static_cast<uint32_t>((n + 7) / 8 * 8) * 2;     // scales: 16-byte granules (buffers are padded)
    uint8_t* st = wsm + s * kStageBytes;
    mbar_expect_tx(&bars[s], pb * (2 + (LSB ? 1 : 0) + (v8 ? 1 : 0)) + 2 * sb);
    bulk_g2s(st, p.kv.k_msb + poff + static_cast<int64_t>(tok) * kPlaneRow, pb, &bars[s], pol);
    if (LSB) bulk_g2s(st + kPlaneTile, p.kv.k_lsb + poff + static_cast<int64_t>(tok) * kPlaneRow, pb, &bars[s], pol);
    bulk_g2s(st + 2 * kPlaneTile, p.kv.v_msb + poff + static_cast<int64_t>(tok) * kPlaneRow, pb, &bars[s], pol);
    if (v8) bulk_g2s(st + 3 * kPlaneTile, p.kv.v_lsb + poff + static_cast<int64_t>(tok) * kPlaneRow, pb, &bars[s], pol);
    bulk_g2s(st + 4 * kPlaneTile, p.kv.k_scale + soff + tok, sb, &bars[s], pol);
    bulk_g2s(st + 4 * kPlaneTile + kScaleTile, p.kv.v_scale + soff + tok, sb, &bars[s], pol);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kQStages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
#pragma unroll 1
    for (int i = 0; i < min(nt, kQStages); ++i) issue(i);
  }
  __syncwarp();

  // query: fp16 RoPE at its position, kept as the fp16 values themselves (FHFMA operands); lane holds dims c*8.. and 64+c*8..
  __half2 qlr[4], qhr[4];
  {
    int64_t pos = p.q_pos ? p.q_pos[bi] : p.kv_len - 1;
    pos = max(static_cast<int64_t>(0), min(pos, static_cast<int64_t>(p.rope_len - 1)));
    const __half* qr = p.q + static_cast<int64_t>(row_ok ? row : 0) * D;
    uint4 a = *reinterpret_cast<const uint4*>(qr + c * 8), b = *reinterpret_cast<const uint4*>(qr + 64 + c * 8);
    uint4 cs = __ldg(reinterpret_cast<const uint4*>(p.rope + pos * D + c * 8));
    uint4 sn = __ldg(reinterpret_cast<const uint4*>(p.rope + pos * D + 64 + c * 8));
    const __half2* a2 = reinterpret_cast<const __half2*>(&a); const __half2* b2 = reinterpret_cast<const __half2*>(&b);
    const __half2* c2 = reinterpret_cast<const __half2*>(&cs); const __half2* s2 = reinterpret_cast<const __half2*>(&sn);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      qlr[i] = __hsub2_rn(__hmul2_rn(a2[i], c2[i]), __hmul2_rn(b2[i], s2[i]));
      qhr[i] = __hadd2_rn(__hmul2_rn(b2[i], c2[i]), __hmul2_rn(a2[i], s2[i]));
    }
  }
  float m = -CUDART_INF_F, l = 0.f, accl[8], acch[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) accl[i] = acch[i] = 0.f;
  __half* lg = p.logits ? p.logits + static_cast<int64_t>(row_ok ? row : 0) * p.logits_stride : nullptr;

#pragma unroll 1
  for (int i = 0; i < nt; ++i) {
    const int s = i % kQStages;
    const int tok0 = t0 + i * kQTile;
    const int ntok = min(kQTile, t1 - tok0);
    const uint8_t* st = wsm + s * kStageBytes;
    mbar_wait(&bars[s], (i / kQStages) & 1);
#pragma unroll 1
    for (int sub = 0; sub < kQTile; sub += 16) {
      if (sub >= ntok) break;
      float sc[4], vs[4], d[4];
      // RoPE rows of the 4 passes first (independent global loads), then the 4 independent dot-product chains
      uint4 cs4[4], sn4[4];
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int jr = min(tok0 + sub + ps * 4 + g, p.rope_len - 1);
        cs4[ps] = __ldg(reinterpret_cast<const uint4*>(p.rope + static_cast<int64_t>(jr) * D + c * 8));
        sn4[ps] = __ldg(reinterpret_cast<const uint4*>(p.rope + static_cast<int64_t>(jr) * D + 64 + c * 8));
      }
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int tl = sub + ps * 4 + g;
        const uint32_t wml = *reinterpret_cast<const uint32_t*>(st + tl * kPlaneRow + c * 4);
        const uint32_t wmh = *reinterpret_cast<const uint32_t*>(st + tl * kPlaneRow + 32 + c * 4);
        uint32_t wll = 0, wlh = 0;
        if (LSB) {
          wll = *reinterpret_cast<const uint32_t*>(st + kPlaneTile + tl * kPlaneRow + c * 4);
          wlh = *reinterpret_cast<const uint32_t*>(st + kPlaneTile + tl * kPlaneRow + 32 + c * 4);
        }
        __half2 kl[4], kh[4];
        deq8<LSB>(wml, wll, kl);
        deq8<LSB>(wmh, wlh, kh);
        const __half2* c2 = reinterpret_cast<const __half2*>(&cs4[ps]);
        const __half2* s2 = reinterpret_cast<const __half2*>(&sn4[ps]);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __half2 rl = __hsub2_rn(__hmul2_rn(kl[e], c2[e]), __hmul2_rn(kh[e], s2[e]));
          const __half2 rh = __hadd2_rn(__hmul2_rn(kh[e], c2[e]), __hmul2_rn(kl[e], s2[e]));
          d0 = fhfma_lo(rl, qlr[e], d0); d0 = fhfma_hi(rl, qlr[e], d0);      // exact fp16 x fp16 products, fp32 sums
          d1 = fhfma_lo(rh, qhr[e], d1); d1 = fhfma_hi(rh, qhr[e], d1);
        }
        d[ps] = d0 + d1;
      }
#pragma unroll
      for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) d[ps] += __shfl_xor_sync(0xffffffffu, d[ps], o);
      }
#pragma unroll
      for (int ps = 0; ps < 4; ++ps) {
        const int tl = sub + ps * 4 + g;
        const int j = tok0 + tl;
        const bool valid = tl < ntok;
        const float ks = valid ? __half2float(*reinterpret_cast<const __half*>(st + 4 * kPlaneTile + tl * 2)) : 0.f;
        vs[ps] = valid ? __half2float(*reinterpret_cast<const __half*>(st + 4 * kPlaneTile + kScaleTile + tl * 2)) : 0.f;
        const __half s16 = __float2half_rn(q_div_by_const(__half2float(__float2half_rn(d[ps] * ks)), p.sqrt_d, p.rsqrt_d));
        if (valid && c == 0 && lg) lg[j] = s16;
        sc[ps] = valid ? __half2float(s16) : -CUDART_INF_F;
      }
      float mx = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
      const float m_new = fmaxf(m, mx);
      if (m_new > -CUDART_INF_F) {
        if (m_new != m) {
          const float corr = fast_exp2((m - m_new) * kLog2eQ);
          l *= corr;
#pragma unroll
          for (int e = 0; e < 8; ++e) { accl[e] *= corr; acch[e] *= corr; }
          m = m_new;
        }
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const int tl = sub + ps * 4 + g;
          const bool valid = tl < ntok;
          const float pr = valid ? fast_exp2((sc[ps] - m) * kLog2eQ) : 0.f;
          l += pr;
          const __half2 pw = __float2half2_rn(pr * vs[ps]);      // probability x value scale, fp16 operand of the FHFMAs
          // invalid rows of a partial stage hold stale bytes: every nibble pattern still dequantises to a finite
          // integer, and pw is exactly 0 for them
          const uint32_t wml = *reinterpret_cast<const uint32_t*>(st + 2 * kPlaneTile + tl * kPlaneRow + c * 4);
          const uint32_t wmh = *reinterpret_cast<const uint32_t*>(st + 2 * kPlaneTile + tl * kPlaneRow + 32 + c * 4);
          __half2 vl[4], vh[4];
          if (v8) {
            const uint32_t wll = *reinterpret_cast<const uint32_t*>(st + 3 * kPlaneTile + tl * kPlaneRow + c * 4);
            const uint32_t wlh = *reinterpret_cast<const uint32_t*>(st + 3 * kPlaneTile + tl * kPlaneRow + 32 + c * 4);
            deq8<true>(wml, wll, vl);
            deq8<true>(wmh, wlh, vh);
          } else {
            deq8<false>(wml, 0, vl);
            deq8<false>(wmh, 0, vh);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            accl[2 * e] = fhfma_lo(vl[e], pw, accl[2 * e]); accl[2 * e + 1] = fhfma_hi(vl[e], pw, accl[2 * e + 1]);
            acch[2 * e] = fhfma_lo(vh[e], pw, acch[2 * e]); acch[2 * e + 1] = fhfma_hi(vh[e], pw, acch[2 * e + 1]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0 && i + kQStages < nt) issue(i + kQStages);
  }

  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    accl[e] += __shfl_xor_sync(0xffffffffu, accl[e], 8);  accl[e] += __shfl_xor_sync(0xffffffffu, accl[e], 16);
    acch[e] += __shfl_xor_sync(0xffffffffu, acch[e], 8);  acch[e] += __shfl_xor_sync(0xffffffffu, acch[e], 16);
  }
  if (live && g == 0) {
    float* rec = p.part + (static_cast<int64_t>(row) * p.n_splits + split) * D;
    *reinterpret_cast<float4*>(rec + c * 8) = make_float4(accl[0], accl[1], accl[2], accl[3]);
    *reinterpret_cast<float4*>(rec + c * 8 + 4) = make_float4(accl[4], accl[5], accl[6], accl[7]);
    *reinterpret_cast<float4*>(rec + 64 + c * 8) = make_float4(acch[0], acch[1], acch[2], acch[3]);
    *reinterpret_cast<float4*>(rec + 64 + c * 8 + 4) = make_float4(acch[4], acch[5], acch[6], acch[7]);
    if (c == 0) p.part_ml[static_cast<int64_t>(row) * p.n_splits + split] = make_float2(m, l);
  }
}

// ---- merge kernel: one CTA of D threads per row (thread = output dim); pass 1 also takes the requant decision -------
// The split weights are computed once per row by the first warps and shared through shared memory; every thread then
// walks the row's partials with 16 independent coalesced loads in flight (the old one-warp-per-row loop was a chain of
// ~10 dependent L2 round trips: 21 us for 32 rows).
constexpr int kQMergeMaxSplits = 512;
template <bool LSB>
__global__ void __launch_bounds__(128) quant_merge_kernel(const QuantDecodeParams p) {
  constexpr int D = 128;
  __shared__ float s_w[kQMergeMaxSplits];
  __shared__ float s_red[8];
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (LSB && p.need[row] == 0) return;                  // pass 2 touches replayed rows only
  const float* rec = p.part + static_cast<int64_t>(row) * p.n_splits * D;
  const float2* ml = p.part_ml + static_cast<int64_t>(row) * p.n_splits;
  float2 mine[kQMergeMaxSplits / 128];
  float M = -CUDART_INF_F;
#pragma unroll
  for (int k = 0; k < kQMergeMaxSplits / 128; ++k) {
    const int s = tid + k * 128;
    mine[k] = s < p.n_splits ? __ldcg(ml + s) : make_float2(-CUDART_INF_F, 0.f);
    M = fmaxf(M, mine[k].x);
  }
  M = warp_max(M);
  if (lane == 0) s_red[warp] = M;
  __syncthreads();
  M = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  float L = 0.f;
#pragma unroll
  for (int k = 0; k < kQMergeMaxSplits / 128; ++k) {
    const int s = tid + k * 128;
    const float w = (M > -CUDART_INF_F) ? fast_exp2((mine[k].x - M) * kLog2eQ) : 0.f;
    if (s < p.n_splits) s_w[s] = w;
    L = fmaf(mine[k].y, w, L);
  }
  L = warp_sum(L);
  if (lane == 0) s_red[4 + warp] = L;
  __syncthreads();
  L = s_red[4] + s_red[5] + s_red[6] + s_red[7];
  float o = 0.f;
  constexpr int kU = 16;
#pragma unroll 1
  for (int s0 = 0; s0 < p.n_splits; s0 += kU) {
    float v[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) v[u] = __ldcg(rec + static_cast<int64_t>(min(s0 + u, p.n_splits - 1)) * D + tid);
#pragma unroll
    for (int u = 0; u < kU; ++u) o = fmaf(v[u], (s0 + u < p.n_splits) ? s_w[s0 + u] : 0.f, o);
  }
  const float inv = L > 0.f ? 1.f / L : 0.f;
  p.out[static_cast<int64_t>(row) * D + tid] = __float2half_rn(o * inv);
  // RequantDecision (RequantDecision.scala:72): the row's largest probability is exp(M - M) / L = 1 / L
  if (!LSB && tid == 0) p.need[row] = (inv < p.threshold) ? 1 : 0;
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
    *reinterpret_cast<uint32_t*>((which ? p.msb[1] : p.msb[0]) + poff + (lane >> 1) * 4) = wm;
    *reinterpret_cast<uint32_t*>((which ? p.lsb[1] : p.lsb[0]) + poff + (lane >> 1) * 4) = wl;
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
  return static_cast<size_t>(kQMaxPartials) * (head_dim + 2) * sizeof(float);
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
  p.sqrt_d = static_cast<float>(sqrt(static_cast<double>(a->head_dim))); p.rsqrt_d = 1.0f / p.sqrt_d;
  p.threshold = a->requant_threshold;

  const int sms = sm_count();
  if (sms <= 0) return SPATTEN_ERR_CUDA;
  p.rpc = p.rows >= 8 ? 8 : (p.rows >= 4 ? 4 : (p.rows >= 2 ? 2 : 1));   // rows per CTA; the other warps split tokens
  p.spc = kQWarps / p.rpc;
  p.n_rg = (p.rows + p.rpc - 1) / p.rpc;
  // ranges: fill the SMs once; every warp gets a multiple of kQTile tokens (scale rows are fetched in 16-byte granules)
  int best = 1; long best_cost = -1;
  const int max_ranges = std::max(1, std::min({(a->kv_len + p.spc * 32 - 1) / (p.spc * 32),
                                               kQMaxPartials / (p.n_rg * p.rpc * p.spc), 256 / p.spc}));
  for (int nr = 1; nr <= max_ranges; ++nr) {
    const long waves = (static_cast<long>(p.n_rg) * nr + sms - 1) / sms;
    const int tps = ((a->kv_len + nr * p.spc - 1) / (nr * p.spc) + kQTile - 1) / kQTile * kQTile;
    const long cost = waves * (tps + 48);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = nr; }
  }
  p.n_splits = best * p.spc;
  p.tok_per_split = ((a->kv_len + p.n_splits - 1) / p.n_splits + kQTile - 1) / kQTile * kQTile;
  SPATTEN_CHECK_ARG(a->workspace_bytes >= spatten_decode_quant_workspace_bytes(a->batch, a->num_heads, a->head_dim),
                    "spatten_decode_attn_quant: workspace too small");
  p.part_ml = static_cast<float2*>(a->workspace);
  p.part = reinterpret_cast<float*>(static_cast<uint8_t*>(a->workspace) + static_cast<size_t>(kQMaxPartials) * sizeof(float2));

  constexpr size_t smem = static_cast<size_t>(kQWarps) * kQStages * (4 * kQTile * 64 + 2 * kQTile * 2) + kQWarps * kQStages * 8;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    cudaError_t e1 = cudaFuncSetAttribute(decode_attn_q4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaError_t e2 = cudaFuncSetAttribute(decode_attn_q4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e1 != cudaSuccess || e2 != cudaSuccess) return cuda_fail(e1 != cudaSuccess ? e1 : e2, "cudaFuncSetAttribute(decode_attn_q4_kernel)");
    configured[dev] = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = p.n_rg * best, mgrid = p.rows;        // merge: one CTA of D threads per row
  decode_attn_q4_kernel<false><<<grid, kQThreads, smem, st>>>(p);           // pass 1: MSB planes
  quant_merge_kernel<false><<<mgrid, 128, 0, st>>>(p);                       //   + RequantDecision
  if (a->requant_threshold > 0.f) {
    decode_attn_q4_kernel<true><<<grid, kQThreads, smem, st>>>(p);          // pass 2: replay flagged rows with the LSB plane
    quant_merge_kernel<true><<<mgrid, 128, 0, st>>>(p);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "decode_attn_q4 launch");
}
