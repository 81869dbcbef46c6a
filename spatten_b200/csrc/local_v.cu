// "Local V pruning" in ONE launch -- spatten_local_v_attn().
//
// The RTL pipeline order (SpAttenController.scala:347-749): DotProduct over ALL cached keys -> Softmax -> TopK
// (target = val_fetch_num, :554; skipped when val_fetch_num >= key_fetch_num, :591) -> FetchValMat for the top-k indexes
// only (:668-726) -> MultiplyValue (:749).  The reference's Python path never runs this (README.md:11), so the
// arithmetic is the restatement oracle/spatten_oracle.py::local_value_attention ("parity unpinned").
//
// One thread-block cluster (4 or 8 CTAs, fewer for many rows) owns one (sequence, head) row and keeps everything between
// the stages on chip:
//   1. DotProduct   each CTA streams its slice of the row's K rows HBM -> shared memory with 1-D bulk copies (per-warp
//                   3-stage ring, as the decode kernels), RoPE-on-K in registers, FHFMA dot, fp16 logit formed exactly
//                   as the reference does (:111-113); the logits stay in shared memory and are exported once.
//   2. TopK         8-bit radix select on order-preserving fp16 keys, MSB first: every CTA histograms its slice, the
//                   256-bin histograms are summed across the cluster through distributed shared memory (ld.shared::cluster),
//                   two cluster barriers per digit.  Rule = the RTL engine's: keep every score above the k-th, then the
//                   lowest-index ones equal to it (TopK.scala:159-160, 209-211); the survivors are compacted in index
//                   order with ballot prefix sums (ZeroEliminator / PrefixSum analogue) and exported.
//   3. MultiplyValue  every CTA gathers only the surviving V rows of its slice (whole-row 256-byte loads, 16 in flight per
//                   warp), softmax weights from the logits already on chip, fp32 accumulate; the CTA partials (acc, m, l)
//                   are merged by CTA 0 of the cluster through distributed shared memory.
// HBM traffic = N keys + keep values per row (48 MiB per layer at N = 4096, keep = 2048, 32 heads), nothing else.
#include <math_constants.h>

#include <algorithm>
#include <cmath>

#include "../../include/spatten_b200.h"
#include "common.cuh"

namespace spatten {

constexpr int kLvWarps = 16;         // 4 per scheduler: the DotProduct stage is latency bound with fewer
constexpr int kLvThreads = kLvWarps * 32;
constexpr int kLvStages = 2;
constexpr int kLvTile = 16;          // tokens per stage: 4 passes of 4 keys x 8 lanes
constexpr int kLvPasses = 4;
constexpr int kLvMaxTok = 8192;      // tokens of one row a CTA can hold (logits + survivor list in shared memory)

struct LocalVParams {
  const __half* q;
  const __half* k_cache;
  const __half* v_cache;
  int64_t stride_b, stride_h;
  const __half* rope;
  const int64_t* q_pos;
  __half* out;
  __half* logits;
  int64_t logits_stride;
  int32_t* keep_idx;
  int64_t keep_stride;
  int H, n_rep, n, keep, rope_len, cluster;
  float sqrt_d, rsqrt_d;
};

// what the CTAs of a cluster read from each other (same offset in every CTA's shared memory)
template <int D>
struct LvShared {
  unsigned hist[256];
  unsigned cnt_gt, cnt_eq;
  float m, l;
  float acc[D];
};

__device__ __forceinline__ void lv_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t lv_ld_dsmem_u32(const void* local, unsigned cta) {
  uint32_t raddr, v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local)), "r"(cta));
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(raddr) : "memory");
  return v;
}
__device__ __forceinline__ float lv_ld_dsmem_f32(const void* local, unsigned cta) {
  return __uint_as_float(lv_ld_dsmem_u32(local, cta));
}
// order-preserving key of an fp16 score; -0 and +0 compare equal
__device__ __forceinline__ uint32_t lv_key(__half x) {
  uint32_t u = __half_as_ushort(x);
  if ((u & 0x7fffu) == 0) u = 0;
  return (u & 0x8000u) ? (~u & 0xffffu) : (u | 0x8000u);
}

template <int D>
__global__ void __launch_bounds__(kLvThreads, 1) local_v_kernel(const LocalVParams p) {
  constexpr int E2 = D / 16, NH2 = E2 / 2, VB = E2 * 2;
  using Vec = typename VecT<VB>::type;
  using HV = HalfVec<NH2>;
  constexpr int kRowBytes = D * 2;
  constexpr int kTileBytes = kLvTile * kRowBytes;             // K only
  constexpr int kWarpSmem = kLvStages * kTileBytes;
  constexpr int LPR = D / 4;                                   // lanes per V row (4 dims = 8 bytes per lane)
  constexpr int RPW = 32 / LPR;                                // V rows per warp-wide load

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  uint64_t* bars_all = reinterpret_cast<uint64_t*>(smem + kLvWarps * kWarpSmem);
  LvShared<D>* sh = reinterpret_cast<LvShared<D>*>(bars_all + kLvWarps * kLvStages);
  __half* s_logit = reinterpret_cast<__half*>(sh + 1);
  unsigned short* s_keep = reinterpret_cast<unsigned short*>(s_logit + kLvMaxTok);
  __shared__ unsigned s_wsum[kLvWarps], s_wgt[kLvWarps], s_weq[kLvWarps];
  __shared__ unsigned s_digit, s_kk;
  __shared__ float s_red[kLvWarps];
  __shared__ float s_acc[kLvWarps][D];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 3, c = lane & 7;
  const int C = p.cluster;
  const int row = blockIdx.x / C;
  const unsigned rank = blockIdx.x % C;                        // == %cluster_ctarank for 1-D clusters
  const int bi = row / p.H, h = row % p.H, kvh = h / p.n_rep;
  uint8_t* wsm = ring + warp * kWarpSmem;
  uint64_t* bars = bars_all + warp * kLvStages;

  // this CTA's slice of the row, in whole 16-token tiles
  const int T = (p.n + kLvTile - 1) / kLvTile;
  const int tb = static_cast<int>(static_cast<long long>(rank) * T / C), te = static_cast<int>(static_cast<long long>(rank + 1) * T / C);
  const int tok_begin = tb * kLvTile, tok_end = min(te * kLvTile, p.n);
  const int ntok_cta = max(0, tok_end - tok_begin);
  const __half* Kc = p.k_cache + bi * p.stride_b + kvh * p.stride_h;
  const __half* Vc = p.v_cache + bi * p.stride_b + kvh * p.stride_h;
  __half* lg = p.logits + static_cast<int64_t>(row) * p.logits_stride;

  // ---- stage 1: DotProduct over every key of the slice ----------------------------------------------------------------
  const uint64_t pol = policy_evict_first();
  const int my_tiles = (te - tb > warp) ? (te - tb - warp + kLvWarps - 1) / kLvWarps : 0;     // tiles tb + warp, + 8, ...
  auto issue = [&](int i) {
    const int tile = tb + warp + i * kLvWarps;
    const int tok = tile * kLvTile;
    const uint32_t bytes = static_cast<uint32_t>(min(kLvTile, p.n - tok)) * kRowBytes;
    const int s = i % kLvStages;
    mbar_expect_tx(&bars[s], bytes);
    bulk_g2s(wsm + s * kTileBytes, Kc + static_cast<int64_t>(tok) * D, bytes, &bars[s], pol);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kLvStages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
#pragma unroll 1
    for (int i = 0; i < min(my_tiles, kLvStages); ++i) issue(i);
  }
  __syncwarp();
  HV qlr, qhr;
  {
    int64_t pos = p.q_pos ? p.q_pos[bi] : p.n - 1;
    pos = max(static_cast<int64_t>(0), min(pos, static_cast<int64_t>(p.rope_len - 1)));
    const __half* qr = p.q + static_cast<int64_t>(row) * D;
    HV qa, qb, qcs, qsn;
    *reinterpret_cast<Vec*>(&qa) = *reinterpret_cast<const Vec*>(qr + c * E2);
    *reinterpret_cast<Vec*>(&qb) = *reinterpret_cast<const Vec*>(qr + D / 2 + c * E2);
    *reinterpret_cast<Vec*>(&qcs) = __ldg(reinterpret_cast<const Vec*>(p.rope + pos * D + c * E2));
    *reinterpret_cast<Vec*>(&qsn) = __ldg(reinterpret_cast<const Vec*>(p.rope + pos * D + D / 2 + c * E2));
#pragma unroll
    for (int e = 0; e < NH2; ++e) {
      qlr.h[e] = __hsub2_rn(__hmul2_rn(qa.h[e], qcs.h[e]), __hmul2_rn(qb.h[e], qsn.h[e]));
      qhr.h[e] = __hadd2_rn(__hmul2_rn(qb.h[e], qcs.h[e]), __hmul2_rn(qa.h[e], qsn.h[e]));
    }
  }
#pragma unroll 1
  for (int i = 0; i < my_tiles; ++i) {
    const int s = i % kLvStages;
    const int tok0 = (tb + warp + i * kLvWarps) * kLvTile;
    const int ntok = min(kLvTile, p.n - tok0);
    const uint8_t* Ks = wsm + s * kTileBytes;
    HV cs[kLvPasses], sn[kLvPasses];
#pragma unroll
    for (int ps = 0; ps < kLvPasses; ++ps) {
      const int j = min(tok0 + ps * 4 + g, p.rope_len - 1);
      const __half* r = p.rope + static_cast<int64_t>(j) * D + c * E2;
      *reinterpret_cast<Vec*>(&cs[ps]) = __ldg(reinterpret_cast<const Vec*>(r));
      *reinterpret_cast<Vec*>(&sn[ps]) = __ldg(reinterpret_cast<const Vec*>(r + D / 2));
    }
    if (!mbar_try_wait(&bars[s], (i / kLvStages) & 1)) {
      const long long tw = clock64();
      while (!mbar_try_wait(&bars[s], (i / kLvStages) & 1)) {
        if (clock64() - tw > 4000000000ll) __trap();
      }
    }
    float d[kLvPasses];
#pragma unroll
    for (int ps = 0; ps < kLvPasses; ++ps) {
      const int tl = min(ps * 4 + g, ntok - 1);             // rows past the end of the row re-read the last valid one
      HV kl, kh;
      *reinterpret_cast<Vec*>(&kl) = *reinterpret_cast<const Vec*>(Ks + tl * kRowBytes + c * VB);
      *reinterpret_cast<Vec*>(&kh) = *reinterpret_cast<const Vec*>(Ks + tl * kRowBytes + D + c * VB);
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int e = 0; e < NH2; ++e) {
        const __half2 rl = __hsub2_rn(__hmul2_rn(kl.h[e], cs[ps].h[e]), __hmul2_rn(kh.h[e], sn[ps].h[e]));
        const __half2 rh = __hadd2_rn(__hmul2_rn(kh.h[e], cs[ps].h[e]), __hmul2_rn(kl.h[e], sn[ps].h[e]));
        d0 = fhfma_lo(rl, qlr.h[e], d0);
        d0 = fhfma_hi(rl, qlr.h[e], d0);
        d1 = fhfma_lo(rh, qhr.h[e], d1);
        d1 = fhfma_hi(rh, qhr.h[e], d1);
      }
      d[ps] = d0 + d1;
    }
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
      for (int ps = 0; ps < kLvPasses; ++ps) d[ps] += __shfl_xor_sync(0xffffffffu, d[ps], o);
    }
#pragma unroll
    for (int ps = 0; ps < kLvPasses; ++ps) {
      const int tl = ps * 4 + g;
      if (tl < ntok && c == 0) {
        const __half s16 = __float2half_rn(div_by_const(__half2float(__float2half_rn(d[ps])), p.sqrt_d, p.rsqrt_d));
        s_logit[tok0 + tl - tok_begin] = s16;
        lg[tok0 + tl] = s16;
      }
    }
    __syncwarp();
    if (lane == 0 && i + kLvStages < my_tiles) issue(i + kLvStages);
  }
  __syncthreads();

  // ---- stage 2: TopK -- radix select of the keep-th largest logit over the whole row (all CTAs of the cluster) --------
  const bool prune = p.keep < p.n;
  uint32_t kth = 0;
  unsigned need_eq = 0;
  if (prune) {
    uint32_t prefix = 0;
    unsigned kk = p.keep;
#pragma unroll 1
    for (int shift = 8; shift >= 0; shift -= 8) {
      if (tid < 256) sh->hist[tid] = 0;
      __syncthreads();
      for (int i0 = 0; i0 < ntok_cta; i0 += kLvThreads) {      // warp-uniform trip count: match_any needs the whole warp
        const int i = i0 + tid;
        bool in = i < ntok_cta;
        uint32_t key = 0;
        if (in) {
          key = lv_key(s_logit[i]);
          in = shift == 8 || (key >> 8) == prefix;
        }
        const uint32_t digit = (key >> shift) & 255u;
        // attention logits crowd into a few exponent buckets: one atomic per distinct digit per warp
        const unsigned peers = __match_any_sync(0xffffffffu, in ? digit : (256u + lane));
        if (in && (__ffs(peers) - 1) == lane) atomicAdd(&sh->hist[digit], static_cast<unsigned>(__popc(peers)));
      }
      __syncthreads();
      lv_cluster_sync();
      // thread t < 256 owns bucket 255 - t: total over the cluster, then a descending cumulative count
      unsigned cnt = 0, incl = 0;
      if (tid < 256) {
        for (int cta = 0; cta < C; ++cta) cnt += lv_ld_dsmem_u32(&sh->hist[255 - tid], cta);
        incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        if (lane == 31) s_wsum[warp] = incl;
      }
      __syncthreads();
      if (tid < 256) {
        unsigned base = 0;
        for (int w = 0; w < warp; ++w) base += s_wsum[w];
        incl += base;
        const unsigned excl = incl - cnt;
        if (excl < kk && kk <= incl) { s_digit = 255 - tid; s_kk = kk - excl; }
      }
      __syncthreads();
      prefix = (prefix << 8) | s_digit;
      kk = s_kk;
      lv_cluster_sync();                                       // nobody still reads this round's histograms
    }
    kth = prefix;
    need_eq = kk;
  }
  // survivors of this slice, in index order; ties at the k-th value go to the lowest indexes of the ROW
  unsigned loc_gt = 0, loc_eq = 0;
  for (int i = tid; i < ntok_cta; i += kLvThreads) {
    const uint32_t key = lv_key(s_logit[i]);
    loc_gt += (!prune || key > kth) ? 1u : 0u;
    loc_eq += (prune && key == kth) ? 1u : 0u;
  }
  loc_gt = __reduce_add_sync(0xffffffffu, loc_gt);
  loc_eq = __reduce_add_sync(0xffffffffu, loc_eq);
  if (lane == 0) { s_wgt[warp] = loc_gt; s_weq[warp] = loc_eq; }
  __syncthreads();
  if (tid == 0) {
    unsigned a = 0, b = 0;
    for (int w = 0; w < kLvWarps; ++w) { a += s_wgt[w]; b += s_weq[w]; }
    sh->cnt_gt = a;
    sh->cnt_eq = b;
  }
  __syncthreads();
  lv_cluster_sync();
  unsigned gt_before = 0, eq_before = 0;
  for (unsigned cta = 0; cta < rank; ++cta) {
    gt_before += lv_ld_dsmem_u32(&sh->cnt_gt, cta);
    eq_before += lv_ld_dsmem_u32(&sh->cnt_eq, cta);
  }
  const unsigned eq_take_before = min(eq_before, need_eq);     // equal-valued survivors that lower slices already took
  int32_t* keep_out = p.keep_idx + static_cast<int64_t>(row) * p.keep_stride;
  unsigned run_gt = 0, run_eq = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int base = 0; base < ntok_cta; base += kLvThreads) {
    const int i = base + tid;
    const bool in = i < ntok_cta;
    const uint32_t key = in ? lv_key(s_logit[i]) : 0u;
    const bool gt = in && (!prune || key > kth), eq = in && prune && key == kth;
    const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { s_wgt[warp] = __popc(bg); s_weq[warp] = __popc(be); }
    __syncthreads();
    unsigned gt_b = run_gt + __popc(bg & lt_mask), eq_b = run_eq + __popc(be & lt_mask);
    unsigned tot_gt = 0, tot_eq = 0;
#pragma unroll
    for (int w = 0; w < kLvWarps; ++w) {
      const unsigned a = s_wgt[w], b = s_weq[w];
      if (w < warp) { gt_b += a; eq_b += b; }
      tot_gt += a; tot_eq += b;
    }
    const unsigned eq_room = need_eq - eq_take_before;         // equal-valued survivors this slice may still take
    if (gt || (eq && eq_b < eq_room)) {
      const unsigned local_pos = gt_b + min(eq_b, eq_room);
      s_keep[local_pos] = static_cast<unsigned short>(i);
      keep_out[gt_before + eq_take_before + local_pos] = tok_begin + i;
    }
    run_gt += tot_gt;
    run_eq += tot_eq;
    __syncthreads();
  }
  const int nk = static_cast<int>(run_gt + min(run_eq, need_eq - eq_take_before));   // survivors in this slice

  // ---- stage 3: MultiplyValue over the survivors of the slice ----------------------------------------------------------
  float mx = -CUDART_INF_F;
  for (int i = tid; i < nk; i += kLvThreads) mx = fmaxf(mx, __half2float(s_logit[s_keep[i]]));
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < kLvWarps; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();

  const int gr = lane / LPR, cl = lane % LPR;                  // V row of the warp-wide load, 4-dim chunk inside the row
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float lsum = 0.f;
  constexpr int kU = 16;                                       // V rows in flight per (warp, row group): 4 KiB per warp, 64 KiB per SM
#pragma unroll 1
  for (int e0 = warp * RPW + gr; e0 < nk; e0 += kLvWarps * RPW * kU) {
    uint2 v[kU];
    float pr[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int e = e0 + u * kLvWarps * RPW;
      pr[u] = 0.f;
      v[u] = make_uint2(0u, 0u);
      if (e < nk) {
        const int t = s_keep[e];
        pr[u] = fast_exp2((__half2float(s_logit[t]) - mx) * kLog2e);
        v[u] = __ldg(reinterpret_cast<const uint2*>(Vc + static_cast<int64_t>(tok_begin + t) * D + cl * 4));
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&v[u].x));
      const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&v[u].y));
      acc[0] = fmaf(pr[u], f0.x, acc[0]);
      acc[1] = fmaf(pr[u], f0.y, acc[1]);
      acc[2] = fmaf(pr[u], f1.x, acc[2]);
      acc[3] = fmaf(pr[u], f1.y, acc[3]);
      lsum += pr[u];
    }
  }
  if (RPW == 2) {
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16);
    lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
  }
  if (gr == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) s_acc[warp][cl * 4 + k] = acc[k];
  }
  if (lane == 0) s_red[warp] = lsum;                           // every lane of a row group summed the same weights
  __syncthreads();
  if (tid < D) {
    float o = 0.f;
#pragma unroll
    for (int w = 0; w < kLvWarps; ++w) o += s_acc[w][tid];
    sh->acc[tid] = o;
  }
  if (tid == 0) {
    float l = 0.f;
    for (int w = 0; w < kLvWarps; ++w) l += s_red[w];
    sh->m = mx;
    sh->l = l;
  }
  __syncthreads();
  lv_cluster_sync();
  if (rank == 0 && tid < D) {
    float M = -CUDART_INF_F;
    for (int cta = 0; cta < C; ++cta) M = fmaxf(M, lv_ld_dsmem_f32(&sh->m, cta));
    float L = 0.f, o = 0.f;
    for (int cta = 0; cta < C; ++cta) {
      const float mc = lv_ld_dsmem_f32(&sh->m, cta);
      const float w = (mc > -CUDART_INF_F) ? fast_exp2((mc - M) * kLog2e) : 0.f;
      L = fmaf(lv_ld_dsmem_f32(&sh->l, cta), w, L);
      o = fmaf(lv_ld_dsmem_f32(&sh->acc[tid], cta), w, o);
    }
    p.out[static_cast<int64_t>(row) * D + tid] = __float2half_rn(L > 0.f ? o / L : 0.f);
  }
  lv_cluster_sync();                                           // peers' shared memory stays alive until CTA 0 has read it
}

template <int D>
constexpr size_t lv_smem_bytes() {
  return static_cast<size_t>(kLvWarps) * kLvStages * kLvTile * D * 2 + kLvWarps * kLvStages * 8 + sizeof(LvShared<D>) +
         kLvMaxTok * 2 + kLvMaxTok * 2;
}

template <int D>
static int launch_local_v(const LocalVParams& p, int rows, cudaStream_t st) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(local_v_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(lv_smem_bytes<D>()));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(local_v_kernel)");
    configured[dev] = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(rows * p.cluster);
  cfg.blockDim = dim3(kLvThreads);
  cfg.dynamicSmemBytes = lv_smem_bytes<D>();
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, local_v_kernel<D>, p);
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "local_v_kernel launch");
}

}  // namespace spatten

using namespace spatten;

extern "C" int spatten_local_v_attn(const SpattenLocalVArgs* a, void* stream) {
  SPATTEN_CHECK_ARG(a != nullptr && a->struct_size == sizeof(SpattenLocalVArgs), "spatten_local_v_attn: bad args/struct_size");
  SPATTEN_CHECK_ARG(a->batch > 0 && a->num_heads > 0 && a->num_kv_heads > 0 && a->num_heads % a->num_kv_heads == 0,
                    "spatten_local_v_attn: bad head counts");
  if (a->head_dim != 64 && a->head_dim != 128) {
    set_error("spatten_local_v_attn: head_dim %d unsupported (64 or 128)", a->head_dim);
    return SPATTEN_ERR_UNSUPPORTED;
  }
  SPATTEN_CHECK_ARG(a->kv_len > 0 && a->keep > 0 && a->rope_len >= a->kv_len, "spatten_local_v_attn: bad kv_len / keep / rope_len");
  SPATTEN_CHECK_ARG(a->q && a->k_cache && a->v_cache && a->rope && a->out && a->logits && a->keep_idx,
                    "spatten_local_v_attn: NULL pointer");
  const int keep = std::min(a->keep, a->kv_len);
  SPATTEN_CHECK_ARG(a->logits_stride >= a->kv_len && a->keep_stride >= keep, "spatten_local_v_attn: stride too small");
  SPATTEN_CHECK_ARG(a->cache_stride_h >= static_cast<int64_t>(a->kv_len) * a->head_dim && a->cache_stride_h % 8 == 0 &&
                        a->cache_stride_b % 8 == 0, "spatten_local_v_attn: bad cache strides");
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k_cache) |
                       reinterpret_cast<uintptr_t>(a->v_cache) | reinterpret_cast<uintptr_t>(a->rope);
  SPATTEN_CHECK_ARG((al & 15) == 0, "spatten_local_v_attn: q / k_cache / v_cache / rope must be 16-byte aligned");
  const int rows = a->batch * a->num_heads;
  const int sms = sm_count();
  if (sms <= 0) return SPATTEN_ERR_CUDA;
  // CTAs per row: as many as keep the SMs busy in one wave, and enough that a slice fits the on-chip logit buffer
  int cluster = rows * 8 <= sms ? 8 : (rows * 4 <= sms ? 4 : (rows * 2 <= sms ? 2 : 1));
  while (cluster < 8 && (a->kv_len + cluster - 1) / cluster + kLvTile > kLvMaxTok) cluster *= 2;
  if ((a->kv_len + cluster - 1) / cluster + kLvTile > kLvMaxTok) {
    set_error("spatten_local_v_attn: kv_len %d too long for the on-chip logit buffer (max %d)", a->kv_len, 8 * (kLvMaxTok - kLvTile));
    return SPATTEN_ERR_UNSUPPORTED;
  }
  LocalVParams p{};
  p.q = static_cast<const __half*>(a->q);
  p.k_cache = static_cast<const __half*>(a->k_cache);
  p.v_cache = static_cast<const __half*>(a->v_cache);
  p.stride_b = a->cache_stride_b; p.stride_h = a->cache_stride_h;
  p.rope = static_cast<const __half*>(a->rope);
  p.q_pos = a->q_pos;
  p.out = static_cast<__half*>(a->out);
  p.logits = static_cast<__half*>(a->logits);
  p.logits_stride = a->logits_stride;
  p.keep_idx = a->keep_idx;
  p.keep_stride = a->keep_stride;
  p.H = a->num_heads; p.n_rep = a->num_heads / a->num_kv_heads; p.n = a->kv_len; p.keep = keep; p.rope_len = a->rope_len;
  p.cluster = cluster;
  p.sqrt_d = static_cast<float>(sqrt(static_cast<double>(a->head_dim)));
  p.rsqrt_d = 1.0f / p.sqrt_d;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return a->head_dim == 128 ? launch_local_v<128>(p, rows, st) : launch_local_v<64>(p, rows, st);
}
