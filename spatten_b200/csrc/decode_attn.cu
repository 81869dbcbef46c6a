// Fused decode attention over the (pruned) KV cache -- spatten_decode_attn().
//
// One launch per layer-step does what llama_pos_shift_attention_forward does between the projections
// for q_len == 1 (reference spatten_llm/pos_shift/modify_llama.py:86-144): RoPE(Q), KV append, RoPE(K at
// slot index) in registers, Q.K^T/sqrt(D), logit export, mask, fp32 softmax, P.V, and the split-KV merge.
//
// Work decomposition (B200: 148 SMs, HBM-bound GEMV-class scan):
//   * a CTA is 8 independent warps; warp w owns one (batch, head) row and one contiguous token range,
//     streams that range's K and V rows HBM -> shared memory with 1-D bulk async copies (TMA engine,
//     cp.async.bulk + mbarrier complete_tx) through a private 3-stage ring of 16-token tiles, and keeps its
//     own online softmax state (m, l, acc[D]) in registers.  No CTA-wide barrier in the streaming loop.
//   * the appended token is folded into the tile loop (dropped into the tile in shared memory and into the
//     cache slot), so there is no append kernel and no second code path.
//   * two ways to merge the splits of a row, picked per launch by the host (see spatten_decode_attn below):
//       MODE 1 "cluster": all splits of a row live in one thread-block cluster (4 or 8 CTAs); every warp pushes
//         its (acc, m, l) slices into the finishing CTAs' shared memory (DSMEM), one cluster barrier, local merge.
//       MODE 0 "ticket": grid = (row groups) x (token ranges) sized to one wave; the 8 warps of a CTA take the
//         same token range of 8 rows (RoPE rows hit L1); partials go to a workspace and a per-row atomic ticket
//         elects the last warp, which merges with all partial loads in flight at once.
//   * optional: programmatic dependent launch with the K/V prefetch ahead of griddepcontrol.wait; fused output
//     exchange for head sharding (stores into every peer GPU's gather buffer + flag; consumer wait folded in
//     before the query load); token / head importance accumulation; head mask; scores-only mode.
//
// Tensor cores are deliberately not used here: Llama-2 7B/13B are MHA, so each KV row meets exactly one
// query row (M = 1); a tcgen05 tile would be >= 63/64 padding and the roofline is HBM bytes, not flops.
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <stdio.h>

#include "../../include/spatten_b200.h"
#include "common.cuh"

namespace spatten {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
// A pass = 4 keys x 8 lanes; the passes of one tile are independent (ILP).  Two tile shapes are built:
//   P = 4: 16-token tiles, 3 stages -> 8 warps * 3 * 8 KiB = 192 KiB in flight per SM, one CTA per SM;
//   P = 2:  8-token tiles, 3 stages -> 96 KiB per CTA and <= 128 registers, so TWO CTAs fit on an SM: with
//           programmatic dependent launch the next layer's kernel becomes resident beside this one and its
//           preamble + first K/V tiles overlap this kernel's merge tail.
constexpr int kStages = 3;
constexpr int kMaxPartials = 8192;  // (row, split) partial records the workspace holds
constexpr int kMaxSplits = 256;     // splits per row the merge handles
constexpr int kMaxPeers = 8;        // GPUs of one NVSwitch domain that can receive the fused output exchange
constexpr int kMaxClusterSize = 8;  // CTAs per cluster in cluster mode: 8 (portable maximum) or 4
constexpr size_t kTicketBytes = static_cast<size_t>(kMaxPartials) * sizeof(unsigned int);  // per-row tickets

struct DecodeParams {
  const __half* q;
  const __half* k_new;
  const __half* v_new;
  __half* k_cache;
  __half* v_cache;
  int64_t stride_b, stride_h;
  const __half* rope;
  const int64_t* q_pos;
  const __half* mask;
  const uint8_t* head_mask;
  __half* out;
  __half* logits;
  int64_t logits_stride;
  float* importance;
  int64_t imp_stride;
  float* head_importance;
  float* part;             // [rows][n_splits][D]   un-normalised partial outputs
  float2* part_ml;         // [rows][n_splits]      (running max, running sum)
  unsigned int* counters;  // [rows] per-row tickets
  int H, H_kv, n_rep, kv_len, n_total, append, rope_len;
  int q_len;                        // query rows per sequence (1 = decode; > 1 = prefill rows, append == 0)
  int64_t logits_stride_b, logits_stride_q;
  int rows, rpc, spc, n_rg, n_ranges, n_splits, tok_per_split;
  float sqrt_d, rsqrt_d;
  int pdl;                 // launched with programmatic stream serialization (SPATTEN_DECODE_PDL)
  int scores_only;         // SPATTEN_DECODE_SCORES_ONLY: stream K only, export logits, no softmax / P.V / out
  // Fused exchange (head sharding): the attention output is stored straight into every peer GPU's gather buffer
  // (P2P over NVLink) instead of `out`; the last warp to finish bumps one flag word on every peer.
  __half* peer_out[kMaxPeers];
  unsigned int* peer_flag[kMaxPeers];
  const unsigned int* wait_flags;   // optional: before reading q, wait until these wait_n local flag words (one per
  unsigned int* wait_expected;      // producing rank) reached *wait_expected + 1 -- the previous layer's exchange
  unsigned int* wait_timeout;
  int wait_n;
  unsigned int* done;      // local completion counter (self-resetting)
  unsigned int done_target;
  int n_peers;
  int ds_shift;            // cluster mode: log2(D * rpc / cluster)
  int passes;              // tile shape: 4 (16-token tiles, 1 CTA/SM) or 2 (8-token tiles, 2 CTAs/SM)
  int cluster;             // 0: global partials + per-row tickets; 4 or 8: every split of a row lives in one
                           // thread-block cluster of that size and partials are merged through distributed shared memory
  unsigned long long* dbg;  // optional [grid][8 warps][8] globaltimer stamps (tools/microbench.py --timeline)
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store a float into the shared memory of CTA `cta` of this cluster (same layout in every CTA)
__device__ __forceinline__ void st_dsmem(float* local, unsigned cta, float v) {
  unsigned raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local)), "r"(cta));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}

// Store one piece of the attention output: locally, or -- fused exchange -- into every peer's gather buffer.
template <typename T>
__device__ __forceinline__ void store_out(const DecodeParams& p, int64_t elem_off, const T& val) {
  if (p.n_peers == 0) {
    *reinterpret_cast<T*>(p.out + elem_off) = val;
  } else {
#pragma unroll 1
    for (int r = 0; r < p.n_peers; ++r) *reinterpret_cast<T*>(p.peer_out[r] + elem_off) = val;
  }
}
// Called after the output stores by every participant (a warp; in cluster mode, where all warps of all CTAs take
// part, a whole CTA): system-scope release, count, and the last one publishes "this rank's heads are in your
// buffer" to every peer.
template <bool kWholeCta>
__device__ __forceinline__ void signal_peers(const DecodeParams& p, int lane) {
  if (p.n_peers == 0) return;
  if (kWholeCta) {
    __syncthreads();                       // the CTA's stores are ordered before thread 0's fence (cumulativity)
    if (threadIdx.x != 0) return;
  } else {
    __syncwarp();
  }
  if (lane == 0) {
    __threadfence_system();
    const unsigned prev = atomicAdd(p.done, 1u);
    if (prev == p.done_target - 1) {
      atomicExch(p.done, 0u);
      if (p.wait_n > 0) atomicAdd(p.wait_expected, 1u);   // every CTA has long read it: arm the next launch's wait
      __threadfence_system();
#pragma unroll 1
      for (int r = 0; r < p.n_peers; ++r) atomicAdd_system(p.peer_flag[r], 1u);
    }
  }
}

template <int D, int MODE, bool DBG, int kPasses>   // MODE 0: global partials + tickets (or one split); MODE 1: cluster / DSMEM merge
__global__ void __launch_bounds__(kThreads, kPasses == 2 ? 2 : 1) decode_attn_kernel(const DecodeParams p) {
  constexpr int kTile = 4 * kPasses;  // tokens per pipeline stage (per warp)
  constexpr int E2 = D / 16;   // halves per lane in each half of the row (8 lanes cover D/2)
  constexpr int NH2 = E2 / 2;  // half2 per lane per half-row
  constexpr int VB = E2 * 2;   // bytes per lane per half-row: 16 (D=128) or 8 (D=64)
  using Vec = typename VecT<VB>::type;
  using HV = HalfVec<NH2>;
  constexpr int kRowBytes = D * 2;
  constexpr int kTileBytes = kTile * kRowBytes;
  constexpr int kWarpSmem = kStages * 2 * kTileBytes;
#define STAMP(k)                                                                                             \
  do {                                                                                                       \
    if (DBG && p.dbg && lane == 0) p.dbg[(static_cast<size_t>(blockIdx.x) * kWarps + warp) * 8 + (k)] = gtime(); \
  } while (0)

  extern __shared__ __align__(128) uint8_t smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 3, c = lane & 7;
  uint8_t* wsm = smem + warp * kWarpSmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWarps * kWarpSmem) + warp * kStages;

  // Cluster mode writes into peer CTAs' shared memory at the end.  That is only legal once every CTA of the cluster
  // has started, so everyone arrives on the cluster barrier here and waits for it right before the first push --
  // by then the arrival is seconds-old in GPU terms and the wait is free (compute-sanitizer racecheck clean).
  if constexpr (MODE == 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

  // Programmatic dependent launch: let the next kernel in the stream start its own prologue early ...
  if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // ... and, unless the caller vouches that the cached K/V rows and the RoPE table are older than any
  // kernel still in flight, wait for the previous kernel right here.
  if (p.pdl != 2) asm volatile("griddepcontrol.wait;" ::: "memory");

  // MODE 0: blockIdx = range * n_rg + row group.   MODE 1: blockIdx = cluster (= row group) * cluster size + CTA rank.
  const int rg = MODE ? blockIdx.x / p.cluster : blockIdx.x % p.n_rg;
  const int range = MODE ? blockIdx.x % p.cluster : blockIdx.x / p.n_rg;
  const int row = rg * p.rpc + (warp % p.rpc);
  const int split = range * p.spc + warp / p.rpc;
  const bool row_ok = row < p.rows;
  const int bq = row_ok ? row / p.H : 0;      // (sequence, query) pair: indexes q, q_pos, mask, out
  const int h = row_ok ? row % p.H : 0;
  const int bi = bq / p.q_len;                 // sequence: indexes the KV cache
  const int kvh = h / p.n_rep;
  const bool live = row_ok && (p.head_mask == nullptr || p.head_mask[h] != 0);

  // the cached tokens are cut into n_splits ranges of tok_per_split; the appended token rides with the last one
  int t0 = split * p.tok_per_split;
  int t1c = min(t0 + p.tok_per_split, p.kv_len);
  if (!live || t0 >= t1c) { t0 = 0; t1c = 0; }
  const int last_split = p.kv_len > 0 ? (p.kv_len - 1) / p.tok_per_split : 0;
  const bool owns_new = live && p.append && split == last_split;   // token index kv_len (== t1c) is ours
  // Who stores the appended K/V row of a kv head into the cache (modify_llama.py:95-100)?  The first LIVE query head of
  // the GQA group -- and when head pruning has switched the whole group off, the warp that owns the dead row of the
  // group's first head copies it straight from k_new/v_new, so later steps never find a hole in the cache.
  bool new_writer = (h % p.n_rep) == 0;
  if (p.head_mask != nullptr && row_ok && p.append && split == last_split) {
    int first_live = -1;
    for (int r = p.n_rep - 1; r >= 0; --r)
      if (p.head_mask[kvh * p.n_rep + r] != 0) first_live = kvh * p.n_rep + r;
    new_writer = h == first_live;
    if (!live && first_live < 0 && (h % p.n_rep) == 0) {
      constexpr int kChunksA = D * 2 / 16;
#pragma unroll 1
      for (int w = lane / kChunksA; w < 2; w += 32 / kChunksA) {
        const int ch = lane % kChunksA;
        const __half* src = (w ? p.v_new : p.k_new) + (static_cast<int64_t>(bi) * p.H_kv + kvh) * D + ch * 8;
        __half* dst = (w ? p.v_cache : p.k_cache) + bi * p.stride_b + kvh * p.stride_h +
                      static_cast<int64_t>(p.kv_len) * D + ch * 8;
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
      }
    }
  }
  if (owns_new && t1c == 0) t0 = p.kv_len;                         // kv_len == 0: the range is just the new token
  const int t_end = owns_new ? p.kv_len + 1 : t1c;
  const int nt = t_end > t0 ? (t_end - t0 + kTile - 1) / kTile : 0;
  STAMP(0);

  const __half* Kc = p.k_cache + bi * p.stride_b + kvh * p.stride_h;
  const __half* Vc = p.v_cache + bi * p.stride_b + kvh * p.stride_h;

  // RoPE rows of one tile: cos | sin for the 4 passes of this lane's key group.  Rows are shared by all
  // heads, so the other warps of the CTA that work on the same token range hit L1.
  auto load_table = [&](int tok0, HV (&cs)[kPasses], HV (&sn)[kPasses]) {
#pragma unroll
    for (int ps = 0; ps < kPasses; ++ps) {
      const int j = min(tok0 + ps * 4 + g, p.rope_len - 1);
      const __half* r = p.rope + static_cast<int64_t>(j) * D + c * E2;
      *reinterpret_cast<Vec*>(&cs[ps]) = __ldg(reinterpret_cast<const Vec*>(r));
      *reinterpret_cast<Vec*>(&sn[ps]) = __ldg(reinterpret_cast<const Vec*>(r + D / 2));
    }
  };

  // ---- issue every independent global load first: query, its RoPE row, RoPE rows of tile 0 --------------
  HV qa, qb, qcs, qsn;
  auto load_query = [&]() {
    int64_t pos = p.q_pos ? p.q_pos[bq] : p.kv_len;
    pos = max(static_cast<int64_t>(0), min(pos, static_cast<int64_t>(p.rope_len - 1)));
    const __half* qr = p.q + static_cast<int64_t>(row_ok ? row : 0) * D;
    *reinterpret_cast<Vec*>(&qa) = *reinterpret_cast<const Vec*>(qr + c * E2);
    *reinterpret_cast<Vec*>(&qb) = *reinterpret_cast<const Vec*>(qr + D / 2 + c * E2);
    *reinterpret_cast<Vec*>(&qcs) = __ldg(reinterpret_cast<const Vec*>(p.rope + pos * D + c * E2));
    *reinterpret_cast<Vec*>(&qsn) = __ldg(reinterpret_cast<const Vec*>(p.rope + pos * D + D / 2 + c * E2));
  };
  const bool late_query = p.pdl == 2 || p.wait_n > 0;
  if (!late_query) load_query();
  HV cs[kPasses], sn[kPasses];
  if (nt > 0) load_table(t0, cs, sn);

  // ---- start the K/V stream ---------------------------------------------------------------------------
  const uint64_t pol = policy_evict_first();
  auto issue = [&](int i) {
    const int s = i % kStages;
    const int tok = t0 + i * kTile;
    const int ncopy = min(kTile, t1c - tok);            // rows that come from the cache (the new token does not)
    if (ncopy <= 0) return;
    const uint32_t bytes = static_cast<uint32_t>(ncopy) * kRowBytes;
    mbar_expect_tx(&bars[s], p.scores_only ? bytes : 2 * bytes);
    bulk_g2s(wsm + s * 2 * kTileBytes, Kc + static_cast<int64_t>(tok) * D, bytes, &bars[s], pol);
    if (!p.scores_only)
      bulk_g2s(wsm + s * 2 * kTileBytes + kTileBytes, Vc + static_cast<int64_t>(tok) * D, bytes, &bars[s], pol);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
#pragma unroll 1
    for (int i = 0; i < min(nt, kStages); ++i) issue(i);
  }
  __syncwarp();
  if (late_query) {
    // K/V prefetch overlapped the previous kernel; everything it may have produced is read after this
    if (p.pdl == 2) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p.wait_n > 0) {
      // fused exchange, consumer side: the query depends on every rank's heads of the previous layer
      const unsigned want = *reinterpret_cast<volatile const unsigned*>(p.wait_expected) + 1;
      if (lane < p.wait_n) {
        const long long tw = clock64();
        unsigned v;
        do {
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.wait_flags + lane) : "memory");
          if (clock64() - tw > 4000000000ll) { atomicExch(p.wait_timeout, 1u); break; }
        } while (static_cast<int>(v - want) < 0);
      }
      __syncwarp();
    }
    load_query();
  }

  // ---- rotate the query at its position id (modify_llama.py:92), fp16 op-for-op; kept as fp16 (exactly the values
  //      the reference forms) and fed to the mixed-precision FMA below -----------------------------------------------
  HV qlr, qhr;
#pragma unroll
  for (int i = 0; i < NH2; ++i) {
    qlr.h[i] = __hsub2_rn(__hmul2_rn(qa.h[i], qcs.h[i]), __hmul2_rn(qb.h[i], qsn.h[i]));
    qhr.h[i] = __hadd2_rn(__hmul2_rn(qb.h[i], qcs.h[i]), __hmul2_rn(qa.h[i], qsn.h[i]));
  }
  STAMP(1);   // query rotated

  float m = -CUDART_INF_F, l = 0.f;
  float accl[E2], acch[E2];
#pragma unroll
  for (int i = 0; i < E2; ++i) accl[i] = acch[i] = 0.f;

  __half* lg = p.logits ? p.logits + bi * p.logits_stride_b + h * p.logits_stride + (bq % p.q_len) * p.logits_stride_q
                        : nullptr;
  const __half* mk = p.mask ? p.mask + static_cast<int64_t>(bq) * p.n_total : nullptr;

  // ---- stream the tokens of the range ---------------------------------------------------------------------
#pragma unroll 1
  for (int i = 0; i < nt; ++i) {
    const int s = i % kStages;
    const int tok0 = t0 + i * kTile;
    const int ntok = min(kTile, t_end - tok0);          // valid tokens of the tile (incl. the appended one)
    const int ncopy = max(0, min(kTile, t1c - tok0));   // of which this many were bulk-copied from the cache
    uint8_t* Ks = wsm + s * 2 * kTileBytes;
    uint8_t* Vs = Ks + kTileBytes;

    // RoPE rows of the NEXT tile: in flight while this tile is consumed
    HV csn[kPasses], snn[kPasses];
    if (i + 1 < nt) load_table(tok0 + kTile, csn, snn);

    if (ntok < kTile || ncopy < ntok) {
      // zero the rows the copy does not write, so masked lanes multiply finite numbers
#pragma unroll 1
      for (int o = ncopy * kRowBytes + lane * 16; o < kTileBytes; o += 32 * 16) {
        *reinterpret_cast<uint4*>(Ks + o) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(Vs + o) = make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
    }
    if (ncopy > 0) mbar_wait(&bars[s], (i / kStages) & 1);
    if (i == 0) STAMP(2);   // first tile landed
    if (ncopy < ntok) {
      // the appended token (modify_llama.py:95-100): its un-rotated K/V row goes into the tile and into slot kv_len
      constexpr int kChunks = kRowBytes / 16;           // 16-byte chunks per row: 16 (D=128) or 8 (D=64)
      const int r = p.kv_len - tok0;
      const int which = lane / kChunks, ch = lane % kChunks;   // D=128: lanes 0-15 copy K, 16-31 copy V
#pragma unroll 1
      for (int w = which; w < 2; w += 32 / kChunks) {
        const __half* src = (w ? p.v_new : p.k_new) + (static_cast<int64_t>(bi) * p.H_kv + kvh) * D + ch * 8;
        const uint4 val = *reinterpret_cast<const uint4*>(src);
        *reinterpret_cast<uint4*>((w ? Vs : Ks) + r * kRowBytes + ch * 16) = val;
        if (new_writer) {                                // one writer per kv head
          __half* dst = (w ? p.v_cache : p.k_cache) + bi * p.stride_b + kvh * p.stride_h +
                        static_cast<int64_t>(p.kv_len) * D + ch * 8;
          *reinterpret_cast<uint4*>(dst) = val;
        }
      }
      __syncwarp();
    }

    // scores of the tile: rotated key rounded to fp16 after every elementwise op (:27), products summed in fp32
    float d[kPasses];
#pragma unroll
    for (int ps = 0; ps < kPasses; ++ps) {
      const int tl = ps * 4 + g;
      HV kl, kh;
      *reinterpret_cast<Vec*>(&kl) = *reinterpret_cast<const Vec*>(Ks + tl * kRowBytes + c * VB);
      *reinterpret_cast<Vec*>(&kh) = *reinterpret_cast<const Vec*>(Ks + tl * kRowBytes + D + c * VB);
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int e = 0; e < NH2; ++e) {
        const __half2 rl = __hsub2_rn(__hmul2_rn(kl.h[e], cs[ps].h[e]), __hmul2_rn(kh.h[e], sn[ps].h[e]));
        const __half2 rh = __hadd2_rn(__hmul2_rn(kh.h[e], cs[ps].h[e]), __hmul2_rn(kl.h[e], sn[ps].h[e]));
        d0 = fhfma_lo(rl, qlr.h[e], d0);   // FHFMA: fp16 x fp16 product is exact in fp32, so these sums are bit-identical
        d0 = fhfma_hi(rl, qlr.h[e], d0);   // to fmaf(float(q), float(k), d) -- without the 2 x 64 conversions per tile
        d1 = fhfma_lo(rh, qhr.h[e], d1);
        d1 = fhfma_hi(rh, qhr.h[e], d1);
      }
      d[ps] = d0 + d1;
    }
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
      for (int ps = 0; ps < kPasses; ++ps) d[ps] += __shfl_xor_sync(0xffffffffu, d[ps], o);
    }
    // fp16 logit exactly as the reference forms it: fp16(matmul) / sqrt(D) -> fp16 (:111-113); mask (:132)
    float sc[kPasses];
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int ps = 0; ps < kPasses; ++ps) {
      const int tl = ps * 4 + g;
      const int j = tok0 + tl;
      const bool valid = tl < ntok;
      __half s16 = __float2half_rn(div_by_const(__half2float(__float2half_rn(d[ps])), p.sqrt_d, p.rsqrt_d));
      if (valid && c == 0 && lg) lg[j] = s16;
      if (mk && valid) s16 = __hadd_rn(s16, mk[j]);
      sc[ps] = valid ? __half2float(s16) : -CUDART_INF_F;
      mx = fmaxf(mx, sc[ps]);
    }
    // online softmax (:135): m is warp-uniform; l and acc are per key group until the end
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    const float m_new = fmaxf(m, mx);
    if (m_new > -CUDART_INF_F && !p.scores_only) {
      if (m_new != m) {                         // warp-uniform: rescale only when the running max moved
        const float corr = fast_exp2((m - m_new) * kLog2e);
        l *= corr;
#pragma unroll
        for (int e = 0; e < E2; ++e) { accl[e] *= corr; acch[e] *= corr; }
        m = m_new;
      }
      // probabilities enter P.V as fp16, as in the reference (softmax(...).to(fp16) @ V, :135-138); the running sum
      // uses the same rounded values, so the final normalisation is exact for what was accumulated
      __half2 ph[kPasses];
#pragma unroll
      for (int ps = 0; ps < kPasses; ++ps) {
        ph[ps] = __float2half2_rn(fast_exp2((sc[ps] - m) * kLog2e));
        l += __low2float(ph[ps]);
      }
      // P.V (:138)
#pragma unroll
      for (int ps = 0; ps < kPasses; ++ps) {
        const int tl = ps * 4 + g;
        HV vl, vh;
        *reinterpret_cast<Vec*>(&vl) = *reinterpret_cast<const Vec*>(Vs + tl * kRowBytes + c * VB);
        *reinterpret_cast<Vec*>(&vh) = *reinterpret_cast<const Vec*>(Vs + tl * kRowBytes + D + c * VB);
#pragma unroll
        for (int e = 0; e < NH2; ++e) {
          accl[2 * e] = fhfma_lo(vl.h[e], ph[ps], accl[2 * e]);
          accl[2 * e + 1] = fhfma_hi(vl.h[e], ph[ps], accl[2 * e + 1]);
          acch[2 * e] = fhfma_lo(vh.h[e], ph[ps], acch[2 * e]);
          acch[2 * e + 1] = fhfma_hi(vh.h[e], ph[ps], acch[2 * e + 1]);
        }
      }
    }

    __syncwarp();                                  // every lane is done reading stage s
    if (lane == 0 && i + kStages < nt) issue(i + kStages);
#pragma unroll
    for (int ps = 0; ps < kPasses; ++ps) { cs[ps] = csn[ps]; sn[ps] = snn[ps]; }
  }
  STAMP(3);   // streaming loop done
  if (p.scores_only) {
    if constexpr (MODE == 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");   // pairs with the entry arrive
    return;
  }

  // ---- merge the 4 key groups of the warp -------------------------------------------------------------
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
  for (int i = 0; i < E2; ++i) {
    accl[i] += __shfl_xor_sync(0xffffffffu, accl[i], 8);
    accl[i] += __shfl_xor_sync(0xffffffffu, accl[i], 16);
    acch[i] += __shfl_xor_sync(0xffffffffu, acch[i], 8);
    acch[i] += __shfl_xor_sync(0xffffffffu, acch[i], 16);
  }

  if constexpr (MODE == 1) {
    // ---- cluster mode: partials never leave the chip ----------------------------------------------------
    // Row `rl` of the cluster is finished by the cluster/rpc CTAs of rank slice*rpc + rl, each owning a DS-wide
    // slice of D.  Every warp PUSHES its partial into the receive buffers of those CTAs (st.shared::cluster,
    // fire and forget), one cluster barrier publishes them, and each CTA then merges from its own shared memory
    // and exits -- no remote loads, no second barrier.
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");   // every CTA of the cluster is running (entry arrive)
    const int ds_sh = p.ds_shift;                       // DS = D * rpc / cluster is a power of two
    const int DS = 1 << ds_sh;                          // dims per merging CTA
    float* recv = reinterpret_cast<float*>(smem + kWarps * kWarpSmem + kWarps * kStages * 8);  // [n_splits][DS + 2]
    {
      const int rl = warp % p.rpc;
      // after the xor-8/16 butterflies every key group holds the full sums: group g pushes 2 of the lane's 8+8 dims
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int e = 0; e < E2; ++e) {                                     // static register indices only
          if (e / (E2 / 4) == g) {
            const int dim = half * (D / 2) + c * E2 + e;
            const float val = half ? acch[e] : accl[e];
            const unsigned dst_cta = (dim >> ds_sh) * p.rpc + rl;
            st_dsmem(recv + split * (DS + 2) + (dim & (DS - 1)), dst_cta, val);
          }
        }
      }
      if (lane < p.cluster / p.rpc) {                                      // (m, l) to every CTA that merges this row
        const unsigned dst_cta = lane * p.rpc + rl;
        st_dsmem(recv + split * (DS + 2) + DS, dst_cta, m);
        st_dsmem(recv + split * (DS + 2) + DS + 1, dst_cta, l);
      }
    }
    STAMP(4);
    cluster_sync_all();
    STAMP(5);
    {
      const int rl = range % p.rpc, slice = range / p.rpc;
      const int mrow = rg * p.rpc + rl;
      const int DW = DS / kWarps;                        // dims per warp: 2..16
      if (mrow < p.rows) {
        // lanes = splits: lane t owns split t (and t + 32)
        float mj[2], lj[2], wj[2];
        float M = -CUDART_INF_F;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int j = lane + 32 * t;
          const bool ok = j < p.n_splits;
          mj[t] = ok ? recv[j * (DS + 2) + DS] : -CUDART_INF_F;
          lj[t] = ok ? recv[j * (DS + 2) + DS + 1] : 0.f;
          M = fmaxf(M, mj[t]);
        }
        M = warp_max(M);
        float L = 0.f;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          wj[t] = (M > -CUDART_INF_F) ? fast_exp2((mj[t] - M) * kLog2e) : 0.f;
          L = fmaf(lj[t], wj[t], L);
        }
        L = warp_sum(L);
        const float inv = L > 0.f ? 1.f / L : 0.f;
        float habs = 0.f;
#pragma unroll 1
        for (int d0 = 0; d0 < DW; d0 += 4) {
          float o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int dl = warp * DW + d0 + i;                             // dim inside the slice
            o[i] = 0.f;
            if (d0 + i < DW) {
              o[i] = wj[0] * recv[min(lane, p.n_splits - 1) * (DS + 2) + dl];
              if (p.n_splits > 32) o[i] = fmaf(wj[1], recv[min(lane + 32, p.n_splits - 1) * (DS + 2) + dl], o[i]);
            }
          }
#pragma unroll
          for (int sft = 16; sft > 0; sft >>= 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] += __shfl_xor_sync(0xffffffffu, o[i], sft);
          }
          if (lane == 0) {                                 // the butterflies left all four sums in every lane
            const __half2 x01 = __floats2half2_rn(o[0] * inv, o[1] * inv), x23 = __floats2half2_rn(o[2] * inv, o[3] * inv);
            const int64_t off = static_cast<int64_t>(mrow) * D + slice * DS + warp * DW + d0;
            if (d0 + 4 <= DW) {                              // one 8-byte store per peer instead of four 2-byte ones
              uint2 pk;
              pk.x = *reinterpret_cast<const uint32_t*>(&x01);
              pk.y = *reinterpret_cast<const uint32_t*>(&x23);
              store_out(p, off, pk);
            } else {
              const __half xs[4] = {__low2half(x01), __high2half(x01), __low2half(x23), __high2half(x23)};
              for (int i = 0; d0 + i < DW; ++i) store_out(p, off + i, xs[i]);
            }
            const float2 f01 = __half22float2(x01), f23 = __half22float2(x23);
            habs += fabsf(f01.x) + fabsf(f01.y);
            if (d0 + 2 < DW) habs += fabsf(f23.x) + fabsf(f23.y);
          }
        }
        if (p.head_importance) {
          if (lane == 0) atomicAdd(p.head_importance + (mrow % p.H), habs);
        }
        if (slice == 0 && p.importance && p.logits && M > -CUDART_INF_F) {
          const int mbq = mrow / p.H;
          const __half* lgr = p.logits + (mbq / p.q_len) * p.logits_stride_b + (mrow % p.H) * p.logits_stride +
                              (mbq % p.q_len) * p.logits_stride_q;
          const __half* mkr = p.mask ? p.mask + static_cast<int64_t>(mbq) * p.n_total : nullptr;
          float* imp = p.importance + static_cast<int64_t>(mrow) * p.imp_stride;
#pragma unroll 1
          for (int j = threadIdx.x; j < p.n_total; j += kThreads) {
            __half s16 = __ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(lgr) + j));
            if (mkr) s16 = __hadd_rn(s16, mkr[j]);
            imp[j] += fast_exp2((__half2float(s16) - M) * kLog2e) * inv;
          }
        }
      }
    }
    STAMP(6);
    signal_peers<true>(p, lane);                         // every warp of every CTA reaches this point
    return;
  } else {
    if (p.n_splits == 1) {
      // single split: this warp holds the whole row
      if (row_ok) {
        const float inv = l > 0.f ? 1.f / l : 0.f;
        float habs = 0.f;
        if (g == 0) {
          HV ol, oh;
#pragma unroll
          for (int i = 0; i < NH2; ++i) {
            ol.h[i] = __floats2half2_rn(accl[2 * i] * inv, accl[2 * i + 1] * inv);
            oh.h[i] = __floats2half2_rn(acch[2 * i] * inv, acch[2 * i + 1] * inv);
            const float2 a = __half22float2(ol.h[i]), b = __half22float2(oh.h[i]);
            habs += fabsf(a.x) + fabsf(a.y) + fabsf(b.x) + fabsf(b.y);
          }
          store_out(p, static_cast<int64_t>(row) * D + c * E2, *reinterpret_cast<Vec*>(&ol));
          store_out(p, static_cast<int64_t>(row) * D + D / 2 + c * E2, *reinterpret_cast<Vec*>(&oh));
        }
        if (p.head_importance) {
          habs = warp_sum(habs);
          if (lane == 0) atomicAdd(p.head_importance + h, habs);
        }
        if (p.importance && lg && live) {
          __syncwarp();
          float* imp = p.importance + static_cast<int64_t>(row) * p.imp_stride;
#pragma unroll 1
          for (int j = lane; j < p.n_total; j += 32) {
            __half s16 = lg[j];
            if (mk) s16 = __hadd_rn(s16, mk[j]);
            imp[j] += fast_exp2((__half2float(s16) - m) * kLog2e) * inv;
          }
        }
        signal_peers<false>(p, lane);                    // one count per row
      }
      return;
    }

    if (row_ok && g == 0) {
      float* rec = p.part + (static_cast<int64_t>(row) * p.n_splits + split) * D;
#pragma unroll
      for (int i = 0; i < E2; i += 4) {
        *reinterpret_cast<float4*>(rec + c * E2 + i) = make_float4(accl[i], accl[i + 1], accl[i + 2], accl[i + 3]);
        *reinterpret_cast<float4*>(rec + D / 2 + c * E2 + i) = make_float4(acch[i], acch[i + 1], acch[i + 2], acch[i + 3]);
      }
      if (c == 0) p.part_ml[static_cast<int64_t>(row) * p.n_splits + split] = make_float2(m, l);
    }

    // ---- the last warp to finish a row merges that row's partials (per-row ticket) -----------------------
    STAMP(4);   // partial written
    __syncwarp();
    unsigned prev = 0;
    if (lane == 0 && row_ok) {
      __threadfence();                                   // release: the warp's partial is visible device-wide
      prev = atomicAdd(p.counters + row, 1u);
    }
    prev = __shfl_sync(0xffffffffu, prev, 0);
    STAMP(5);   // ticket taken
    if (!row_ok || prev != static_cast<unsigned>(p.n_splits - 1)) return;
    __threadfence();                                     // acquire side

    // the K/V ring is dead now: reuse the warp's slice for the split weights
    float* wts = reinterpret_cast<float*>(wsm);
    static_assert(kWarpSmem >= kMaxSplits * 4, "ring slice too small for the merge weights");
    const float* rec = p.part + static_cast<int64_t>(row) * p.n_splits * D;
    const float2* ml = p.part_ml + static_cast<int64_t>(row) * p.n_splits;
    constexpr int PER = D / 32;   // output dims per lane: 4 (D=128) or 2 (D=64)
    using PV_t = typename VecT<PER * 4>::type;
    // partial loads in flight per lane: one batch covers the usual <= 40 splits (the 128-register build takes two)
    constexpr int kU = kPasses == 2 ? 20 : 40;
    PV_t v[kU];
    // everything the merge needs is requested at once: the (m, l) pairs and the first batch of partials
    float2 mls[kMaxSplits / 32];
#pragma unroll
    for (int k = 0; k < kMaxSplits / 32; ++k) {
      const int sidx = lane + 32 * k;
      mls[k] = sidx < p.n_splits ? __ldcg(ml + sidx) : make_float2(-CUDART_INF_F, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u)
      v[u] = __ldcg(reinterpret_cast<const PV_t*>(rec + static_cast<int64_t>(min(u, p.n_splits - 1)) * D + lane * PER));
    float M = -CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < kMaxSplits / 32; ++k) M = fmaxf(M, mls[k].x);
    M = warp_max(M);
    float L = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxSplits / 32; ++k) {
      const int sidx = lane + 32 * k;
      const float w = (M > -CUDART_INF_F) ? fast_exp2((mls[k].x - M) * kLog2e) : 0.f;
      if (sidx < p.n_splits) wts[sidx] = w;
      L = fmaf(mls[k].y, w, L);
    }
    L = warp_sum(L);
    __syncwarp();
    STAMP(7);   // (m, l) of every split arrived
    float o[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) o[i] = 0.f;
#pragma unroll 1
    for (int s0 = 0; s0 < p.n_splits; s0 += kU) {
      if (s0 > 0) {
#pragma unroll
        for (int u = 0; u < kU; ++u)
          v[u] = __ldcg(reinterpret_cast<const PV_t*>(rec + static_cast<int64_t>(min(s0 + u, p.n_splits - 1)) * D + lane * PER));
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const float w = (s0 + u < p.n_splits) ? wts[s0 + u] : 0.f;
        const float* vf = reinterpret_cast<const float*>(&v[u]);
#pragma unroll
        for (int i = 0; i < PER; ++i) o[i] = fmaf(vf[i], w, o[i]);
      }
    }
    const float inv = L > 0.f ? 1.f / L : 0.f;
    __half oh16[PER];
    float habs = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      oh16[i] = __float2half_rn(o[i] * inv);
      habs += fabsf(__half2float(oh16[i]));
    }
    if (PER == 4) store_out(p, static_cast<int64_t>(row) * D + lane * PER, *reinterpret_cast<uint2*>(oh16));
    else store_out(p, static_cast<int64_t>(row) * D + lane * PER, *reinterpret_cast<uint32_t*>(oh16));
    if (p.head_importance) {
      habs = warp_sum(habs);
      if (lane == 0) atomicAdd(p.head_importance + h, habs);
    }
    if (p.importance && p.logits && M > -CUDART_INF_F) {
      float* imp = p.importance + static_cast<int64_t>(row) * p.imp_stride;
#pragma unroll 1
      for (int j = lane; j < p.n_total; j += 32) {
        __half s16 = __ushort_as_half(__ldcg(reinterpret_cast<const unsigned short*>(lg) + j));
        if (mk) s16 = __hadd_rn(s16, mk[j]);
        imp[j] += fast_exp2((__half2float(s16) - M) * kLog2e) * inv;
      }
    }
    STAMP(6);   // merged
    if (lane == 0) p.counters[row] = 0;         // self-cleaning ticket for the next launch
    signal_peers<false>(p, lane);               // one count per row
  }
#undef STAMP
}

// ---- host side ------------------------------------------------------------------------------------------

static void choose_split(int rows, int n_total, int sms, int& rpc, int& spc, int& n_rg, int& n_ranges) {
  rpc = rows >= 8 ? 8 : (rows >= 4 ? 4 : (rows >= 2 ? 2 : 1));
  spc = kWarps / rpc;
  n_rg = (rows + rpc - 1) / rpc;
  const int kMinTok = 16;        // do not cut ranges thinner than this per warp
  const int kOverheadTok = 24;   // fixed per-CTA cost (prologue, merge) in token-equivalents
  int max_ranges = (n_total + spc * kMinTok - 1) / (spc * kMinTok);
  const int rows_pad = n_rg * rpc;
  max_ranges = std::min(max_ranges, std::max(1, kMaxPartials / (rows_pad * spc)));
  max_ranges = std::max(1, std::min(max_ranges, kMaxSplits / spc));
  long best_cost = -1;
  n_ranges = 1;
  for (int nr = 1; nr <= max_ranges; ++nr) {
    const long ctas = static_cast<long>(n_rg) * nr;
    const long waves = (ctas + sms - 1) / sms;
    const int tps = (n_total + nr * spc - 1) / (nr * spc);
    const long cost = waves * (tps + kOverheadTok);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; n_ranges = nr; }
  }
}

// K/V rings + mbarriers + (cluster mode) the receive buffer for pushed partials: n_splits * (DS + 2) floats with
// n_splits * DS == 8 * D and n_splits <= 64
template <int D>
constexpr size_t decode_smem_bytes(int passes) {
  return static_cast<size_t>(kWarps) * kStages * 2 * (4 * passes) * D * 2 + kWarps * kStages * 8 +
         (static_cast<size_t>(kWarps) * D + 2 * 64) * sizeof(float);
}

using DecodeKernel = void (*)(const DecodeParams);

template <int D>
static DecodeKernel decode_kernel(int mode, bool dbg, int passes) {
  if (passes == 2) return mode ? decode_attn_kernel<D, 1, false, 2> : decode_attn_kernel<D, 0, false, 2>;
  if (D == 128 && dbg) return mode ? decode_attn_kernel<128, 1, true, 4> : decode_attn_kernel<128, 0, true, 4>;
  return mode ? decode_attn_kernel<D, 1, false, 4> : decode_attn_kernel<D, 0, false, 4>;
}

template <int D>
static int configure_decode(DecodeKernel k, int slot, int passes) {
  static bool configured[64][8] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !configured[dev][slot]) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(decode_smem_bytes<D>(passes)));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(decode_attn_kernel)");
    configured[dev][slot] = true;
  }
  return SPATTEN_OK;
}

// How many clusters of `cs` CTAs of the cluster-mode kernel the device can hold at once (GPC-granular).
template <int D>
static int max_active_clusters(int cs) {
  static int cached[64][kMaxClusterSize + 1] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && cached[dev][cs] != 0) return cached[dev][cs];
  DecodeKernel k = decode_kernel<D>(1, false, 4);
  if (configure_decode<D>(k, 1, 4) != SPATTEN_OK) return 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cs * 64);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = decode_smem_bytes<D>(4);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess) { cudaGetLastError(); n = -1; }
  if (dev < 64) cached[dev][cs] = n > 0 ? n : -1;
  return n;
}

template <int D>
static int launch_decode(const DecodeParams& p, cudaStream_t st) {
  const bool dbg = p.dbg != nullptr && D == 128 && p.passes == 4;
  const int mode = p.cluster ? 1 : 0;
  const size_t smem = decode_smem_bytes<D>(p.passes);
  DecodeKernel k = decode_kernel<D>(mode, dbg, p.passes);
  if (int rc = configure_decode<D>(k, mode + (dbg ? 2 : 0) + (p.passes == 2 ? 4 : 0), p.passes)) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.cluster ? p.n_rg * p.cluster : p.n_rg * p.n_ranges);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (p.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (p.cluster) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = p.cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, p);
  if (e != cudaSuccess) return cuda_fail(e, "decode_attn_kernel launch");
  return SPATTEN_OK;
}

}  // namespace spatten

using namespace spatten;

static unsigned long long* g_dbg = nullptr;
// Debug hook (not part of the public header): device buffer of [grid][8][8] u64 receiving %globaltimer stamps.
// Debug hook: how many 8-CTA clusters of the decode kernel the current device holds at once.
extern "C" int spatten_debug_max_clusters(int head_dim) {
  return (head_dim == 128 ? max_active_clusters<128>(8) : max_active_clusters<64>(8)) * 1000 +
         (head_dim == 128 ? max_active_clusters<128>(4) : max_active_clusters<64>(4));
}
extern "C" void spatten_debug_set_timeline(unsigned long long* dev_buf) { g_dbg = dev_buf; }

extern "C" size_t spatten_decode_workspace_bytes(int32_t batch, int32_t num_heads, int32_t head_dim) {
  if (batch <= 0 || num_heads <= 0 || head_dim <= 0) return 0;
  // fixed layout, independent of batch/heads, so one zero-initialised workspace serves every shape:
  // [tickets: kTicketBytes][(m, l): kMaxPartials float2][acc: kMaxPartials * head_dim floats]
  return kTicketBytes + static_cast<size_t>(kMaxPartials) * (head_dim + 2) * sizeof(float);
}

extern "C" int spatten_decode_attn(const SpattenDecodeArgs* a, void* stream) {
  SPATTEN_CHECK_ARG(a != nullptr, "spatten_decode_attn: args is NULL");
  SPATTEN_CHECK_ARG(a->struct_size == sizeof(SpattenDecodeArgs),
                    "spatten_decode_attn: struct_size %u != %zu (ABI mismatch)", a->struct_size,
                    sizeof(SpattenDecodeArgs));
  SPATTEN_CHECK_ARG(a->batch > 0 && a->num_heads > 0 && a->num_kv_heads > 0,
                    "spatten_decode_attn: batch/num_heads/num_kv_heads must be positive");
  SPATTEN_CHECK_ARG(a->num_heads % a->num_kv_heads == 0, "spatten_decode_attn: num_heads %d not a multiple of num_kv_heads %d",
                    a->num_heads, a->num_kv_heads);
  if (a->head_dim != 64 && a->head_dim != 128) {
    set_error("spatten_decode_attn: head_dim %d unsupported (64 or 128)", a->head_dim);
    return SPATTEN_ERR_UNSUPPORTED;
  }
  SPATTEN_CHECK_ARG(a->kv_len >= 0 && (a->append == 0 || a->append == 1), "spatten_decode_attn: bad kv_len/append");
  const int n_total = a->kv_len + a->append;
  const int q_len = a->q_len > 1 ? a->q_len : 1;
  SPATTEN_CHECK_ARG(q_len == 1 || (a->append == 0 && a->q_pos != nullptr && a->importance == nullptr),
                    "spatten_decode_attn: q_len > 1 needs append == 0, q_pos, and no importance accumulation");
  SPATTEN_CHECK_ARG(q_len == 1 || !a->logits || a->logits_stride_q >= n_total,
                    "spatten_decode_attn: logits_stride_q < kv_len");
  SPATTEN_CHECK_ARG(static_cast<int64_t>(a->batch) * q_len * a->num_heads < (1ll << 31), "spatten_decode_attn: too many rows");
  SPATTEN_CHECK_ARG(n_total > 0, "spatten_decode_attn: nothing to attend over (kv_len 0 and append 0)");
  SPATTEN_CHECK_ARG(a->rope_len >= n_total, "spatten_decode_attn: rope_len %d < kv_len+append %d", a->rope_len, n_total);
  SPATTEN_CHECK_ARG(a->q && a->k_cache && a->v_cache && a->rope, "spatten_decode_attn: NULL q/k_cache/v_cache/rope");
  SPATTEN_CHECK_ARG(a->out || a->n_peers > 0 || (a->flags & SPATTEN_DECODE_SCORES_ONLY), "spatten_decode_attn: NULL out");
  SPATTEN_CHECK_ARG(!a->append || (a->k_new && a->v_new), "spatten_decode_attn: append=1 needs k_new and v_new");
  SPATTEN_CHECK_ARG(a->cache_stride_h >= static_cast<int64_t>(n_total) * a->head_dim,
                    "spatten_decode_attn: cache_stride_h %lld too small for %d tokens", (long long)a->cache_stride_h, n_total);
  SPATTEN_CHECK_ARG(a->cache_stride_b >= a->cache_stride_h * a->num_kv_heads || a->batch == 1,
                    "spatten_decode_attn: cache_stride_b too small");
  SPATTEN_CHECK_ARG(a->cache_stride_h % 8 == 0 && a->cache_stride_b % 8 == 0,
                    "spatten_decode_attn: cache strides must be multiples of 8 elements (16-byte rows)");
  uintptr_t al = reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k_cache) |
                       reinterpret_cast<uintptr_t>(a->v_cache) | reinterpret_cast<uintptr_t>(a->rope) |
                       reinterpret_cast<uintptr_t>(a->out) | reinterpret_cast<uintptr_t>(a->k_new) |
                       reinterpret_cast<uintptr_t>(a->v_new);
  SPATTEN_CHECK_ARG((al & 15) == 0, "spatten_decode_attn: q/k_new/v_new/k_cache/v_cache/rope/out must be 16-byte aligned");
  SPATTEN_CHECK_ARG(!a->logits || a->logits_stride >= n_total, "spatten_decode_attn: logits_stride < kv_len+append");
  SPATTEN_CHECK_ARG(!a->importance || (a->logits && a->importance_stride >= n_total),
                    "spatten_decode_attn: importance needs logits and importance_stride >= kv_len+append");

  DecodeParams p{};
  p.q = static_cast<const __half*>(a->q);
  p.k_new = static_cast<const __half*>(a->k_new);
  p.v_new = static_cast<const __half*>(a->v_new);
  p.k_cache = static_cast<__half*>(a->k_cache);
  p.v_cache = static_cast<__half*>(a->v_cache);
  p.stride_b = a->cache_stride_b;
  p.stride_h = a->cache_stride_h;
  p.rope = static_cast<const __half*>(a->rope);
  p.q_pos = a->q_pos;
  p.mask = static_cast<const __half*>(a->mask);
  p.head_mask = a->head_mask;
  p.out = static_cast<__half*>(a->out);
  p.logits = static_cast<__half*>(a->logits);
  p.logits_stride = a->logits_stride;
  p.q_len = q_len;
  p.logits_stride_b = a->logits_stride_b ? a->logits_stride_b : a->logits_stride * a->num_heads;
  p.logits_stride_q = q_len > 1 ? a->logits_stride_q : 0;
  p.importance = a->importance;
  p.imp_stride = a->importance_stride;
  p.head_importance = a->head_importance;
  p.H = a->num_heads;
  p.H_kv = a->num_kv_heads;
  p.n_rep = a->num_heads / a->num_kv_heads;
  p.kv_len = a->kv_len;
  p.n_total = n_total;
  p.append = a->append;
  p.rope_len = a->rope_len;
  p.rows = a->batch * q_len * a->num_heads;
  p.sqrt_d = static_cast<float>(sqrt(static_cast<double>(a->head_dim)));
  p.rsqrt_d = 1.0f / p.sqrt_d;

  p.scores_only = (a->flags & SPATTEN_DECODE_SCORES_ONLY) ? 1 : 0;
  SPATTEN_CHECK_ARG(!p.scores_only || a->logits != nullptr, "spatten_decode_attn: SCORES_ONLY needs a logits buffer");
  p.pdl = 0;
  if (a->flags & SPATTEN_DECODE_PDL) p.pdl = 1;
  if ((a->flags & SPATTEN_DECODE_PDL_KV_STABLE) && (a->flags & SPATTEN_DECODE_PDL) && a->head_mask == nullptr) p.pdl = 2;
  p.dbg = g_dbg;

  const int sms = sm_count();
  if (sms <= 0) return SPATTEN_ERR_CUDA;
  choose_split(p.rows, n_total, sms, p.rpc, p.spc, p.n_rg, p.n_ranges);
  p.n_splits = p.n_ranges * p.spc;
  p.cluster = 0;
  // Cluster mode: all splits of a row inside one thread-block cluster (8 or 4 CTAs), merged over DSMEM -- no
  // global partials, no tickets.  Taken when some (cluster size, rows per CTA) fills >= 3/4 of the SMs in ONE
  // wave of clusters and every warp still gets >= 16 cached tokens; the candidate with the most CTAs wins, ties
  // go to more rows per CTA (more RoPE-row sharing).  Big-row-count shapes stay in ticket mode.
  if (!(a->flags & SPATTEN_DECODE_NO_CLUSTER)) {
    int best_ctas = 0;
    for (int cs = kMaxClusterSize; cs >= 4; cs >>= 1) {
      const int max_cl = a->head_dim == 128 ? max_active_clusters<128>(cs) : max_active_clusters<64>(cs);
      for (int rpc = 8; rpc >= 1 && max_cl > 0; rpc >>= 1) {
        const int n_cl = (p.rows + rpc - 1) / rpc;
        const int spc = kWarps / rpc;
        const int ctas = n_cl * cs;
        if (rpc > cs) continue;                          // every row is finished by cs / rpc >= 1 CTAs of its cluster
        if (n_cl > max_cl || ctas > sms || ctas * 4 < sms * 3 || ctas <= best_ctas) continue;
        if (a->kv_len < cs * spc * 16) continue;
        best_ctas = ctas;
        p.cluster = cs;
        p.rpc = rpc; p.spc = spc; p.n_rg = n_cl; p.n_ranges = cs;
        p.n_splits = cs * spc;
        p.ds_shift = 0;
        while ((1 << p.ds_shift) < a->head_dim * rpc / cs) ++p.ds_shift;
      }
    }
  }
  p.tok_per_split = (std::max(a->kv_len, 1) + p.n_splits - 1) / p.n_splits;
  // The small-footprint build is opt-in: measured on B200 (profiles/r1_microbench_decode_v11.jsonl) two co-resident
  // CTAs of consecutive PDL launches compete more than they overlap (config 2: 14.5 us vs 10.5 us per layer).
  p.passes = (a->flags & SPATTEN_DECODE_SMALL_TILES) ? 2 : 4;
  p.n_peers = 0;
  if (a->n_peers > 0 && !p.scores_only) {
    SPATTEN_CHECK_ARG(a->n_peers <= kMaxPeers && a->done_counter != nullptr, "spatten_decode_attn: bad n_peers / done_counter");
    for (int r = 0; r < a->n_peers; ++r) {
      SPATTEN_CHECK_ARG(a->peer_out[r] && a->peer_flags[r], "spatten_decode_attn: NULL peer_out / peer_flags entry");
      p.peer_out[r] = static_cast<__half*>(a->peer_out[r]);
      p.peer_flag[r] = static_cast<unsigned int*>(a->peer_flags[r]);
    }
    p.n_peers = a->n_peers;
    p.done = static_cast<unsigned int*>(a->done_counter);
    p.done_target = p.cluster ? static_cast<unsigned>(p.n_rg) * p.cluster : static_cast<unsigned>(p.rows);
    p.wait_n = 0;
    if (a->wait_n > 0) {
      SPATTEN_CHECK_ARG(a->wait_n <= 32 && a->wait_flags && a->wait_expected && a->wait_timeout,
                        "spatten_decode_attn: wait_n needs wait_flags, wait_expected and wait_timeout");
      p.wait_flags = static_cast<const unsigned int*>(a->wait_flags);
      p.wait_expected = static_cast<unsigned int*>(a->wait_expected);
      p.wait_timeout = static_cast<unsigned int*>(a->wait_timeout);
      p.wait_n = a->wait_n;
    }
  }
  if (p.n_splits > 1 && !p.cluster) {
    const size_t need = spatten_decode_workspace_bytes(a->batch, a->num_heads, a->head_dim);
    if (a->workspace == nullptr || a->workspace_bytes < need) {
      set_error("spatten_decode_attn: workspace %zu bytes < required %zu", a->workspace_bytes, need);
      return SPATTEN_ERR_WORKSPACE;
    }
    // split rows are bounded by the partial budget (rows * n_splits <= kMaxPartials, n_splits >= 2)
    const size_t counters = kTicketBytes;
    p.counters = static_cast<unsigned int*>(a->workspace);
    p.part_ml = reinterpret_cast<float2*>(static_cast<uint8_t*>(a->workspace) + counters);
    p.part = reinterpret_cast<float*>(static_cast<uint8_t*>(a->workspace) + counters +
                                      static_cast<size_t>(kMaxPartials) * sizeof(float2));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return a->head_dim == 128 ? launch_decode<128>(p, st) : launch_decode<64>(p, st);
}

// Same call with HOST buffers for the per-step tensors: copies q / k_new / v_new from (pinned) host memory into
// the device staging buffers named by `args`, launches, copies `out` back and synchronises the stream.
extern "C" int spatten_decode_attn_host(const SpattenDecodeArgs* a, const void* q_host, const void* k_new_host,
                                        const void* v_new_host, void* out_host, void* stream) {
  SPATTEN_CHECK_ARG(a != nullptr && a->struct_size == sizeof(SpattenDecodeArgs), "spatten_decode_attn_host: bad args");
  SPATTEN_CHECK_ARG(q_host && out_host && a->q && a->out, "spatten_decode_attn_host: NULL q/out");
  SPATTEN_CHECK_ARG(!a->append || (k_new_host && v_new_host && a->k_new && a->v_new),
                    "spatten_decode_attn_host: append needs k_new/v_new host and staging buffers");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int q_len = a->q_len > 1 ? a->q_len : 1;
  const size_t qb = static_cast<size_t>(a->batch) * q_len * a->num_heads * a->head_dim * sizeof(__half);
  const size_t kb = static_cast<size_t>(a->batch) * a->num_kv_heads * a->head_dim * sizeof(__half);
  // one packed copy when the caller laid q | k_new | v_new out back to back on both sides (saves two driver calls)
  const uint8_t* qh8 = static_cast<const uint8_t*>(q_host);
  const uint8_t* qd8 = static_cast<const uint8_t*>(a->q);
  const bool packed = a->append && k_new_host == qh8 + qb && v_new_host == qh8 + qb + kb &&
                      a->k_new == qd8 + qb && a->v_new == qd8 + qb + kb;
  cudaError_t e = cudaMemcpyAsync(const_cast<void*>(a->q), q_host, packed ? qb + 2 * kb : qb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && a->append && !packed) e = cudaMemcpyAsync(const_cast<void*>(a->k_new), k_new_host, kb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && a->append && !packed) e = cudaMemcpyAsync(const_cast<void*>(a->v_new), v_new_host, kb, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(e, "spatten_decode_attn_host: H2D copy");
  const int rc = spatten_decode_attn(a, stream);
  if (rc != SPATTEN_OK) return rc;
  e = cudaMemcpyAsync(out_host, a->out, qb, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "spatten_decode_attn_host: D2H copy / sync");
  return SPATTEN_OK;
}

// ---- consumer side of the fused exchange ------------------------------------------------------------------------
namespace spatten {
__global__ void wait_flags_kernel(const unsigned int* flags, int n, unsigned int* expected, unsigned int* timed_out) {
  __shared__ unsigned int s_exp;
  if (threadIdx.x == 0) s_exp = *expected + 1;
  __syncthreads();
  const unsigned int want = s_exp;
  if (threadIdx.x < n) {
    const long long t0 = clock64();
    unsigned int v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
      if (clock64() - t0 > 4000000000ll) { atomicExch(timed_out, 1u); break; }   // ~2 s: never wedge the GPU
    } while (static_cast<int>(v - want) < 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) *expected = want;
}
}  // namespace spatten

extern "C" int spatten_wait_flags(const void* flags, int32_t n, void* expected_counter, void* timeout_flag, void* stream) {
  SPATTEN_CHECK_ARG(flags && expected_counter && timeout_flag && n > 0 && n <= 32, "spatten_wait_flags: bad arguments");
  spatten::wait_flags_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const unsigned int*>(flags), n, static_cast<unsigned int*>(expected_counter),
      static_cast<unsigned int*>(timeout_flag));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : spatten::cuda_fail(e, "wait_flags_kernel launch");
}
