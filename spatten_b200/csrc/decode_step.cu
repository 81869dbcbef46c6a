// Decode STEP kernel -- spatten_decode_step(): the attention core of every layer of one decode token in ONE
// persistent launch.
//
// Per layer the function is the one spatten_decode_attn computes (reference
// spatten_llm/pos_shift/modify_llama.py:86-144 for q_len == 1: RoPE(Q) :92, KV append :95-100, RoPE(K at slot
// index) :103-104, Q.K^T/sqrt(D) :111-113, logit stash :116-119, fp32 softmax :135, P.V :138); the per-tile
// arithmetic below is the same instruction sequence, so exported logits are bit-identical between the two.
//
// Why a second kernel.  At BASELINE config 2 one layer is 32 MiB = 5.2 us of HBM time; a per-layer launch pays
// launch + first-byte latency + merge tail (about 5 us) every layer and cannot overlap them
// with the next launch because one CTA owns the SM's whole shared memory.  Here one CTA per SM stays resident for
// the whole step and the K/V stream never stops:
//   * work list of a layer = (live row, 16-token tile) pairs in row-major order; it is cut into W equal contiguous
//     ranges (W = warps in the grid), rotated by one warp per layer so the 3-vs-4-tile remainder evens out;
//   * every warp owns a 3-stage ring of 16-token K+V tiles fed by 1-D bulk copies (cp.async.bulk -> mbarrier
//     complete_tx).  The ISSUE cursor runs three tiles ahead of the CONSUME cursor and simply walks on into the next
//     layer's range, so layer l+1's rows are landing in shared memory while layer l is merged;
//   * a range may cover pieces of several rows ("segments").  A finished segment is reduce-scattered across the
//     warp's 4 key groups (12 shuffles), written as one fp32 partial, and a per-(layer, row) ticket elects the last
//     warp of the row, which merges the row's partials, stores the output row and bumps the layer's flag;
//   * SPATTEN_STEP_CHAIN: the query of layer l+1 is read only after layer l's flag says every row is stored (the
//     decoder's dependency); only the query load and the arithmetic wait, the K/V prefetch does not.
//   * head sharding: the merged row is stored into every peer GPU's gather buffer by the whole warp (8-byte
//     vectors) and lanes 0..N-1 each release one peer's per-layer flag -- no grid-wide counter, no serial loop.
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "../../include/spatten_b200.h"
#include "common.cuh"

namespace spatten {

constexpr int kStepMaxPeers = 8;
constexpr int kStepRot = 1;          // work-range rotation per layer (warps)
constexpr int kStepCtrlWords = 64;   // workspace header: [0] steps completed, [1] exit ticket, [2] wait timed out
constexpr int kStepSmemLayers = 128; // layer descriptors staged in shared memory when the model has at most this many

struct StepParams {
  const SpattenStepLayer* layers;
  int L, b, H, H_kv, n_rep, kv_len, n_total, append, rope_len;
  int64_t stride_b, stride_h, logits_stride, imp_stride;
  const __half* rope;
  const int64_t* q_pos;
  const uint8_t* head_mask;
  float* head_importance;
  int tpr;                 // tiles per row
  int W;                   // warps in the grid
  int slots;               // partial slots per layer
  unsigned int rows;       // b * H: flag increments per layer per step
  unsigned int* ctrl;
  unsigned int* tickets;   // [L][rows]
  float2* part_ml;         // [L][slots]
  float* part;             // [L][slots][D]
  int n_out;               // 0: layers[l].out; > 0: every peer's gather buffer
  __half* peer_out[kStepMaxPeers];
  int64_t peer_layer_stride;
  int n_flag;              // flag words bumped per finished row
  unsigned int* flag_ptr[kStepMaxPeers];
  int64_t flag_layer_stride;
  const unsigned int* wait_flags;   // [L][wait_n]
  int wait_n;
  int chain;               // wait for layer l-1's flags before reading layer l's query
  int final_wait;          // wait for the last layer's flags before the kernel exits
  int sys_scope;           // flags live in peer memory
  float sqrt_d, rsqrt_d;
};

// Work-list cursor.  Every lane keeps the same copy; set_layer() ballots, so it is called warp-uniformly.
struct Cursor {
  int l;                   // layer (>= L: exhausted)
  int t, t_end;            // this warp's tile range in the layer's live tile list
  int j;                   // tile inside the current row
  int i;                   // live row index
  int bi, h;               // sequence and head of the current row
  int first;               // 1 until the first tile of the range has been passed
  int nlh;                 // live heads of the layer
  unsigned long long hm;   // live head bits of the layer (valid when head_mask != nullptr)
};

__device__ __forceinline__ int nth_set_bit64(unsigned long long m, int n) {
  const unsigned lo = static_cast<unsigned>(m);
  const int cl = __popc(lo);
  if (n < cl) return __fns(lo, 0, n + 1);
  return 32 + __fns(static_cast<unsigned>(m >> 32), 0, n - cl + 1);
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* ptr, int sys) {
  unsigned v;
  if (sys) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  else asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}

// arm the stage's mbarrier and start the K and V bulk copies -- executed by the lane for which `pred` is set, without
// a divergent branch
__device__ __forceinline__ void issue_kv_copies(bool pred, uint32_t k_dst, uint32_t v_dst, const void* k_src, const void* v_src,
                                                uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t"
      "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %2;\n\t"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%3], [%4], %5, [%1], %6;\n\t"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%7], [%8], %5, [%1], %6;\n\t}"
      ::"r"(static_cast<uint32_t>(pred)), "r"(bar), "r"(2 * bytes), "r"(k_dst), "l"(k_src), "r"(bytes), "l"(policy), "r"(v_dst),
        "l"(v_src)
      : "memory");
}

template <int D, int WARPS, int STAGES, int PASSES>
__global__ void __launch_bounds__(WARPS * 32, 1) decode_step_kernel(const StepParams p) {
  constexpr int kTile = 4 * PASSES;   // tokens per stage: PASSES passes of 4 keys x 8 lanes
  constexpr int E2 = D / 16;   // halves per lane in each half of the row (8 lanes cover D/2)
  constexpr int NH2 = E2 / 2;  // half2 per lane per half-row
  constexpr int VB = E2 * 2;   // bytes per lane per half-row
  constexpr int PER = D / 32;  // output dims a lane owns after the reduce-scatter / in the merge
  using Vec = typename VecT<VB>::type;
  using HV = HalfVec<NH2>;
  using PV_t = typename VecT<PER * 4>::type;
  constexpr int kRowBytes = D * 2;
  constexpr int kTileBytes = kTile * kRowBytes;
  constexpr int kStageBytes = 2 * kTileBytes;
  constexpr int kWarpSmem = STAGES * kStageBytes;

  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 3, c = lane & 7;
  uint8_t* wsm = smem + warp * kWarpSmem;
  const uint32_t wsm_u32 = smem_u32(wsm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WARPS * kWarpSmem) + warp * STAGES;
  const uint32_t bars_u32 = smem_u32(bars);
  const int gw = blockIdx.x * WARPS + warp;

  // layer descriptors: 64 bytes each, read at every segment start / row change -- keep them on chip
  const SpattenStepLayer* lys = p.layers;
  if (p.L <= kStepSmemLayers) {
    uint4* dst = reinterpret_cast<uint4*>(smem + WARPS * kWarpSmem + WARPS * STAGES * 8);
    const uint4* src = reinterpret_cast<const uint4*>(p.layers);
    for (int k = threadIdx.x; k < p.L * 4; k += WARPS * 32) dst[k] = src[k];
    lys = reinterpret_cast<const SpattenStepLayer*>(dst);
    __syncthreads();
  }

  const unsigned epoch = *reinterpret_cast<volatile const unsigned*>(p.ctrl);
  const unsigned flag_target = (epoch + 1u) * p.rows;

  // ---- cursors ------------------------------------------------------------------------------------------------
  auto set_row = [&](Cursor& cu) {                       // (bi, h) of live row cu.i
    cu.bi = cu.i / cu.nlh;
    const int hidx = cu.i - cu.bi * cu.nlh;
    cu.h = p.head_mask ? nth_set_bit64(cu.hm, hidx) : hidx;
  };
  auto set_layer = [&](Cursor& cu, int l) {
    cu.l = l; cu.t = 0; cu.t_end = 0; cu.i = 0; cu.j = 0; cu.first = 1; cu.nlh = p.H; cu.hm = ~0ull; cu.bi = 0; cu.h = 0;
    if (l >= p.L) return;
    if (p.head_mask) {
      const uint8_t* hr = p.head_mask + static_cast<size_t>(l) * p.H;
      const unsigned lo = __ballot_sync(0xffffffffu, lane < p.H && hr[lane] != 0);
      const unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < p.H && hr[lane + 32] != 0);
      cu.hm = static_cast<unsigned long long>(lo) | (static_cast<unsigned long long>(hi) << 32);
      cu.nlh = __popcll(cu.hm);
    }
    const unsigned T = static_cast<unsigned>(p.b * cu.nlh * p.tpr);   // host checked: T * W < 2^32
    const unsigned weff = T < static_cast<unsigned>(p.W) ? T : static_cast<unsigned>(p.W);
    const unsigned v = static_cast<unsigned>(gw + l * kStepRot) % static_cast<unsigned>(p.W);
    if (v < weff) {
      cu.t = static_cast<int>(v * T / weff);
      cu.t_end = static_cast<int>((v + 1) * T / weff);
      cu.i = cu.t / p.tpr;
      cu.j = cu.t - cu.i * p.tpr;
      set_row(cu);
    }
  };
  // slow path of advance(): the row or the range ended
  auto advance_slow = [&](Cursor& cu) {
    if (cu.t >= cu.t_end) {
      do set_layer(cu, cu.l + 1); while (cu.l < p.L && cu.t >= cu.t_end);
    } else {
      cu.j = 0;
      ++cu.i;
      set_row(cu);
    }
  };

  const uint64_t pol = policy_evict_first();
  // issue cursor: addresses of the tile it names
  Cursor ci;
  const __half* ik = nullptr;
  const __half* iv = nullptr;
  auto issue_ptrs = [&]() {
    if (ci.l >= p.L) return;
    const SpattenStepLayer* ly = lys + ci.l;
    const int64_t off = ci.bi * p.stride_b + (ci.h / p.n_rep) * p.stride_h + static_cast<int64_t>(ci.j) * (kTile * D);
    ik = static_cast<const __half*>(ly->k_cache) + off;
    iv = static_cast<const __half*>(ly->v_cache) + off;
  };
  auto issue_next = [&](int s) {                          // copy the issue cursor's tile into stage s, move it on
    if (ci.l >= p.L) return;
    const int ncopy = min(kTile, p.kv_len - ci.j * kTile);
    if (ncopy > 0)
      issue_kv_copies(lane == 0, wsm_u32 + s * kStageBytes, wsm_u32 + s * kStageBytes + kTileBytes, ik, iv,
                      static_cast<uint32_t>(ncopy) * kRowBytes, bars_u32 + s * 8, pol);
    ++ci.t; ++ci.j;
    ik += kTile * D; iv += kTile * D;
    if (ci.j == p.tpr || ci.t >= ci.t_end) { advance_slow(ci); issue_ptrs(); }
  };

  // RoPE rows of one tile: cos | sin for the 4 passes of this lane's key group (L2 resident table)
  auto load_table = [&](int tok0, HV (&cs)[PASSES], HV (&sn)[PASSES]) {
    if (tok0 + kTile <= p.rope_len) {
      const __half* r = p.rope + static_cast<int64_t>(tok0 + g) * D + c * E2;
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        *reinterpret_cast<Vec*>(&cs[ps]) = __ldg(reinterpret_cast<const Vec*>(r + ps * 4 * D));
        *reinterpret_cast<Vec*>(&sn[ps]) = __ldg(reinterpret_cast<const Vec*>(r + ps * 4 * D + D / 2));
      }
    } else {
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int j = min(tok0 + ps * 4 + g, p.rope_len - 1);
        const __half* r = p.rope + static_cast<int64_t>(j) * D + c * E2;
        *reinterpret_cast<Vec*>(&cs[ps]) = __ldg(reinterpret_cast<const Vec*>(r));
        *reinterpret_cast<Vec*>(&sn[ps]) = __ldg(reinterpret_cast<const Vec*>(r + D / 2));
      }
    }
  };

  auto wait_layer = [&](int layer) {
    if (lane < p.wait_n) {
      const unsigned* f = p.wait_flags + static_cast<int64_t>(layer) * p.wait_n + lane;
      const long long tw = clock64();
      while (static_cast<int>(ld_acquire(f, p.sys_scope) - flag_target) < 0) {
        if (clock64() - tw > 4000000000ll) { atomicExch(p.ctrl + 2, 1u); break; }   // ~2 s: never wedge the GPU
      }
    }
    __syncwarp();
  };

  // every lane stored its piece of a row of layer `layer`: publish it
  auto signal_row = [&](int layer) {
    if (p.n_flag == 0) return;
    __syncwarp();
    if (lane < p.n_flag) {
      unsigned int* f = p.flag_ptr[lane] + static_cast<int64_t>(layer) * p.flag_layer_stride;
      if (p.sys_scope) asm volatile("fence.acq_rel.sys;\n\tred.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
      else asm volatile("fence.acq_rel.gpu;\n\tred.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
    }
  };

  // lane owns dims [d0, d0 + PER) of the finished row: normalise, round, store (locally or to every peer), side outputs
  auto emit_row = [&](int layer, int row, int d0, const float (&o)[PER], float inv, float M) {
    __half oh16[PER];
    float habs = 0.f;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      oh16[k] = __float2half_rn(o[k] * inv);
      habs += fabsf(__half2float(oh16[k]));
    }
    const int64_t off = static_cast<int64_t>(row) * D + d0;
    if (p.n_out == 0) {
      __half* dst = static_cast<__half*>(lys[layer].out) + off;
      if (PER == 4) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(oh16);
      else *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(oh16);
    } else {
#pragma unroll 1
      for (int r = 0; r < p.n_out; ++r) {
        __half* dst = p.peer_out[r] + static_cast<int64_t>(layer) * p.peer_layer_stride + off;
        if (PER == 4) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(oh16);
        else *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(oh16);
      }
    }
    if (p.head_importance) {
      habs = warp_sum(habs);
      if (lane == 0) atomicAdd(p.head_importance + row % p.H, habs);
    }
    const SpattenStepLayer* ly = lys + layer;
    if (ly->importance && ly->logits && M > -CUDART_INF_F) {
      const unsigned short* lgr = reinterpret_cast<const unsigned short*>(
          static_cast<const __half*>(ly->logits) + static_cast<int64_t>(row) * p.logits_stride);
      float* imp = ly->importance + static_cast<int64_t>(row) * p.imp_stride;
#pragma unroll 1
      for (int j = lane; j < p.n_total; j += 32) {
        const __half s16 = __ushort_as_half(__ldcg(lgr + j));
        imp[j] += fast_exp2((__half2float(s16) - M) * kLog2e) * inv;
      }
    }
    signal_row(layer);
  };

  // pruned heads of a layer: zero their output rows, keep the KV cache complete, count them as done
  auto dead_rows = [&](int layer, int nlh, unsigned long long hm) {
    const SpattenStepLayer* ly = lys + layer;
    const int ndh = p.H - nlh, n_dead = p.b * ndh;
    const unsigned T = static_cast<unsigned>(p.b * nlh * p.tpr);
    const int weff = static_cast<int>(T < static_cast<unsigned>(p.W) ? T : static_cast<unsigned>(p.W));
    const int v = (gw + layer * kStepRot) % p.W;
    const unsigned long long hall = p.H >= 64 ? ~0ull : ((1ull << p.H) - 1ull);
#pragma unroll 1
    for (int dr = v; dr < n_dead; dr += weff) {
      const int dbi = dr / ndh;
      const int dh = nth_set_bit64(~hm & hall, dr - dbi * ndh);
      const int dkvh = dh / p.n_rep;
      const unsigned long long grp = (hm >> (dkvh * p.n_rep)) & (p.n_rep >= 64 ? ~0ull : ((1ull << p.n_rep) - 1ull));
      if (p.append && grp == 0 && dh == dkvh * p.n_rep) {
        constexpr int kChunks = kRowBytes / 16;
        const int which = lane / kChunks, ch = lane % kChunks;
#pragma unroll 1
        for (int w = which; w < 2; w += 32 / kChunks) {
          const __half* src = static_cast<const __half*>(w ? ly->v_new : ly->k_new) +
                              (static_cast<int64_t>(dbi) * p.H_kv + dkvh) * D + ch * 8;
          __half* dst = static_cast<__half*>(w ? ly->v_cache : ly->k_cache) + dbi * p.stride_b + dkvh * p.stride_h +
                        static_cast<int64_t>(p.kv_len) * D + ch * 8;
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
        }
      }
      float z[PER];
#pragma unroll
      for (int k = 0; k < PER; ++k) z[k] = 0.f;
      emit_row(layer, dbi * p.H + dh, lane * PER, z, 0.f, -CUDART_INF_F);
    }
  };

  // ---- pipeline prologue ----------------------------------------------------------------------------------------
  Cursor cc;
  set_layer(cc, 0);
  while (cc.l < p.L && cc.t >= cc.t_end) set_layer(cc, cc.l + 1);
  ci = cc;
  issue_ptrs();
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncwarp();
#pragma unroll 1
  for (int s = 0; s < STAGES; ++s) issue_next(s);

  unsigned phase = 0;     // bit s = parity the next wait on stage s uses
  int s = 0;
  float ql[E2], qh[E2], accl[E2], acch[E2];
  float m = -CUDART_INF_F, l = 0.f;
  HV cs[PASSES], sn[PASSES];
  if (cc.l < p.L) load_table(cc.j * kTile, cs, sn);
  // segment constants (a segment = this warp's tiles of one row)
  int seg_l = 0, seg_row = 0, seg_i = 0, seg_nlh = p.H;
  bool seg_writer = false;
  const SpattenStepLayer* ly = lys;
  __half* lg = nullptr;

#pragma unroll 1
  while (cc.l < p.L) {
    const int cj = cc.j;
    if (cc.first | (cj == 0)) {
      // ---- segment start ------------------------------------------------------------------------------------------
      seg_l = cc.l; seg_i = cc.i; seg_nlh = cc.nlh;
      const int bi = cc.bi, h = cc.h;
      seg_row = bi * p.H + h;
      ly = lys + seg_l;
      if (cc.first) {
        // the decoder's dependency: layer l's query exists only when every row of layer l-1 has been stored
        if (seg_l > 0 && p.chain) wait_layer(seg_l - 1);
        if (p.head_mask && seg_nlh < p.H) dead_rows(seg_l, seg_nlh, cc.hm);
      }
      seg_writer = (h % p.n_rep) == 0;
      if (p.head_mask) {
        const int kvh = h / p.n_rep;
        const unsigned long long grp = (cc.hm >> (kvh * p.n_rep)) & (p.n_rep >= 64 ? ~0ull : ((1ull << p.n_rep) - 1ull));
        seg_writer = h == kvh * p.n_rep + (__ffsll(static_cast<long long>(grp)) - 1);   // first live head of the kv group
      }
      lg = ly->logits ? static_cast<__half*>(ly->logits) + static_cast<int64_t>(seg_row) * p.logits_stride : nullptr;
      // query of this row, rotated at its position id (modify_llama.py:92), fp16 op-for-op, kept as fp32
      int64_t pos = p.q_pos ? p.q_pos[bi] : p.kv_len;
      pos = max(static_cast<int64_t>(0), min(pos, static_cast<int64_t>(p.rope_len - 1)));
      const __half* qr = static_cast<const __half*>(ly->q) + static_cast<int64_t>(seg_row) * D;
      HV qa, qb, qcs, qsn;
      *reinterpret_cast<Vec*>(&qa) = *reinterpret_cast<const Vec*>(qr + c * E2);
      *reinterpret_cast<Vec*>(&qb) = *reinterpret_cast<const Vec*>(qr + D / 2 + c * E2);
      *reinterpret_cast<Vec*>(&qcs) = __ldg(reinterpret_cast<const Vec*>(p.rope + pos * D + c * E2));
      *reinterpret_cast<Vec*>(&qsn) = __ldg(reinterpret_cast<const Vec*>(p.rope + pos * D + D / 2 + c * E2));
#pragma unroll
      for (int k = 0; k < NH2; ++k) {
        const __half2 rl = __hsub2_rn(__hmul2_rn(qa.h[k], qcs.h[k]), __hmul2_rn(qb.h[k], qsn.h[k]));
        const __half2 rh = __hadd2_rn(__hmul2_rn(qb.h[k], qcs.h[k]), __hmul2_rn(qa.h[k], qsn.h[k]));
        const float2 fl = __half22float2(rl), fh = __half22float2(rh);
        ql[2 * k] = fl.x; ql[2 * k + 1] = fl.y;
        qh[2 * k] = fh.x; qh[2 * k + 1] = fh.y;
      }
      m = -CUDART_INF_F;
      l = 0.f;
#pragma unroll
      for (int k = 0; k < E2; ++k) accl[k] = acch[k] = 0.f;
    }
    const bool seg_end = (cj == p.tpr - 1) | (cc.t == cc.t_end - 1);

    // ---- one tile -------------------------------------------------------------------------------------------------
    const int tok0 = cj * kTile;
    const int ntok = min(kTile, p.n_total - tok0);           // valid tokens (incl. the appended one)
    const int ncopy = max(0, min(kTile, p.kv_len - tok0));   // of which bulk-copied from the cache
    uint8_t* Ks = wsm + s * kStageBytes;
    uint8_t* Vs = Ks + kTileBytes;

    const bool generic_writes = ntok < kTile;                // only the last tile of a row
    if (generic_writes) {
      // zero the rows the copy does not write, so masked lanes multiply finite numbers
#pragma unroll 1
      for (int o = ncopy * kRowBytes + lane * 16; o < kTileBytes; o += 32 * 16) {
        *reinterpret_cast<uint4*>(Ks + o) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(Vs + o) = make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
    }
    if (ncopy > 0) {
      if (!mbar_try_wait(&bars[s], (phase >> s) & 1u)) {
        const long long tw = clock64();
        while (!mbar_try_wait(&bars[s], (phase >> s) & 1u)) {
          if (clock64() - tw > 4000000000ll) __trap();    // a lost copy must fail the launch, not wedge the GPU
        }
      }
      phase ^= 1u << s;
    }
    if (ncopy < ntok) {
      // the appended token (modify_llama.py:95-100): its un-rotated K/V row goes into the tile and into slot kv_len
      constexpr int kChunks = kRowBytes / 16;
      const int r = p.kv_len - tok0;
      const int which = lane / kChunks, ch = lane % kChunks;
      const int bi = seg_row / p.H, kvh = (seg_row - bi * p.H) / p.n_rep;
#pragma unroll 1
      for (int w = which; w < 2; w += 32 / kChunks) {
        const __half* src = static_cast<const __half*>(w ? ly->v_new : ly->k_new) +
                            (static_cast<int64_t>(bi) * p.H_kv + kvh) * D + ch * 8;
        const uint4 val = *reinterpret_cast<const uint4*>(src);
        *reinterpret_cast<uint4*>((w ? Vs : Ks) + r * kRowBytes + ch * 16) = val;
        if (seg_writer) {
          __half* dst = static_cast<__half*>(w ? ly->v_cache : ly->k_cache) + bi * p.stride_b + kvh * p.stride_h +
                        static_cast<int64_t>(p.kv_len) * D + ch * 8;
          *reinterpret_cast<uint4*>(dst) = val;
        }
      }
      __syncwarp();
    }

    // scores of the tile: rotated key rounded to fp16 after every elementwise op (:27), products summed in fp32
    float d[PASSES];
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int tl = ps * 4 + g;
      HV kl, kh;
      *reinterpret_cast<Vec*>(&kl) = *reinterpret_cast<const Vec*>(Ks + tl * kRowBytes + c * VB);
      *reinterpret_cast<Vec*>(&kh) = *reinterpret_cast<const Vec*>(Ks + tl * kRowBytes + D + c * VB);
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int e = 0; e < NH2; ++e) {
        const __half2 rl = __hsub2_rn(__hmul2_rn(kl.h[e], cs[ps].h[e]), __hmul2_rn(kh.h[e], sn[ps].h[e]));
        const __half2 rh = __hadd2_rn(__hmul2_rn(kh.h[e], cs[ps].h[e]), __hmul2_rn(kl.h[e], sn[ps].h[e]));
        const float2 fl = __half22float2(rl), fh = __half22float2(rh);
        d0 = fmaf(ql[2 * e], fl.x, d0);
        d0 = fmaf(ql[2 * e + 1], fl.y, d0);
        d1 = fmaf(qh[2 * e], fh.x, d1);
        d1 = fmaf(qh[2 * e + 1], fh.y, d1);
      }
      d[ps] = d0 + d1;
    }

    // ---- the RoPE registers are free: move the cursor on and fetch the NEXT tile's rows under the rest of this tile --
    const bool was_last_of_range = cc.t == cc.t_end - 1;
    cc.first = 0;
    ++cc.t; ++cc.j;
    if (cc.j == p.tpr || was_last_of_range) advance_slow(cc);
    if (cc.l < p.L) load_table(cc.j * kTile, cs, sn);

#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) d[ps] += __shfl_xor_sync(0xffffffffu, d[ps], o);
    }
    // fp16 logit exactly as the reference forms it: fp16(matmul) / sqrt(D) -> fp16 (:111-113)
    float sc[PASSES], pr[PASSES];
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int tl = ps * 4 + g;
      const bool valid = tl < ntok;
      const __half s16 = __float2half_rn(div_by_const(__half2float(__float2half_rn(d[ps])), p.sqrt_d, p.rsqrt_d));
      if (valid && c == 0 && lg) lg[tok0 + tl] = s16;
      sc[ps] = valid ? __half2float(s16) : -CUDART_INF_F;
      mx = fmaxf(mx, sc[ps]);
    }
    // online softmax (:135): m is warp-uniform; l and acc are per key group until the segment ends
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    const float m_new = fmaxf(m, mx);
    if (m_new > -CUDART_INF_F) {
      if (m_new != m) {
        const float corr = fast_exp2((m - m_new) * kLog2e);
        l *= corr;
#pragma unroll
        for (int e = 0; e < E2; ++e) { accl[e] *= corr; acch[e] *= corr; }
        m = m_new;
      }
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        pr[ps] = fast_exp2((sc[ps] - m) * kLog2e);
        l += pr[ps];
      }
      // P.V (:138)
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int tl = ps * 4 + g;
        HV vl, vh;
        *reinterpret_cast<Vec*>(&vl) = *reinterpret_cast<const Vec*>(Vs + tl * kRowBytes + c * VB);
        *reinterpret_cast<Vec*>(&vh) = *reinterpret_cast<const Vec*>(Vs + tl * kRowBytes + D + c * VB);
#pragma unroll
        for (int e = 0; e < NH2; ++e) {
          const float2 fl = __half22float2(vl.h[e]), fh = __half22float2(vh.h[e]);
          accl[2 * e] = fmaf(pr[ps], fl.x, accl[2 * e]);
          accl[2 * e + 1] = fmaf(pr[ps], fl.y, accl[2 * e + 1]);
          acch[2 * e] = fmaf(pr[ps], fh.x, acch[2 * e]);
          acch[2 * e + 1] = fmaf(pr[ps], fh.y, acch[2 * e + 1]);
        }
      }
    }

    // ---- stage s is free: refill it with the tile STAGES ahead (possibly the next layer's) ---------------------------
    if (generic_writes) fence_proxy_async();         // generic-proxy writes above vs the async-proxy copy below
    __syncwarp();
    issue_next(s);
    s = (s + 1 == STAGES) ? 0 : s + 1;

    if (!seg_end) continue;

    // ---- segment finished: reduce-scatter over the 4 key groups (lanes that differ in bits 3 and 4) -------------
    l += __shfl_xor_sync(0xffffffffu, l, 8);
    l += __shfl_xor_sync(0xffffffffu, l, 16);
    float keep[E2], fin[PER];
    {
      const bool up = (g & 1) != 0;                 // bit 3: keep the upper half-row, send the lower
#pragma unroll
      for (int e = 0; e < E2; ++e) {
        const float send = up ? accl[e] : acch[e];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
        keep[e] = (up ? acch[e] : accl[e]) + recv;
      }
      const bool up2 = (g & 2) != 0;                // bit 4: keep the upper PER of the E2 values
#pragma unroll
      for (int e = 0; e < PER; ++e) {
        const float send = up2 ? keep[e] : keep[e + PER];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
        fin[e] = (up2 ? keep[e + PER] : keep[e]) + recv;
      }
    }
    const int d0 = (g & 1) * (D / 2) + c * E2 + (g >> 1) * PER;   // first of the PER dims this lane now owns

    // splits of this row: the virtual warps whose ranges intersect the row's tiles (32-bit: host checked T * W < 2^32)
    const unsigned T = static_cast<unsigned>(p.b * seg_nlh * p.tpr);
    const unsigned weff = T < static_cast<unsigned>(p.W) ? T : static_cast<unsigned>(p.W);
    const int v = (gw + seg_l * kStepRot) % p.W;
    const unsigned tf = static_cast<unsigned>(seg_i) * p.tpr;
    const int o_first = static_cast<int>(((tf + 1) * weff - 1) / T);
    const int o_last = static_cast<int>(((tf + p.tpr) * weff - 1) / T);
    const int n_seg = o_last - o_first + 1;

    if (n_seg == 1) {
      emit_row(seg_l, seg_row, d0, fin, l > 0.f ? 1.f / l : 0.f, m);
      continue;
    }
    {
      const int64_t slot = static_cast<int64_t>(seg_l) * p.slots + v + seg_i;
      *reinterpret_cast<PV_t*>(p.part + slot * D + d0) = *reinterpret_cast<const PV_t*>(fin);
      if (lane == 0) p.part_ml[slot] = make_float2(m, l);
    }
    __syncwarp();
    unsigned prev = 0;
    unsigned int* ticket = p.tickets + static_cast<int64_t>(seg_l) * p.rows + seg_row;
    if (lane == 0) {
      __threadfence();                                   // release: the partial is visible device-wide
      prev = atomicAdd(ticket, 1u);
    }
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev != static_cast<unsigned>(n_seg - 1)) continue;
    __threadfence();                                     // acquire side

    // ---- last warp of the row: merge its n_seg partials (online over chunks of 32 splits) -----------------------
    {
      const int64_t base = static_cast<int64_t>(seg_l) * p.slots + o_first + seg_i;
      const float* rec = p.part + base * D + lane * PER;
      const float2* ml = p.part_ml + base;
      float M = -CUDART_INF_F, Lp = 0.f;
      float o[PER];
#pragma unroll
      for (int k = 0; k < PER; ++k) o[k] = 0.f;
      constexpr int kU = WARPS > 8 ? 8 : 16;             // partial loads in flight per lane (register budget)
#pragma unroll 1
      for (int c0 = 0; c0 < n_seg; c0 += 32) {
        const int nc = min(32, n_seg - c0);
        const float2 mlj = lane < nc ? __ldcg(ml + c0 + lane) : make_float2(-CUDART_INF_F, 0.f);
        PV_t vv[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u)
          vv[u] = __ldcg(reinterpret_cast<const PV_t*>(rec + static_cast<int64_t>(c0 + min(u, nc - 1)) * D));
        const float Mn = fmaxf(M, warp_max(mlj.x));
        const float corr = (Mn > -CUDART_INF_F) ? fast_exp2((M - Mn) * kLog2e) : 0.f;   // M = -inf -> 0
        const float wj = (Mn > -CUDART_INF_F) ? fast_exp2((mlj.x - Mn) * kLog2e) : 0.f;
        Lp = fmaf(mlj.y, wj, Lp * corr);
#pragma unroll
        for (int k = 0; k < PER; ++k) o[k] *= corr;
        M = Mn;
#pragma unroll 1
        for (int u0 = 0; u0 < nc; u0 += kU) {
          if (u0 > 0) {
#pragma unroll
            for (int u = 0; u < kU; ++u)
              vv[u] = __ldcg(reinterpret_cast<const PV_t*>(rec + static_cast<int64_t>(c0 + min(u0 + u, nc - 1)) * D));
          }
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const float w = __shfl_sync(0xffffffffu, wj, (u0 + u) & 31);
            const float wu = (u0 + u < nc) ? w : 0.f;
            const float* vf = reinterpret_cast<const float*>(&vv[u]);
#pragma unroll
            for (int k = 0; k < PER; ++k) o[k] = fmaf(vf[k], wu, o[k]);
          }
        }
      }
      const float Lsum = warp_sum(Lp);
      if (lane == 0) *ticket = 0;                        // self-cleaning ticket for the next step
      emit_row(seg_l, seg_row, lane * PER, o, Lsum > 0.f ? 1.f / Lsum : 0.f, M);
    }
  }

  // ---- step epilogue ------------------------------------------------------------------------------------------------
  if (p.final_wait && blockIdx.x == 0 && warp == 0) wait_layer(p.L - 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(p.ctrl + 1, 1u);
    if (prev == gridDim.x - 1) {
      p.ctrl[1] = 0;
      __threadfence();
      atomicAdd(p.ctrl, 1u);                             // steps completed: the next launch's epoch
    }
  }
}

template <int D, int WARPS, int STAGES, int PASSES>
constexpr size_t step_smem_bytes() {
  return static_cast<size_t>(WARPS) * STAGES * 2 * (4 * PASSES) * D * 2 + WARPS * STAGES * 8 +
         kStepSmemLayers * sizeof(SpattenStepLayer);
}

// Kernel variants: (warps per CTA, ring stages per warp).  192 KiB of K/V ring per SM either way.
//   0: 8 warps x 3 stages x 16-token tiles     1: 16 warps x 3 stages x 8-token tiles (4 warps per scheduler, <= 128 regs)
constexpr int kStepVariants = 2;
constexpr int kStepMaxWarps = 16;
using StepKernel = void (*)(const StepParams);
struct StepVariant {
  StepKernel kernel;
  int warps, tile;
  size_t smem;
};
template <int D>
static StepVariant step_variant(int v) {
  if (v == 1) return {decode_step_kernel<D, 16, 3, 2>, 16, 8, step_smem_bytes<D, 16, 3, 2>()};
  return {decode_step_kernel<D, 8, 3, 4>, 8, 16, step_smem_bytes<D, 8, 3, 4>()};
}
static int step_variant_choice() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SPATTEN_STEP_VARIANT");      // tuning knob for tools/microbench.py; default 0
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v;
}

template <int D>
static int step_grid(int variant, int& grid) {
  static int cached[64][kStepVariants] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && cached[dev][variant] > 0) { grid = cached[dev][variant]; return SPATTEN_OK; }
  const StepVariant sv = step_variant<D>(variant);
  cudaError_t e = cudaFuncSetAttribute(sv.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sv.smem));
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(decode_step_kernel)");
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sv.kernel, sv.warps * 32, sv.smem);
  if (e != cudaSuccess || per_sm < 1) return cuda_fail(e, "decode_step_kernel does not fit on an SM");
  const int sms = sm_count();
  if (sms <= 0) return SPATTEN_ERR_CUDA;
  grid = sms;                                            // one persistent CTA per SM: all co-resident
  if (dev < 64) cached[dev][variant] = grid;
  return SPATTEN_OK;
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

struct StepWorkspace {
  size_t ctrl, chain, tickets, ml, part, total;
  int slots;
};
static StepWorkspace step_workspace(int L, int rows, int D, int W) {
  StepWorkspace w{};
  w.slots = W + rows;
  w.ctrl = 0;
  w.chain = align256(kStepCtrlWords * 4);
  w.tickets = w.chain + align256(static_cast<size_t>(L) * 4);
  w.ml = w.tickets + align256(static_cast<size_t>(L) * rows * 4);
  w.part = w.ml + align256(static_cast<size_t>(L) * w.slots * sizeof(float2));
  w.total = w.part + align256(static_cast<size_t>(L) * w.slots * D * sizeof(float));
  return w;
}

}  // namespace spatten

using namespace spatten;

extern "C" size_t spatten_decode_step_workspace_bytes(int32_t num_layers, int32_t batch, int32_t num_heads, int32_t head_dim) {
  if (num_layers <= 0 || batch <= 0 || num_heads <= 0 || (head_dim != 64 && head_dim != 128)) return 0;
  const int sms = sm_count();
  if (sms <= 0) return 0;
  return step_workspace(num_layers, batch * num_heads, head_dim, sms * kStepMaxWarps).total;
}

extern "C" int spatten_decode_step(const SpattenStepArgs* a, void* stream) {
  SPATTEN_CHECK_ARG(a != nullptr, "spatten_decode_step: args is NULL");
  SPATTEN_CHECK_ARG(a->struct_size == sizeof(SpattenStepArgs), "spatten_decode_step: struct_size %u != %zu (ABI mismatch)",
                    a->struct_size, sizeof(SpattenStepArgs));
  SPATTEN_CHECK_ARG(a->num_layers > 0 && a->batch > 0 && a->num_heads > 0 && a->num_kv_heads > 0,
                    "spatten_decode_step: num_layers/batch/num_heads/num_kv_heads must be positive");
  SPATTEN_CHECK_ARG(a->num_heads % a->num_kv_heads == 0, "spatten_decode_step: num_heads %d not a multiple of num_kv_heads %d",
                    a->num_heads, a->num_kv_heads);
  if (a->head_dim != 64 && a->head_dim != 128) {
    set_error("spatten_decode_step: head_dim %d unsupported (64 or 128)", a->head_dim);
    return SPATTEN_ERR_UNSUPPORTED;
  }
  if (a->head_mask != nullptr && a->num_heads > 64) {
    set_error("spatten_decode_step: head_mask supports at most 64 heads (got %d)", a->num_heads);
    return SPATTEN_ERR_UNSUPPORTED;
  }
  SPATTEN_CHECK_ARG(a->kv_len >= 0 && (a->append == 0 || a->append == 1), "spatten_decode_step: bad kv_len/append");
  const int n_total = a->kv_len + a->append;
  SPATTEN_CHECK_ARG(n_total > 0, "spatten_decode_step: nothing to attend over (kv_len 0 and append 0)");
  SPATTEN_CHECK_ARG(a->rope_len >= n_total, "spatten_decode_step: rope_len %d < kv_len+append %d", a->rope_len, n_total);
  SPATTEN_CHECK_ARG(a->layers && a->rope, "spatten_decode_step: NULL layers/rope");
  SPATTEN_CHECK_ARG(a->cache_stride_h >= static_cast<int64_t>(n_total) * a->head_dim,
                    "spatten_decode_step: cache_stride_h %lld too small for %d tokens", (long long)a->cache_stride_h, n_total);
  SPATTEN_CHECK_ARG(a->cache_stride_b >= a->cache_stride_h * a->num_kv_heads || a->batch == 1,
                    "spatten_decode_step: cache_stride_b too small");
  SPATTEN_CHECK_ARG(a->cache_stride_h % 8 == 0 && a->cache_stride_b % 8 == 0,
                    "spatten_decode_step: cache strides must be multiples of 8 elements (16-byte rows)");
  SPATTEN_CHECK_ARG((reinterpret_cast<uintptr_t>(a->rope) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->layers) & 15) == 0,
                    "spatten_decode_step: rope and layers must be 16-byte aligned");
  SPATTEN_CHECK_ARG(static_cast<int64_t>(a->batch) * a->num_heads < (1ll << 24), "spatten_decode_step: too many rows");
  SPATTEN_CHECK_ARG(a->n_peers >= 0 && a->n_peers <= kStepMaxPeers, "spatten_decode_step: bad n_peers");

  int grid = 0;
  const int variant = step_variant_choice();
  if (int rc = (a->head_dim == 128 ? step_grid<128>(variant, grid) : step_grid<64>(variant, grid))) return rc;
  const StepVariant sv = a->head_dim == 128 ? step_variant<128>(variant) : step_variant<64>(variant);

  StepParams p{};
  p.layers = a->layers;
  p.L = a->num_layers;
  p.b = a->batch;
  p.H = a->num_heads;
  p.H_kv = a->num_kv_heads;
  p.n_rep = a->num_heads / a->num_kv_heads;
  p.kv_len = a->kv_len;
  p.n_total = n_total;
  p.append = a->append;
  p.rope_len = a->rope_len;
  p.stride_b = a->cache_stride_b;
  p.stride_h = a->cache_stride_h;
  p.logits_stride = a->logits_stride;
  p.imp_stride = a->importance_stride;
  p.rope = static_cast<const __half*>(a->rope);
  p.q_pos = a->q_pos;
  p.head_mask = a->head_mask;
  p.head_importance = a->head_importance;
  p.W = grid * sv.warps;
  p.tpr = (n_total + sv.tile - 1) / sv.tile;
  p.rows = static_cast<unsigned>(a->batch) * a->num_heads;
  p.sqrt_d = static_cast<float>(sqrt(static_cast<double>(a->head_dim)));
  p.rsqrt_d = 1.0f / p.sqrt_d;

  SPATTEN_CHECK_ARG(static_cast<double>(p.rows) * p.tpr * p.W < 4.0e9,
                    "spatten_decode_step: batch * heads * tiles too large for the 32-bit work-list arithmetic");
  // the layout is sized for the widest variant so that one workspace serves both
  const StepWorkspace w = step_workspace(p.L, static_cast<int>(p.rows), a->head_dim, grid * kStepMaxWarps);
  if (a->workspace == nullptr || a->workspace_bytes < w.total) {
    set_error("spatten_decode_step: workspace %zu bytes < required %zu", a->workspace_bytes, w.total);
    return SPATTEN_ERR_WORKSPACE;
  }
  SPATTEN_CHECK_ARG((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0, "spatten_decode_step: workspace must be 16-byte aligned");
  uint8_t* ws = static_cast<uint8_t*>(a->workspace);
  p.slots = w.slots;
  p.ctrl = reinterpret_cast<unsigned int*>(ws + w.ctrl);
  p.tickets = reinterpret_cast<unsigned int*>(ws + w.tickets);
  p.part_ml = reinterpret_cast<float2*>(ws + w.ml);
  p.part = reinterpret_cast<float*>(ws + w.part);

  const bool chain = (a->flags & SPATTEN_STEP_CHAIN) != 0;
  if (a->n_peers > 0) {
    SPATTEN_CHECK_ARG(a->wait_flags != nullptr, "spatten_decode_step: n_peers > 0 needs wait_flags");
    for (int r = 0; r < a->n_peers; ++r) {
      SPATTEN_CHECK_ARG(a->peer_out[r] && a->peer_flags[r], "spatten_decode_step: NULL peer_out / peer_flags entry");
      p.peer_out[r] = static_cast<__half*>(a->peer_out[r]);
      p.flag_ptr[r] = static_cast<unsigned int*>(a->peer_flags[r]);
    }
    p.n_out = p.n_flag = p.wait_n = a->n_peers;
    p.peer_layer_stride = a->peer_layer_stride;
    p.flag_layer_stride = a->peer_flag_layer_stride;
    p.wait_flags = static_cast<const unsigned int*>(a->wait_flags);
    p.chain = chain ? 1 : 0;
    p.final_wait = 1;
    p.sys_scope = 1;
  } else if (chain) {
    p.n_out = 0;
    p.n_flag = p.wait_n = 1;
    p.flag_ptr[0] = reinterpret_cast<unsigned int*>(ws + w.chain);
    p.flag_layer_stride = 1;
    p.wait_flags = p.flag_ptr[0];
    p.chain = 1;
  }

  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(sv.warps * 32);
  cfg.dynamicSmemBytes = sv.smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaLaunchKernelEx(&cfg, sv.kernel, p);
  if (e != cudaSuccess) return cuda_fail(e, "decode_step_kernel launch");
  return SPATTEN_OK;
}
