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
// the whole step and the K/V stream never stops.  Warps are role-specialised:
//   * COMPUTE warps (8 or 12 per CTA) only stream.  The work list of a layer = (live row, 16-token tile) pairs in
//     row-major order; it is cut into equal contiguous CTA ranges (rotated by one CTA per layer), and inside a CTA warp w
//     takes range positions w, w + WARPS, ...  Every warp owns a ring of 16-token K+V tiles fed by 1-D bulk copies
//     (cp.async.bulk -> mbarrier complete_tx); the ISSUE cursor runs STAGES tiles ahead of the CONSUME cursor and simply
//     walks on into the next layer's range, so layer l+1's rows are landing in shared memory while layer l is merged.
//     A range may cover pieces of several rows ("segments"); a finished segment is reduce-scattered across the warp's 4
//     key groups (12 shuffles) into one of the CTA's shared-memory segment buffers.
//   * the QUERY warp rotates the query of every segment into shared memory ahead of the compute warps.
//     SPATTEN_STEP_CHAIN: the query of layer l+1 is read only after layer l's flag says every row is stored (the
//     decoder's dependency); only the query load and the arithmetic wait, the K/V prefetch does not.
//   * the MERGE warp combines the warp partials of a segment, writes one CTA partial, and a per-(layer, row) ticket
//     elects the last CTA of the row, whose merge warp merges the row's partials, stores the output row, adds the
//     importance terms and bumps the layer's flag.  No global round trip of the merge path runs on a streaming warp.
//   * head sharding: the merged row goes into this rank's slot of the LOCAL gather buffer; the merge warp of one
//     designated CTA per layer copies the finished layer to every peer over NVLink and releases one system-scope flag
//     per peer; the launch ends when every peer's flags of every layer have arrived (one acquire fence for all).
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
constexpr int kStepSmemLayers = 80;  // layer descriptors staged in shared memory when the model has at most this many

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
  int self_peer;           // index of this rank in peer_out
  unsigned int* layer_rows; // [L] rows of a layer stored so far on this GPU (head sharding: one system-scope release per layer)
  unsigned long long* dbg; // optional [L][grid][warps][8] globaltimer stamps (tools/microbench.py --step-timeline)
};

// CTA-level partial buffers in shared memory = segments a fast warp may run ahead of the slowest warp of its CTA
__host__ __device__ constexpr int step_seg_bufs(int warps) { return warps <= 8 ? 6 : 3; }

// One layer's share of work for a CTA: a contiguous range of the layer's (live row, tile) list.
struct LayerRange {
  int cs, ce;              // tile range [cs, ce) of this CTA
  int nlh;                 // live heads of the layer
  int vcta;                // virtual CTA id (rotates with the layer)
  unsigned T, geff;        // tiles in the layer, CTAs that take part
  unsigned long long hm;   // live head bits (valid when head_mask != nullptr)
};

__device__ __forceinline__ int nth_set_bit64(unsigned long long m, int n) {
  const unsigned lo = static_cast<unsigned>(m);
  const int cl = __popc(lo);
  if (n < cl) return __fns(lo, 0, n + 1);
  return 32 + __fns(static_cast<unsigned>(m >> 32), 0, n - cl + 1);
}

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* ptr, int sys) {
  unsigned v;
  if (sys) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  else asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
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

__device__ __forceinline__ unsigned long long step_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Shared-memory record of one layer's share of work for this CTA (filled once at kernel start).
struct RangeRec {
  int cs, ce;              // tile range [cs, ce) of this CTA in the layer's (live row, tile) list
  int nlh, vcta;           // live heads of the layer, virtual CTA id (rotates with the layer)
  unsigned T, geff;        // tiles in the layer, CTAs that take part
  unsigned long long hm;   // live head bits (valid when head_mask != nullptr)
};

// Warp roles of a CTA:  WARPS compute warps stream K/V tiles and do nothing else;  one QUERY warp rotates the query of every
// upcoming segment into shared memory (and is the only one that polls the chain / exchange flags);  one MERGE warp
// combines the compute warps' partials of a finished segment on chip, publishes the CTA partial, takes the row ticket,
// merges the row if it was the last CTA, stores the row (locally or into every peer's gather buffer) and releases the
// flags.  Every global round trip of the merge path (measured: 4.8 us per layer when a streaming warp did it, 3.8 us more
// for a system-scope release) is off the warps that keep HBM busy.
template <int D, int WARPS, int STAGES, int PASSES, bool DBG = false>
__global__ void __launch_bounds__((WARPS + 2) * 32, 1) decode_step_kernel(const StepParams p) {
#define STAMP(layer, k)                                                                                                   \
  do {                                                                                                                   \
    if (DBG && p.dbg && lane == 0)                                                                                       \
      p.dbg[((static_cast<size_t>(layer) * gridDim.x + blockIdx.x) * (WARPS + 2) + warp) * 8 + (k)] = step_gtime();      \
  } while (0)
  constexpr int kTile = 4 * PASSES;   // tokens per stage: PASSES passes of 4 keys x 8 lanes
  constexpr int E2 = D / 16;   // halves per lane in each half of the row (8 lanes cover D/2)
  constexpr int NH2 = E2 / 2;  // half2 per lane per half-row
  constexpr int VB = E2 * 2;   // bytes per lane per half-row
  constexpr int PER = D / 32;  // output dims a lane owns after the reduce-scatter / in the merges
  using Vec = typename VecT<VB>::type;
  using HV = HalfVec<NH2>;
  using PV_t = typename VecT<PER * 4>::type;
  constexpr int kRowBytes = D * 2;
  constexpr int kTileBytes = kTile * kRowBytes;
  constexpr int kStageBytes = 2 * kTileBytes;
  constexpr int kWarpSmem = STAGES * kStageBytes;
  constexpr int NB = step_seg_bufs(WARPS);                      // segments in flight per CTA
  constexpr int kSegBufFloats = WARPS * D + WARPS * 2 + D / 2;  // partials, (m, l) pairs, rotated query (fp16 [D])

  extern __shared__ __align__(128) uint8_t smem[];
  // control words per segment buffer: [0] compute warps arrived, [1] uses combined, [2] queries published
  __shared__ unsigned seg_ctl[NB][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 3, c = lane & 7;
  uint8_t* wsm = smem + warp * kWarpSmem;
  const uint32_t wsm_u32 = smem_u32(wsm);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WARPS * kWarpSmem) + warp * STAGES;
  const uint32_t bars_u32 = smem_u32(bars);
  float* segbuf = reinterpret_cast<float*>(smem + WARPS * kWarpSmem + WARPS * STAGES * 8);
  SpattenStepLayer* lys_s = reinterpret_cast<SpattenStepLayer*>(segbuf + NB * kSegBufFloats);
  RangeRec* ranges = reinterpret_cast<RangeRec*>(lys_s + kStepSmemLayers);
  const int G = gridDim.x;

  // ---- one-time setup: layer descriptors, this CTA's range in every layer, barriers ------------------------------------
  const SpattenStepLayer* lys = p.layers;
  if (p.L <= kStepSmemLayers) {
    uint4* dst = reinterpret_cast<uint4*>(lys_s);
    const uint4* src = reinterpret_cast<const uint4*>(p.layers);
    for (int k = threadIdx.x; k < p.L * 4; k += (WARPS + 2) * 32) dst[k] = src[k];
    lys = lys_s;
  }
  auto compute_range = [&](int l) {                            // warp-uniform (ballots)
    RangeRec r;
    r.cs = r.ce = 0; r.nlh = p.H; r.hm = ~0ull; r.vcta = 0; r.T = 0; r.geff = 1;
    if (p.head_mask) {
      const uint8_t* hr = p.head_mask + static_cast<size_t>(l) * p.H;
      const unsigned lo = __ballot_sync(0xffffffffu, lane < p.H && hr[lane] != 0);
      const unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < p.H && hr[lane + 32] != 0);
      r.hm = static_cast<unsigned long long>(lo) | (static_cast<unsigned long long>(hi) << 32);
      r.nlh = __popcll(r.hm);
    }
    r.T = static_cast<unsigned>(p.b * r.nlh * p.tpr);            // host checked: T * G < 2^32
    r.geff = r.T < static_cast<unsigned>(G) ? (r.T ? r.T : 1u) : static_cast<unsigned>(G);
    r.vcta = static_cast<int>((blockIdx.x + static_cast<unsigned>(l) * kStepRot) % static_cast<unsigned>(G));
    if (static_cast<unsigned>(r.vcta) < r.geff && r.T > 0) {
      r.cs = static_cast<int>(static_cast<unsigned>(r.vcta) * r.T / r.geff);
      r.ce = static_cast<int>((static_cast<unsigned>(r.vcta) + 1u) * r.T / r.geff);
    }
    return r;
  };
  const bool ranges_in_smem = p.L <= kStepSmemLayers;
  if (ranges_in_smem) {
    for (int l = warp; l < p.L; l += WARPS + 2) {
      const RangeRec r = compute_range(l);
      if (lane == 0) ranges[l] = r;
    }
  }
  if (threadIdx.x < NB * 4) seg_ctl[threadIdx.x >> 2][threadIdx.x & 3] = 0u;
  if (lane == 0 && warp < WARPS) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto layer_range = [&](int l) -> RangeRec { return ranges_in_smem ? ranges[l] : compute_range(l); };
  auto row_coords = [&](int i, int nlh, unsigned long long hm, int& bi, int& h) {
    bi = i / nlh;
    const int hidx = i - bi * nlh;
    h = p.head_mask ? nth_set_bit64(hm, hidx) : hidx;
  };

  const unsigned epoch = *reinterpret_cast<volatile const unsigned*>(p.ctrl);
  const unsigned flag_target = (epoch + 1u) * p.rows;

  // wait until every producer's flag of `layer` has reached this step's target: relaxed polls, then one acquire fence
  // (`acquire` = false leaves the fence to the caller: a fence.acq_rel.sys costs ~2.6 us, and 32 of them in the final wait
  // of a head-sharded step were 83 us per step)
  auto wait_layer = [&](int layer, bool acquire = true) {
    if (lane < p.wait_n) {
      const unsigned* f = p.wait_flags + static_cast<int64_t>(layer) * p.wait_n + lane;
      const long long tw = clock64();
      while (static_cast<int>(ld_relaxed(f, p.sys_scope) - flag_target) < 0) {
        if (clock64() - tw > 4000000000ll) { atomicExch(p.ctrl + 2, 1u); break; }   // ~2 s: never wedge the GPU
        __nanosleep(20);
      }
    }
    __syncwarp();
    if (acquire) {
      if (p.sys_scope) asm volatile("fence.acq_rel.sys;" ::: "memory");
      else asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
  };
  auto seg_part = [&](unsigned seq) { return segbuf + (seq % NB) * kSegBufFloats; };
  auto seg_ml = [&](unsigned seq) { return reinterpret_cast<float2*>(seg_part(seq) + WARPS * D); };
  auto seg_q = [&](unsigned seq) { return reinterpret_cast<__half*>(seg_part(seq) + WARPS * D + WARPS * 2); };

  if (warp == WARPS) {
    // ================= QUERY warp: rotate the query of every segment of this CTA into shared memory, in order ==========
    unsigned seq = 0;
#pragma unroll 1
    for (int l = 0; l < p.L; ++l) {
      const RangeRec R = layer_range(l);
      if (R.cs >= R.ce) continue;
      const SpattenStepLayer* ly = lys + l;
      // the decoder's dependency: layer l's query exists only when every row of layer l-1 has been stored
      if (l > 0 && p.chain) wait_layer(l - 1);
      STAMP(l, 1);
      const int i0 = R.cs / p.tpr, i1 = (R.ce - 1) / p.tpr;
#pragma unroll 1
      for (int i = i0; i <= i1; ++i, ++seq) {
        volatile unsigned* ctl = seg_ctl[seq % NB];
        const unsigned use = seq / NB;
        while (ctl[1] != use) __nanosleep(40);               // the buffer's previous segment has been combined
        int bi, h;
        row_coords(i, R.nlh, R.hm, bi, h);
        const int row = bi * p.H + h;
        int64_t pos = p.q_pos ? p.q_pos[bi] : p.kv_len;
        pos = max(static_cast<int64_t>(0), min(pos, static_cast<int64_t>(p.rope_len - 1)));
        // lane t rotates dims {2t, 2t+1} (D = 128) of the lower half with their partners in the upper half (:21-28, fp16)
        if (lane < D / 4) {
          const __half* qr = static_cast<const __half*>(ly->q) + static_cast<int64_t>(row) * D;
          const __half2 qa = *reinterpret_cast<const __half2*>(qr + 2 * lane);
          const __half2 qb = *reinterpret_cast<const __half2*>(qr + D / 2 + 2 * lane);
          const __half2 cs = *reinterpret_cast<const __half2*>(p.rope + pos * D + 2 * lane);
          const __half2 sn = *reinterpret_cast<const __half2*>(p.rope + pos * D + D / 2 + 2 * lane);
          __half* qs = seg_q(seq);
          *reinterpret_cast<__half2*>(qs + 2 * lane) = __hsub2_rn(__hmul2_rn(qa, cs), __hmul2_rn(qb, sn));
          *reinterpret_cast<__half2*>(qs + D / 2 + 2 * lane) = __hadd2_rn(__hmul2_rn(qb, cs), __hmul2_rn(qa, sn));
        }
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          ctl[2] = use + 1;                                    // query of segment `seq` is in shared memory
        }
      }
    }
  } else if (warp == WARPS + 1) {
    // ================= MERGE warp: combine finished segments, publish, merge rows, store, signal ======================
    // every lane stored its piece of a row of layer `layer`: publish it to the consumers of the flags
    auto signal_row = [&](int layer) {
      if (p.n_flag == 0) return;
      __syncwarp();
      if (p.sys_scope) {
        // Head sharding: the row went into this rank's slot of the LOCAL gather buffer; count it (GPU scope).  The layer is
        // shipped to the peers by the merge warp of a designated CTA once the count is complete (ship_layer below): a
        // system-scope release per row costs 2.5 us per layer even on this warp, and giving the extra work to whichever
        // warp stores the last row makes the slowest CTA slower still (measured: 9.0 -> 13.0 us per layer).
        if (lane == 0) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          atomicAdd(p.layer_rows + layer, 1u);
        }
      } else if (lane < p.n_flag) {
        unsigned int* f = p.flag_ptr[lane] + static_cast<int64_t>(layer) * p.flag_layer_stride;
        asm volatile("fence.acq_rel.gpu;\n\tred.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(f) : "memory");
      }
    };
    // lane owns dims [lane*PER, lane*PER + PER) of the finished row: normalise, round, store (locally or to every peer)
    auto emit_row = [&](int layer, int row, const float (&o)[PER], float inv, float M) {
      __half oh16[PER];
      float habs = 0.f;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        oh16[k] = __float2half_rn(o[k] * inv);
        habs += fabsf(__half2float(oh16[k]));
      }
      const int64_t off = static_cast<int64_t>(row) * D + lane * PER;
      if (p.n_out == 0) {
        __half* dst = static_cast<__half*>(lys[layer].out) + off;
        if (PER == 4) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(oh16);
        else *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(oh16);
      } else {
        // head sharding: into this rank's slot of the LOCAL gather buffer; the layer's last row ships the slot (signal_row)
        __half* dst = p.peer_out[p.self_peer] + static_cast<int64_t>(layer) * p.peer_layer_stride + off;
        if (PER == 4) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(oh16);
        else *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(oh16);
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
    // ---- shipping (head sharding): layer l is sent by the merge warp of CTA (41 l + 7) mod G, between its own segments ----
    auto ship_owner = [&](int l) { return static_cast<int>((41u * static_cast<unsigned>(l) + 7u) % static_cast<unsigned>(G)); };
    int ship_l = 0;
    auto ship_seek = [&]() { while (ship_l < p.L && ship_owner(ship_l) != static_cast<int>(blockIdx.x)) ++ship_l; };
    if (!p.sys_scope) ship_l = p.L;
    ship_seek();
    auto try_ship = [&](bool block) {
      while (ship_l < p.L) {
        unsigned have = 0;
        if (lane == 0) have = ld_relaxed(p.layer_rows + ship_l, 0);
        have = __shfl_sync(0xffffffffu, have, 0);
        if (static_cast<int>(have - flag_target) < 0) {
          if (!block) return;
          __nanosleep(100);
          continue;
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");     // acquire: every row of the layer is in the local slot
        const uint4* src = reinterpret_cast<const uint4*>(p.peer_out[p.self_peer] + static_cast<int64_t>(ship_l) * p.peer_layer_stride);
        const int n16 = static_cast<int>(p.rows) * D / 8;
#pragma unroll 1
        for (int r = 0; r < p.n_out; ++r) {
          if (r == p.self_peer) continue;
          uint4* dst = reinterpret_cast<uint4*>(p.peer_out[r] + static_cast<int64_t>(ship_l) * p.peer_layer_stride);
#pragma unroll 4
          for (int k = lane; k < n16; k += 32) dst[k] = __ldcg(src + k);
        }
        __syncwarp();
        if (lane < p.n_flag) {
          unsigned int* f = p.flag_ptr[lane] + static_cast<int64_t>(ship_l) * p.flag_layer_stride;
          asm volatile("fence.acq_rel.sys;\n\tred.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(f), "r"(p.rows) : "memory");
        }
        ++ship_l;
        ship_seek();
      }
    };
    unsigned seq = 0;
#pragma unroll 1
    for (int l = 0; l < p.L; ++l) {
      const RangeRec R = layer_range(l);
      const SpattenStepLayer* ly = lys + l;
      if (p.head_mask && R.nlh < p.H && R.T > 0 && static_cast<unsigned>(R.vcta) < R.geff) {
        // pruned heads of the layer: zero their output rows, keep the KV cache complete, count them as done.  The appended
        // rows of fully pruned kv groups come from this layer's projections: same dependency as the query.
        if (l > 0 && p.chain) wait_layer(l - 1);
        const int ndh = p.H - R.nlh, n_dead = p.b * ndh;
        const unsigned long long hall = p.H >= 64 ? ~0ull : ((1ull << p.H) - 1ull);
#pragma unroll 1
        for (int dr = R.vcta; dr < n_dead; dr += static_cast<int>(R.geff)) {
          const int dbi = dr / ndh;
          const int dh = nth_set_bit64(~R.hm & hall, dr - dbi * ndh);
          const int dkvh = dh / p.n_rep;
          const unsigned long long grp = (R.hm >> (dkvh * p.n_rep)) & (p.n_rep >= 64 ? ~0ull : ((1ull << p.n_rep) - 1ull));
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
          emit_row(l, dbi * p.H + dh, z, 0.f, -CUDART_INF_F);
        }
      }
      if (R.cs >= R.ce) continue;
      const int i0 = R.cs / p.tpr, i1 = (R.ce - 1) / p.tpr;
#pragma unroll 1
      for (int i = i0; i <= i1; ++i, ++seq) {
        volatile unsigned* ctl = seg_ctl[seq % NB];
        const unsigned use = seq / NB;
        while (ctl[0] != static_cast<unsigned>(WARPS)) {                     // every compute warp has checked in
          try_ship(false);
          __nanosleep(40);
        }
        __threadfence_block();
        STAMP(l, 4);
        const float* sb = seg_part(seq);
        const float2* sml = seg_ml(seq);
        int bi, h;
        row_coords(i, R.nlh, R.hm, bi, h);
        const int row = bi * p.H + h;
        float M, Lsum;
        float o[PER];
        {
          const float2 mlj = lane < WARPS ? sml[lane] : make_float2(-CUDART_INF_F, 0.f);
          M = warp_max(mlj.x);
          const float wj = (M > -CUDART_INF_F && lane < WARPS) ? fast_exp2((mlj.x - M) * kLog2e) : 0.f;
          Lsum = warp_sum(mlj.y * wj);
#pragma unroll
          for (int e = 0; e < PER; ++e) o[e] = 0.f;
#pragma unroll
          for (int w = 0; w < WARPS; ++w) {
            const float ww = __shfl_sync(0xffffffffu, wj, w);
            if (ww > 0.f) {                                      // warps without a tile in the segment wrote no partial
              const PV_t v = *reinterpret_cast<const PV_t*>(sb + w * D + lane * PER);
              const float* vf = reinterpret_cast<const float*>(&v);
#pragma unroll
              for (int e = 0; e < PER; ++e) o[e] = fmaf(vf[e], ww, o[e]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          ctl[0] = 0;
          __threadfence_block();
          ctl[1] = use + 1;                                      // hand the buffer to segment seq + NB
        }

        // global level: the CTAs whose ranges intersect row i
        const unsigned tf = static_cast<unsigned>(i) * p.tpr;
        const int o_first = static_cast<int>(((tf + 1) * R.geff - 1) / R.T);
        const int o_last = static_cast<int>(((tf + p.tpr) * R.geff - 1) / R.T);
        const int n_grp = o_last - o_first + 1;
        if (n_grp == 1) {
          emit_row(l, row, o, Lsum > 0.f ? 1.f / Lsum : 0.f, M);
          STAMP(l, 7);
          continue;
        }
        {
          const int64_t slot = static_cast<int64_t>(l) * p.slots + R.vcta + i;
          *reinterpret_cast<PV_t*>(p.part + slot * D + lane * PER) = *reinterpret_cast<const PV_t*>(o);
          if (lane == 0) p.part_ml[slot] = make_float2(M, Lsum);
        }
        __syncwarp();
        unsigned prev = 0;
        unsigned int* ticket = p.tickets + static_cast<int64_t>(l) * p.rows + row;
        if (lane == 0) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");     // release: the partial is visible device-wide
          prev = atomicAdd(ticket, 1u);
        }
        prev = __shfl_sync(0xffffffffu, prev, 0);
        STAMP(l, 5);
        if (prev != static_cast<unsigned>(n_grp - 1)) continue;
        asm volatile("fence.acq_rel.gpu;" ::: "memory");       // acquire side

        // last CTA of the row: merge its n_grp partials (a handful: rows / CTAs)
        const int64_t base = static_cast<int64_t>(l) * p.slots + o_first + i;
        const float* rec = p.part + base * D + lane * PER;
        const float2* ml = p.part_ml + base;
        float Mg = -CUDART_INF_F, Lp = 0.f;
        float og[PER];
#pragma unroll
        for (int e = 0; e < PER; ++e) og[e] = 0.f;
        constexpr int kU = 8;                                  // partial loads in flight per lane
#pragma unroll 1
        for (int c0 = 0; c0 < n_grp; c0 += kU) {
          const int nc = min(kU, n_grp - c0);
          const float2 mlj = lane < nc ? __ldcg(ml + c0 + lane) : make_float2(-CUDART_INF_F, 0.f);
          PV_t vv[kU];
#pragma unroll
          for (int u = 0; u < kU; ++u)
            vv[u] = __ldcg(reinterpret_cast<const PV_t*>(rec + static_cast<int64_t>(c0 + min(u, nc - 1)) * D));
          const float Mn = fmaxf(Mg, warp_max(mlj.x));
          const float corr = (Mn > -CUDART_INF_F) ? fast_exp2((Mg - Mn) * kLog2e) : 0.f;   // Mg = -inf -> 0
          const float wj = (Mn > -CUDART_INF_F) ? fast_exp2((mlj.x - Mn) * kLog2e) : 0.f;
          Lp = fmaf(mlj.y, wj, Lp * corr);
#pragma unroll
          for (int e = 0; e < PER; ++e) og[e] *= corr;
          Mg = Mn;
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const float w = __shfl_sync(0xffffffffu, wj, u);
            const float wu = (u < nc) ? w : 0.f;
            const float* vf = reinterpret_cast<const float*>(&vv[u]);
#pragma unroll
            for (int e = 0; e < PER; ++e) og[e] = fmaf(vf[e], wu, og[e]);
          }
        }
        const float Lg = warp_sum(Lp);
        if (lane == 0) *ticket = 0;                            // self-cleaning ticket for the next step
        STAMP(l, 6);
        emit_row(l, row, og, Lg > 0.f ? 1.f / Lg : 0.f, Mg);
        STAMP(l, 7);
      }
    }
    try_ship(true);                                            // whatever this CTA still owes the peers
    // head sharding: the launch completes only when every rank's rows of every layer are in the local gather buffer
    // (chained steps have already waited for all layers but the last)
    if (p.final_wait && blockIdx.x == 0) {
#pragma unroll 1
      for (int l = p.chain ? p.L - 1 : 0; l < p.L; ++l) wait_layer(l, false);          // relaxed polls of every layer's flags ...
      if (p.sys_scope) asm volatile("fence.acq_rel.sys;" ::: "memory");                 // ... and ONE acquire fence for all of them
      else asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
  } else {
    // ================= COMPUTE warps: stream K/V tiles, nothing else =====================================================
    const uint64_t pol = policy_evict_first();
    // issue side: this warp's tile sequence, STAGES tiles ahead of the consumer, walking on across layers.  This warp's
    // tiles of layer l are the range positions (warp + l) % WARPS, + WARPS, ... (the offset rotates with the layer so that
    // the warps taking one tile more than the others change from layer to layer).
    int il = 0, ik = 0, ike = 0, ii = 0, ij = 0;
    bool irow_new = true;
    int inlh = p.H;
    unsigned long long ihm = ~0ull;
    const __half* ikp = nullptr;
    const __half* ivp = nullptr;
    auto issue_enter = [&](int l) {
      const RangeRec r = layer_range(l);
      ik = r.cs + (warp + l) % WARPS; ike = r.ce; inlh = r.nlh; ihm = r.hm;
      if (ik < ike) { ii = ik / p.tpr; ij = ik - ii * p.tpr; irow_new = true; }
    };
    auto issue_seek = [&]() {
      while (il < p.L && ik >= ike) {
        ++il;
        if (il < p.L) issue_enter(il);
      }
    };
    issue_enter(0);
    issue_seek();
    auto issue_next = [&](int s) {                 // copy the issue cursor's tile into stage s, move the cursor on
      if (il >= p.L) return;
      if (irow_new) {
        int bi, h;
        row_coords(ii, inlh, ihm, bi, h);
        const SpattenStepLayer* ly = lys + il;
        const int64_t off = bi * p.stride_b + (h / p.n_rep) * p.stride_h;
        ikp = static_cast<const __half*>(ly->k_cache) + off;
        ivp = static_cast<const __half*>(ly->v_cache) + off;
        irow_new = false;
      }
      const int ncopy = min(kTile, p.kv_len - ij * kTile);
      if (ncopy > 0)
        issue_kv_copies(lane == 0, wsm_u32 + s * kStageBytes, wsm_u32 + s * kStageBytes + kTileBytes,
                        ikp + static_cast<int64_t>(ij) * (kTile * D), ivp + static_cast<int64_t>(ij) * (kTile * D),
                        static_cast<uint32_t>(ncopy) * kRowBytes, bars_u32 + s * 8, pol);
      ik += WARPS; ij += WARPS;
      while (ij >= p.tpr) { ij -= p.tpr; ++ii; irow_new = true; }
      if (ik >= ike) issue_seek();
    };
    // RoPE rows of one tile: cos | sin for the passes of this lane's key group (L2 resident table)
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

#pragma unroll 1
    for (int s = 0; s < STAGES; ++s) issue_next(s);

    unsigned phase = 0;     // bit s = parity the next wait on stage s uses
    int s = 0;
    unsigned seg_seq = 0;   // segments (CTA range x row) this CTA has walked past: selects the shared-memory buffer
    HV cs[PASSES], sn[PASSES];
    int table_for = -1;     // token offset the RoPE registers currently hold (-1: none)

#pragma unroll 1
    for (int l = 0; l < p.L; ++l) {
      const RangeRec R = layer_range(l);
      const SpattenStepLayer* ly = lys + l;
      if (R.cs >= R.ce) continue;
      const int i0 = R.cs / p.tpr, i1 = (R.ce - 1) / p.tpr;
      int k = R.cs + (warp + l) % WARPS;                     // this warp's next tile in the layer's list

#pragma unroll 1
      for (int i = i0; i <= i1; ++i, ++seg_seq) {
        // ---- one segment: the part of row i inside this CTA's range; this warp's tiles of it --------------------------
        const int row_end = min(R.ce, (i + 1) * p.tpr);
        const bool have = k < row_end;
        volatile unsigned* ctl = seg_ctl[seg_seq % NB];
        const unsigned use = seg_seq / NB;
        float fin[PER];
        float m = -CUDART_INF_F, lsum = 0.f;
        if (have) {
          int j = k - i * p.tpr;
          if (table_for != j * kTile) load_table(j * kTile, cs, sn);
          int bi, h;
          row_coords(i, R.nlh, R.hm, bi, h);
          const int row = bi * p.H + h;
          const int kvh = h / p.n_rep;
          bool writer = (h % p.n_rep) == 0;
          if (p.head_mask) {
            const unsigned long long grp = (R.hm >> (kvh * p.n_rep)) & (p.n_rep >= 64 ? ~0ull : ((1ull << p.n_rep) - 1ull));
            writer = h == kvh * p.n_rep + (__ffsll(static_cast<long long>(grp)) - 1);   // first live head of the kv group
          }
          __half* lg = ly->logits ? static_cast<__half*>(ly->logits) + static_cast<int64_t>(row) * p.logits_stride : nullptr;
          STAMP(l, 0);
          // the rotated query of this segment, published by the query warp (which has also passed the chain wait)
          while (ctl[2] != use + 1) __nanosleep(20);
          __threadfence_block();
          HV qlr, qhr;
          {
            const __half* qs = seg_q(seg_seq);
            *reinterpret_cast<Vec*>(&qlr) = *reinterpret_cast<const Vec*>(qs + c * E2);
            *reinterpret_cast<Vec*>(&qhr) = *reinterpret_cast<const Vec*>(qs + D / 2 + c * E2);
          }
          float accl[E2], acch[E2];
#pragma unroll
          for (int e = 0; e < E2; ++e) accl[e] = acch[e] = 0.f;

#pragma unroll 1
          for (; k < row_end; k += WARPS, j += WARPS) {
            // ---- one tile -------------------------------------------------------------------------------------------
            const int tok0 = j * kTile;
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
            if (DBG && k == R.cs + (warp + l) % WARPS) STAMP(l, 2);
            if (ncopy < ntok) {
              // the appended token (modify_llama.py:95-100): its un-rotated K/V row goes into the tile and into slot kv_len
              constexpr int kChunks = kRowBytes / 16;
              const int r = p.kv_len - tok0;
              const int which = lane / kChunks, ch = lane % kChunks;
#pragma unroll 1
              for (int w = which; w < 2; w += 32 / kChunks) {
                const __half* src = static_cast<const __half*>(w ? ly->v_new : ly->k_new) +
                                    (static_cast<int64_t>(bi) * p.H_kv + kvh) * D + ch * 8;
                const uint4 val = *reinterpret_cast<const uint4*>(src);
                *reinterpret_cast<uint4*>((w ? Vs : Ks) + r * kRowBytes + ch * 16) = val;
                if (writer) {
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
                d0 = fhfma_lo(rl, qlr.h[e], d0);             // fp16 x fp16 products are exact in fp32 (FHFMA): same sums
                d0 = fhfma_hi(rl, qlr.h[e], d0);             // as fmaf(float(q), float(k), d) without the conversions
                d1 = fhfma_lo(rh, qhr.h[e], d1);
                d1 = fhfma_hi(rh, qhr.h[e], d1);
              }
              d[ps] = d0 + d1;
            }
            // the RoPE registers are free: fetch the rows of this warp's NEXT tile (same segment, next row of the layer, or
            // the next layer's first tile) under the rest of this one
            {
              int nj = -1;
              if (k + WARPS < R.ce) {
                nj = j + WARPS;
                while (nj >= p.tpr) nj -= p.tpr;
              } else if (l + 1 < p.L) {
                const RangeRec Rn = layer_range(l + 1);
                const int kn = Rn.cs + (warp + l + 1) % WARPS;
                if (kn < Rn.ce) nj = kn % p.tpr;
              }
              table_for = -1;
              if (nj >= 0) { load_table(nj * kTile, cs, sn); table_for = nj * kTile; }
            }

#pragma unroll
            for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
              for (int ps = 0; ps < PASSES; ++ps) d[ps] += __shfl_xor_sync(0xffffffffu, d[ps], o);
            }
            // fp16 logit exactly as the reference forms it: fp16(matmul) / sqrt(D) -> fp16 (:111-113)
            float sc[PASSES];
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
                lsum *= corr;
#pragma unroll
                for (int e = 0; e < E2; ++e) { accl[e] *= corr; acch[e] *= corr; }
                m = m_new;
              }
              // the probabilities enter P.V as fp16, as in the reference (softmax(...).to(fp16) @ V, :135-138); the running
              // sum uses the same rounded values, so the final normalisation is exact for what was accumulated
              __half2 ph[PASSES];
#pragma unroll
              for (int ps = 0; ps < PASSES; ++ps) {
                ph[ps] = __float2half2_rn(fast_exp2((sc[ps] - m) * kLog2e));
                lsum += __low2float(ph[ps]);
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
                  accl[2 * e] = fhfma_lo(vl.h[e], ph[ps], accl[2 * e]);
                  accl[2 * e + 1] = fhfma_hi(vl.h[e], ph[ps], accl[2 * e + 1]);
                  acch[2 * e] = fhfma_lo(vh.h[e], ph[ps], acch[2 * e]);
                  acch[2 * e + 1] = fhfma_hi(vh.h[e], ph[ps], acch[2 * e + 1]);
                }
              }
            }

            // stage s is free: refill it with this warp's tile STAGES ahead (possibly the next layer's)
            if (generic_writes) fence_proxy_async();         // generic-proxy writes above vs the async-proxy copy below
            __syncwarp();
            issue_next(s);
            s = (s + 1 == STAGES) ? 0 : s + 1;
          }
          STAMP(l, 3);

          // ---- this warp's tiles of the segment are done: reduce-scatter over the 4 key groups (lane bits 3 and 4) -----
          lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
          lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
          float keep[E2];
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
          const int d0 = (g & 1) * (D / 2) + c * E2 + (g >> 1) * PER;   // first of the PER dims this lane owns
          *reinterpret_cast<PV_t*>(seg_part(seg_seq) + warp * D + d0) = *reinterpret_cast<const PV_t*>(fin);
        } else {
          // no tile of this segment: the buffer must still be ours before (m, l) goes in
          while (ctl[1] != use) __nanosleep(40);
        }
        // ---- check in at the segment's buffer; the merge warp does the rest ---------------------------------------------
        if (lane == 0) seg_ml(seg_seq)[warp] = make_float2(m, lsum);
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          atomicAdd(const_cast<unsigned*>(&ctl[0]), 1u);
        }
      }
    }
  }

  // ---- step epilogue ------------------------------------------------------------------------------------------------
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
#undef STAMP
}

template <int D, int WARPS, int STAGES, int PASSES>
constexpr size_t step_smem_bytes() {
  return static_cast<size_t>(WARPS) * STAGES * 2 * (4 * PASSES) * D * 2 + WARPS * STAGES * 8 +
         static_cast<size_t>(step_seg_bufs(WARPS)) * (WARPS * D + WARPS * 2 + D / 2) * sizeof(float) +
         kStepSmemLayers * (sizeof(SpattenStepLayer) + sizeof(RangeRec));
}

// Kernel variants: (compute warps per CTA, ring stages per warp); every build adds the query warp and the merge warp.
// 192 KiB (or 128 KiB) of K/V ring per SM.
//   0: 8 warps x 3 stages   1: 12 warps x 2 stages   2: 8 warps x 2 stages   3: variant 1 with %globaltimer stamps
constexpr int kStepVariants = 4;
constexpr int kStepMaxWarps = 12;
using StepKernel = void (*)(const StepParams);
struct StepVariant {
  StepKernel kernel;
  int warps, tile;
  size_t smem;
};
template <int D>
static StepVariant step_variant(int v) {
  if (v == 1) return {decode_step_kernel<D, 12, 2, 4>, 12, 16, step_smem_bytes<D, 12, 2, 4>()};
  if (v == 2) return {decode_step_kernel<D, 8, 2, 4>, 8, 16, step_smem_bytes<D, 8, 2, 4>()};
  if (v == 3) return {decode_step_kernel<D, 12, 2, 4, true>, 12, 16, step_smem_bytes<D, 12, 2, 4>()};   // timeline build
  return {decode_step_kernel<D, 8, 3, 4>, 8, 16, step_smem_bytes<D, 8, 3, 4>()};
}
// Measured on B200 (profiles/r2_step_variants.jsonl): with few tiles per warp and layer (BASELINE config 2: 27.7
// tiles per SM) twelve compute warps hide the segment-boundary latencies better; long rows stream best from eight warps
// with two-stage rings.  SPATTEN_STEP_VARIANT=<n> overrides the choice (tools/microbench.py).
static int step_variant_choice(long tiles_per_layer, int sms) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("SPATTEN_STEP_VARIANT");
    forced = e ? atoi(e) : -1;
    if (forced < 0 || forced >= kStepVariants) forced = -1;
  }
  if (forced >= 0) return forced;
  // end of round 2 (role-specialised build): 8 x 2 also wins for very short layers (config 2, 27.7 tiles per SM: 6.93 vs 7.03 us
  // sustained, 6.20 vs 6.40 burst); 12 x 2 keeps the middle (n = 4096: 10.15 vs 10.6 us)
  return (tiles_per_layer >= 64L * sms || tiles_per_layer < 36L * sms) ? 2 : 1;
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
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sv.kernel, (sv.warps + 2) * 32, sv.smem);
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

static unsigned long long* g_step_dbg = nullptr;
// Debug hook (not part of the public header): device buffer [L][grid][warps][8] u64 for %globaltimer stamps; selects
// the instrumented build of the kernel while set.
extern "C" void spatten_debug_set_step_timeline(unsigned long long* dev_buf) { g_step_dbg = dev_buf; }

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
  const long tiles16 = static_cast<long>(a->batch) * a->num_heads * ((n_total + 15) / 16);
  const int sms_now = sm_count();
  if (sms_now <= 0) return SPATTEN_ERR_CUDA;
  const int variant = g_step_dbg ? 3 : step_variant_choice(tiles16, sms_now);
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
  p.dbg = g_step_dbg;

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
    SPATTEN_CHECK_ARG(a->self_peer >= 0 && a->self_peer < a->n_peers, "spatten_decode_step: self_peer out of range");
    SPATTEN_CHECK_ARG((reinterpret_cast<uintptr_t>(a->peer_out[a->self_peer]) & 15) == 0 && a->peer_layer_stride % 8 == 0,
                      "spatten_decode_step: gather buffers must be 16-byte aligned");
    p.self_peer = a->self_peer;
    p.n_out = p.n_flag = p.wait_n = a->n_peers;
    p.peer_layer_stride = a->peer_layer_stride;
    p.flag_layer_stride = a->peer_flag_layer_stride;
    p.wait_flags = static_cast<const unsigned int*>(a->wait_flags);
    p.chain = chain ? 1 : 0;
    p.final_wait = 1;
    p.sys_scope = 1;
    p.layer_rows = reinterpret_cast<unsigned int*>(ws + w.chain);
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
  cfg.blockDim = dim3((sv.warps + 2) * 32);
  cfg.dynamicSmemBytes = sv.smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaLaunchKernelEx(&cfg, sv.kernel, p);
  if (e != cudaSuccess) return cuda_fail(e, "decode_step_kernel launch");
  return SPATTEN_OK;
}

// The step with HOST buffers for the per-step tensors: one packed host -> device copy of every layer's q | k_new | v_new
// (the caller lays them out so that `in_dev` is what the layer descriptors point into), the persistent launch, one
// device -> host copy of every layer's output, stream synchronised before returning.
extern "C" int spatten_decode_step_host(const SpattenStepArgs* a, const void* in_host, void* in_dev, size_t in_bytes,
                                        const void* out_dev, void* out_host, size_t out_bytes, void* stream) {
  SPATTEN_CHECK_ARG(a != nullptr && in_host && in_dev && out_dev && out_host && in_bytes > 0 && out_bytes > 0,
                    "spatten_decode_step_host: NULL argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemcpyAsync(in_dev, in_host, in_bytes, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return cuda_fail(e, "spatten_decode_step_host: H2D copy");
  const int rc = spatten_decode_step(a, stream);
  if (rc != SPATTEN_OK) return rc;
  e = cudaMemcpyAsync(out_host, out_dev, out_bytes, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return cuda_fail(e, "spatten_decode_step_host: D2H copy / sync");
  return SPATTEN_OK;
}
