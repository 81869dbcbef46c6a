// KV-cache token pruning on the device -- spatten_score_reduce / spatten_topk_select / spatten_kv_gather.
//
// Replaces SpAttenKVCache.apply_token_pruning (reference spatten_llm/kv_cache_token_pruning.py:42-96):
//   :51     importance_score = attn_scores.sum(0).sum(1)            -> score_reduce_kernel
//   :59-63  topk(score[:, start : S-recent+num_coming], important) -> topk_select_kernel
//           .sort() + start                                           (radix select, ascending output)
//   :65-93  bool-mask gather of K and V + cat(start | important | recent) -> kv_gather_kernel
// with no host synchronisation (the reference does one .cpu() per layer, :66) and no Python loop inside
// a layer.  The selection is the RTL top-k engine's rule (TopK.scala:159-160,209-211; QuickSelect.scala):
// keep every score strictly above the k-th value, then the lowest-index elements equal to it.
#include <algorithm>

#include "../../include/spatten_b200.h"
#include "common.cuh"

namespace spatten {

// ---- order-preserving integer keys -----------------------------------------------------------------------
// key(a) < key(b)  <=>  a < b for non-NaN floats; -0 and +0 map to the same key (they compare equal in topk).
__device__ __forceinline__ uint32_t score_key(__half x) {
  uint32_t u = __half_as_ushort(x);
  if ((u & 0x7fffu) == 0) u = 0;
  return (u & 0x8000u) ? (~u & 0xffffu) : (u | 0x8000u);
}
__device__ __forceinline__ uint32_t score_key(float x) {
  uint32_t u = __float_as_uint(x);
  if ((u & 0x7fffffffu) == 0) u = 0;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
template <typename T>
struct KeyBits;
template <>
struct KeyBits<__half> {
  static constexpr int value = 16;
};
template <>
struct KeyBits<float> {
  static constexpr int value = 32;
};

constexpr int kSelThreads = 512;
constexpr int kSelWarps = kSelThreads / 32;

struct SelectParams {
  const void* score;
  int64_t score_stride;
  int32_t* keep;
  int64_t keep_stride;
  int seq_len, start, important, win_begin, win_end, tail_begin;
  const SpattenPruneLayer* layers;   // optional (spatten_prune_layers): row r = head r % rows_per_layer of layer r / rows_per_layer
  int rows_per_layer;
};

// One CTA per score row.  Radix select (8-bit digits, MSB first) finds the k-th largest key; a second,
// index-ordered pass emits the survivors in ascending index order with warp-ballot prefix sums.
template <typename T>
__global__ void __launch_bounds__(kSelThreads) topk_select_kernel(const SelectParams p) {
  constexpr int kBits = KeyBits<T>::value;
  __shared__ unsigned hist[256];
  __shared__ unsigned s_wsum[kSelWarps];
  __shared__ unsigned s_digit, s_kk;
  __shared__ unsigned s_wgt[kSelWarps], s_weq[kSelWarps];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T* sc = p.layers == nullptr
                    ? static_cast<const T*>(p.score) + static_cast<int64_t>(blockIdx.x) * p.score_stride
                    : static_cast<const T*>(p.layers[blockIdx.x / p.rows_per_layer].score) +
                          static_cast<int64_t>(blockIdx.x % p.rows_per_layer) * p.score_stride;
  int32_t* keep = p.keep + static_cast<int64_t>(blockIdx.x) * p.keep_stride;
  const int n = p.win_end - p.win_begin;
  const int n_pad = (n + kSelThreads - 1) / kSelThreads * kSelThreads;

  uint32_t prefix = 0;            // digits fixed so far (high bits of the k-th key)
  unsigned kk = p.important;      // rank still to resolve inside the current bucket
  for (int shift = kBits - 8; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int j = tid; j < n_pad; j += kSelThreads) {
      bool in = j < n;
      uint32_t key = 0;
      if (in) {
        key = score_key(sc[p.win_begin + j]);
        in = (shift + 8 >= kBits) || ((key >> (shift + 8)) == prefix);
      }
      const uint32_t digit = (key >> shift) & 255u;
      // warp-aggregated histogram update: one atomic per distinct digit per warp
      const unsigned peers = __match_any_sync(0xffffffffu, in ? digit : (256u + lane));
      if (in && (__ffs(peers) - 1) == lane) atomicAdd(&hist[digit], __popc(peers));
    }
    __syncthreads();
    // descending cumulative count over the 256 buckets: thread t owns bucket 255 - t
    unsigned cnt = 0, incl = 0;
    if (tid < 256) {
      cnt = hist[255 - tid];
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
    __syncthreads();
  }
  const uint32_t kth = prefix;     // key of the k-th largest score
  const unsigned need_eq = kk;     // how many elements equal to it survive (lowest index first)

  // untouched windows: [0, start) and [tail_begin, seq_len)
  for (int j = tid; j < p.start; j += kSelThreads) keep[j] = j;
  for (int j = p.tail_begin + tid; j < p.seq_len; j += kSelThreads)
    keep[p.start + p.important + (j - p.tail_begin)] = j;

  // ordered compaction of the middle window
  unsigned run_gt = 0, run_eq = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int base = 0; base < n_pad; base += kSelThreads) {
    const int j = base + tid;
    uint32_t key = 0;
    const bool in = j < n;
    if (in) key = score_key(sc[p.win_begin + j]);
    const bool gt = in && key > kth, eq = in && key == kth;
    const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { s_wgt[warp] = __popc(bg); s_weq[warp] = __popc(be); }
    __syncthreads();
    unsigned gt_before = run_gt + __popc(bg & lt_mask), eq_before = run_eq + __popc(be & lt_mask);
    unsigned tot_gt = 0, tot_eq = 0;
#pragma unroll
    for (int w = 0; w < kSelWarps; ++w) {
      const unsigned a = s_wgt[w], b = s_weq[w];
      if (w < warp) { gt_before += a; eq_before += b; }
      tot_gt += a; tot_eq += b;
    }
    if (gt || (eq && eq_before < need_eq))
      keep[p.start + gt_before + min(eq_before, need_eq)] = p.win_begin + j;
    run_gt += tot_gt;
    run_eq += tot_eq;
    __syncthreads();
  }
}

// ---- attn_scores [b, H, q, S] -> importance [H, S]: sum over batch, then over query, each rounded to T ----
struct ReduceParams {
  const void* in;
  void* out;
  int b, H, q, S;
  int64_t in_stride_b, in_stride_h, in_stride_q, out_stride;
};

__device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ void from_f(float v, __half& o) { o = __float2half_rn(v); }
__device__ __forceinline__ void from_f(float v, float& o) { o = v; }

template <typename T>
__global__ void score_reduce_kernel(const ReduceParams p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y;
  if (j >= p.S) return;
  const T* in = static_cast<const T*>(p.in) + h * p.in_stride_h + j;
  float outer = 0.f;
  for (int qi = 0; qi < p.q; ++qi) {
    float inner = 0.f;
    for (int bi = 0; bi < p.b; ++bi) inner += to_f(in[bi * p.in_stride_b + qi * p.in_stride_q]);
    T r;
    from_f(inner, r);            // .sum(0) materialises a tensor of the score dtype (:51)
    outer += to_f(r);
  }
  T r;
  from_f(outer, r);              // .sum(1)
  static_cast<T*>(p.out)[h * p.out_stride + j] = r;
}

// ---- gather surviving rows of K and V ---------------------------------------------------------------------
constexpr int kGatherThreads = 256;
constexpr int kGatherTileRows = 64;

struct GatherParams {
  const uint8_t* src[2];
  uint8_t* dst[2];
  int64_t src_stride_b, src_stride_h, dst_stride_b, dst_stride_h;   // bytes
  const int32_t* keep;
  int64_t keep_stride;
  int num_heads, keep_total, row_bytes, in_place;
  const SpattenPruneLayer* layers;   // optional (spatten_prune_layers): blockIdx.z = layer
};

// grid.x = batch * heads * 2 (K then V); grid.y = row tiles (out of place) or 1 (in place: tiles are walked in
// ascending order by one CTA, which is safe because keep[] is strictly ascending so dst row <= src row and a
// tile's sources are read into registers before any of its destinations are written).
template <int kMaxPer>    // 16-byte chunks per thread per 64-row tile: 4 covers rows of <= 256 bytes (fp16 head_dim 128) at half the registers
__global__ void __launch_bounds__(kGatherThreads, kMaxPer <= 4 ? 6 : 3) kv_gather_kernel(const GatherParams p) {
  const int which = blockIdx.x & 1;
  const int bh = blockIdx.x >> 1;
  const int h = bh % p.num_heads, b = bh / p.num_heads;
  const uint8_t* src0 = which ? p.src[1] : p.src[0];
  uint8_t* dst0 = which ? p.dst[1] : p.dst[0];
  const int32_t* keep = p.keep + h * p.keep_stride;
  if (p.layers != nullptr) {
    const SpattenPruneLayer& ly = p.layers[blockIdx.z];
    src0 = static_cast<const uint8_t*>(which ? ly.src_v : ly.src_k);
    dst0 = static_cast<uint8_t*>(which ? ly.dst_v : ly.dst_k);
    keep += static_cast<int64_t>(blockIdx.z) * p.num_heads * p.keep_stride;
  }
  const uint8_t* src = src0 + b * p.src_stride_b + h * p.src_stride_h;
  uint8_t* dst = dst0 + b * p.dst_stride_b + h * p.dst_stride_h;
  const int cpr = p.row_bytes >> 4;                       // 16-byte chunks per row
  const int chunks = kGatherTileRows * cpr;
  const int n_tiles = (p.keep_total + kGatherTileRows - 1) / kGatherTileRows;
  for (int t = p.in_place ? 0 : blockIdx.y; t < n_tiles; t += p.in_place ? 1 : gridDim.y) {
    uint4 v[kMaxPer];
    int drow[kMaxPer];
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
      const int ch = threadIdx.x + i * kGatherThreads;
      drow[i] = -1;
      if (ch < chunks) {
        const int r = t * kGatherTileRows + ch / cpr;
        if (r < p.keep_total) {
          const int srow = keep[r];
          if (!(p.in_place && srow == r)) {
            drow[i] = r;
            v[i] = __ldcs(reinterpret_cast<const uint4*>(src + static_cast<int64_t>(srow) * p.row_bytes + (ch % cpr) * 16));   // read once
          }
        }
      }
    }
    if (p.in_place) __syncthreads();
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
      const int ch = threadIdx.x + i * kGatherThreads;
      if (drow[i] >= 0)
        __stcs(reinterpret_cast<uint4*>(dst + static_cast<int64_t>(drow[i]) * p.row_bytes + (ch % cpr) * 16), v[i]);
    }
    if (p.in_place) __syncthreads();
  }
}

}  // namespace spatten

using namespace spatten;

extern "C" int spatten_prune_plan(int32_t seq_len, int32_t num_coming, int32_t start, int32_t important,
                                  int32_t recent, int32_t* keep_total, int32_t* win_begin, int32_t* win_end) {
  SPATTEN_CHECK_ARG(seq_len >= 0 && start >= 0 && recent >= 0 && num_coming >= 0,
                    "spatten_prune_plan: negative size");
  if (static_cast<int64_t>(seq_len) + num_coming <= static_cast<int64_t>(start) + important + recent) return 1;
  SPATTEN_CHECK_ARG(important > 0, "spatten_prune_plan: important_size must be > 0 (the reference raises TypeError)");
  const int64_t e = static_cast<int64_t>(seq_len) - recent + num_coming;   // > start + important here
  const int wb = start, we = static_cast<int>(std::min<int64_t>(e, seq_len));
  SPATTEN_CHECK_ARG(we - wb >= important, "selected index k out of range (window %d < important %d)", we - wb, important);
  if (keep_total) *keep_total = start + important + (seq_len - we);
  if (win_begin) *win_begin = wb;
  if (win_end) *win_end = we;
  return SPATTEN_OK;
}

extern "C" int spatten_topk_select(const SpattenSelectArgs* a, void* stream) {
  SPATTEN_CHECK_ARG(a != nullptr && a->struct_size == sizeof(SpattenSelectArgs), "spatten_topk_select: bad args/struct_size");
  SPATTEN_CHECK_ARG(a->num_rows > 0 && a->score && a->keep_idx, "spatten_topk_select: NULL pointer or no rows");
  SPATTEN_CHECK_ARG(a->score_dtype == SPATTEN_DTYPE_F16 || a->score_dtype == SPATTEN_DTYPE_F32,
                    "spatten_topk_select: score_dtype must be F16 or F32");
  int keep_total = 0, wb = 0, we = 0;
  const int rc = spatten_prune_plan(a->seq_len, a->num_coming, a->start, a->important, a->recent, &keep_total, &wb, &we);
  if (rc < 0) return rc;
  SPATTEN_CHECK_ARG(rc == 0, "spatten_topk_select: cache already fits (seq_len + num_coming <= cache size); nothing to select");
  SPATTEN_CHECK_ARG(a->score_stride >= a->seq_len && a->keep_stride >= keep_total, "spatten_topk_select: stride too small");
  SelectParams p{a->score, a->score_stride, a->keep_idx, a->keep_stride, a->seq_len, a->start, a->important, wb, we, we,
                 nullptr, 1};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->score_dtype == SPATTEN_DTYPE_F16) topk_select_kernel<__half><<<a->num_rows, kSelThreads, 0, st>>>(p);
  else topk_select_kernel<float><<<a->num_rows, kSelThreads, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "topk_select_kernel launch");
}

extern "C" int spatten_score_reduce(const void* scores, int32_t dtype, int32_t b, int32_t H, int32_t q, int32_t S,
                                    int64_t stride_b, int64_t stride_h, int64_t stride_q, void* out,
                                    int64_t out_stride, void* stream) {
  SPATTEN_CHECK_ARG(scores && out && b > 0 && H > 0 && q > 0 && S > 0, "spatten_score_reduce: bad arguments");
  SPATTEN_CHECK_ARG(dtype == SPATTEN_DTYPE_F16 || dtype == SPATTEN_DTYPE_F32, "spatten_score_reduce: dtype must be F16 or F32");
  ReduceParams p{scores, out, b, H, q, S, stride_b, stride_h, stride_q, out_stride};
  dim3 grid((S + 255) / 256, H);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SPATTEN_DTYPE_F16) score_reduce_kernel<__half><<<grid, 256, 0, st>>>(p);
  else score_reduce_kernel<float><<<grid, 256, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "score_reduce_kernel launch");
}

extern "C" int spatten_kv_gather(const SpattenGatherArgs* a, void* stream) {
  SPATTEN_CHECK_ARG(a != nullptr && a->struct_size == sizeof(SpattenGatherArgs), "spatten_kv_gather: bad args/struct_size");
  SPATTEN_CHECK_ARG(a->batch > 0 && a->num_heads > 0 && a->keep_total > 0, "spatten_kv_gather: bad shape");
  SPATTEN_CHECK_ARG(a->row_bytes > 0 && a->row_bytes % 16 == 0 && a->row_bytes <= 512,
                    "spatten_kv_gather: row_bytes %d must be a multiple of 16 and <= 512", a->row_bytes);
  SPATTEN_CHECK_ARG(a->src_k && a->src_v && a->dst_k && a->dst_v && a->keep_idx, "spatten_kv_gather: NULL pointer");
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->src_k) | reinterpret_cast<uintptr_t>(a->src_v) |
                       reinterpret_cast<uintptr_t>(a->dst_k) | reinterpret_cast<uintptr_t>(a->dst_v) |
                       static_cast<uintptr_t>(a->src_stride_b | a->src_stride_h | a->dst_stride_b | a->dst_stride_h);
  SPATTEN_CHECK_ARG((al & 15) == 0, "spatten_kv_gather: pointers and byte strides must be 16-byte aligned");
  const bool ipk = a->src_k == a->dst_k, ipv = a->src_v == a->dst_v;
  SPATTEN_CHECK_ARG(ipk == ipv, "spatten_kv_gather: K and V must both be in place or both out of place");
  SPATTEN_CHECK_ARG(!ipk || (a->src_stride_b == a->dst_stride_b && a->src_stride_h == a->dst_stride_h),
                    "spatten_kv_gather: in-place gather needs identical strides");
  GatherParams p{};
  p.src[0] = static_cast<const uint8_t*>(a->src_k);
  p.src[1] = static_cast<const uint8_t*>(a->src_v);
  p.dst[0] = static_cast<uint8_t*>(a->dst_k);
  p.dst[1] = static_cast<uint8_t*>(a->dst_v);
  p.src_stride_b = a->src_stride_b; p.src_stride_h = a->src_stride_h;
  p.dst_stride_b = a->dst_stride_b; p.dst_stride_h = a->dst_stride_h;
  p.keep = a->keep_idx; p.keep_stride = a->keep_stride;
  p.num_heads = a->num_heads; p.keep_total = a->keep_total; p.row_bytes = a->row_bytes; p.in_place = ipk ? 1 : 0;
  const int n_tiles = (a->keep_total + kGatherTileRows - 1) / kGatherTileRows;
  dim3 grid(a->batch * a->num_heads * 2, ipk ? 1 : std::min(n_tiles, 64));
  if (a->row_bytes <= 256) kv_gather_kernel<4><<<grid, kGatherThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else kv_gather_kernel<8><<<grid, kGatherThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "kv_gather_kernel launch");
}

// Every layer of a cache in two launches: one select over num_layers * num_heads score rows, one gather over
// (layer, sequence, head, K|V, row tile) -- the Python loop over layers of the reference (kv_cache_token_pruning.py:55-70)
// without per-layer launches, allocations or synchronisation.
extern "C" int spatten_prune_layers(const SpattenPruneLayersArgs* a, void* stream) {
  SPATTEN_CHECK_ARG(a != nullptr && a->struct_size == sizeof(SpattenPruneLayersArgs), "spatten_prune_layers: bad args/struct_size");
  SPATTEN_CHECK_ARG(a->num_layers > 0 && a->batch > 0 && a->num_heads > 0 && a->layers && a->keep_idx,
                    "spatten_prune_layers: bad shape or NULL pointer");
  SPATTEN_CHECK_ARG(a->num_layers <= 65535, "spatten_prune_layers: too many layers");
  SPATTEN_CHECK_ARG(a->score_dtype == SPATTEN_DTYPE_F16 || a->score_dtype == SPATTEN_DTYPE_F32,
                    "spatten_prune_layers: score_dtype must be F16 or F32");
  SPATTEN_CHECK_ARG(a->row_bytes > 0 && a->row_bytes % 16 == 0 && a->row_bytes <= 512,
                    "spatten_prune_layers: row_bytes %d must be a multiple of 16 and <= 512", a->row_bytes);
  SPATTEN_CHECK_ARG(((a->src_stride_b | a->src_stride_h | a->dst_stride_b | a->dst_stride_h) & 15) == 0,
                    "spatten_prune_layers: byte strides must be multiples of 16");
  int keep_total = 0, wb = 0, we = 0;
  const int rc = spatten_prune_plan(a->seq_len, a->num_coming, a->start, a->important, a->recent, &keep_total, &wb, &we);
  if (rc < 0) return rc;
  SPATTEN_CHECK_ARG(rc == 0, "spatten_prune_layers: cache already fits; nothing to select");
  SPATTEN_CHECK_ARG(a->score_stride >= a->seq_len && a->keep_stride >= keep_total, "spatten_prune_layers: stride too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SelectParams sp{nullptr, a->score_stride, a->keep_idx, a->keep_stride, a->seq_len, a->start, a->important, wb, we, we,
                  a->layers, a->num_heads};
  const int rows = a->num_layers * a->num_heads;
  if (a->score_dtype == SPATTEN_DTYPE_F16) topk_select_kernel<__half><<<rows, kSelThreads, 0, st>>>(sp);
  else topk_select_kernel<float><<<rows, kSelThreads, 0, st>>>(sp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "topk_select_kernel launch");
  GatherParams gp{};
  gp.src_stride_b = a->src_stride_b; gp.src_stride_h = a->src_stride_h;
  gp.dst_stride_b = a->dst_stride_b; gp.dst_stride_h = a->dst_stride_h;
  gp.keep = a->keep_idx; gp.keep_stride = a->keep_stride;
  gp.num_heads = a->num_heads; gp.keep_total = keep_total; gp.row_bytes = a->row_bytes; gp.in_place = 0;
  gp.layers = a->layers;
  const int n_tiles = (keep_total + kGatherTileRows - 1) / kGatherTileRows;
  dim3 grid(a->batch * a->num_heads * 2, std::min(n_tiles, 64), a->num_layers);
  if (gp.row_bytes <= 256) kv_gather_kernel<4><<<grid, kGatherThreads, 0, st>>>(gp);
  else kv_gather_kernel<8><<<grid, kGatherThreads, 0, st>>>(gp);
  e = cudaGetLastError();
  return e == cudaSuccess ? SPATTEN_OK : cuda_fail(e, "kv_gather_kernel launch");
}
