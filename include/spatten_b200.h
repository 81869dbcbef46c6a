/*
 * spatten_b200 -- C ABI of the B200-native sparse-attention decode path.
 *
 * Drop-in boundary for mit-han-lab/spatten's `spatten_llm` hot path.  The reference has no FFI of
 * its own (pure PyTorch), so every entry point below cites the reference *Python*
 * interface it replaces; INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / CUDA types in any signature (`stream` is a cudaStream_t
 *     passed as void*; NULL = legacy default stream).
 *   - all data pointers are DEVICE pointers owned by the caller unless the name ends in `_host`.
 *   - the library allocates nothing after the first call on a device (it caches device properties).
 *   - every function returns SPATTEN_OK (0) or a negative SPATTEN_ERR_* code and never throws;
 *     `spatten_last_error()` returns a thread-local human-readable message for the last failure.
 *   - fp16 = IEEE binary16 (CUDA __half), row-major, innermost dimension contiguous.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns SPATTEN_ERR_CUDA.
 */
#ifndef SPATTEN_B200_H_
#define SPATTEN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPATTEN_ABI_VERSION 1

#define SPATTEN_OK 0
#define SPATTEN_ERR_INVALID (-1)      /* bad argument (NULL pointer, shape, alignment) */
#define SPATTEN_ERR_UNSUPPORTED (-2)  /* valid request the kernels do not implement (e.g. head_dim) */
#define SPATTEN_ERR_CUDA (-3)         /* CUDA runtime error, message carries cudaGetErrorString */
#define SPATTEN_ERR_WORKSPACE (-4)    /* workspace too small / not initialised */

int spatten_abi_version(void);
const char* spatten_last_error(void);
/* SM count and compute capability of the current device. */
int spatten_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Decode attention: one new token per sequence against the (pruned) KV cache.
 *
 * Replaces the attention core of `llama_pos_shift_attention_forward` for q_len == 1
 * (spatten_llm/pos_shift/modify_llama.py:86-144): RoPE(Q at position_ids) :92, KV append :95-100
 * (keys cached UN-rotated), RoPE(K at cache-slot index) :103-104, Q.K^T/sqrt(D) :111, logits stash
 * `self.attn_scores` :116-119, additive mask :127-132, fp32 softmax :135, P.V :138 -- in ONE kernel
 * launch, without the O(S) cat / re-rotation / clone passes the reference pays every step.
 *
 * Numerics follow the reference op-for-op where it matters for the pruned index set: rotated keys
 * and queries are rounded to fp16 after each elementwise op exactly as torch does (:27), the dot
 * product accumulates in fp32 and is rounded to fp16, then divided by sqrt(D) (IEEE division) and
 * rounded to fp16 again (:111-113); softmax runs in fp32 on those fp16 logits (:135).
 * ---------------------------------------------------------------------------------------------- */
/* Launch with programmatic dependent launch (PDL): the kernel lets its successor in the stream start
 * early and itself waits (griddepcontrol.wait) for its predecessor before touching any argument. */
#define SPATTEN_DECODE_PDL 1u
/* With SPATTEN_DECODE_PDL: the caller guarantees that the cached rows [0, kv_len) of k_cache/v_cache and
 * the `rope` table were NOT written by any kernel that can still be in flight when this one starts (true in
 * a decoder: a layer's cache was last written one whole decode step earlier).  The kernel then streams the
 * first K/V tiles into shared memory BEFORE waiting for its predecessor, hiding launch + HBM latency. */
#define SPATTEN_DECODE_PDL_KV_STABLE 2u
/* Never use the thread-block-cluster (DSMEM merge) variant; always global partials + tickets. */
#define SPATTEN_DECODE_NO_CLUSTER 4u
/* DotProduct stage only: stream K, export the logits, skip softmax / P.V / out (V is not read; v_new is still
 * appended).  First step of the RTL-ordered path DotProduct -> TopK -> MultiplyValue over survivors
 * (SpAttenController.scala:347,588,668-749); see spatten_sparse_pv. */
#define SPATTEN_DECODE_SCORES_ONLY 8u
/* Use the 8-token-tile build (96 KiB shared memory, <= 128 registers: two CTAs per SM) instead of the default
 * 16-token tiles (192 KiB, one CTA per SM).  Opt-in: on B200 the co-resident CTAs of consecutive PDL launches were
 * measured to compete more than they overlap. */
#define SPATTEN_DECODE_SMALL_TILES 16u

typedef struct SpattenDecodeArgs {
  uint32_t struct_size;   /* = sizeof(SpattenDecodeArgs), ABI guard */
  int32_t batch;          /* b */
  int32_t num_heads;      /* H   (query heads) */
  int32_t num_kv_heads;   /* H_kv (H % H_kv == 0; repeat_kv :108-109 is resolved by indexing) */
  int32_t head_dim;       /* D: 64 or 128 */
  int32_t kv_len;         /* tokens already in the cache (same for every sequence and head, as in the
                             reference where the cache is one [b,H_kv,S,D] tensor) */
  int32_t rope_len;       /* rows in `rope`; must be >= kv_len + 1 and > every q_pos */
  int32_t append;         /* 1: write k_new/v_new into slot kv_len and attend over kv_len+1 tokens
                             (reference behaviour); 0: attend over the kv_len cached tokens only */
  uint32_t flags;         /* SPATTEN_DECODE_* bits above */
  int32_t q_len;          /* query rows per sequence.  1 (or 0) = decode.  > 1 = a prefill / chunk: every query
                             row is scored against ALL kv_len cached tokens (the reference exports the full
                             pre-mask [q, kv] logit matrix, :111-119, and masks additively, :127-132); requires
                             append == 0, i.e. the caller has already stored the new K/V rows in the cache */

  const void* q;          /* fp16 [b, q_len, H, D]  un-rotated queries (q_proj output viewed, NOT transposed) */
  const void* k_new;      /* fp16 [b, H_kv, D]   un-rotated key   (append == 1) */
  const void* v_new;      /* fp16 [b, H_kv, D]   value            (append == 1) */
  void* k_cache;          /* fp16 [b, H_kv, >=kv_len+append, D] un-rotated keys, strides below */
  void* v_cache;          /* fp16 same layout */
  int64_t cache_stride_b; /* elements between sequences  */
  int64_t cache_stride_h; /* elements between kv heads (= capacity * D for a preallocated cache) */
  const void* rope;       /* fp16 [rope_len, D]: row p = cos(p*theta_i) i<D/2 | sin(p*theta_i) i<D/2,
                             i.e. the first halves of HF-4.33 LlamaRotaryEmbedding's cos/sin rows (:89) */
  const int64_t* q_pos;   /* [b, q_len] position_ids of the queries (:92); NULL = kv_len (decode only) */
  const void* mask;       /* optional fp16 additive mask [b, q_len, kv_len+append] (:127-132); NULL = none */
  const uint8_t* head_mask; /* optional [H]; 0 = pruned head: no KV read, output row zeroed */

  void* out;              /* fp16 [b, q_len, H, D] == attn_output[b, q_len, H*D] before o_proj (:146-147) */
  void* logits;           /* optional fp16 attn_scores[b, h, q, :] (:116-119), strides below */
  int64_t logits_stride;  /* elements between heads   (>= kv_len+append) */
  int64_t logits_stride_b;/* elements between sequences; 0 = num_heads * logits_stride (decode layout) */
  int64_t logits_stride_q;/* elements between query rows; ignored when q_len <= 1 */
  float* importance;      /* optional fp32 [b, H, importance_stride]: += softmax probability
                             (cumulative token importance, README.md:21) */
  int64_t importance_stride;
  float* head_importance; /* optional fp32 [H]: += sum_d |out[b,h,d]| (head importance) */

  /* Fused exchange for head sharding (n_peers > 0): the output rows are stored into every peer GPU's gather
   * buffer over NVLink (peer_out[r] = address, mapped in THIS process, of this rank's slot in rank r's buffer;
   * the list includes the rank itself) instead of `out`, and when the whole grid has finished, one flag word per
   * peer (peer_flags[r], in rank r's memory) is incremented with system-scope release.  The consumer calls
   * spatten_wait_flags.  No separate collective kernel runs. */
  int32_t n_peers;        /* 0 = write `out` only */
  void* peer_out[8];
  void* peer_flags[8];
  void* done_counter;     /* local u32, zero-initialised once (self-resetting) */
  /* Consumer side folded into the kernel (needs n_peers > 0): after its K/V prefetch is in flight and before it
   * reads q, every warp waits until the wait_n local flag words at wait_flags (the PREVIOUS layer's exchange, one
   * word per producing rank) have reached *wait_expected + 1; the kernel increments *wait_expected when it is
   * done.  Saves the separate spatten_wait_flags launch between layers. */
  int32_t wait_n;
  const void* wait_flags;
  void* wait_expected;    /* local u32 call counter, zero-initialised once */
  void* wait_timeout;     /* local u32, set to 1 if a peer never reports (~2 s) */

  void* workspace;        /* >= spatten_decode_workspace_bytes(...) bytes, zero-filled ONCE by the caller
                             before first use (the kernel restores its counters to zero) */
  size_t workspace_bytes;
} SpattenDecodeArgs;

size_t spatten_decode_workspace_bytes(int32_t batch, int32_t num_heads, int32_t head_dim);
int spatten_decode_attn(const SpattenDecodeArgs* args, void* stream);

/* Consumer side of the fused exchange: block the stream until each of the n flag words (one per producing rank,
 * in local memory) has been incremented once more than at the previous call.  `expected_counter` is a local u32
 * (zero-initialised once) that counts calls, so the wait is CUDA-graph replayable; `timeout_flag` (u32) is set
 * to 1 if a peer does not show up within ~2 s (the kernel then returns instead of wedging the GPU). */
int spatten_wait_flags(const void* flags, int32_t n, void* expected_counter, void* timeout_flag, void* stream);

/* The same step with HOST buffers for the per-step tensors (what a caller that lives on the CPU side of an
 * FFI hands over): q_host / k_new_host / v_new_host are copied into the device staging buffers args->q /
 * args->k_new / args->v_new, the kernel runs, args->out is copied to out_host and the stream is
 * synchronised before returning.  The KV cache, RoPE table and workspace stay resident on the device (they
 * are state, like the reference's past_key_values).  Pinned host memory makes the copies asynchronous. */
int spatten_decode_attn_host(const SpattenDecodeArgs* args, const void* q_host, const void* k_new_host,
                             const void* v_new_host, void* out_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decode STEP: the attention core of every layer of one decode token in ONE persistent launch.
 *
 * Same per-layer function as spatten_decode_attn (modify_llama.py:86-144 for q_len == 1, no mask -- HF 4.33
 * builds an all-zero mask for a decode step), applied to `num_layers` layers that share one shape.  One CTA
 * per SM stays resident for the whole step; every layer's (row, 16-token tile) work list is cut into equal
 * contiguous ranges over all warps of the grid, and each warp's bulk-copy ring runs ahead ACROSS layer
 * boundaries, so the K/V stream of layer l+1 is already in flight while layer l is being merged -- the
 * per-launch latency (launch, first byte, merge tail) that bounds a 5 us per-layer kernel is paid once per
 * step instead of once per layer.  The K/V cache rows of a layer do not depend on the previous layer; only the
 * query does, and that dependency is kept when asked for:
 *
 *   SPATTEN_STEP_CHAIN  a warp reads the query (and k_new / v_new) of layer l+1 only after EVERY row of layer l
 *                       has been merged and stored (per-layer completion flags, release/acquire at GPU scope) --
 *                       the ordering a decoder imposes (the next layer's q is a function of all heads of this
 *                       one).  K/V prefetch of layer l+1 still overlaps.  With n_peers > 0 the chain runs over the
 *                       fused exchange instead: layer l+1 waits until all n_peers ranks have delivered layer l.
 *
 * Split-KV partials go through a workspace and per-(layer, row) tickets; the merge warp of a row's last CTA merges it.
 * Outputs: layers[l].out (or, n_peers > 0, every peer's gather buffer), layers[l].logits, optional importance /
 * head_importance as in spatten_decode_attn.  head_mask is per layer: u8 [num_layers, num_heads] (<= 64 heads);
 * pruned heads cost no KV traffic and no work slot, their output rows are zeroed, and the appended K/V row is
 * still written (by the first live head of the kv group, or by the zeroing warp when the whole group is pruned).
 * ---------------------------------------------------------------------------------------------- */
#define SPATTEN_STEP_CHAIN 1u
/* Historical: an earlier head-sharded build split the work differently from the local one, and this flag made a local
 * launch sum in the sharded order so that the fused exchange could be checked bit for bit.  Both now share one work split
 * (the flag is accepted and has no effect). */
#define SPATTEN_STEP_SHARDED_LAYOUT 2u

typedef struct SpattenStepLayer {   /* one per layer, array in DEVICE memory */
  const void* q;          /* fp16 [b, H, D] un-rotated */
  const void* k_new;      /* fp16 [b, H_kv, D] (append == 1) */
  const void* v_new;
  void* k_cache;          /* fp16 [b, H_kv, >= kv_len + append, D] */
  void* v_cache;
  void* out;              /* fp16 [b, H, D] (n_peers == 0) */
  void* logits;           /* optional fp16 [b, H, logits_stride] */
  float* importance;      /* optional fp32 [b, H, importance_stride], += softmax probability (needs logits) */
} SpattenStepLayer;

typedef struct SpattenStepArgs {
  uint32_t struct_size;
  int32_t num_layers, batch, num_heads, num_kv_heads, head_dim;
  int32_t kv_len, rope_len, append;     /* as in SpattenDecodeArgs, the same for every layer */
  uint32_t flags;                       /* SPATTEN_STEP_* */
  const SpattenStepLayer* layers;       /* DEVICE pointer to num_layers descriptors */
  int64_t cache_stride_b, cache_stride_h;   /* elements */
  int64_t logits_stride, importance_stride; /* elements between heads */
  const void* rope;                     /* fp16 [rope_len, D] */
  const int64_t* q_pos;                 /* [b] or NULL = kv_len */
  const uint8_t* head_mask;             /* optional [num_layers, num_heads] */
  float* head_importance;               /* optional fp32 [num_heads], += sum_d |out| over all layers */
  /* fused exchange (head sharding), as in SpattenDecodeArgs but addressed per layer:
   * layer l's rows go to peer_out[r] + l * peer_layer_stride (elements) on every peer r, and every finished row
   * adds 1 to the u32 at peer_flags[r] + l * peer_flag_layer_stride (words).  wait_flags is THIS rank's own flag
   * array [num_layers][n_peers] (the words the peers increment); the kernel waits on it between layers
   * (SPATTEN_STEP_CHAIN) and, for the last layer, before it exits -- when the launch completes, every rank's rows
   * of every layer are in the local gather buffer. */
  int32_t n_peers;
  int32_t self_peer;                    /* index of THIS rank in peer_out / peer_flags: rows are stored there first, and the
                                           warp that finishes a layer copies the layer's slot to the other peers in one go
                                           (one system-scope release per layer instead of NVLink stores per row) */
  void* peer_out[8];
  int64_t peer_layer_stride;
  void* peer_flags[8];
  int64_t peer_flag_layer_stride;
  const void* wait_flags;
  void* workspace;                      /* >= spatten_decode_step_workspace_bytes(...), zero-filled ONCE */
  size_t workspace_bytes;               /* u32 word 0: steps completed; word 2: set to 1 if a flag wait timed out */
} SpattenStepArgs;
size_t spatten_decode_step_workspace_bytes(int32_t num_layers, int32_t batch, int32_t num_heads, int32_t head_dim);
int spatten_decode_step(const SpattenStepArgs* args, void* stream);
/* The step with HOST buffers: `in_host` (pinned) holds every layer's q | k_new | v_new exactly as the device region
 * `in_dev` that the layer descriptors point into; one host->device copy, the launch, one device->host copy of
 * `out_bytes` from `out_dev` (every layer's output) to `out_host`, stream synchronised before returning. */
int spatten_decode_step_host(const SpattenStepArgs* args, const void* in_host, void* in_dev, size_t in_bytes,
                             const void* out_dev, void* out_host, size_t out_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Prefill attention (q_len > 1) on the tcgen05 tensor cores.  Same reference lines as spatten_decode_attn
 * (modify_llama.py:86-144); the full pre-mask logit matrix is an output (`self.attn_scores`, :116-119).
 * The new K/V rows must already be in the cache (rows [kv_len - q_len, kv_len) for a plain prefill).  head_dim == 128.
 *
 * Fused path (csrc/prefill_fused.cu), taken when `scratch_bytes` >= spatten_prefill_scratch_bytes(...): Q and K are
 * rotated once and Q/K/V re-tiled into tensor-core operand images in `probs` (used as plain scratch); ONE
 * warp-specialised kernel then computes S = rope(Q).rope(K)^T per 128x128 tile, writes the fp16 logits (:111-113) once,
 * keeps the row max / row sum, recomputes S for the tiles the mask leaves alive, and accumulates O += P.V with the
 * accumulator in tensor memory, plus the column sums of P (`importance`).  The additive mask is classified on the
 * device: the standard causal mask (0 on and below the diagonal shifted by kv_len - q_len, <= -60000 above) becomes a
 * predicate and fully masked tiles are skipped; any other mask is applied element by element (:127-132).
 * Legacy path (csrc/prefill_umma.cu; `scratch_bytes` == 0 or SPATTEN_PREFILL_LEGACY): the reference's own three steps
 * as three launches; `probs` must then be scratch of the same shape and strides as `logits`.
 * ---------------------------------------------------------------------------------------------- */
#define SPATTEN_PREFILL_CAUSAL 1u   /* mask == NULL means "standard causal mask" instead of "no mask" */
#define SPATTEN_PREFILL_LEGACY 2u   /* force the three-launch path */
typedef struct SpattenPrefillArgs {
  uint32_t struct_size;
  int32_t batch, num_heads, num_kv_heads, head_dim, q_len, kv_len, rope_len;
  const void* q;            /* fp16 [b, q_len, H, D] un-rotated */
  const void* k_cache;      /* fp16 [b, H_kv, >= kv_len, D] un-rotated */
  const void* v_cache;
  int64_t cache_stride_b, cache_stride_h;   /* elements */
  const void* rope;         /* fp16 [rope_len, D] */
  const int64_t* q_pos;     /* [b, q_len] */
  const void* mask;         /* optional fp16 additive [b, q_len, kv_len] */
  void* logits;             /* out fp16 attn_scores[b, h, q, :] */
  int64_t logits_stride_b, logits_stride_h, logits_stride_q;   /* elements */
  void* probs;              /* scratch (16-byte aligned): operand images (fused) or fp16 probabilities with the logits' strides (legacy) */
  void* out;                /* fp16 [b, q_len, H, D] */
  float* importance;        /* optional fp32 [b, H, importance_stride]: += sum over the q_len query rows of the softmax
                               probabilities (column sums of P: the cumulative token importance of README.md:21, SURVEY.md
                               8f rank 2) */
  int64_t importance_stride;/* elements between heads (>= kv_len) */
  uint32_t flags;           /* SPATTEN_PREFILL_* */
  uint32_t reserved0;
  int64_t scratch_bytes;    /* size of `probs` in bytes; 0 = legacy contract (same shape as logits) */
} SpattenPrefillArgs;
/* bytes of scratch the fused path needs for these shapes */
int64_t spatten_prefill_scratch_bytes(int32_t batch, int32_t num_heads, int32_t num_kv_heads, int32_t q_len, int32_t kv_len);
int spatten_prefill_attn(const SpattenPrefillArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * KV-cache token pruning.  Replaces SpAttenKVCache.apply_token_pruning
 * (spatten_llm/kv_cache_token_pruning.py:42-96) with three device-side steps and no host sync
 * (the reference synchronises once per layer through `.cpu()`, :66).
 * ---------------------------------------------------------------------------------------------- */
#define SPATTEN_DTYPE_F16 0
#define SPATTEN_DTYPE_F32 1

/* Host-only arithmetic of :45-47, :59, :78-80.  Returns 1 when seq_len + num_coming already fits in
 * start+important+recent (the reference returns past_key_values unchanged, :46-47), SPATTEN_OK (0) when
 * pruning applies -- then *keep_total = start + important + max(0, seq_len - win_end) and the candidate
 * window is [win_begin, win_end) = [start, min(seq_len, seq_len - recent + num_coming)) -- or
 * SPATTEN_ERR_INVALID for important <= 0 (reference: TypeError at :63) and for a window shorter than
 * `important` (reference: RuntimeError "selected index k out of range" from torch.topk, :60). */
int spatten_prune_plan(int32_t seq_len, int32_t num_coming, int32_t start, int32_t important, int32_t recent,
                       int32_t* keep_total, int32_t* win_begin, int32_t* win_end);

/* importance_score = attn_scores.sum(0).sum(1) (:51): [b, H, q, S] -> [H, S], summed over batch then
 * over query in fp32 and rounded to the score dtype after each of the two sums, as torch does.
 * Strides in elements; the innermost (S) dimension is contiguous. */
int spatten_score_reduce(const void* scores, int32_t dtype, int32_t b, int32_t H, int32_t q, int32_t S,
                         int64_t stride_b, int64_t stride_h, int64_t stride_q, void* out, int64_t out_stride,
                         void* stream);

/* Per score row: the kept token list  [0,start) U top-`important` of score[win_begin:win_end) U
 * [win_end, seq_len)  in ascending order (:59-63 topk + sort, :74-93 the three cat pieces).
 * Ties at the k-th value are broken lowest-index-first (the RTL engine's rule, TopK.scala:159-160,
 * 209-211); torch.topk leaves the choice unspecified.  One CTA per row: 8-bit radix select on
 * order-preserving keys with warp-aggregated histograms, then a ballot/prefix-sum compaction. */
typedef struct SpattenSelectArgs {
  uint32_t struct_size;
  int32_t num_rows;       /* independent rows (heads of one layer, or layers*heads) */
  int32_t seq_len;        /* S */
  int32_t num_coming, start, important, recent;
  int32_t score_dtype;    /* SPATTEN_DTYPE_F16 / _F32 */
  const void* score;      /* [num_rows, score_stride] */
  int64_t score_stride;   /* elements, >= seq_len */
  int32_t* keep_idx;      /* out [num_rows, keep_stride], first keep_total entries written */
  int64_t keep_stride;    /* elements, >= keep_total */
} SpattenSelectArgs;
int spatten_topk_select(const SpattenSelectArgs* args, void* stream);

/* dst[b, h, p, :] = src[b, h, keep_idx[h, p], :] for K and V (:65-93; the bool-mask gather plus the two
 * cat()s).  The same per-head index list is applied to every sequence of the batch (the reference sums the
 * score over the batch, :51, and itself only runs for batch 1, :68).  Byte strides, rows of `row_bytes`
 * contiguous bytes (multiple of 16, <= 512) -- dtype-agnostic, so fp16, fp32 and packed int4/int8 planes
 * all go through it.  src == dst (same strides) compacts IN PLACE. */
typedef struct SpattenGatherArgs {
  uint32_t struct_size;
  int32_t batch, num_heads, keep_total, row_bytes;
  const void* src_k;
  const void* src_v;
  int64_t src_stride_b, src_stride_h;   /* bytes */
  void* dst_k;
  void* dst_v;
  int64_t dst_stride_b, dst_stride_h;   /* bytes */
  const int32_t* keep_idx;              /* [num_heads, keep_stride] ascending */
  int64_t keep_stride;
} SpattenGatherArgs;
int spatten_kv_gather(const SpattenGatherArgs* args, void* stream);

/* Every layer of a cache in two launches (the reference loops over layers in Python, :55-70): per-layer pointers come
 * from a DEVICE array; all layers share shapes and strides.  keep_idx: [num_layers, num_heads, keep_stride].  Out of
 * place (fresh destination buffers, as the reference's cat()). */
typedef struct SpattenPruneLayer {
  const void* score;      /* [num_heads, score_stride] importance of this layer (score_reduce output, or attn_scores
                             itself when batch == q_len == 1) */
  const void* src_k;
  const void* src_v;
  void* dst_k;
  void* dst_v;
} SpattenPruneLayer;
typedef struct SpattenPruneLayersArgs {
  uint32_t struct_size;
  int32_t num_layers, batch, num_heads, seq_len, num_coming, start, important, recent, score_dtype, row_bytes;
  const SpattenPruneLayer* layers;          /* DEVICE pointer */
  int64_t score_stride;                     /* elements */
  int64_t src_stride_b, src_stride_h, dst_stride_b, dst_stride_h;   /* bytes */
  int32_t* keep_idx;
  int64_t keep_stride;
} SpattenPruneLayersArgs;
int spatten_prune_layers(const SpattenPruneLayersArgs* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * RTL-named stages that the reference Python never executes (README.md:11,21; "parity unpinned":
 * checked against the self-consistent restatements in oracle/spatten_oracle.py).
 * ---------------------------------------------------------------------------------------------- */

/* "Local V pruning": MultiplyValue over the top-k survivors only (SpAttenController.scala:554,588-591,
 * 668-726,749).  For every (b,h) row: p = softmax(logits[row, keep_idx[row, :]]) in fp32,
 * out[row, :] = sum_i p_i * V[row, keep_idx[row, i], :].  Only the kept V rows are read from HBM.
 * logits: fp16 [rows, logits_stride] (spatten_decode_attn with SPATTEN_DECODE_SCORES_ONLY),
 * keep_idx: int32 [rows, keep_stride] ascending (spatten_topk_select with start = recent = 0),
 * v_cache: fp16 [batch, num_kv_heads, >= max index + 1, head_dim]; rows = batch * num_heads. */
typedef struct SpattenSparsePvArgs {
  uint32_t struct_size;
  int32_t batch, num_heads, num_kv_heads, head_dim, keep;
  const void* logits;
  int64_t logits_stride;
  const int32_t* keep_idx;
  int64_t keep_stride;
  const void* v_cache;
  int64_t cache_stride_b, cache_stride_h;   /* elements */
  void* out;                                 /* fp16 [batch, num_heads, head_dim] */
} SpattenSparsePvArgs;
int spatten_sparse_pv(const SpattenSparsePvArgs* args, void* stream);

/* The same RTL-ordered path in ONE launch: DotProduct over all kv_len cached keys -> TopK(keep) -> MultiplyValue over the
 * survivors, one thread-block cluster per (sequence, head) row, logits and the survivor list kept in (distributed)
 * shared memory between the stages.  The query's own K/V row must already be in the cache (append it first);
 * kv_len <= 65408.  Outputs: out fp16 [batch, num_heads, head_dim]; logits fp16 [batch*num_heads, logits_stride] (all
 * kv_len scores); keep_idx int32 [batch*num_heads, keep_stride] ascending (min(keep, kv_len) entries). */
typedef struct SpattenLocalVArgs {
  uint32_t struct_size;
  int32_t batch, num_heads, num_kv_heads, head_dim, kv_len, rope_len, keep;
  const void* q;                             /* fp16 [batch, num_heads, head_dim] un-rotated */
  const void* k_cache;                       /* fp16 [batch, num_kv_heads, >= kv_len, head_dim] un-rotated */
  const void* v_cache;
  int64_t cache_stride_b, cache_stride_h;    /* elements */
  const void* rope;                          /* fp16 [rope_len, head_dim] */
  const int64_t* q_pos;                      /* [batch] or NULL = kv_len - 1 */
  void* out;
  void* logits;
  int64_t logits_stride;
  int32_t* keep_idx;
  int64_t keep_stride;
} SpattenLocalVArgs;
int spatten_local_v_attn(const SpattenLocalVArgs* args, void* stream);

/* Progressive quantisation of the KV cache (README.md:11 "KV progressive quantization"; BASELINE config 4).
 * Per (batch, kv head): K and V each stored as an MSB and an LSB 4-bit plane, u8 [tokens, head_dim/2], of the
 * offset-128 int8 code u = clamp(rint(x / scale), -127, 127) + 128 (msb = u >> 4, lsb = u & 15), plus one fp16
 * scale = max|x| / 127 per token (MatrixFetcher.scala:48-51 profiles (4,1)/(8,1); :344-347 MSB left-aligned, LSB
 * right-aligned).  Nibble order inside every 8-dim group: [d0,d2,d4,d6,d1,d3,d5,d7]; the sixteen
 * 8-dim groups of a row are stored in the order [0,1,8,9, 2,3,10,11, 4,5,12,13, 6,7,14,15] (a 16-byte chunk = 16 dims +
 * their RoPE partners).  Token capacity of every
 * buffer must be a multiple of 8.  head_dim == 128 only. */
typedef struct SpattenQuantKV {
  void* k_msb; void* k_lsb; void* k_scale;
  void* v_msb; void* v_lsb; void* v_scale;
  int64_t plane_stride_b, plane_stride_h;   /* bytes between sequences / kv heads of a plane */
  int64_t scale_stride_b, scale_stride_h;   /* elements between sequences / kv heads of a scale array */
} SpattenQuantKV;

/* Quantise `n_tokens` fp16 rows per (batch, kv head) of K and V (strides in elements) into slots
 * [slot0, slot0 + n_tokens) of the planes: the cache write path that precedes the decode kernel. */
int spatten_kv_quantize(const void* k, const void* v, int64_t stride_b, int64_t stride_h, int64_t stride_t,
                        int32_t batch, int32_t num_kv_heads, int32_t n_tokens, int32_t head_dim,
                        const SpattenQuantKV* dst, int32_t slot0, void* stream);

/* Decode attention over the quantised cache with the RTL's MSB-first / requant-replay control flow:
 * pass 1 scores with the MSB plane only (bucket mid-points) and runs softmax and P.V; RequantDecision
 * (RequantDecision.scala:47-72) flags every (batch, head) row whose largest softmax probability is below
 * `requant_threshold`; pass 2 replays exactly those rows with MSB+LSB keys (SpAttenController.scala:140-175).
 * Values use the MSB plane (v_bits = 4) or both planes (v_bits = 8).  The query's token must already be in the
 * cache (spatten_kv_quantize); q_pos NULL = kv_len - 1.  requant_threshold <= 0 skips pass 2. */
typedef struct SpattenQuantDecodeArgs {
  uint32_t struct_size;
  int32_t batch, num_heads, num_kv_heads, head_dim, kv_len, rope_len, v_bits;
  float requant_threshold;
  SpattenQuantKV kv;
  const void* q;            /* fp16 [b, H, D] un-rotated */
  const void* rope;         /* fp16 [rope_len, D] (as spatten_decode_attn) */
  const int64_t* q_pos;     /* [b] or NULL */
  void* out;                /* fp16 [b, H, D] */
  void* logits;             /* optional fp16 [b*H, logits_stride] */
  int64_t logits_stride;
  uint8_t* need_requant;    /* out [b*H]: 1 = row was replayed with the LSB plane */
  void* workspace;          /* >= spatten_decode_quant_workspace_bytes(...), zero-filled ONCE by the caller before first
                               use (split partials + per-row tickets; the kernel restores the tickets to zero) */
  size_t workspace_bytes;
} SpattenQuantDecodeArgs;
size_t spatten_decode_quant_workspace_bytes(int32_t batch, int32_t num_heads, int32_t head_dim);
int spatten_decode_attn_quant(const SpattenQuantDecodeArgs* args, void* stream);

/* Head pruning (README.md:21 "head pruning to remove unessential heads"): mask[h] = 1 for the `keep_heads`
 * heads with the largest cumulative importance (fp32 [num_heads], accumulated by spatten_decode_attn's
 * `head_importance` output), ties -> lowest head index; 0 elsewhere.  The mask is what spatten_decode_attn
 * takes as `head_mask`: pruned heads read no KV at all. */
int spatten_head_select(const float* head_importance, int32_t num_heads, int32_t keep_heads, uint8_t* head_mask,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPATTEN_B200_H_ */
