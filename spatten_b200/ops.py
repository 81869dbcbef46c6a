"""Tensor-level wrappers over the C ABI: torch supplies device memory and the current stream only."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

_workspaces = {}


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("spatten_b200: tensors must live on a CUDA device (no CPU path exists)")


def decode_workspace(device, batch: int, num_heads: int, head_dim: int) -> torch.Tensor:
    """Zero-initialised scratch for split-KV partials + tickets, one per (device, stream)."""
    lib = _lib.load()
    need = lib.spatten_decode_workspace_bytes(batch, num_heads, head_dim)
    key = (torch.device(device).index, _stream_ptr(device))
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def rope_table(head_dim: int, length: int, device, base: float = 10000.0) -> torch.Tensor:
    """fp16 [length, D]: row p = cos(p*theta_i) | sin(p*theta_i), i < D/2.

    Values are those of transformers-4.33 LlamaRotaryEmbedding (reference call site
    pos_shift/modify_llama.py:89): fp32 outer(arange, inv_freq) -> cos/sin -> cast to fp16.  Built on
    the host so that every platform rounds the same fp32 cos/sin the pinned oracle uses.
    """
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    freqs = torch.einsum("i,j->ij", torch.arange(length, dtype=torch.float32), inv_freq)
    tab = torch.cat((freqs.cos(), freqs.sin()), dim=-1).to(torch.float16)
    return tab.to(device).contiguous()


def make_decode_args(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, kv_len: int, rope: torch.Tensor,
                     k_new: Optional[torch.Tensor] = None, v_new: Optional[torch.Tensor] = None,
                     q_pos: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                     head_mask: Optional[torch.Tensor] = None, logits: Optional[torch.Tensor] = None,
                     importance: Optional[torch.Tensor] = None, head_importance: Optional[torch.Tensor] = None,
                     out: Optional[torch.Tensor] = None, flags: int = 0):
    """Validate tensors and fill a `SpattenDecodeArgs`.  Returns (args, out, keepalive).

    Callers that launch the same step repeatedly (bench, CUDA-graph capture) build the struct once and
    call `lib.spatten_decode_attn(byref(args), stream)` directly.
    """
    _lib.load()
    _require_cuda(q, k_cache, v_cache, rope, k_new, v_new, q_pos, mask, head_mask, logits, importance,
                  head_importance, out)
    if q.dim() == 3:
        b, H, D = q.shape
        q_len = 1
    else:
        b, q_len, H, D = q.shape
    H_kv = k_cache.shape[1]
    if q.dtype != torch.float16 or k_cache.dtype != torch.float16 or v_cache.dtype != torch.float16:
        raise TypeError("spatten_b200.decode_attn: fp16 tensors required")
    if k_cache.stride(3) != 1 or k_cache.stride(2) != D or v_cache.stride() != k_cache.stride():
        raise ValueError("spatten_b200.decode_attn: cache rows must be contiguous [.., S, D] with equal K/V strides")
    if not q.is_contiguous():
        q = q.contiguous()
    append = k_new is not None
    n_total = int(kv_len) + int(append)
    if append:
        k_new, v_new = k_new.contiguous(), v_new.contiguous()
    if out is None:
        out = None if (flags & _lib.DECODE_SCORES_ONLY) else torch.empty(q.shape, dtype=torch.float16, device=q.device)
    elif not out.is_contiguous() or out.numel() != q.numel():
        raise ValueError("out must be contiguous with q's shape")
    a = _lib.DecodeArgs()
    a.struct_size = C.sizeof(_lib.DecodeArgs)
    a.batch, a.num_heads, a.num_kv_heads, a.head_dim = b, H, H_kv, D
    a.kv_len, a.rope_len, a.append, a.flags, a.q_len = int(kv_len), rope.shape[0], int(append), int(flags), q_len
    a.q, a.k_cache, a.v_cache = q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr()
    a.k_new = k_new.data_ptr() if append else None
    a.v_new = v_new.data_ptr() if append else None
    a.cache_stride_b, a.cache_stride_h = k_cache.stride(0), k_cache.stride(1)
    a.rope = rope.data_ptr()
    if q_pos is not None:
        if q_pos.dtype != torch.int64 or q_pos.numel() != b * q_len:
            raise TypeError("q_pos must be int64 with b*q_len entries (HF position_ids)")
        q_pos = q_pos.contiguous()
        a.q_pos = q_pos.data_ptr()
    if mask is not None:
        if mask.dtype != torch.float16 or mask.numel() != b * q_len * n_total or not mask.is_contiguous():
            raise ValueError("mask must be contiguous fp16 [b, q_len, kv_len+append]")
        a.mask = mask.data_ptr()
    if head_mask is not None:
        if head_mask.dtype != torch.uint8 or head_mask.numel() != H or not head_mask.is_contiguous():
            raise ValueError("head_mask must be contiguous uint8 [H] (the kernel reads one byte per head)")
        a.head_mask = head_mask.data_ptr()
    a.out = out.data_ptr() if out is not None else None
    if logits is not None:
        if logits.dtype != torch.float16 or logits.stride(-1) != 1:
            raise ValueError("logits must be fp16 with a contiguous last dimension")
        a.logits, a.logits_stride_b, a.logits_stride = logits.data_ptr(), logits.stride(0), logits.stride(1)
        if q_len > 1:
            a.logits_stride_q = logits.stride(2)
    if importance is not None:
        if (importance.dtype != torch.float32 or importance.dim() != 3 or importance.shape[:2] != (b, H)
                or importance.shape[2] < n_total or importance.stride(2) != 1
                or importance.stride(0) != H * importance.stride(1)):
            raise ValueError("importance must be fp32 [b, H, >= kv_len+append] with contiguous tokens and (b, H) rows "
                             "one stride apart (the kernel indexes row * stride)")
        if logits is None:
            raise ValueError("importance needs a logits buffer (probabilities are recomputed from the exported logits)")
        a.importance, a.importance_stride = importance.data_ptr(), importance.stride(1)
    if head_importance is not None:
        if head_importance.dtype != torch.float32 or head_importance.numel() != H or not head_importance.is_contiguous():
            raise ValueError("head_importance must be contiguous fp32 [H]")
        a.head_importance = head_importance.data_ptr()
    ws = decode_workspace(q.device, b * q_len, H, D)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    keep = (q, k_new, v_new, q_pos, mask, ws)      # tensors the struct points at but the caller may not hold
    return a, out, keep


def decode_attn(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, kv_len: int, rope: torch.Tensor,
                **kw) -> torch.Tensor:
    """Attention of new-token queries against the KV cache -- `spatten_decode_attn`.

    q: fp16 [b,H,D] (decode) or [b,q_len,H,D] (prefill rows; the new K/V must already be in the cache
    and k_new must be None); k_cache/v_cache [b,H_kv,cap,D] fp16 (any batch/head strides, rows
    contiguous); kv_len tokens valid; k_new/v_new [b,H_kv,D] appended at slot kv_len when given.
    q_pos int64 [b] / [b,q_len]; mask fp16 [b,n] / [b,q_len,n] additive.
    Returns out fp16 with q's shape.  `logits` receives attn_scores: [b,H,>=n] (decode) or
    [b,H,q_len,>=n]; `importance` fp32 [b,H,>=n] and `head_importance` fp32 [H] accumulate in place.
    Keyword arguments as in `make_decode_args`.
    """
    a, out, _keep = make_decode_args(q, k_cache, v_cache, kv_len, rope, **kw)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().spatten_decode_attn(C.byref(a), C.c_void_p(_stream_ptr(q.device))))
    return out


class DecodeStep:
    """One decode token through the attention core of EVERY layer in one persistent launch -- `spatten_decode_step`.

    q / k_new / v_new: fp16 [L, b, H(_kv), D]; k_caches / v_caches: per-layer fp16 [b, H_kv, cap, D] (equal strides);
    out fp16 [L, b, H, D]; logits optional fp16 [L, b, H, >= kv_len+append]; importance optional fp32 like logits;
    head_mask optional u8 [L, H]; head_importance optional fp32 [H].  `chain=True` keeps the decoder's ordering:
    layer l+1's query is read only after every row of layer l has been stored.  The struct is built once;
    `launch()` enqueues the step on the current (or given) stream and is CUDA-graph capturable.
    """

    def __init__(self, q, k_caches, v_caches, kv_len: int, rope, k_new=None, v_new=None, q_pos=None, out=None,
                 logits=None, importance=None, head_mask=None, head_importance=None, chain: bool = True,
                 sharded_layout: bool = False):
        lib = _lib.load()
        _require_cuda(q, rope, k_new, v_new, q_pos, out, logits, importance, head_mask, head_importance, *k_caches, *v_caches)
        L, b, H, D = q.shape
        H_kv = k_caches[0].shape[1]
        if len(k_caches) != L or len(v_caches) != L:
            raise ValueError("DecodeStep: one K and one V cache per layer")
        if q.dtype != torch.float16 or not q.is_contiguous():
            raise TypeError("DecodeStep: q must be contiguous fp16 [L, b, H, D]")
        st = k_caches[0].stride()
        for t in list(k_caches) + list(v_caches):
            if t.dtype != torch.float16 or t.stride() != st or t.stride(3) != 1 or t.stride(2) != D or t.shape[1] != H_kv:
                raise ValueError("DecodeStep: caches must be fp16 [b, H_kv, cap, D] with contiguous rows and equal strides")
        append = k_new is not None
        if append and not (k_new.is_contiguous() and v_new.is_contiguous() and k_new.shape == (L, b, H_kv, D)
                           and v_new.shape == k_new.shape and k_new.dtype == torch.float16 and v_new.dtype == torch.float16):
            raise ValueError("DecodeStep: k_new / v_new must be contiguous fp16 [L, b, H_kv, D]")
        n_total = int(kv_len) + int(append)
        dev = q.device
        if out is None:
            out = torch.empty((L, b, H, D), dtype=torch.float16, device=dev)
        elif out.shape != q.shape or not out.is_contiguous() or out.dtype != torch.float16:
            raise ValueError("DecodeStep: out must be contiguous fp16 [L, b, H, D]")
        for name, t, dt in (("logits", logits, torch.float16), ("importance", importance, torch.float32)):
            if t is not None and (t.dtype != dt or t.dim() != 4 or t.shape[:3] != (L, b, H) or t.stride(3) != 1
                                  or t.shape[3] < n_total or t.stride(1) != H * t.stride(2)):
                raise ValueError(f"DecodeStep: {name} must be {dt} [L, b, H, >= kv_len+append], heads x tokens contiguous")
        if importance is not None and logits is None:
            raise ValueError("DecodeStep: importance needs logits")
        if head_mask is not None and (head_mask.dtype != torch.uint8 or head_mask.shape != (L, H) or not head_mask.is_contiguous()):
            raise ValueError("DecodeStep: head_mask must be contiguous uint8 [L, H]")
        if head_importance is not None and (head_importance.dtype != torch.float32 or head_importance.numel() != H
                                            or not head_importance.is_contiguous()):
            raise ValueError("DecodeStep: head_importance must be contiguous fp32 [H]")
        if q_pos is not None and (q_pos.dtype != torch.int64 or q_pos.numel() != b or not q_pos.is_contiguous()):
            raise TypeError("DecodeStep: q_pos must be contiguous int64 [b]")
        table = torch.zeros((L, 8), dtype=torch.int64)
        for l in range(L):
            table[l, 0] = q[l].data_ptr()
            table[l, 1] = k_new[l].data_ptr() if append else 0
            table[l, 2] = v_new[l].data_ptr() if append else 0
            table[l, 3], table[l, 4] = k_caches[l].data_ptr(), v_caches[l].data_ptr()
            table[l, 5] = out[l].data_ptr()
            table[l, 6] = logits[l].data_ptr() if logits is not None else 0
            table[l, 7] = importance[l].data_ptr() if importance is not None else 0
        self.table = table.to(dev)
        need = lib.spatten_decode_step_workspace_bytes(L, b, H, D)
        self.workspace = torch.zeros(need, dtype=torch.uint8, device=dev)
        a = _lib.StepArgs()
        a.struct_size = C.sizeof(_lib.StepArgs)
        a.num_layers, a.batch, a.num_heads, a.num_kv_heads, a.head_dim = L, b, H, H_kv, D
        a.kv_len, a.rope_len, a.append = int(kv_len), rope.shape[0], int(append)
        a.flags = (_lib.STEP_CHAIN if chain else 0) | (_lib.STEP_SHARDED_LAYOUT if sharded_layout else 0)
        a.layers = self.table.data_ptr()
        a.cache_stride_b, a.cache_stride_h = st[0], st[1]
        a.logits_stride = logits.stride(2) if logits is not None else 0
        a.importance_stride = importance.stride(2) if importance is not None else 0
        a.rope = rope.data_ptr()
        a.q_pos = q_pos.data_ptr() if q_pos is not None else None
        a.head_mask = head_mask.data_ptr() if head_mask is not None else None
        a.head_importance = head_importance.data_ptr() if head_importance is not None else None
        a.workspace, a.workspace_bytes = self.workspace.data_ptr(), self.workspace.numel()
        self.args, self.out, self.device = a, out, dev
        self._keep = (q, k_new, v_new, list(k_caches), list(v_caches), rope, q_pos, logits, importance, head_mask,
                      head_importance)

    def launch(self, stream_ptr: Optional[int] = None):
        sp = _stream_ptr(self.device) if stream_ptr is None else stream_ptr
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().spatten_decode_step(C.byref(self.args), C.c_void_p(sp)))
        return self.out

    def steps_completed(self) -> int:
        return int(self.workspace[:4].view(torch.int32).item())

    def timed_out(self) -> bool:
        return bool(self.workspace[8:12].view(torch.int32).item())


def prefill_attn(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, kv_len: int, rope: torch.Tensor,
                 q_pos: torch.Tensor, mask: Optional[torch.Tensor] = None, logits: Optional[torch.Tensor] = None,
                 importance: Optional[torch.Tensor] = None, causal: bool = False, legacy: bool = False, _debug: int = 0):
    """Prefill / chunk attention on the tcgen05 tensor cores -- `spatten_prefill_attn` (head_dim 128).

    q fp16 [b,q_len,H,D] un-rotated; the new K/V rows are already in k_cache/v_cache [b,H_kv,cap,D]; kv_len rows
    valid; q_pos int64 [b,q_len]; mask fp16 additive [b,q_len,kv_len] (a standard causal mask is recognised on the
    device and turned into tile skipping); `causal=True` with mask=None applies the causal mask without a mask tensor.
    Returns (out [b,q_len,H,D], logits [b,H,q_len,kv_len] == attn_scores).  `importance` fp32 [b,H,>=kv_len]
    accumulates the column sums of the softmax probabilities (cumulative token importance) in place.
    `legacy=True` runs the three-launch path (csrc/prefill_umma.cu) instead of the fused kernel.
    """
    lib = _lib.load()
    _require_cuda(q, k_cache, v_cache, rope, q_pos, mask, logits)
    b, q_len, H, D = q.shape
    H_kv = k_cache.shape[1]
    if q.dtype != torch.float16 or k_cache.dtype != torch.float16 or v_cache.dtype != torch.float16:
        raise TypeError("spatten_b200.prefill_attn: fp16 tensors required")
    if k_cache.stride(3) != 1 or k_cache.stride(2) != D or v_cache.stride() != k_cache.stride():
        raise ValueError("spatten_b200.prefill_attn: cache rows must be contiguous with equal K/V strides")
    q = q.contiguous()
    if logits is None:
        logits = torch.empty((b, H, q_len, kv_len), dtype=torch.float16, device=q.device)
    if logits.stride(3) != 1:
        raise ValueError("logits last dimension must be contiguous")
    if legacy:
        probs = torch.empty_strided(logits.shape, logits.stride(), dtype=torch.float16, device=q.device)
    else:       # fused path: the scratch only holds the rotated / re-tiled operand images
        probs = torch.empty(int(lib.spatten_prefill_scratch_bytes(b, H, H_kv, q_len, int(kv_len))), dtype=torch.uint8, device=q.device)
    out = torch.empty((b, q_len, H, D), dtype=torch.float16, device=q.device)
    a = _lib.PrefillArgs()
    a.struct_size = C.sizeof(_lib.PrefillArgs)
    a.batch, a.num_heads, a.num_kv_heads, a.head_dim = b, H, H_kv, D
    a.q_len, a.kv_len, a.rope_len = q_len, int(kv_len), rope.shape[0]
    a.q, a.k_cache, a.v_cache = q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr()
    a.cache_stride_b, a.cache_stride_h = k_cache.stride(0), k_cache.stride(1)
    a.rope = rope.data_ptr()
    if q_pos.dtype != torch.int64 or q_pos.numel() != b * q_len:
        raise TypeError("q_pos must be int64 [b, q_len]")
    q_pos = q_pos.contiguous()
    a.q_pos = q_pos.data_ptr()
    if mask is not None:
        if mask.dtype != torch.float16 or mask.numel() != b * q_len * kv_len or not mask.is_contiguous():
            raise ValueError("mask must be contiguous fp16 [b, q_len, kv_len]")
        a.mask = mask.data_ptr()
    a.logits, a.probs, a.out = logits.data_ptr(), probs.data_ptr(), out.data_ptr()
    if importance is not None:
        if (importance.dtype != torch.float32 or importance.dim() != 3 or importance.shape[:2] != (b, H)
                or importance.shape[2] < kv_len or importance.stride(2) != 1 or importance.stride(0) != H * importance.stride(1)):
            raise ValueError("importance must be fp32 [b, H, >= kv_len] with contiguous tokens")
        _require_cuda(importance)
        a.importance, a.importance_stride = importance.data_ptr(), importance.stride(1)
    a.logits_stride_b, a.logits_stride_h, a.logits_stride_q = logits.stride(0), logits.stride(1), logits.stride(2)
    a.flags = (_lib.PREFILL_CAUSAL if causal else 0) | (_lib.PREFILL_LEGACY if legacy else 0) | ((_debug & 7) << 8)
    a.scratch_bytes = 0 if legacy else probs.numel()
    if causal and mask is not None:
        raise ValueError("causal=True takes no mask tensor")
    with torch.cuda.device(q.device):
        _lib.check(lib.spatten_prefill_attn(C.byref(a), C.c_void_p(_stream_ptr(q.device))))
    return out, logits


# ------------------------------------------------------------------------------------------------------
# token pruning
# ------------------------------------------------------------------------------------------------------

def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float16:
        return _lib.DTYPE_F16
    if t.dtype == torch.float32:
        return _lib.DTYPE_F32
    raise TypeError(f"spatten_b200: scores must be fp16 or fp32, got {t.dtype}")


def prune_plan(seq_len: int, num_coming: int, start: int, important: int, recent: int):
    """Host arithmetic of the reference's prune (kv_cache_token_pruning.py:45-47,59,78-80).

    Returns None when the cache already fits, else (keep_total, win_begin, win_end).  Raises what the
    reference raises: TypeError for important_size == 0 (:63), RuntimeError when the candidate window
    is shorter than `important` (torch.topk, :60).
    """
    lib = _lib.load()
    kt, wb, we = C.c_int32(), C.c_int32(), C.c_int32()
    rc = lib.spatten_prune_plan(seq_len, num_coming, start, important, recent, C.byref(kt), C.byref(wb), C.byref(we))
    if rc == 1:
        return None
    if rc != 0:
        msg = lib.spatten_last_error().decode()
        if "important_size" in msg:
            raise TypeError(msg)
        raise RuntimeError(msg)
    return kt.value, wb.value, we.value


def score_reduce(attn_scores: torch.Tensor) -> torch.Tensor:
    """[b,H,q,S] -> [H,S] == attn_scores.sum(0).sum(1) (kv_cache_token_pruning.py:51)."""
    lib = _lib.load()
    _require_cuda(attn_scores)
    b, H, q, S = attn_scores.shape
    if attn_scores.stride(3) != 1:
        attn_scores = attn_scores.contiguous()
    if b == 1 and q == 1:
        return attn_scores[0, :, 0, :]          # both sums are over singleton dims: exact identity
    out = torch.empty((H, S), dtype=attn_scores.dtype, device=attn_scores.device)
    with torch.cuda.device(attn_scores.device):
        _lib.check(lib.spatten_score_reduce(attn_scores.data_ptr(), _dtype_code(attn_scores), b, H, q, S,
                                            attn_scores.stride(0), attn_scores.stride(1), attn_scores.stride(2),
                                            out.data_ptr(), out.stride(0), C.c_void_p(_stream_ptr(out.device))))
    return out


def topk_select(score: torch.Tensor, seq_len: int, num_coming: int, start: int, important: int, recent: int,
                keep_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """score [rows, >=seq_len] fp16/fp32 -> int32 [rows, keep_total] ascending kept token indices."""
    lib = _lib.load()
    _require_cuda(score, keep_idx)
    plan = prune_plan(seq_len, num_coming, start, important, recent)
    if plan is None:
        raise ValueError("topk_select: cache already fits, nothing to select")
    keep_total = plan[0]
    if score.stride(-1) != 1:
        score = score.contiguous()
    rows = score.shape[0]
    if keep_idx is None:
        keep_idx = torch.empty((rows, keep_total), dtype=torch.int32, device=score.device)
    a = _lib.SelectArgs()
    a.struct_size = C.sizeof(_lib.SelectArgs)
    a.num_rows, a.seq_len = rows, seq_len
    a.num_coming, a.start, a.important, a.recent = num_coming, start, important, recent
    a.score_dtype = _dtype_code(score)
    a.score, a.score_stride = score.data_ptr(), score.stride(0)
    a.keep_idx, a.keep_stride = keep_idx.data_ptr(), keep_idx.stride(0)
    with torch.cuda.device(score.device):
        _lib.check(lib.spatten_topk_select(C.byref(a), C.c_void_p(_stream_ptr(score.device))))
    return keep_idx


def kv_gather(k: torch.Tensor, v: torch.Tensor, keep_idx: torch.Tensor, k_out: Optional[torch.Tensor] = None,
              v_out: Optional[torch.Tensor] = None, capacity: Optional[int] = None):
    """K/V [b,H,S,D] (rows contiguous) -> [b,H,keep_total,D] gathered by keep_idx [H,keep_total].

    k_out/v_out None: fresh tensors are returned, allocated with `capacity` rows per head (>= keep_total)
    and sliced to keep_total, so later decode steps can append in place.  k_out is k: in-place compaction.
    """
    lib = _lib.load()
    _require_cuda(k, v, keep_idx, k_out, v_out)
    b, H, S, D = k.shape
    keep_total = keep_idx.shape[1]
    if keep_idx.dtype != torch.int32 or keep_idx.shape[0] != H:
        raise ValueError("keep_idx must be int32 [H, keep_total]")
    es = k.element_size()
    for t in (k, v):
        if t.stride(3) != 1 or t.stride(2) != D:
            raise ValueError("kv_gather: K/V rows must be contiguous")
    if k_out is None:
        cap = max(keep_total, capacity or keep_total)
        k_out = torch.empty((b, H, cap, D), dtype=k.dtype, device=k.device)[:, :, :keep_total]
        v_out = torch.empty((b, H, cap, D), dtype=v.dtype, device=v.device)[:, :, :keep_total]
    a = _lib.GatherArgs()
    a.struct_size = C.sizeof(_lib.GatherArgs)
    a.batch, a.num_heads, a.keep_total, a.row_bytes = b, H, keep_total, D * es
    a.src_k, a.src_v = k.data_ptr(), v.data_ptr()
    a.src_stride_b, a.src_stride_h = k.stride(0) * es, k.stride(1) * es
    if v.stride() != k.stride() or v_out.stride() != k_out.stride():
        raise ValueError("kv_gather: K and V must share strides")
    a.dst_k, a.dst_v = k_out.data_ptr(), v_out.data_ptr()
    a.dst_stride_b, a.dst_stride_h = k_out.stride(0) * es, k_out.stride(1) * es
    a.keep_idx, a.keep_stride = keep_idx.data_ptr(), keep_idx.stride(0)
    with torch.cuda.device(k.device):
        _lib.check(lib.spatten_kv_gather(C.byref(a), C.c_void_p(_stream_ptr(k.device))))
    return k_out, v_out


def prune_layers(past_key_values, scores, seq_len: int, num_coming: int, start: int, important: int, recent: int,
                 capacity: int):
    """All layers of a cache in two launches -- `spatten_prune_layers`.

    past_key_values: list of (K, V) [b,H,S,D] with identical shapes, dtypes and strides; scores: list of [H, >=S] rows of
    one dtype and row stride.  Returns (list of (K', V') capacity buffers [b,H,capacity,D] whose first keep_total rows are
    the kept tokens, keep_idx int32 [L,H,keep_total]) or None when the layers are not uniform (caller falls back to the
    per-layer calls).
    """
    lib = _lib.load()
    L = len(past_key_values)
    k0, v0 = past_key_values[0]
    b, H, S, D = k0.shape
    plan = prune_plan(seq_len, num_coming, start, important, recent)
    keep_total = plan[0]
    s0 = scores[0]
    uniform = all(k.shape == k0.shape and v.shape == k0.shape and k.stride() == k0.stride() and v.stride() == k0.stride()
                  and k.dtype == k0.dtype and v.dtype == k0.dtype and k.is_cuda and k.stride(3) == 1 and k.stride(2) == D
                  for k, v in past_key_values) and \
        all(s.dim() == 2 and s.shape[0] == H and s.dtype == s0.dtype and s.stride() == s0.stride() and s.stride(1) == 1
            for s in scores)
    es = k0.element_size()
    if not uniform or (D * es) % 16 or D * es > 512:
        return None
    dev = k0.device
    big = torch.empty((L, 2, b, H, capacity, D), dtype=k0.dtype, device=dev)
    keep_idx = torch.empty((L, H, keep_total), dtype=torch.int32, device=dev)
    dst0, dstep = big.data_ptr(), big.stride(0) * es
    table_dev = torch.tensor([[s.data_ptr(), k.data_ptr(), v.data_ptr(), dst0 + l * dstep, dst0 + l * dstep + big.stride(1) * es]
                              for l, ((k, v), s) in enumerate(zip(past_key_values, scores))], dtype=torch.int64).to(dev)
    a = _lib.PruneLayersArgs()
    a.struct_size = C.sizeof(_lib.PruneLayersArgs)
    a.num_layers, a.batch, a.num_heads, a.seq_len = L, b, H, seq_len
    a.num_coming, a.start, a.important, a.recent = num_coming, start, important, recent
    a.score_dtype, a.row_bytes = _dtype_code(s0), D * es
    a.layers, a.score_stride = table_dev.data_ptr(), s0.stride(0)
    a.src_stride_b, a.src_stride_h = k0.stride(0) * es, k0.stride(1) * es
    a.dst_stride_b, a.dst_stride_h = big.stride(2) * es, big.stride(3) * es
    a.keep_idx, a.keep_stride = keep_idx.data_ptr(), keep_idx.stride(1)
    with torch.cuda.device(dev):
        _lib.check(lib.spatten_prune_layers(C.byref(a), C.c_void_p(_stream_ptr(dev))))
    table_dev.record_stream(torch.cuda.current_stream(dev))
    return [(big[l, 0], big[l, 1]) for l in range(L)], keep_idx


# ------------------------------------------------------------------------------------------------------
# RTL-ordered stages (parity unpinned): local V pruning and head pruning
# ------------------------------------------------------------------------------------------------------

def sparse_pv(logits: torch.Tensor, keep_idx: torch.Tensor, v_cache: torch.Tensor, num_heads: int) -> torch.Tensor:
    """out[b,h,:] = softmax(logits[b,h,keep]) @ V[b,h,keep,:]  -- `spatten_sparse_pv`.

    logits fp16 [b*H, >=n] (or [b,H,>=n]); keep_idx int32 [b*H, keep] ascending; v_cache fp16 [b,H_kv,cap,D].
    """
    lib = _lib.load()
    _require_cuda(logits, keep_idx, v_cache)
    b, H_kv, _, D = v_cache.shape
    lg = logits.reshape(b * num_heads, logits.shape[-1]) if logits.dim() == 3 else logits
    if lg.stride(-1) != 1 or keep_idx.dtype != torch.int32 or keep_idx.shape[0] != b * num_heads:
        raise ValueError("sparse_pv: logits rows must be contiguous; keep_idx must be int32 [b*H, keep]")
    out = torch.empty((b, num_heads, D), dtype=torch.float16, device=v_cache.device)
    a = _lib.SparsePvArgs()
    a.struct_size = C.sizeof(_lib.SparsePvArgs)
    a.batch, a.num_heads, a.num_kv_heads, a.head_dim, a.keep = b, num_heads, H_kv, D, keep_idx.shape[1]
    a.logits, a.logits_stride = lg.data_ptr(), lg.stride(0)
    a.keep_idx, a.keep_stride = keep_idx.data_ptr(), keep_idx.stride(0)
    a.v_cache, a.cache_stride_b, a.cache_stride_h = v_cache.data_ptr(), v_cache.stride(0), v_cache.stride(1)
    a.out = out.data_ptr()
    with torch.cuda.device(v_cache.device):
        _lib.check(lib.spatten_sparse_pv(C.byref(a), C.c_void_p(_stream_ptr(v_cache.device))))
    return out


def local_v_attn(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, kv_len: int, rope: torch.Tensor,
                 keep_v: int, k_new: Optional[torch.Tensor] = None, v_new: Optional[torch.Tensor] = None,
                 q_pos: Optional[torch.Tensor] = None, fused: Optional[bool] = None):
    """RTL stage order (SpAttenController.scala:347-749): DotProduct over ALL cached keys -> TopK(keep_v) ->
    MultiplyValue over the survivors only.  Returns (out [b,H,D], logits [b,H,n], keep_idx int32 [b*H, keep_v]
    ascending).  HBM traffic: N keys + keep_v values per head.

    One launch (`spatten_local_v_attn`, a thread-block cluster per row) when the new token is already in the cache
    (k_new is None) and a row slice fits on chip; otherwise -- or with fused=False -- the three-launch path
    (scores-only decode kernel -> `spatten_topk_select` -> `spatten_sparse_pv`).
    """
    b, H, D = q.shape
    n = int(kv_len) + (1 if k_new is not None else 0)
    logits = torch.empty((b, H, n), dtype=torch.float16, device=q.device)
    if fused is None:
        fused = k_new is None and n <= 60000 and keep_v >= 1
    if fused:
        if k_new is not None:
            raise ValueError("local_v_attn(fused=True): append the new K/V row to the cache first")
        lib = _lib.load()
        _require_cuda(q, k_cache, v_cache, rope, q_pos)
        if q.dtype != torch.float16 or k_cache.dtype != torch.float16 or v_cache.dtype != torch.float16:
            raise TypeError("spatten_b200.local_v_attn: fp16 tensors required")
        if k_cache.stride(3) != 1 or k_cache.stride(2) != D or v_cache.stride() != k_cache.stride():
            raise ValueError("spatten_b200.local_v_attn: cache rows must be contiguous with equal K/V strides")
        q = q.contiguous()
        keep_n = min(int(keep_v), n)
        keep = torch.empty((b * H, keep_n), dtype=torch.int32, device=q.device)
        out = torch.empty((b, H, D), dtype=torch.float16, device=q.device)
        a = _lib.LocalVArgs()
        a.struct_size = C.sizeof(_lib.LocalVArgs)
        a.batch, a.num_heads, a.num_kv_heads, a.head_dim = b, H, k_cache.shape[1], D
        a.kv_len, a.rope_len, a.keep = n, rope.shape[0], keep_n
        a.q, a.k_cache, a.v_cache = q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr()
        a.cache_stride_b, a.cache_stride_h = k_cache.stride(0), k_cache.stride(1)
        a.rope = rope.data_ptr()
        if q_pos is not None:
            if q_pos.dtype != torch.int64 or q_pos.numel() != b:
                raise TypeError("q_pos must be int64 [b]")
            q_pos = q_pos.contiguous()
            a.q_pos = q_pos.data_ptr()
        a.out, a.logits, a.logits_stride = out.data_ptr(), logits.data_ptr(), logits.stride(1)
        a.keep_idx, a.keep_stride = keep.data_ptr(), keep.stride(0)
        with torch.cuda.device(q.device):
            _lib.check(lib.spatten_local_v_attn(C.byref(a), C.c_void_p(_stream_ptr(q.device))))
        return out, logits, keep
    a, _, _keep = make_decode_args(q, k_cache, v_cache, kv_len, rope, k_new=k_new, v_new=v_new, q_pos=q_pos,
                                   logits=logits, flags=_lib.DECODE_SCORES_ONLY)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().spatten_decode_attn(C.byref(a), C.c_void_p(_stream_ptr(q.device))))
    rows = logits.view(b * H, n)
    if keep_v >= n:
        keep = torch.arange(n, dtype=torch.int32, device=q.device).expand(b * H, n).contiguous()
    else:
        keep = topk_select(rows, n, 0, 0, keep_v, 0)
    return sparse_pv(rows, keep, v_cache, H), logits, keep


def head_select(head_importance: torch.Tensor, keep_heads: int, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """u8 [H] mask of the `keep_heads` most important heads (ties -> lowest index) -- `spatten_head_select`."""
    lib = _lib.load()
    _require_cuda(head_importance, mask)
    if head_importance.dtype != torch.float32 or not head_importance.is_contiguous():
        raise TypeError("head_importance must be contiguous fp32 [H]")
    H = head_importance.numel()
    if mask is None:
        mask = torch.empty(H, dtype=torch.uint8, device=head_importance.device)
    with torch.cuda.device(mask.device):
        _lib.check(lib.spatten_head_select(head_importance.data_ptr(), H, int(keep_heads), mask.data_ptr(),
                                           C.c_void_p(_stream_ptr(mask.device))))
    return mask
