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


def decode_attn(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, kv_len: int, rope: torch.Tensor,
                k_new: Optional[torch.Tensor] = None, v_new: Optional[torch.Tensor] = None,
                q_pos: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                head_mask: Optional[torch.Tensor] = None, logits: Optional[torch.Tensor] = None,
                importance: Optional[torch.Tensor] = None, head_importance: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, flags: int = 0) -> torch.Tensor:
    """One decode step of attention for every (batch, head) row -- `spatten_decode_attn`.

    q [b,H,D] fp16 un-rotated; k_cache/v_cache [b,H_kv,cap,D] fp16 (any batch/head strides, rows
    contiguous); kv_len tokens valid; k_new/v_new [b,H_kv,D] appended at slot kv_len when given.
    Returns out [b,H,D] fp16.  `logits` [b,H,>=n] receives attn_scores; `importance` fp32 [b,H,>=n]
    and `head_importance` fp32 [H] are accumulated in place.
    """
    lib = _lib.load()
    _require_cuda(q, k_cache, v_cache, rope, k_new, v_new, q_pos, mask, head_mask, logits, importance,
                  head_importance, out)
    b, H, D = q.shape
    H_kv = k_cache.shape[1]
    if q.dtype != torch.float16 or k_cache.dtype != torch.float16 or v_cache.dtype != torch.float16:
        raise TypeError("spatten_b200.decode_attn: fp16 tensors required")
    if k_cache.stride(3) != 1 or k_cache.stride(2) != D or v_cache.stride() != k_cache.stride():
        raise ValueError("spatten_b200.decode_attn: cache rows must be contiguous [.., S, D] with equal K/V strides")
    if not q.is_contiguous():
        q = q.contiguous()
    append = k_new is not None
    if append:
        k_new, v_new = k_new.contiguous(), v_new.contiguous()
    if out is None:
        out = torch.empty((b, H, D), dtype=torch.float16, device=q.device)
    a = _lib.DecodeArgs()
    a.struct_size = C.sizeof(_lib.DecodeArgs)
    a.batch, a.num_heads, a.num_kv_heads, a.head_dim = b, H, H_kv, D
    a.kv_len, a.rope_len, a.append, a.flags = int(kv_len), rope.shape[0], int(append), int(flags)
    a.q, a.k_cache, a.v_cache = q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr()
    a.k_new = k_new.data_ptr() if append else None
    a.v_new = v_new.data_ptr() if append else None
    a.cache_stride_b, a.cache_stride_h = k_cache.stride(0), k_cache.stride(1)
    a.rope = rope.data_ptr()
    if q_pos is not None:
        if q_pos.dtype != torch.int64:
            raise TypeError("q_pos must be int64 (HF position_ids)")
        a.q_pos = q_pos.contiguous().data_ptr()
    if mask is not None:
        if mask.dtype != torch.float16 or mask.shape != (b, kv_len + int(append)) or not mask.is_contiguous():
            raise ValueError("mask must be contiguous fp16 [b, kv_len+append]")
        a.mask = mask.data_ptr()
    if head_mask is not None:
        a.head_mask = head_mask.data_ptr()
    a.out = out.data_ptr()
    if logits is not None:
        a.logits, a.logits_stride = logits.data_ptr(), logits.stride(1)
    if importance is not None:
        a.importance, a.importance_stride = importance.data_ptr(), importance.stride(1)
    if head_importance is not None:
        a.head_importance = head_importance.data_ptr()
    ws = decode_workspace(q.device, b, H, D)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    with torch.cuda.device(q.device):
        _lib.check(lib.spatten_decode_attn(C.byref(a), C.c_void_p(_stream_ptr(q.device))))
    return out
