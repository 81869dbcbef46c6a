"""Progressively-quantised KV cache (BASELINE config 4): int4 MSB + int4 LSB planes and fp16 per-token scales.

Host-side container over `spatten_kv_quantize` / `spatten_decode_attn_quant`.  The technique is named by the
reference (README.md:11) but only modelled in its RTL, so the arithmetic is the restatement in the test oracle
(parity unpinned).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .ops import _require_cuda, _stream_ptr


class QuantKVCache:
    """Planes for one layer: K/V MSB + LSB u8 [b, H_kv, cap, D/2] and fp16 scales [b, H_kv, cap]; D == 128."""

    def __init__(self, batch: int, num_kv_heads: int, capacity: int, head_dim: int = 128, device="cuda"):
        if head_dim != 128:
            raise ValueError("QuantKVCache: head_dim must be 128")
        cap = (capacity + 7) // 8 * 8
        self.batch, self.num_kv_heads, self.capacity, self.head_dim, self.length = batch, num_kv_heads, cap, head_dim, 0
        mk = lambda: torch.zeros((batch, num_kv_heads, cap, head_dim // 2), dtype=torch.uint8, device=device)
        sc = lambda: torch.zeros((batch, num_kv_heads, cap), dtype=torch.float16, device=device)
        self.k_msb, self.k_lsb, self.v_msb, self.v_lsb = mk(), mk(), mk(), mk()
        self.k_scale, self.v_scale = sc(), sc()
        self._ws = None

    def _struct(self) -> _lib.QuantKV:
        s = _lib.QuantKV()
        s.k_msb, s.k_lsb, s.k_scale = self.k_msb.data_ptr(), self.k_lsb.data_ptr(), self.k_scale.data_ptr()
        s.v_msb, s.v_lsb, s.v_scale = self.v_msb.data_ptr(), self.v_lsb.data_ptr(), self.v_scale.data_ptr()
        s.plane_stride_b, s.plane_stride_h = self.k_msb.stride(0), self.k_msb.stride(1)
        s.scale_stride_b, s.scale_stride_h = self.k_scale.stride(0), self.k_scale.stride(1)
        return s

    def append(self, k: torch.Tensor, v: torch.Tensor) -> None:
        """Quantise fp16 k, v [b, H_kv, n, D] (any strides, rows contiguous) into slots [length, length+n)."""
        lib = _lib.load()
        _require_cuda(k, v)
        b, h, n, d = k.shape
        if (b, h, d) != (self.batch, self.num_kv_heads, self.head_dim) or v.shape != k.shape:
            raise ValueError("append: shape mismatch")
        if self.length + n > self.capacity:
            raise ValueError("append: capacity exceeded")
        if k.dtype != torch.float16 or k.stride(3) != 1 or v.stride() != k.stride():
            k, v = k.contiguous(), v.contiguous()
        st = self._struct()
        with torch.cuda.device(k.device):
            _lib.check(lib.spatten_kv_quantize(k.data_ptr(), v.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
                                               b, h, n, d, C.byref(st), self.length,
                                               C.c_void_p(_stream_ptr(k.device))))
        self.length += n

    def decode(self, q: torch.Tensor, rope: torch.Tensor, threshold: float = 0.0, v_bits: int = 4,
               q_pos: Optional[torch.Tensor] = None, logits: Optional[torch.Tensor] = None):
        """q fp16 [b, H, D] (its K/V already appended).  Returns (out [b,H,D], need_requant u8 [b*H])."""
        lib = _lib.load()
        _require_cuda(q, rope, q_pos, logits)
        b, H, D = q.shape
        out = torch.empty_like(q)
        need = torch.empty(b * H, dtype=torch.uint8, device=q.device)      # pass 1 writes every row's decision
        nbytes = lib.spatten_decode_quant_workspace_bytes(b, H, D)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.zeros(nbytes, dtype=torch.uint8, device=q.device)      # tickets start at zero; self-resetting
        a = _lib.QuantDecodeArgs()
        a.struct_size = C.sizeof(_lib.QuantDecodeArgs)
        a.batch, a.num_heads, a.num_kv_heads, a.head_dim = b, H, self.num_kv_heads, D
        a.kv_len, a.rope_len, a.v_bits, a.requant_threshold = self.length, rope.shape[0], v_bits, float(threshold)
        a.kv = self._struct()
        q = q.contiguous()
        a.q, a.rope, a.out = q.data_ptr(), rope.data_ptr(), out.data_ptr()
        if q_pos is not None:
            q_pos = q_pos.contiguous()
            a.q_pos = q_pos.data_ptr()
        if logits is not None:
            lg = logits.view(b * H, logits.shape[-1])
            a.logits, a.logits_stride = lg.data_ptr(), lg.stride(0)
        a.need_requant = need.data_ptr()
        a.workspace, a.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        with torch.cuda.device(q.device):
            _lib.check(lib.spatten_decode_attn_quant(C.byref(a), C.c_void_p(_stream_ptr(q.device))))
        return out, need
