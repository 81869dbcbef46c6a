"""Patched `LlamaAttention.forward` -- the B200 drop-in for the reference's
`llama_pos_shift_attention_forward` (spatten_llm/pos_shift/modify_llama.py:31-168).

Same signature, same return triple, same `self.attn_scores` side channel (:116-119), same shape
ValueErrors (:121-131).  The projections stay `nn.Linear` (cuBLAS, as in the reference); everything
between them -- RoPE(Q), KV append, RoPE(K at slot index), Q.K^T/sqrt(D), logit stash, mask, softmax,
P.V -- is ONE launch of the CUDA library (`spatten_decode_attn`).  There is no eager fallback: a
non-CUDA or non-fp16 input raises.

The module is expected to expose the transformers-4.33 `LlamaAttention` attribute contract the
reference reads: hidden_size, num_heads, num_key_value_heads, head_dim, config.pretraining_tp,
q_proj/k_proj/v_proj/o_proj and (optionally) rotary_emb.base.
"""
from __future__ import annotations

import types
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from .. import _lib, ops
from ..cache_buffer import CacheBuffer

__all__ = ["enable_llama_pos_shift_attention", "llama_pos_shift_attention_forward"]

_GROW = 256           # minimum spare rows when a cache buffer has to be (re)allocated


def _grow(past: Optional[torch.Tensor], b, H_kv, D, have, need, device) -> CacheBuffer:
    """CacheBuffer holding the first `have` rows of `past` with capacity >= need (amortised doubling)."""
    cb = CacheBuffer.of(past)
    if cb is not None and cb.buf.shape[2] >= need and cb.buf.shape[:2] == (b, H_kv) and cb.buf.device == device:
        return cb
    cap = max(need + _GROW, 2 * have)
    buf = torch.empty((b, H_kv, cap, D), dtype=torch.float16, device=device)
    if past is not None and have:
        buf[:, :, :have].copy_(past)
    return CacheBuffer(buf, have)


def _token_importance(self, bsz: int, H: int, need: int):
    """Optional cumulative token importance (README.md:21): a module attribute `spatten_token_importance`, fp32
    [b, H, capacity], set by the caller; every forward adds its softmax probabilities (decode: one row; prefill: the
    column sums over the query rows).  None = not tracked (the reference Python path prunes on last-step logits only)."""
    imp = getattr(self, "spatten_token_importance", None)
    if imp is None:
        return None
    if imp.dtype != torch.float32 or imp.dim() != 3 or imp.shape[0] != bsz or imp.shape[1] != H or imp.shape[2] < need:
        raise ValueError(f"spatten_token_importance must be fp32 [{bsz}, {H}, >= {need}], got {tuple(imp.shape)} {imp.dtype}")
    return imp


def _rope_for(self, device, need: int) -> torch.Tensor:
    tab = getattr(self, "_spatten_rope", None)
    if tab is None or tab.shape[0] < need or tab.device != device:
        base = float(getattr(getattr(self, "rotary_emb", None), "base", 10000.0))
        n = max(need, 2048 if tab is None else 2 * tab.shape[0])
        tab = ops.rope_table(self.head_dim, n, device, base)
        self._spatten_rope = tab
    return tab


def llama_pos_shift_attention_forward(
    self,
    hidden_states: torch.Tensor,
    attention_mask: Optional[torch.Tensor] = None,
    position_ids: Optional[torch.LongTensor] = None,
    past_key_value: Optional[Tuple[torch.Tensor]] = None,
    output_attentions: bool = False,
    use_cache: bool = False,
    padding_mask: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[Tuple[torch.Tensor]]]:
    if not hidden_states.is_cuda:
        raise RuntimeError("spatten_b200: the patched forward runs on CUDA only (no CPU fallback)")
    if hidden_states.dtype != torch.float16:
        raise TypeError("spatten_b200: the CUDA path computes in fp16 (the reference GPU dtype, utils.py:62)")
    _lib.load()
    bsz, q_len, _ = hidden_states.size()
    H, H_kv, D = self.num_heads, self.num_key_value_heads, self.head_dim
    dev = hidden_states.device

    # projections (:43-74).  pretraining_tp > 1 only re-associates the same matmuls (:43-69); the
    # fused nn.Linear call is numerically the tp == 1 branch.
    q = self.q_proj(hidden_states).view(bsz, q_len, H, D)
    k_new = self.k_proj(hidden_states).view(bsz, q_len, H_kv, D)
    v_new = self.v_proj(hidden_states).view(bsz, q_len, H_kv, D)

    past_len = past_key_value[0].shape[-2] if past_key_value is not None else 0
    kv_seq_len = past_len + q_len                                                    # :86-88
    if position_ids is None:
        position_ids = torch.arange(past_len, kv_seq_len, device=dev).unsqueeze(0)
    # transformers-4.33 LlamaModel builds default position_ids of shape [1, q_len]; the reference broadcasts them
    # through cos[position_ids] (:23-26)
    position_ids = position_ids.to(torch.int64).expand(bsz, q_len).contiguous()
    rope = _rope_for(self, dev, kv_seq_len)

    mask = None
    if attention_mask is not None:                                                   # :127-131
        if attention_mask.size() != (bsz, 1, q_len, kv_seq_len):
            raise ValueError(
                f"Attention mask should be of size {(bsz, 1, q_len, kv_seq_len)}, but is {attention_mask.size()}")
        mask = attention_mask.to(torch.float16).reshape(bsz, q_len, kv_seq_len).contiguous()

    past_k = past_key_value[0] if past_key_value is not None else None
    past_v = past_key_value[1] if past_key_value is not None else None
    kcb = _grow(past_k, bsz, H_kv, D, past_len, kv_seq_len, dev)
    vcb = _grow(past_v, bsz, H_kv, D, past_len, kv_seq_len, dev)
    if vcb.buf.stride() != kcb.buf.stride():               # keep K and V congruent for the kernel
        vb2 = torch.empty_strided(kcb.buf.shape, kcb.buf.stride(), dtype=torch.float16, device=dev)
        vb2[:, :, :past_len].copy_(vcb.buf[:, :, :past_len])
        vcb = CacheBuffer(vb2, past_len)
    kbuf, vbuf = kcb.buf, vcb.buf

    if q_len == 1:
        scores = torch.empty((bsz, H, 1, kv_seq_len), dtype=torch.float16, device=dev)
        out = ops.decode_attn(q.view(bsz, H, D), kbuf, vbuf, past_len, rope,
                              k_new=k_new.view(bsz, H_kv, D), v_new=v_new.view(bsz, H_kv, D),
                              q_pos=position_ids.reshape(bsz), mask=mask,
                              logits=scores.view(bsz, H, kv_seq_len),
                              head_mask=getattr(self, "spatten_head_mask", None),
                              head_importance=getattr(self, "spatten_head_importance", None),
                              importance=_token_importance(self, bsz, H, kv_seq_len))
    else:
        # prefill / chunk: store the new K/V rows (:95-100), then score every query row against the cache
        kbuf[:, :, past_len:kv_seq_len].copy_(k_new.transpose(1, 2))
        vbuf[:, :, past_len:kv_seq_len].copy_(v_new.transpose(1, 2))
        scores = torch.empty((bsz, H, q_len, kv_seq_len), dtype=torch.float16, device=dev)
        head_mask = getattr(self, "spatten_head_mask", None)
        if D == 128 and head_mask is None and q_len >= 16:
            # dense tiles: Q.K^T and P.V on the tcgen05 tensor cores (spatten_prefill_attn)
            out, _ = ops.prefill_attn(q, kbuf, vbuf, kv_seq_len, rope, position_ids.reshape(bsz, q_len), mask, scores,
                                      importance=_token_importance(self, bsz, H, kv_seq_len))
        else:
            # short chunks / other head sizes / pruned heads: one decode-kernel row per (query, head)
            out = ops.decode_attn(q, kbuf, vbuf, kv_seq_len, rope, q_pos=position_ids.reshape(bsz, q_len), mask=mask,
                                  logits=scores, head_mask=head_mask)
    self.attn_scores = scores                                                        # :116-119 (pre-mask logits)

    attn_output = self.o_proj(out.view(bsz, q_len, H * D))                           # :146-163
    attn_weights = None
    if output_attentions:                                                            # :135-137, :165-166
        s = scores if attention_mask is None else scores + attention_mask.to(torch.float16)
        attn_weights = F.softmax(s, dim=-1, dtype=torch.float32).to(torch.float16)
    # the rows are in the buffers either way; without use_cache the views are simply not handed out
    present = (kcb.view(kv_seq_len), vcb.view(kv_seq_len)) if use_cache else None        # :100
    return attn_output, attn_weights, present


def _is_llama_attention(module) -> bool:
    if getattr(module, "_spatten_llama_attention", False):
        return True
    try:
        from transformers.models.llama.modeling_llama import LlamaAttention
    except Exception:
        return False
    return isinstance(module, LlamaAttention)


def enable_llama_pos_shift_attention(model) -> int:
    """Rebind `.forward` of every LlamaAttention under `model` (reference :171-181).  Returns the count."""
    n = 0
    for module in model.modules():
        if _is_llama_attention(module):
            for attr in ("num_heads", "num_key_value_heads", "head_dim", "q_proj", "k_proj", "v_proj", "o_proj"):
                if not hasattr(module, attr):
                    raise AttributeError(
                        f"{type(module).__name__} lacks `{attr}`: the patched forward needs the transformers-4.33 "
                        "LlamaAttention contract the reference was written against (README.md:46)")
            module.forward = types.MethodType(llama_pos_shift_attention_forward, module)
            n += 1
    return n
