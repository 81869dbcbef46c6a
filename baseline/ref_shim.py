"""Bind the UNMODIFIED reference functions (spatten_llm/pos_shift/modify_llama.py, kv_cache_token_pruning.py) to a
shim `self` that supplies the transformers-4.33 `LlamaAttention` attribute contract they read.

The reference sources are imported from `baseline/_ref/` (a copy of /root/reference/spatten_llm made by
`baseline/install_ref.sh`; git-ignored, but shipped to the GPU box) or, in the build container, straight from
/root/reference.  Nothing here is part of the product: it is used by tests/golden/make_golden.py (golden vectors) and by
bench.py's reference arm / `eager_gpu` baseline.
"""
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    """Returns (llama_pos_shift_attention_forward, SpAttenKVCache, where) or raises ImportError."""
    for root in (os.path.join(_HERE, "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(root, "spatten_llm")):
            if root not in sys.path:
                sys.path.insert(0, root)
            from spatten_llm.kv_cache_token_pruning import SpAttenKVCache
            from spatten_llm.pos_shift.modify_llama import llama_pos_shift_attention_forward
            return llama_pos_shift_attention_forward, SpAttenKVCache, root
    raise ImportError("reference spatten_llm not found under baseline/_ref or /root/reference")


class Rotary433(torch.nn.Module):
    """transformers 4.33 LlamaRotaryEmbedding (restated; its source is not on disk)."""

    def __init__(self, dim, max_position_embeddings=2048, base=10000):
        super().__init__()
        self.dim, self.base = dim, base
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)
        self._set(max_position_embeddings)

    def _set(self, seq_len):
        self.max_seq_len_cached = seq_len
        t = torch.arange(seq_len, dtype=self.inv_freq.dtype, device=self.inv_freq.device)
        freqs = torch.einsum("i,j->ij", t, self.inv_freq)
        emb = torch.cat((freqs, freqs), dim=-1)
        self.cos_cached = emb.cos()[None, None, :, :].to(torch.float32)
        self.sin_cached = emb.sin()[None, None, :, :].to(torch.float32)

    def forward(self, x, seq_len=None):
        if seq_len > self.max_seq_len_cached:
            self._set(seq_len)
        return (self.cos_cached[:, :, :seq_len, ...].to(dtype=x.dtype),
                self.sin_cached[:, :, :seq_len, ...].to(dtype=x.dtype))


class Const(torch.nn.Module):
    """Stands in for a projection when the case is specified after the projections."""

    def __init__(self, value):
        super().__init__()
        self.value = value

    def forward(self, _):
        return self.value


class Linear(torch.nn.Module):
    def __init__(self, w):
        super().__init__()
        self.weight = w

    def forward(self, x):
        return torch.nn.functional.linear(x, self.weight)


class ShimAttention(torch.nn.Module):
    """`forward` is bound to the reference function by `make_shim_class`."""

    def __init__(self, H, H_kv, D):
        super().__init__()
        self.config = types.SimpleNamespace(pretraining_tp=1)
        self.hidden_size, self.num_heads, self.num_key_value_heads = H * D, H, H_kv
        self.head_dim, self.num_key_value_groups = D, H // H_kv
        self.rotary_emb = Rotary433(D)




def make_shim_class():
    """ShimAttention subclass whose forward IS the reference's llama_pos_shift_attention_forward, unmodified."""
    fwd, _, _ = import_reference()
    return type("RefAttention", (ShimAttention,), {"forward": fwd})


def core_module(H, H_kv, D, q, k_new, v_new, device="cpu"):
    """A RefAttention whose projections return the given post-projection tensors [b, heads, q_len, D] (the attention
    core in isolation: what bench.py times and what the core-level goldens were generated with)."""
    m = make_shim_class()(H, H_kv, D).to(device)
    m.rotary_emb._set(m.rotary_emb.max_seq_len_cached)       # rebuild the cos/sin cache on `device`
    b, _, q_len, _ = q.shape
    flat = lambda x: x.transpose(1, 2).reshape(b, q_len, -1)
    m.q_proj, m.k_proj, m.v_proj = Const(flat(q)), Const(flat(k_new)), Const(flat(v_new))
    m.o_proj = torch.nn.Identity()
    return m
