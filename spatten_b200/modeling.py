"""Llama-shaped attention modules with random weights (no checkpoints, no network).

`LlamaAttention433` carries exactly the attributes the reference forward reads from transformers-4.33's
LlamaAttention, so `enable_spatten_llm` can patch it like a real model.  Used by
tests, bench.py and smoke(); it is scaffolding around the hot path, not part of it.
"""
from __future__ import annotations

import math
import types

import torch
from torch import nn


class LlamaAttention433(nn.Module):
    _spatten_llama_attention = True

    def __init__(self, hidden_size: int, num_heads: int, num_key_value_heads: int | None = None,
                 rope_base: float = 10000.0, dtype=torch.float16, device=None, seed: int = 0,
                 head_slice: tuple[int, int] | None = None):
        super().__init__()
        self.config = types.SimpleNamespace(pretraining_tp=1, model_type="llama")
        self.hidden_size = hidden_size
        self.total_heads = num_heads
        self.head_dim = hidden_size // num_heads
        kv = num_key_value_heads or num_heads
        # head_slice = (first, count): tensor-parallel shard holding `count` heads (column-parallel q/k/v,
        # row-parallel o_proj); the caller all-reduces the output.
        h0, hc = head_slice if head_slice is not None else (0, num_heads)
        self.num_heads = hc
        self.num_key_value_heads = kv * hc // num_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.rotary_emb = types.SimpleNamespace(base=rope_base, dim=self.head_dim)
        D, g = self.head_dim, torch.Generator(device="cpu").manual_seed(seed)
        sc = 1.0 / math.sqrt(hidden_size)

        def lin(out_f, in_f, rows=None, cols=None):
            w = torch.randn((out_f, in_f), generator=g) * sc
            if rows is not None:
                w = w[rows[0]:rows[1]]
            if cols is not None:
                w = w[:, cols[0]:cols[1]]
            m = nn.Linear(w.shape[1], w.shape[0], bias=False, dtype=dtype, device=device)
            with torch.no_grad():
                m.weight.copy_(w.to(dtype))
            return m

        kv0, kvc = kv * h0 // num_heads, self.num_key_value_heads
        self.q_proj = lin(num_heads * D, hidden_size, rows=(h0 * D, (h0 + hc) * D))
        self.k_proj = lin(kv * D, hidden_size, rows=(kv0 * D, (kv0 + kvc) * D))
        self.v_proj = lin(kv * D, hidden_size, rows=(kv0 * D, (kv0 + kvc) * D))
        self.o_proj = lin(hidden_size, num_heads * D, cols=(h0 * D, (h0 + hc) * D))

    def forward(self, *a, **k):   # replaced by enable_llama_pos_shift_attention
        raise RuntimeError("LlamaAttention433.forward: call enable_spatten_llm(model, ...) first")


class AttentionStack(nn.Module):
    """L attention layers of one Llama shape: the object `enable_spatten_llm` is pointed at."""

    def __init__(self, num_layers: int, hidden_size: int, num_heads: int, num_key_value_heads=None,
                 dtype=torch.float16, device=None, seed: int = 0, head_slice=None):
        super().__init__()
        self.config = types.SimpleNamespace(model_type="llama", hidden_size=hidden_size,
                                            num_attention_heads=num_heads, num_hidden_layers=num_layers)
        self.layers = nn.ModuleList([
            LlamaAttention433(hidden_size, num_heads, num_key_value_heads, dtype=dtype, device=device,
                              seed=seed + 7919 * l, head_slice=head_slice) for l in range(num_layers)])
