"""Cascade head pruning (README.md:21).  Not part of the reference's Python path -- the
repository only states the technique and its workload traces show head counts shrinking (16 -> 13 -> 10) -- so the
semantics are the restatement in oracle/spatten_oracle.py (head_importance = sum |attention output|, keep the
top-k heads, ties -> lowest index), labelled parity-unpinned.

Usage: `pruner = SpAttenHeadPruner(model, keep_heads)` after `enable_spatten_llm`; every decode step then
accumulates per-layer head importance inside the attention kernel; `pruner.update_masks()` (e.g. once per chat
round, next to apply_token_pruning) recomputes the masks on the device.  Masked heads read no KV at all and
output zeros, so their o_proj contribution vanishes.
"""
from __future__ import annotations

import torch

from . import ops
from .pos_shift.modify_llama import _is_llama_attention


class SpAttenHeadPruner:
    def __init__(self, model, keep_heads: int):
        self.keep_heads = keep_heads
        self.modules = [m for m in model.modules() if _is_llama_attention(m)]
        for m in self.modules:
            dev = m.q_proj.weight.device
            if keep_heads > m.num_heads:
                raise ValueError(f"keep_heads {keep_heads} > num_heads {m.num_heads}")
            m.spatten_head_importance = torch.zeros(m.num_heads, dtype=torch.float32, device=dev)
            m.spatten_head_mask = torch.ones(m.num_heads, dtype=torch.uint8, device=dev)

    def update_masks(self):
        """Recompute every layer's head mask from its cumulative importance (no host sync)."""
        for m in self.modules:
            ops.head_select(m.spatten_head_importance, self.keep_heads, m.spatten_head_mask)

    def reset(self):
        for m in self.modules:
            m.spatten_head_importance.zero_()
            m.spatten_head_mask.fill_(1)
