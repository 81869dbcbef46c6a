"""`SpAttenKVCache` -- same constructor, attributes and `apply_token_pruning` contract as the reference
(spatten_llm/kv_cache_token_pruning.py:24-96), executed by the CUDA library with zero host syncs.

Differences that are extensions, not behaviour changes:
  * batch > 1 is accepted (the reference fails with IndexError at :68); the per-head index set, which the
    reference already derives from the batch-summed score (:51), is applied to every sequence.
  * ties at the k-th score are resolved lowest-index-first (torch.topk leaves them unspecified).
  * returned tensors carry spare capacity behind their visible length so that the following decode steps
    append in place instead of reallocating (the reference re-cats the whole cache every step).
"""
from __future__ import annotations

import torch

from . import ops
from .cache_buffer import CacheBuffer


class SpAttenKVCache:
    def __init__(self, start_size=4, recent_size=128, important_size=128, k_seq_dim=2, v_seq_dim=2):
        print(f"SpAttenKVCache: keep start: {start_size}, keep recent: {recent_size}, "
              f"keep important: {important_size}")                                   # banner as :32
        if k_seq_dim != 2 or v_seq_dim != 2:
            raise ValueError("spatten_b200 caches are [batch, heads, seq, head_dim]: k_seq_dim == v_seq_dim == 2 "
                             "(what enable_spatten_llm passes for llama, enable_spatten_llm.py:7)")
        self.start_size, self.recent_size, self.important_size = start_size, recent_size, important_size
        self.cache_size = start_size + important_size + recent_size                  # :36
        self.k_seq_dim, self.v_seq_dim = k_seq_dim, v_seq_dim
        self.importance_score = None
        self.keep_indices = None        # per layer int32 [H, keep_total] of the last prune (device)

    def apply_token_pruning(self, past_key_values, num_coming, attn_score_all):
        """past_key_values: list of (K, V) [b,H,S,D] CUDA tensors; attn_score_all: list of [b,H,q,S].

        Returns the list unchanged when seq_len + num_coming fits (:46-47), None for None (:43-44),
        else a list of [K', V'] with S' = start + important + max(0, recent - num_coming) (:72-96).
        """
        if past_key_values is None:
            return None
        seq_len = past_key_values[0][0].size(self.k_seq_dim)
        plan = ops.prune_plan(seq_len, int(num_coming), self.start_size, self.important_size, self.recent_size)
        if plan is None:
            return past_key_values
        keep_total = plan[0]
        # :51 -- both sums run over singleton dims after a decode step (b = q = 1): an exact identity, no launch
        self.importance_score = [s[0, :, 0, :] if (s.shape[0] == 1 and s.shape[2] == 1 and s.stride(3) == 1 and s.is_cuda)
                                 else ops.score_reduce(s) for s in attn_score_all]
        for (k, _), score in zip(past_key_values, self.importance_score):
            if score.shape[0] != k.shape[1]:
                raise ValueError(f"score heads {score.shape[0]} != cache heads {k.shape[1]} "
                                 "(the reference's mask gather needs num_heads == num_key_value_heads)")
        # room for the tokens the caller announced plus some slack, so decode appends in place
        capacity = keep_total + int(num_coming) + 64
        # all layers in two launches when they share shapes and strides (any HF cache does) ...
        batched = ops.prune_layers(past_key_values, self.importance_score, seq_len, int(num_coming), self.start_size,
                                   self.important_size, self.recent_size, capacity)
        if batched is not None:
            bufs, keep_idx = batched
            self.keep_indices = list(keep_idx.unbind(0))
            # fresh buffers of ours: the patched forward may append behind these views in place
            return [[CacheBuffer(kb, keep_total).view(keep_total), CacheBuffer(vb, keep_total).view(keep_total)]
                    for kb, vb in bufs]
        # ... else layer by layer
        self.keep_indices = []
        out = []
        for (k, v), score in zip(past_key_values, self.importance_score):
            keep = ops.topk_select(score, seq_len, int(num_coming), self.start_size, self.important_size,
                                   self.recent_size)
            self.keep_indices.append(keep)
            k2, v2 = ops.kv_gather(k, v, keep, capacity=capacity)
            out.append([CacheBuffer(k2._base, keep_total).view(keep_total), CacheBuffer(v2._base, keep_total).view(keep_total)])
        return out
