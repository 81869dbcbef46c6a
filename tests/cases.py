"""Shared case definitions: the golden generator (tests/golden/make_golden.py, run once in the build
container against the unmodified reference) and the parity tests build identical inputs from here."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

from detgen import det_normal


@dataclass(frozen=True)
class FwdCase:
    name: str
    b: int
    H: int
    H_kv: int
    D: int
    past: int          # cached tokens before this call
    q_len: int         # new tokens in this call (1 = decode)
    dtype: str         # "fp16" | "fp32"
    seed: int
    full: bool = False  # True: hidden + q/k/v/o weights (whole patched forward); False: core (q,k,v given)
    mask: bool = True   # pass an HF-4.33-style additive mask


FWD_CASES = [
    # BASELINE.json configs[0]: GPT-2-small dims, one layer, seq 128
    FwdCase("c1_g2s_prefill_fp32", 1, 12, 12, 64, 0, 128, "fp32", 101, full=True),
    FwdCase("c1_g2s_decode_fp32", 1, 12, 12, 64, 128, 1, "fp32", 102, full=True),
    FwdCase("c1_g2s_decode_fp16", 1, 12, 12, 64, 128, 1, "fp16", 103, full=True),
    # Llama head_dim, ragged lengths, core level
    FwdCase("d128_decode_fp16", 1, 8, 8, 128, 515, 1, "fp16", 104),
    FwdCase("d128_decode_nomask_fp16", 1, 8, 8, 128, 64, 1, "fp16", 105, mask=False),
    FwdCase("d128_decode_b2_fp16", 2, 8, 8, 128, 300, 1, "fp16", 106),
    FwdCase("d128_decode_b3_long_fp16", 3, 4, 4, 128, 2049, 1, "fp16", 107),
    FwdCase("gqa_decode_fp16", 1, 8, 2, 128, 257, 1, "fp16", 108),
    FwdCase("d64_decode_fp16", 2, 12, 12, 64, 333, 1, "fp16", 109),
    FwdCase("d128_first_token_fp16", 1, 4, 4, 128, 0, 1, "fp16", 110),
    FwdCase("d128_prefill_fp16", 1, 4, 4, 128, 0, 96, "fp16", 111),
    FwdCase("d128_chunk_prefill_fp16", 1, 4, 4, 128, 64, 33, "fp16", 112),
    FwdCase("d64_prefill_b2_fp16", 2, 4, 4, 64, 0, 70, "fp16", 113),
]


def torch_dtype(s: str) -> torch.dtype:
    return {"fp16": torch.float16, "fp32": torch.float32}[s]


def causal_mask(b: int, q_len: int, past: int, dtype: torch.dtype) -> torch.Tensor:
    """HF 4.33 `_make_causal_mask` + past columns: [b,1,q,kv], 0 where visible, finfo.min elsewhere."""
    kv = past + q_len
    m = torch.full((q_len, q_len), torch.finfo(dtype).min, dtype=dtype)
    cond = torch.arange(q_len)
    m.masked_fill_(cond < (cond + 1).view(q_len, 1), 0)
    m = torch.cat([torch.zeros(q_len, past, dtype=dtype), m], dim=-1)
    return m[None, None].expand(b, 1, q_len, kv).contiguous()


def build_fwd_inputs(c: FwdCase) -> dict:
    dt = torch_dtype(c.dtype)
    hidden_size = c.H * c.D
    d = {}
    if c.full:
        d["hidden"] = det_normal((c.b, c.q_len, hidden_size), c.seed, 1.0, dt)
        sc = 1.0 / math.sqrt(hidden_size)
        d["wq"] = det_normal((c.H * c.D, hidden_size), c.seed + 1000, sc, dt)
        d["wk"] = det_normal((c.H_kv * c.D, hidden_size), c.seed + 2000, sc, dt)
        d["wv"] = det_normal((c.H_kv * c.D, hidden_size), c.seed + 3000, sc, dt)
        d["wo"] = det_normal((hidden_size, hidden_size), c.seed + 4000, sc, dt)
    else:
        d["q"] = det_normal((c.b, c.H, c.q_len, c.D), c.seed, 1.0, dt)
        d["k_new"] = det_normal((c.b, c.H_kv, c.q_len, c.D), c.seed + 1000, 1.0, dt)
        d["v_new"] = det_normal((c.b, c.H_kv, c.q_len, c.D), c.seed + 2000, 1.0, dt)
    if c.past > 0:
        d["past_k"] = det_normal((c.b, c.H_kv, c.past, c.D), c.seed + 5000, 1.0, dt)
        d["past_v"] = det_normal((c.b, c.H_kv, c.past, c.D), c.seed + 6000, 1.0, dt)
    else:
        d["past_k"] = d["past_v"] = None
    d["position_ids"] = torch.arange(c.past, c.past + c.q_len).unsqueeze(0).expand(c.b, c.q_len).contiguous()
    d["mask"] = causal_mask(c.b, c.q_len, c.past, dt) if c.mask else None
    return d


@dataclass(frozen=True)
class PruneCase:
    name: str
    L: int             # layers
    H: int
    D: int
    S: int             # cached tokens
    q_len: int         # query rows in the stashed attn_scores (1 after a decode step)
    start: int
    important: int
    recent: int
    num_coming: int
    dtype: str
    seed: int


PRUNE_CASES = [
    # configs[0]: seq 128 -> 64 (4 + 30 + 30)
    PruneCase("c1_g2s_128_to_64", 1, 12, 64, 128, 1, 4, 30, 30, 0, "fp16", 201),
    PruneCase("c1_g2s_128_to_64_fp32", 1, 12, 64, 128, 1, 4, 30, 30, 0, "fp32", 202),
    # the demo's defaults: start 0, important 150, recent 150, num_coming = prompt+64 (run_spatten_llama.py:72,134-136)
    PruneCase("demo_defaults", 2, 8, 128, 700, 1, 0, 150, 150, 90, "fp16", 203),
    # num_coming larger than recent: suffix disappears, candidate window runs to the end
    PruneCase("coming_gt_recent", 2, 4, 64, 300, 1, 3, 40, 20, 20, "fp16", 204),
    PruneCase("coming_shrinks_recent", 1, 4, 64, 257, 1, 2, 50, 40, 33, "fp16", 205),
    # scores stashed by a prefill (q_len > 1): summed over the query dim in the score dtype (:51)
    PruneCase("qsum_prefill_scores", 1, 4, 64, 200, 5, 4, 60, 60, 0, "fp32", 206),
    # 50 % prune at a Llama-ish head_dim, bigger window
    PruneCase("d128_2049_to_1024", 1, 8, 128, 2049, 1, 4, 510, 510, 0, "fp16", 207),
    # identity: fits already (:46-47)
    PruneCase("noop_fits", 1, 4, 64, 100, 1, 4, 50, 50, 0, "fp16", 208),
]


def build_prune_inputs(c: PruneCase) -> dict:
    dt = torch_dtype(c.dtype)
    past, scores = [], []
    for l in range(c.L):
        k = det_normal((1, c.H, c.S, c.D), c.seed + 10 * l + 1, 1.0, dt)
        v = det_normal((1, c.H, c.S, c.D), c.seed + 10 * l + 2, 1.0, dt)
        # tie-free scores: a per-head permutation of distinct fp16-exact values plus structure
        s = det_normal((1, c.H, c.q_len, c.S), c.seed + 10 * l + 3, 2.0, torch.float32)
        scores.append(s.to(dt))
        past.append((k, v))
    return {"past": past, "scores": scores}


@dataclass(frozen=True)
class ChainCase:
    """Decode step -> prune on the step's OWN attn_scores (run_spatten_llama.py:74-79): the keep set is derived from
    the logits the forward itself produced, so logit parity and selection parity are tested as one chain."""
    name: str
    H: int
    D: int
    past: int          # cached tokens before the decode step (the step appends one)
    start: int
    important: int
    recent: int
    num_coming: int
    seed: int


CHAIN_CASES = [
    ChainCase("g2s_129_to_64", 12, 64, 128, 4, 30, 30, 0, 301),              # configs[0] dims
    ChainCase("d128_516_to_204", 8, 128, 515, 4, 100, 100, 0, 302),
    ChainCase("demo_defaults_700", 8, 128, 699, 0, 150, 150, 90, 303),       # run_spatten_llama.py:134-136 defaults
    ChainCase("c2_4096_to_2048", 32, 128, 4095, 4, 1022, 1022, 0, 304),      # BASELINE config 2, full size
]


def build_chain_inputs(c: ChainCase) -> dict:
    dt = torch.float16
    return {
        "q": det_normal((1, c.H, 1, c.D), c.seed, 1.0, dt),
        "k_new": det_normal((1, c.H, 1, c.D), c.seed + 1000, 1.0, dt),
        "v_new": det_normal((1, c.H, 1, c.D), c.seed + 2000, 1.0, dt),
        "past_k": det_normal((1, c.H, c.past, c.D), c.seed + 5000, 1.0, dt),
        "past_v": det_normal((1, c.H, c.past, c.D), c.seed + 6000, 1.0, dt),
        "position_ids": torch.full((1, 1), c.past, dtype=torch.int64),
        "mask": torch.zeros((1, 1, 1, c.past + 1), dtype=dt),
    }
