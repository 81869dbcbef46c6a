"""CPU oracle for the spatten_llm sparse-attention decode path.  TEST INFRASTRUCTURE ONLY.

This file is a torch-CPU *restatement* of the reference algorithm.  It is the checker
for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``spatten_b200/`` imports it, and the product path raises when the CUDA library is
missing instead of falling back to this code.

Pinning:
  * PINNED against the reference itself: ``rotary_tables``, ``rope_single``,
    ``attn_forward_core``, ``attn_forward`` and ``token_prune`` are checked against
    outputs of the *unmodified* reference functions (imported from /root/reference in the
    build container by ``tests/golden/make_golden.py``) stored as fixtures under
    ``tests/golden/*.npz`` -- see ``tests/test_oracle_golden.py``.
  * PARITY UNPINNED (the reference has no executable arithmetic for them; restated from the
    RTL / README semantics cited at each function and labelled "self-consistent"):
    ``softmax_importance``, ``local_value_attention`` (Mode B), ``head_importance`` /
    ``head_prune_mask``, ``quantize_kv_planes`` / ``requant_decision`` /
    ``attn_core_quantized``.

All `file:line` citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

# ----------------------------------------------------------------------------------------------
# Third-party pieces the reference calls (transformers==4.33.0, README.md:46).  Not on disk in
# 4.33 form; restated from the published definitions and anchored on the reference call sites.
# ----------------------------------------------------------------------------------------------


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """transformers `rotate_half` (call site pos_shift/modify_llama.py:27): cat(-x[D/2:], x[:D/2])."""
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


def repeat_kv(x: torch.Tensor, n_rep: int) -> torch.Tensor:
    """transformers `repeat_kv` (call sites modify_llama.py:108-109)."""
    if n_rep == 1:
        return x
    b, h, s, d = x.shape
    return x[:, :, None, :, :].expand(b, h, n_rep, s, d).reshape(b, h * n_rep, s, d)


def rotary_tables(head_dim: int, seq_len: int, dtype: torch.dtype, base: float = 10000.0):
    """LlamaRotaryEmbedding.forward(x, seq_len) of transformers 4.33 (call site modify_llama.py:89).

    inv_freq = 1/base^(arange(0,D,2)/D); freqs = outer(arange(S), inv_freq) in fp32;
    emb = cat(freqs, freqs); cos/sin cached in fp32 and cast to x.dtype on return,
    shape [1, 1, seq_len, D].
    """
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2).float() / head_dim))
    t = torch.arange(seq_len, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos = emb.cos()[None, None, :, :].to(dtype)
    sin = emb.sin()[None, None, :, :].to(dtype)
    return cos, sin


# ----------------------------------------------------------------------------------------------
# a2: apply_rotary_pos_emb_single  (modify_llama.py:21-28)
# ----------------------------------------------------------------------------------------------


def rope_single(x, cos, sin, position_ids):
    """x*cos[pos] + rotate_half(x)*sin[pos], every op rounded in x.dtype (modify_llama.py:23-27)."""
    c = cos.squeeze(1).squeeze(0)[position_ids].unsqueeze(1)
    s = sin.squeeze(1).squeeze(0)[position_ids].unsqueeze(1)
    return (x * c) + (rotate_half(x) * s)


# ----------------------------------------------------------------------------------------------
# a1: llama_pos_shift_attention_forward  (modify_llama.py:31-168)
# ----------------------------------------------------------------------------------------------


def attn_forward_core(
    q: torch.Tensor,            # [b, H, q_len, D]   un-rotated query (after q_proj + view/transpose)
    k_new: torch.Tensor,        # [b, H_kv, q_len, D] un-rotated new keys
    v_new: torch.Tensor,        # [b, H_kv, q_len, D]
    past_k: Optional[torch.Tensor],  # [b, H_kv, S_past, D] un-rotated cached keys, or None
    past_v: Optional[torch.Tensor],
    position_ids: torch.Tensor,      # [b, q_len] int64
    attention_mask: Optional[torch.Tensor] = None,  # [b, 1, q_len, kv] additive
    rope_base: float = 10000.0,
):
    """The attention core between the projections (modify_llama.py:86-144).

    Returns (attn_out [b,H,q_len,D], attn_scores [b,H,q_len,kv], (K_cache, V_cache)).
    Follows the reference op by op *including* the O(S) costs it pays every step
    (cat :97-98, RoPE over the whole cache :103-104, logits clone :117), so timing it is a
    fair CPU baseline of the reference path.
    """
    b, H, q_len, D = q.shape
    H_kv = k_new.shape[1]
    kv_seq_len = k_new.shape[-2] + (past_k.shape[-2] if past_k is not None else 0)   # :86-88
    cos, sin = rotary_tables(D, kv_seq_len, v_new.dtype, rope_base)                    # :89
    q_rot = rope_single(q, cos, sin, position_ids)                                     # :92
    if past_k is not None:                                                             # :95-98
        k_all = torch.cat([past_k, k_new], dim=2)
        v_all = torch.cat([past_v, v_new], dim=2)
    else:
        k_all, v_all = k_new, v_new
    cache = (k_all, v_all)                                                             # :100 (un-rotated K)
    key_pos = torch.arange(kv_seq_len).unsqueeze(0)                                    # :103
    k_rot = rope_single(k_all, cos, sin, key_pos)                                      # :104
    k_rot = repeat_kv(k_rot, H // H_kv)                                                # :108
    v_rep = repeat_kv(v_all, H // H_kv)                                                # :109
    logits = torch.matmul(q_rot, k_rot.transpose(2, 3)) / math.sqrt(D)                 # :111-113
    attn_scores = logits.detach().clone()                                              # :116-119
    if attention_mask is not None:                                                     # :127-132
        if attention_mask.size() != (b, 1, q_len, kv_seq_len):
            raise ValueError(
                f"Attention mask should be of size {(b, 1, q_len, kv_seq_len)}, but is {attention_mask.size()}")
        logits = logits + attention_mask
    probs = torch.nn.functional.softmax(logits, dim=-1, dtype=torch.float32).to(q_rot.dtype)  # :135-137
    out = torch.matmul(probs, v_rep)                                                   # :138
    return out, attn_scores, cache


def attn_forward(
    hidden: torch.Tensor,       # [b, q_len, hidden]
    wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor, wo: torch.Tensor,   # nn.Linear weights [out, in]
    num_heads: int, num_kv_heads: int,
    past_k=None, past_v=None, position_ids=None, attention_mask=None, rope_base: float = 10000.0,
):
    """Whole patched forward, pretraining_tp == 1 branch (modify_llama.py:41, :72-84, :146-147, :163).

    Returns (attn_output [b,q_len,hidden], attn_scores, (K_cache, V_cache)).
    """
    F = torch.nn.functional
    b, q_len, hidden_size = hidden.shape
    D = hidden_size // num_heads
    q = F.linear(hidden, wq).view(b, q_len, num_heads, D).transpose(1, 2)
    k = F.linear(hidden, wk).view(b, q_len, num_kv_heads, D).transpose(1, 2)
    v = F.linear(hidden, wv).view(b, q_len, num_kv_heads, D).transpose(1, 2)
    out, scores, cache = attn_forward_core(q, k, v, past_k, past_v, position_ids, attention_mask, rope_base)
    out = out.transpose(1, 2).contiguous().reshape(b, q_len, hidden_size)
    return F.linear(out, wo), scores, cache


# ----------------------------------------------------------------------------------------------
# a3/a4: SpAttenKVCache.apply_token_pruning  (kv_cache_token_pruning.py:24-96)
# ----------------------------------------------------------------------------------------------


def importance_from_scores(attn_scores: torch.Tensor) -> torch.Tensor:
    """[b,H,q,kv] -> [H,kv]: sum over batch then over query, each in the tensor's dtype (:51)."""
    return attn_scores.sum(0).sum(1)


def select_keep_indices(score: torch.Tensor, seq_len: int, start: int, important: int, recent: int,
                        num_coming: int) -> torch.Tensor:
    """Per-head ascending index list of the `important` kept middle tokens (:59-63).

    Tie rule.  torch.topk's choice among equal scores is implementation-defined;
    the oracle fixes the RTL rule instead -- every score strictly above the k-th value, then the
    lowest-index elements equal to it (TopK.scala:159-160, 209-211) -- which selects the same
    *multiset of scores* as torch.topk always, and the same index set whenever the k-th and
    (k+1)-th candidate scores differ.
    """
    cand = score[:, start: seq_len - recent + num_coming]
    H, n = cand.shape
    if important > n:
        raise RuntimeError("selected index k out of range")        # what torch.topk raises (:60)
    # stable descending sort == (score desc, index asc)
    order = torch.sort(cand.float(), dim=-1, descending=True, stable=True).indices
    keep = order[:, :important].sort(dim=-1).values + start
    return keep


def boundary_margin(score: torch.Tensor, seq_len: int, start: int, important: int, recent: int,
                    num_coming: int) -> torch.Tensor:
    """Per head: (k-th largest) - ((k+1)-th largest) candidate score; 0 means a boundary tie."""
    cand = score[:, start: seq_len - recent + num_coming].float()
    srt = torch.sort(cand, dim=-1, descending=True).values
    if important >= cand.shape[1]:
        return torch.full((cand.shape[0],), float("inf"))
    return srt[:, important - 1] - srt[:, important]


def token_prune(past_key_values, num_coming: int, attn_score_all: Sequence[torch.Tensor],
                start: int, important: int, recent: int):
    """apply_token_pruning (:42-96).  Returns (new_past, keep_indices_per_layer).

    Kept set per head: [0,start) U top-`important` of score[start : S-recent+num_coming) U
    [S-recent+num_coming, S); result length start+important+max(0,recent-num_coming).
    """
    if past_key_values is None:                                    # :43-44
        return None, None
    seq_len = past_key_values[0][0].size(2)                        # :45
    if seq_len + num_coming <= start + important + recent:         # :46-47
        return past_key_values, None
    if important <= 0:
        raise TypeError("important_size must be > 0 (reference fails at :63 on None)")
    out, keeps = [], []
    tail0 = seq_len - recent + num_coming
    for (k, v), s in zip(past_key_values, attn_score_all):
        if k.shape[0] != 1:
            raise IndexError("reference prune path requires batch 1 (squeeze at :68)")
        score = importance_from_scores(s)                          # :51
        keep = select_keep_indices(score, seq_len, start, important, recent, num_coming)
        keeps.append(keep)
        H, D = k.shape[1], k.shape[3]
        gidx = keep[None, :, :, None].expand(1, H, important, D)
        k_imp = torch.gather(k, 2, gidx)                           # :65-69 (mask gather == ascending gather)
        v_imp = torch.gather(v, 2, gidx)
        out.append([
            torch.cat([k[:, :, :start], k_imp, k[:, :, tail0:seq_len]], dim=2),    # :74-83
            torch.cat([v[:, :, :start], v_imp, v[:, :, tail0:seq_len]], dim=2),    # :84-93
        ])
    return out, keeps


# ----------------------------------------------------------------------------------------------
# a6: RTL-named stages with no executable reference arithmetic ("parity unpinned").
# ----------------------------------------------------------------------------------------------


def softmax_importance(attn_scores: torch.Tensor, attention_mask=None) -> torch.Tensor:
    """Cumulative-importance increment: softmax probabilities summed over batch and query -> [H, kv] fp32.

    README.md:21 ("cascade token pruning ... according to cumulative token importance"); the RTL
    softmax is a stub (Softmax.scala:30-34) so the arithmetic is the Python path's fp32 softmax
    (modify_llama.py:135).  Self-consistent, not reference-pinned.
    """
    s = attn_scores if attention_mask is None else attn_scores + attention_mask
    p = torch.softmax(s.float(), dim=-1)
    return p.sum(0).sum(1)


def local_value_attention(q_rot, k_rot, v, keep_v: int):
    """Mode B / "local V pruning": score every key, keep top-`keep_v` keys per (b,h) row by score,
    multiply values of survivors only, softmax renormalised over survivors.

    RTL order DotProduct -> Softmax -> TopK(target=val_fetch_num) -> FetchValMat(top-k) -> MultiplyValue
    (SpAttenController.scala:347,368,554,588-591,668-726,749; skip top-k when val_fetch_num >= key_fetch_num :591).
    q_rot [b,H,1,D], k_rot [b,H,S,D] (already rotated), v [b,H,S,D].
    Returns (out [b,H,1,D], logits fp16-rounded [b,H,1,S], keep_idx [b,H,keep_v] ascending).
    Self-consistent, not reference-pinned.
    """
    D = q_rot.shape[-1]
    logits = torch.matmul(q_rot, k_rot.transpose(2, 3)) / math.sqrt(D)
    b, H, _, S = logits.shape
    if keep_v >= S:
        idx = torch.arange(S).expand(b, H, S)
    else:
        order = torch.sort(logits[:, :, 0, :].float(), dim=-1, descending=True, stable=True).indices
        idx = order[..., :keep_v].sort(dim=-1).values
    sel = torch.gather(logits[:, :, 0, :], 2, idx).float()
    p = torch.softmax(sel, dim=-1)
    vsel = torch.gather(v, 2, idx[..., None].expand(b, H, idx.shape[-1], D)).float()
    out = torch.einsum("bhk,bhkd->bhd", p, vsel)[:, :, None, :].to(q_rot.dtype)
    return out, logits, idx


def head_importance(attn_out: torch.Tensor) -> torch.Tensor:
    """Head importance increment = sum |attention output| per head over batch, query and D -> [H] fp32.

    README.md:21 ("head pruning to remove unessential heads"); magnitude definition is the SpAtten
    paper's, not in the repository.  Self-consistent, not reference-pinned.
    """
    return attn_out.float().abs().sum(dim=(0, 2, 3))


def head_prune_mask(cum_head_importance: torch.Tensor, keep_heads: int) -> torch.Tensor:
    """Keep the `keep_heads` most important heads; ties -> lowest head index.  Returns u8 [H]."""
    H = cum_head_importance.numel()
    order = torch.sort(cum_head_importance.float(), descending=True, stable=True).indices
    mask = torch.zeros(H, dtype=torch.uint8)
    mask[order[:keep_heads]] = 1
    return mask


# nibble position p (bits 4p..4p+3 of each little-endian u32 that covers 8 dims) holds dim _NIB_ORDER[p] of the group
_NIB_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]
# word position w of a 128-dim row holds the 8-dim group _WORD_ORDER[w]: every 16-byte chunk kq carries dims
# [16kq, 16kq+16) followed by their RoPE partners [64+16kq, 64+16kq+16), which is what one lane of the MMA kernel fetches
_WORD_ORDER = [0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15]


def _pack_nibbles(n: torch.Tensor) -> torch.Tensor:
    """int nibbles [.., D] (dim order) -> u8 [.., D/2] in the plane's fetch order."""
    *lead, D = n.shape
    g = n.reshape(*lead, D // 8, 8)[..., _NIB_ORDER]            # now indexed by nibble position
    if D == 128:
        g = g[..., _WORD_ORDER, :]
    return (g[..., 0::2] | (g[..., 1::2] << 4)).reshape(*lead, D // 2).to(torch.uint8)


def _unpack_nibbles(p: torch.Tensor) -> torch.Tensor:
    """u8 [.., D/2] -> int32 nibbles [.., D] in dim order."""
    *lead, half = p.shape
    lo, hi = (p & 15).to(torch.int32), (p >> 4).to(torch.int32)
    pos = torch.stack((lo, hi), dim=-1).reshape(*lead, half // 4, 8)       # nibble positions 0..7 of each group
    if half == 64:
        pos = pos[..., [_WORD_ORDER.index(c) for c in range(16)], :]
    inv = [_NIB_ORDER.index(d) for d in range(8)]
    return pos[..., inv].reshape(*lead, half * 2)


def quantize_kv_planes(x: torch.Tensor):
    """fp16 [.., S, D] -> (msb u8 [.., S, D/2], lsb u8 [.., S, D/2], scale fp16 [.., S]).

    Symmetric int8 per (row) quantisation q = clamp(rint(x/scale), -127, 127), scale = max|x|/127
    (fp16-rounded), stored as an unsigned byte u = q + 128 split MSB-first into two 4-bit planes:
    msb = u >> 4, lsb = u & 15 (MatrixFetcher.scala:48-51 profiles (4,1)/(8,1); MSB fetch
    left-aligned :344-345, LSB refetch right-aligned :346-347; masked low-4-bit write
    SpAttenController.scala:230-232).  Nibble order inside every 8-dim group is [d0,d2,d4,d6,d1,d3,d5,d7] and
    the 8-dim groups of a 128-dim row are ordered _WORD_ORDER (a fetch-friendly interleave, cf. the RTL's
    bit-reversed layout BMR.scala:59-65).
    MSB-only dequant uses the bucket midpoint: x ~ ((msb<<4 | 8) - 128) * scale.
    Self-consistent, not reference-pinned.
    """
    xf = x.float()
    amax = xf.abs().amax(dim=-1)
    scale = (amax / 127.0).to(torch.float16)
    sf = scale.float().clamp_min(1e-8)
    q = torch.clamp(torch.round(xf / sf[..., None]), -127, 127).to(torch.int32)
    u = q + 128
    return _pack_nibbles(u >> 4), _pack_nibbles(u & 15), scale


def dequant_ints(msb: torch.Tensor, lsb: Optional[torch.Tensor]) -> torch.Tensor:
    """Planes -> the signed integer code [.., S, D] (int32): 16*m + l - 128, l = 8 (mid-point) without the LSB plane."""
    m = _unpack_nibbles(msb)
    l = _unpack_nibbles(lsb) if lsb is not None else torch.full_like(m, 8)
    return m * 16 + l - 128


def dequant_planes(msb: torch.Tensor, lsb: Optional[torch.Tensor], scale: torch.Tensor) -> torch.Tensor:
    """Inverse of quantize_kv_planes -> fp32 [.., S, D].  lsb=None => MSB-only (midpoint) dequant."""
    return dequant_ints(msb, lsb).float() * scale.float()[..., None]


def requant_decision(max_prob: torch.Tensor, threshold: float) -> torch.Tensor:
    """need_requant := max(row) < threshold  (RequantDecision.scala:47-72).  bool per (b,h) row."""
    return max_prob < threshold


def attn_core_quantized(q_rot, k_msb, k_lsb, k_scale, v_msb, v_lsb, v_scale, rope_cos, rope_sin,
                        threshold: float, v_bits: int = 4):
    """Progressive-quantisation decode attention (config 4), one query row per (b,h).

    Pass 1 scores with MSB-only keys; if the largest softmax probability of the row is below
    `threshold` (flat distribution => quantisation error matters, RequantDecision.scala:72) the row
    refetches the LSB plane and rescoring uses full 8-bit keys (SpAttenController.scala:140-175 replay);
    a replayed row recomputes its output from the 8-bit logits.
    Keys are stored un-rotated; RoPE is applied to the integer code (exact in fp16) with the same fp16
    op sequence as the fp16 path (modify_llama.py:27), the dot with the rotated query accumulates in fp32
    and is multiplied by the token's scale, then rounded to fp16 and divided by sqrt(D) like :111-113.
    Values use the MSB plane only (v_bits=4) or both planes (v_bits=8).
    rope_cos / rope_sin: fp16 [S, D] (HF tables).
    Returns (out fp16 [b,H,1,D], logits fp16 [b,H,1,S], need_requant bool [b,H]).
    Self-consistent, not reference-pinned.
    """
    D = q_rot.shape[-1]
    cos, sin = rope_cos.to(torch.float16), rope_sin.to(torch.float16)

    def scores(k_int):
        kh = k_int.to(torch.float16)                                       # exact: |code| <= 128
        kr = (kh * cos) + (rotate_half(kh) * sin)                          # fp16, rounded after every op
        s = torch.einsum("bhd,bhsd->bhs", q_rot[:, :, 0, :].float(), kr.float()) * k_scale.float()
        return (s.to(torch.float16).float() / math.sqrt(D)).to(torch.float16)

    s1 = scores(dequant_ints(k_msb, None))
    p1 = torch.softmax(s1.float(), dim=-1)
    need = requant_decision(p1.amax(dim=-1), threshold) if threshold > 0 else torch.zeros_like(p1[..., 0], dtype=torch.bool)
    s2 = scores(dequant_ints(k_msb, k_lsb))
    logits = torch.where(need[..., None], s2, s1)
    p = torch.softmax(logits.float(), dim=-1)
    vd = dequant_planes(v_msb, v_lsb if v_bits == 8 else None, v_scale)
    out = torch.einsum("bhs,bhsd->bhd", p, vd)[:, :, None, :].to(torch.float16)
    return out, logits[:, :, None, :], need
