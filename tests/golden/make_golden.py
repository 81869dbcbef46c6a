"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run once in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference forward (spatten_llm/pos_shift/modify_llama.py:31-168) was written against
transformers==4.33.0 (README.md:46); the installed 5.5 LlamaAttention lacks the attributes it
reads, so it is called through a shim `self` that supplies the 4.33 attribute contract
(SURVEY.md App. A.3): hidden_size, num_heads, num_key_value_heads, head_dim,
num_key_value_groups, config.pretraining_tp, q/k/v/o_proj, rotary_emb(x, seq_len=).
The reference code itself is not modified or copied; only its outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))           # tests/
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))   # repo root (baseline/)
sys.path.insert(0, "/root/reference")

from spatten_llm.pos_shift.modify_llama import llama_pos_shift_attention_forward  # noqa: E402
from spatten_llm.kv_cache_token_pruning import SpAttenKVCache                      # noqa: E402

from cases import (CHAIN_CASES, FWD_CASES, PRUNE_CASES, build_chain_inputs, build_fwd_inputs,  # noqa: E402
                   build_prune_inputs)


from baseline.ref_shim import Const, Linear, Rotary433, ShimAttention as _ShimBase     # noqa: E402


class ShimAttention(_ShimBase):
    forward = llama_pos_shift_attention_forward          # the reference function, unmodified


def run_fwd(c):
    d = build_fwd_inputs(c)
    m = ShimAttention(c.H, c.H_kv, c.D)
    if c.full:
        m.q_proj, m.k_proj, m.v_proj, m.o_proj = Linear(d["wq"]), Linear(d["wk"]), Linear(d["wv"]), Linear(d["wo"])
        hidden = d["hidden"]
    else:
        flat = lambda x: x.transpose(1, 2).reshape(c.b, c.q_len, -1)
        m.q_proj, m.k_proj, m.v_proj = Const(flat(d["q"])), Const(flat(d["k_new"])), Const(flat(d["v_new"]))
        m.o_proj = torch.nn.Identity()
        hidden = torch.zeros(c.b, c.q_len, c.H * c.D, dtype=d["q"].dtype)
    past = (d["past_k"], d["past_v"]) if d["past_k"] is not None else None
    with torch.no_grad():
        out, w, new_past = m(hidden, attention_mask=d["mask"], position_ids=d["position_ids"],
                             past_key_value=past, use_cache=True)
    assert w is None
    res = {"out": out, "attn_scores": m.attn_scores}
    # the cache must be the un-rotated keys incl. the new ones (modify_llama.py:100); store a checksum row
    res["cache_k_last"] = new_past[0][:, :, -1, :]
    res["cache_v_last"] = new_past[1][:, :, -1, :]
    res["cache_len"] = torch.tensor(new_past[0].shape[2])
    return {k: v.numpy() for k, v in res.items()}          # native dtype (fp16 stays fp16)


def run_prune(c):
    d = build_prune_inputs(c)
    cache = SpAttenKVCache(start_size=c.start, recent_size=c.recent, important_size=c.important,
                           k_seq_dim=2, v_seq_dim=2)
    # index-coded twin cache: K[..., 0] = token index, so the kept index set can be read back
    coded = []
    for (k, v) in d["past"]:
        kk = k.float().clone()
        kk[..., 0] = torch.arange(c.S, dtype=torch.float32)[None, None, :]
        coded.append((kk, v.float()))
    new_coded = cache.apply_token_pruning(coded, c.num_coming, [s.clone() for s in d["scores"]])
    new_past = cache.apply_token_pruning(d["past"], c.num_coming, [s.clone() for s in d["scores"]])
    res = {}
    for l in range(c.L):
        res[f"kept_idx_{l}"] = new_coded[l][0][0, :, :, 0].to(torch.int32).numpy()      # [H, S']
        k2, v2 = new_past[l]
        if c.S <= 300 and new_past is not d["past"]:
            res[f"k_{l}"], res[f"v_{l}"] = k2.numpy(), v2.numpy()
        # position-weighted checksums pin the gathered bytes of the large cases
        w = torch.arange(1, k2.shape[2] + 1, dtype=torch.float64)[None, None, :, None]
        res[f"k_wsum_{l}"] = (k2.double() * w).sum(dim=(0, 2)).numpy()                  # [H, D]
        res[f"v_wsum_{l}"] = (v2.double() * w).sum(dim=(0, 2)).numpy()
        res[f"shape_{l}"] = np.array(k2.shape)
    res["identity"] = np.array(new_past is d["past"])
    return res


def run_chain(c):
    """Reference decode step, then the reference prune on the attn_scores that step stashed (the demo's call pattern,
    run_spatten_llama.py:74-79).  Stores the scores, the kept index set, and per head the gap between the k-th and
    (k+1)-th candidate score (0 = a tie at the boundary, where torch.topk's choice is implementation-defined)."""
    d = build_chain_inputs(c)
    m = ShimAttention(c.H, c.H, c.D)
    flat = lambda x: x.transpose(1, 2).reshape(1, 1, -1)
    m.q_proj, m.k_proj, m.v_proj = Const(flat(d["q"])), Const(flat(d["k_new"])), Const(flat(d["v_new"]))
    m.o_proj = torch.nn.Identity()
    with torch.no_grad():
        out, _, new_past = m(torch.zeros(1, 1, c.H * c.D, dtype=torch.float16), attention_mask=d["mask"],
                             position_ids=d["position_ids"], past_key_value=(d["past_k"], d["past_v"]), use_cache=True)
    scores = m.attn_scores                                        # [1, H, 1, S]
    S = c.past + 1
    cache = SpAttenKVCache(start_size=c.start, recent_size=c.recent, important_size=c.important, k_seq_dim=2, v_seq_dim=2)
    kk = new_past[0].float().clone()
    kk[..., 0] = torch.arange(S, dtype=torch.float32)[None, None, :]
    coded = cache.apply_token_pruning([(kk, new_past[1].float())], c.num_coming, [scores.clone()])
    kept = coded[0][0][0, :, :, 0].to(torch.int32)                # [H, keep_total]
    cand = scores[0, :, 0, c.start: S - c.recent + c.num_coming].float()
    srt = torch.sort(cand, dim=-1, descending=True).values
    return {"out": out.numpy(), "attn_scores": scores.numpy(), "kept_idx": kept.numpy(),
            "kth": srt[:, c.important - 1].numpy(), "next": srt[:, c.important].numpy()}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)       # one thread => one summation order for the stored logits
    only = sys.argv[1] if len(sys.argv) > 1 else None            # e.g. `make_golden.py chain`: (re)generate one family
    if only == "chain":
        for c in CHAIN_CASES:
            np.savez_compressed(os.path.join(HERE, f"chain_{c.name}.npz"), **run_chain(c))
            print("chain", c.name)
        return
    for c in FWD_CASES:
        np.savez_compressed(os.path.join(HERE, f"fwd_{c.name}.npz"), **run_fwd(c))
        print("fwd", c.name)
    for c in CHAIN_CASES:
        np.savez_compressed(os.path.join(HERE, f"chain_{c.name}.npz"), **run_chain(c))
        print("chain", c.name)
    for c in PRUNE_CASES:
        np.savez_compressed(os.path.join(HERE, f"prune_{c.name}.npz"), **run_prune(c))
        print("prune", c.name)


if __name__ == "__main__":
    main()
