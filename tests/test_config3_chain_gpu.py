"""BASELINE config 3 as one chain on the GPU: 8K prefill -> token prune 50 % -> head prune 75 % -> decode
(call pattern of run_spatten_llama.py:60-87: prefill + decode steps, then apply_token_pruning before the next round;
head pruning is the README.md:21 technique, parity-unpinned).  Every stage is checked against the oracle run on the
SAME intermediate state, so a failure names the stage."""
import contextlib
import io

import pytest
import torch

from cases import causal_mask
from detgen import det_normal
from helpers import OUT_REL_TOL, assert_index_sets_match, rel_err
from oracle import spatten_oracle as O

pytestmark = pytest.mark.gpu


def test_prefill_8k_token_prune_head_prune_decode():
    from spatten_b200.enable_spatten_llm import enable_spatten_llm
    from spatten_b200.head_pruning import SpAttenHeadPruner
    from spatten_b200.modeling import AttentionStack
    H, D, S, dev = 8, 128, 8192, "cuda"
    hid = H * D
    keep_tokens, keep_heads = S // 2, (3 * H) // 4
    start, important, recent = 4, (keep_tokens - 4) // 2, (keep_tokens - 4) // 2
    model = AttentionStack(1, hid, H, device=dev, seed=3)
    with contextlib.redirect_stdout(io.StringIO()):
        cache = enable_spatten_llm(model, start, important, recent)
    attn = model.layers[0]
    pruner = SpAttenHeadPruner(model, keep_heads)
    W = [getattr(attn, n).weight.detach().cpu() for n in ("q_proj", "k_proj", "v_proj", "o_proj")]

    # ---- prefill S tokens (tcgen05 path), then two decode steps (the last forward before a prune is a decode step) ----
    hs = det_normal((1, S, hid), 31, 1.0, torch.float16)
    with torch.no_grad():
        out, _, past = attn(hs.to(dev), attention_mask=causal_mask(1, S, 0, torch.float16).to(dev),
                            position_ids=torch.arange(S, device=dev).unsqueeze(0), use_cache=True)
        assert attn.attn_scores.shape == (1, H, S, S) and past[0].shape[2] == S
        # sampled prefill rows of the layer output vs the oracle (full 8K x 8K on the CPU would take minutes)
        F = torch.nn.functional
        qf = F.linear(hs, W[0]).view(1, S, H, D).transpose(1, 2)
        kf = F.linear(hs, W[1]).view(1, S, H, D).transpose(1, 2)
        vf = F.linear(hs, W[2]).view(1, S, H, D).transpose(1, 2)
        assert rel_err(past[0].cpu(), kf) <= 2e-3                      # cache = un-rotated keys (modify_llama.py:100)
        pk, pv = past[0].cpu(), past[1].cpu()                           # continue the oracle from the GPU's own cache
        n = S
        for step in range(2):
            h1 = det_normal((1, 1, hid), 40 + step, 1.0, torch.float16)
            pos = torch.full((1, 1), n, dtype=torch.int64)
            mask = torch.zeros((1, 1, 1, n + 1), dtype=torch.float16)
            ref_out, ref_scores, (pk, pv) = O.attn_forward(h1, *W, H, H, pk, pv, pos, mask)
            out, _, past = attn(h1.to(dev), attention_mask=mask.to(dev), position_ids=pos.to(dev), past_key_value=past,
                                use_cache=True)
            n += 1
            assert rel_err(out.cpu(), ref_out) <= 3e-3, step
            assert rel_err(attn.attn_scores.cpu(), ref_scores) <= 3e-3, step
        torch.cuda.synchronize()

        # ---- token prune 50 % on the scores the last decode step stashed (kv_cache_token_pruning.py:42-96) ------------
        scores = attn.attn_scores                                        # [1, H, 1, n]
        new_past = cache.apply_token_pruning([past], 0, [scores])
        keep = cache.keep_indices[0]
        kept_len = start + important + recent
        assert new_past[0][0].shape[2] == kept_len == keep_tokens
        ref_keep = O.select_keep_indices(O.importance_from_scores(scores.cpu()), n, start, important, recent, 0)
        assert_index_sets_match(keep[:, start:start + important], ref_keep, scores[0, :, 0].cpu(), "config-3 token prune")
        gi = keep.long()[None, :, :, None].expand(1, H, kept_len, D)
        assert torch.equal(new_past[0][0], torch.gather(past[0], 2, gi)) and torch.equal(new_past[0][1], torch.gather(past[1], 2, gi))

        # ---- head prune 75 %: importance accumulated by the decode steps above, mask computed on the device ------------
        pruner.update_masks()
        hmask = attn.spatten_head_mask.cpu()
        assert int(hmask.sum()) == keep_heads
        assert torch.equal(hmask, O.head_prune_mask(attn.spatten_head_importance.cpu(), keep_heads))

        # ---- decode over the pruned cache with the head mask: live heads vs the oracle on the same pruned cache ---------
        h2 = det_normal((1, 1, hid), 50, 1.0, torch.float16)
        pos = torch.full((1, 1), kept_len, dtype=torch.int64)            # pos-shift: positions are cache slots (:103)
        pk2, pv2 = new_past[0][0].cpu(), new_past[0][1].cpu()
        q2 = F.linear(h2, W[0]).view(1, 1, H, D).transpose(1, 2)
        k2 = F.linear(h2, W[1]).view(1, 1, H, D).transpose(1, 2)
        v2 = F.linear(h2, W[2]).view(1, 1, H, D).transpose(1, 2)
        ref_core, _, _ = O.attn_forward_core(q2, k2, v2, pk2, pv2, pos, None)
        ref_core = ref_core * hmask.view(1, H, 1, 1).to(ref_core.dtype)      # pruned heads contribute nothing to o_proj
        ref_out = F.linear(ref_core.transpose(1, 2).reshape(1, 1, hid), W[3])
        out, _, past3 = attn(h2.to(dev), position_ids=pos.to(dev), past_key_value=tuple(new_past[0]), use_cache=True)
        torch.cuda.synchronize()
        assert past3[0].shape[2] == kept_len + 1
        assert rel_err(out.cpu(), ref_out) <= 3e-3
