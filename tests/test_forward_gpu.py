"""GPU parity of the patched LlamaAttention.forward (module API) and of prefill rows through the C ABI."""
import pytest
import torch

from cases import FWD_CASES, build_fwd_inputs
from helpers import OUT_REL_TOL, assert_logits_close, load_golden, rel_err
from oracle import spatten_oracle as O

pytestmark = pytest.mark.gpu

PREFILL_CASES = [c for c in FWD_CASES if c.q_len > 1 and c.dtype == "fp16"]


@pytest.mark.parametrize("c", PREFILL_CASES, ids=lambda c: c.name)
def test_prefill_rows_match_reference_golden(c):
    from spatten_b200 import ops
    d, g = build_fwd_inputs(c), load_golden("fwd", c.name)
    dev, kv = "cuda", c.past + c.q_len
    kc = torch.zeros((c.b, c.H_kv, kv + 3, c.D), dtype=torch.float16, device=dev)
    vc = torch.zeros_like(kc)
    if c.past:
        kc[:, :, :c.past], vc[:, :, :c.past] = d["past_k"].to(dev), d["past_v"].to(dev)
    kc[:, :, c.past:kv], vc[:, :, c.past:kv] = d["k_new"].to(dev), d["v_new"].to(dev)
    logits = torch.empty((c.b, c.H, c.q_len, kv), dtype=torch.float16, device=dev)
    q = d["q"].transpose(1, 2).contiguous().to(dev)                      # [b, q, H, D]
    out = ops.decode_attn(q, kc, vc, kv, ops.rope_table(c.D, kv, dev), q_pos=d["position_ids"].to(dev),
                          mask=d["mask"][:, 0].contiguous().to(dev), logits=logits)
    torch.cuda.synchronize()
    assert_logits_close(logits.cpu(), g["attn_scores"], c.name)
    ref = g["out"].view(c.b, c.q_len, c.H, c.D)                          # reference returns [b, q, H*D]
    assert rel_err(out.cpu(), ref) <= OUT_REL_TOL


def _module_from(c, d, dev):
    from spatten_b200.modeling import LlamaAttention433
    m = LlamaAttention433(c.H * c.D, c.H, c.H_kv, device=dev)
    with torch.no_grad():
        m.q_proj.weight.copy_(d["wq"]); m.k_proj.weight.copy_(d["wk"])
        m.v_proj.weight.copy_(d["wv"]); m.o_proj.weight.copy_(d["wo"])
    return m


def test_patched_forward_c1_decode_matches_reference_golden():
    """BASELINE config 1 (GPT-2-small dims, seq 128) decode step through enable_spatten_llm + forward."""
    from spatten_b200.enable_spatten_llm import enable_spatten_llm
    c = next(x for x in FWD_CASES if x.name == "c1_g2s_decode_fp16")
    d, g = build_fwd_inputs(c), load_golden("fwd", c.name)
    dev = "cuda"
    m = _module_from(c, d, dev)
    holder = torch.nn.Module()
    holder.attn = m
    holder.config = m.config
    enable_spatten_llm(holder, 4, 30, 30)
    out, w, past = m(d["hidden"].to(dev), attention_mask=d["mask"].to(dev), position_ids=d["position_ids"].to(dev),
                     past_key_value=(d["past_k"].to(dev), d["past_v"].to(dev)), use_cache=True)
    torch.cuda.synchronize()
    assert w is None and past[0].shape[2] == c.past + 1
    # the projections run on cuBLAS here and on CPU in the reference: tolerance, not bit equality
    assert rel_err(out.cpu(), g["out"]) <= 2e-3
    assert rel_err(m.attn_scores.cpu(), g["attn_scores"]) <= 2e-3
    assert m.attn_scores.shape == (c.b, c.H, 1, c.past + 1)


def test_patched_forward_prefill_then_decode_chain_vs_oracle():
    """Prefill 70 tokens, then 5 decode steps, cache carried as (K, V) views exactly like greedy_generate
    (run_spatten_llama.py:18-57) carries past_key_values; compared step by step with the oracle chain."""
    from spatten_b200.modeling import LlamaAttention433
    from spatten_b200.pos_shift.modify_llama import enable_llama_pos_shift_attention
    from cases import causal_mask
    from detgen import det_normal
    H, D, hid, dev = 4, 128, 512, "cuda"
    m = LlamaAttention433(hid, H, device=dev, seed=5)
    holder = torch.nn.Module(); holder.attn = m
    assert enable_llama_pos_shift_attention(holder) == 1
    W = [getattr(m, n).weight.detach().cpu() for n in ("q_proj", "k_proj", "v_proj", "o_proj")]
    past_gpu, pk, pv, n = None, None, None, 0
    for step, q_len in enumerate([70, 1, 1, 1, 1, 1]):
        hs = det_normal((1, q_len, hid), 900 + step, 1.0, torch.float16)
        pos = torch.arange(n, n + q_len).unsqueeze(0)
        mask = causal_mask(1, q_len, n, torch.float16)
        ref_out, ref_scores, (pk, pv) = O.attn_forward(hs, *W, H, H, pk, pv, pos, mask)
        out, _, past_gpu = m(hs.to(dev), attention_mask=mask.to(dev), position_ids=pos.to(dev),
                             past_key_value=past_gpu, use_cache=True)
        n += q_len
        assert past_gpu[0].shape[2] == n
        assert rel_err(out.cpu(), ref_out) <= 3e-3, step
        assert rel_err(m.attn_scores.cpu(), ref_scores) <= 3e-3, step
        # cache holds un-rotated keys (modify_llama.py:100)
        assert rel_err(past_gpu[0].cpu(), pk) <= 2e-3


def test_patched_forward_rejects_bad_mask_and_cpu():
    from spatten_b200.modeling import LlamaAttention433
    from spatten_b200.pos_shift.modify_llama import enable_llama_pos_shift_attention
    m = LlamaAttention433(256, 4, device="cuda")
    holder = torch.nn.Module(); holder.attn = m
    enable_llama_pos_shift_attention(holder)
    hs = torch.zeros(1, 1, 256, dtype=torch.float16, device="cuda")
    with pytest.raises(ValueError, match="Attention mask should be of size"):
        m(hs, attention_mask=torch.zeros(1, 1, 1, 7, dtype=torch.float16, device="cuda"))
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(hs.cpu())
    with pytest.raises(TypeError):
        m(hs.float())
