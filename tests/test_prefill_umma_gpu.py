"""GPU parity of the tcgen05 prefill path (spatten_prefill_attn) vs reference goldens, the oracle and the row-wise path."""
import pytest
import torch

from cases import FWD_CASES, build_fwd_inputs, causal_mask
from detgen import det_normal
from helpers import OUT_REL_TOL, assert_logits_close, load_golden, rel_err
from oracle import spatten_oracle as O

pytestmark = pytest.mark.gpu

CASES = [c for c in FWD_CASES if c.q_len > 1 and c.dtype == "fp16" and c.D == 128]


def _run(c, d):
    from spatten_b200 import ops
    dev, kv = "cuda", c.past + c.q_len
    kc = torch.zeros((c.b, c.H_kv, kv + 5, c.D), dtype=torch.float16, device=dev)
    vc = torch.zeros_like(kc)
    if c.past:
        kc[:, :, :c.past], vc[:, :, :c.past] = d["past_k"].to(dev), d["past_v"].to(dev)
    kc[:, :, c.past:kv], vc[:, :, c.past:kv] = d["k_new"].to(dev), d["v_new"].to(dev)
    q = d["q"].transpose(1, 2).contiguous().to(dev)
    mask = d["mask"][:, 0].contiguous().to(dev) if d["mask"] is not None else None
    out, logits = ops.prefill_attn(q, kc, vc, kv, ops.rope_table(c.D, kv, dev), d["position_ids"].to(dev), mask)
    torch.cuda.synchronize()
    return out.cpu(), logits.cpu()


@pytest.mark.parametrize("c", CASES, ids=lambda c: c.name)
def test_prefill_umma_matches_reference_golden(c):
    d, g = build_fwd_inputs(c), load_golden("fwd", c.name)
    out, logits = _run(c, d)
    assert_logits_close(logits, g["attn_scores"], c.name)
    exact = (logits.view(torch.int16) == g["attn_scores"].view(torch.int16)).float().mean().item()
    assert exact > 0.99
    assert rel_err(out, g["out"].view(c.b, c.q_len, c.H, c.D)) <= OUT_REL_TOL


@pytest.mark.parametrize("b,H,H_kv,past,q_len", [(1, 8, 8, 0, 300), (2, 4, 4, 130, 257), (1, 8, 2, 0, 128), (1, 2, 2, 1000, 129)])
def test_prefill_umma_matches_oracle_and_rowwise(b, H, H_kv, past, q_len):
    from spatten_b200 import ops
    D, kv, dev = 128, past + q_len, "cuda"
    seed = 4000 + past + q_len
    q = det_normal((b, H, q_len, D), seed, 1.0, torch.float16)
    kn = det_normal((b, H_kv, q_len, D), seed + 1, 1.0, torch.float16)
    vn = det_normal((b, H_kv, q_len, D), seed + 2, 1.0, torch.float16)
    pk = det_normal((b, H_kv, past, D), seed + 3, 1.0, torch.float16) if past else None
    pv = det_normal((b, H_kv, past, D), seed + 4, 1.0, torch.float16) if past else None
    pos = torch.arange(past, kv).unsqueeze(0).expand(b, q_len).contiguous()
    mask = causal_mask(b, q_len, past, torch.float16)
    ref_out, ref_scores, (k_all, v_all) = O.attn_forward_core(q, kn, vn, pk, pv, pos, mask)
    kc, vc = k_all.to(dev).contiguous(), v_all.to(dev).contiguous()
    q4 = q.transpose(1, 2).contiguous().to(dev)
    rope = ops.rope_table(D, kv, dev)
    out, logits = ops.prefill_attn(q4, kc, vc, kv, rope, pos.to(dev), mask[:, 0].contiguous().to(dev))
    torch.cuda.synchronize()
    assert_logits_close(logits.cpu(), ref_scores)
    assert rel_err(out.cpu(), ref_out.transpose(1, 2)) <= OUT_REL_TOL
    # the row-wise CUDA path (one decode-kernel row per query) computes the same function
    lg2 = torch.empty_like(logits)
    out2 = ops.decode_attn(q4, kc, vc, kv, rope, q_pos=pos.to(dev), mask=mask[:, 0].contiguous().to(dev), logits=lg2)
    assert (logits == lg2).float().mean().item() > 0.99
    assert rel_err(out.cpu(), out2.cpu()) <= 2e-3


def test_prefill_umma_full_size_8192_sampled_rows_vs_oracle():
    """BASELINE config 3's prefill length (S = 8192): sampled query rows of the tcgen05 path against the oracle's op
    sequence (RoPE in fp16, fp16 logits = fp16(q.k)/sqrt(D), additive causal mask, fp32 softmax -> fp16, P.V)."""
    import math
    from spatten_b200 import ops
    b, H, D, S, dev = 1, 2, 128, 8192, "cuda"
    q = det_normal((b, H, S, D), 7101, 1.0, torch.float16)
    k = det_normal((b, H, S, D), 7102, 1.0, torch.float16)
    v = det_normal((b, H, S, D), 7103, 1.0, torch.float16)
    pos = torch.arange(S).unsqueeze(0)
    mask = causal_mask(b, S, 0, torch.float16)
    out, logits = ops.prefill_attn(q.transpose(1, 2).contiguous().to(dev), k.to(dev), v.to(dev), S, ops.rope_table(D, S, dev),
                                   pos.to(dev), mask[:, 0].contiguous().to(dev))
    torch.cuda.synchronize()
    rows = torch.tensor([0, 1, 127, 128, 129, 1000, 4095, 4096, 4097, 6143, 8000, 8190, 8191])
    cos, sin = O.rotary_tables(D, S, torch.float16)
    q_rot = O.rope_single(q[:, :, rows], cos, sin, rows.unsqueeze(0))
    k_rot = O.rope_single(k, cos, sin, torch.arange(S).unsqueeze(0))
    ref_scores = torch.matmul(q_rot, k_rot.transpose(2, 3)) / math.sqrt(D)
    assert_logits_close(logits[:, :, rows.to(dev)].cpu(), ref_scores, "prefill 8192")
    probs = torch.softmax((ref_scores + mask[:, :, rows]).float(), dim=-1).to(torch.float16)
    ref_out = torch.matmul(probs, v)                                          # [b, H, rows, D]
    assert rel_err(out[:, rows.to(dev)].cpu(), ref_out.transpose(1, 2)) <= OUT_REL_TOL


@pytest.mark.parametrize("b,H,past,q_len", [(1, 4, 0, 300), (2, 2, 130, 257), (1, 2, 0, 2048)])
def test_prefill_importance_column_sums(b, H, past, q_len):
    """importance[b,h,j] += sum_q softmax(S + mask)[b,h,q,j] == oracle.softmax_importance per sequence (parity-unpinned
    restatement of README.md:21's cumulative token importance); accumulates across calls."""
    from spatten_b200 import ops
    D, kv, dev = 128, past + q_len, "cuda"
    seed = 8800 + past + q_len
    q = det_normal((b, H, q_len, D), seed, 1.0, torch.float16)
    k = det_normal((b, H, kv, D), seed + 1, 1.0, torch.float16)
    v = det_normal((b, H, kv, D), seed + 2, 1.0, torch.float16)
    pos = torch.arange(past, kv).unsqueeze(0).expand(b, q_len).contiguous()
    mask = causal_mask(b, q_len, past, torch.float16)
    imp = torch.zeros((b, H, kv + 3), dtype=torch.float32, device=dev)
    args = (q.transpose(1, 2).contiguous().to(dev), k.to(dev), v.to(dev), kv, ops.rope_table(D, kv, dev), pos.to(dev),
            mask[:, 0].contiguous().to(dev))
    _, logits = ops.prefill_attn(*args, importance=imp)
    torch.cuda.synchronize()
    ref = torch.stack([O.softmax_importance(logits[i:i + 1].cpu(), mask[i:i + 1]) for i in range(b)])     # [b, H, kv]
    assert torch.allclose(imp[..., :kv].cpu(), ref, rtol=2e-3, atol=2e-4)
    assert (imp[..., kv:] == 0).all()
    assert abs(float(imp.sum()) - b * H * q_len) / (b * H * q_len) < 1e-3      # every row of P sums to one
    ops.prefill_attn(*args, importance=imp)
    torch.cuda.synchronize()
    assert torch.allclose(imp[..., :kv].cpu(), 2 * ref, rtol=2e-3, atol=4e-4)
