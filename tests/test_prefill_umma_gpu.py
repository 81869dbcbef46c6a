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


# ---- fused single-kernel path (csrc/prefill_fused.cu) vs the three-launch path and the oracle -------------------------

def _rand_case(b, H, H_kv, past, q_len, seed):
    D, kv = 128, past + q_len
    q = det_normal((b, H, q_len, D), seed, 1.0, torch.float16)
    k = det_normal((b, H_kv, kv, D), seed + 1, 1.0, torch.float16)
    v = det_normal((b, H_kv, kv, D), seed + 2, 1.0, torch.float16)
    pos = torch.arange(past, kv).unsqueeze(0).expand(b, q_len).contiguous()
    return q, k, v, pos, kv


@pytest.mark.parametrize("b,H,H_kv,past,q_len", [(1, 4, 4, 0, 128), (1, 4, 4, 0, 300), (2, 4, 2, 130, 257), (1, 2, 2, 1000, 129),
                                                 (1, 2, 1, 0, 1024), (1, 1, 1, 77, 16)])
def test_fused_prefill_equals_three_launch_path(b, H, H_kv, past, q_len):
    """Same logits bit for bit (both form fp16(fp16(acc)/sqrt(D)) from the same tensor-core accumulation order), same
    output within the parity bar, same importance column sums."""
    from spatten_b200 import ops
    dev = "cuda"
    q, k, v, pos, kv = _rand_case(b, H, H_kv, past, q_len, 9100 + past + q_len)
    mask = causal_mask(b, q_len, past, torch.float16)[:, 0].contiguous().to(dev)
    args = (q.transpose(1, 2).contiguous().to(dev), k.to(dev), v.to(dev), kv, ops.rope_table(128, kv, dev), pos.to(dev), mask)
    imp_f = torch.zeros((b, H, kv + 5), dtype=torch.float32, device=dev)
    imp_l = torch.zeros_like(imp_f)
    out_f, lg_f = ops.prefill_attn(*args, importance=imp_f)
    out_l, lg_l = ops.prefill_attn(*args, importance=imp_l, legacy=True)
    torch.cuda.synchronize()
    assert torch.equal(lg_f.view(torch.int16), lg_l.view(torch.int16))
    assert rel_err(out_f.cpu(), out_l.cpu()) <= OUT_REL_TOL
    assert torch.allclose(imp_f, imp_l, rtol=2e-3, atol=2e-4)
    assert (imp_f[..., kv:] == 0).all()


def test_fused_prefill_causal_flag_equals_causal_mask_tensor():
    from spatten_b200 import ops
    dev = "cuda"
    b, H, H_kv, past, q_len = 2, 4, 4, 200, 333
    q, k, v, pos, kv = _rand_case(b, H, H_kv, past, q_len, 9300)
    mask = causal_mask(b, q_len, past, torch.float16)[:, 0].contiguous().to(dev)
    base = (q.transpose(1, 2).contiguous().to(dev), k.to(dev), v.to(dev), kv, ops.rope_table(128, kv, dev), pos.to(dev))
    out_m, lg_m = ops.prefill_attn(*base, mask)
    out_c, lg_c = ops.prefill_attn(*base, None, causal=True)
    torch.cuda.synchronize()
    assert torch.equal(lg_m, lg_c) and torch.equal(out_m, out_c)


@pytest.mark.parametrize("kind", ["none", "padding", "random"])
def test_fused_prefill_generic_masks_vs_oracle(kind):
    """A mask that is not the standard causal one takes the element-wise path (:127-132): no mask, a left-padding mask
    on top of the causal one, and an arbitrary 0 / -inf / finite pattern."""
    import math
    from spatten_b200 import ops
    dev = "cuda"
    b, H, past, q_len, D = 2, 2, 60, 200, 128
    q, k, v, pos, kv = _rand_case(b, H, H, past, q_len, 9400)
    if kind == "none":
        mask = None
    else:
        mask = causal_mask(b, q_len, past, torch.float16).clone()
        if kind == "padding":
            mask[1, :, :, :37] = torch.finfo(torch.float16).min
        else:
            r = det_normal((b, 1, q_len, kv), 9401, 1.0, torch.float32)
            mask = torch.where(r > 0.8, torch.full_like(mask, float("-inf")), mask)
            mask = torch.where((r < -0.8) & (mask == 0), torch.full_like(mask, -3.0), mask)
            mask[:, :, :, 0] = 0                                    # no fully masked rows
    cos, sin = O.rotary_tables(D, kv, torch.float16)
    q_rot = O.rope_single(q, cos, sin, pos)
    k_rot = O.rope_single(k, cos, sin, torch.arange(kv).unsqueeze(0).expand(b, kv))
    ref_scores = torch.matmul(q_rot, k_rot.transpose(2, 3)) / math.sqrt(D)
    s = ref_scores if mask is None else ref_scores + mask
    probs32 = torch.softmax(s.float(), dim=-1)
    ref_out = torch.matmul(probs32.to(torch.float16), v)
    imp = torch.zeros((b, H, kv), dtype=torch.float32, device=dev)
    out, logits = ops.prefill_attn(q.transpose(1, 2).contiguous().to(dev), k.to(dev), v.to(dev), kv, ops.rope_table(D, kv, dev),
                                   pos.to(dev), None if mask is None else mask[:, 0].contiguous().to(dev), importance=imp)
    torch.cuda.synchronize()
    assert_logits_close(logits.cpu(), ref_scores, kind)
    assert rel_err(out.cpu(), ref_out.transpose(1, 2)) <= OUT_REL_TOL
    assert torch.allclose(imp.cpu(), probs32.sum(2), rtol=2e-3, atol=2e-4)


def test_fused_prefill_unaligned_logit_strides():
    """kv_len not a multiple of 8 (scalar logit stores) and a logits buffer with a padded row pitch."""
    from spatten_b200 import ops
    dev = "cuda"
    b, H, past, q_len = 1, 2, 3, 250
    q, k, v, pos, kv = _rand_case(b, H, H, past, q_len, 9500)
    mask = causal_mask(b, q_len, past, torch.float16)[:, 0].contiguous().to(dev)
    args = (q.transpose(1, 2).contiguous().to(dev), k.to(dev), v.to(dev), kv, ops.rope_table(128, kv, dev), pos.to(dev), mask)
    out_a, lg_a = ops.prefill_attn(*args)
    big = torch.full((b, H, q_len, kv + 11), 7.0, dtype=torch.float16, device=dev)
    out_b, lg_b = ops.prefill_attn(*args, logits=big[..., :kv])
    out_l, lg_l = ops.prefill_attn(*args, legacy=True)
    torch.cuda.synchronize()
    assert torch.equal(lg_a, lg_l) and torch.equal(lg_b, lg_l) and (big[..., kv:] == 7.0).all()
    assert torch.equal(out_a, out_b) and rel_err(out_a.cpu(), out_l.cpu()) <= OUT_REL_TOL
