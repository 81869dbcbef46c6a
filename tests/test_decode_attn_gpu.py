"""GPU parity: spatten_decode_attn (through the C ABI) vs the oracle and the reference goldens."""
import math

import pytest
import torch

from cases import FWD_CASES, build_fwd_inputs
from detgen import det_normal
from helpers import OUT_REL_TOL, assert_logits_close, load_golden, rel_err
from oracle import spatten_oracle as O

pytestmark = pytest.mark.gpu

DECODE_CASES = [c for c in FWD_CASES if c.q_len == 1 and c.dtype == "fp16"]


def run_kernel(q, k_new, v_new, past_k, past_v, position_ids, mask, cap_pad=5, want_importance=False,
               head_mask=None, flags=0):
    from spatten_b200 import ops
    dev = "cuda"
    b, H, _, D = q.shape
    H_kv = k_new.shape[1]
    past = 0 if past_k is None else past_k.shape[2]
    cap = past + 1 + cap_pad
    kc = torch.full((b, H_kv, cap, D), float("nan"), dtype=torch.float16, device=dev)
    vc = torch.full((b, H_kv, cap, D), float("nan"), dtype=torch.float16, device=dev)
    if past:
        kc[:, :, :past] = past_k.to(dev)
        vc[:, :, :past] = past_v.to(dev)
    rope = ops.rope_table(D, past + 1, dev)
    logits = torch.full((b, H, past + 1 + 3), float("nan"), dtype=torch.float16, device=dev)
    imp = torch.zeros((b, H, past + 1), dtype=torch.float32, device=dev) if want_importance else None
    himp = torch.zeros(H, dtype=torch.float32, device=dev) if want_importance else None
    m = None if mask is None else mask[:, 0, 0, :].contiguous().to(dev)
    out = ops.decode_attn(q[:, :, 0].contiguous().to(dev), kc, vc, past, rope,
                          k_new=k_new[:, :, 0].contiguous().to(dev), v_new=v_new[:, :, 0].contiguous().to(dev),
                          q_pos=position_ids[:, 0].contiguous().to(dev), mask=m, logits=logits,
                          importance=imp, head_importance=himp, flags=flags,
                          head_mask=None if head_mask is None else head_mask.to(dev))
    torch.cuda.synchronize()
    return out.cpu(), logits[:, :, :past + 1].cpu(), kc.cpu(), vc.cpu(), imp, himp


@pytest.mark.parametrize("c", DECODE_CASES, ids=lambda c: c.name)
def test_decode_matches_reference_golden(c):
    d, g = build_fwd_inputs(c), load_golden("fwd", c.name)
    F = torch.nn.functional
    if c.full:
        q = F.linear(d["hidden"], d["wq"]).view(c.b, 1, c.H, c.D).transpose(1, 2)
        k = F.linear(d["hidden"], d["wk"]).view(c.b, 1, c.H_kv, c.D).transpose(1, 2)
        v = F.linear(d["hidden"], d["wv"]).view(c.b, 1, c.H_kv, c.D).transpose(1, 2)
    else:
        q, k, v = d["q"], d["k_new"], d["v_new"]
    out, logits, kc, vc, _, _ = run_kernel(q, k, v, d["past_k"], d["past_v"], d["position_ids"], d["mask"])
    # attn_scores: [b,H,1,kv]
    assert_logits_close(logits, g["attn_scores"][:, :, 0, :], c.name)
    exact = (logits.view(torch.int16) == g["attn_scores"][:, :, 0, :].view(torch.int16)).float().mean().item()
    assert exact > 0.99, f"only {exact:.4f} of logits bit-identical to the reference"
    if c.full:
        mine = F.linear(out.reshape(c.b, 1, c.H * c.D), d["wo"])
        ref = g["out"]
    else:
        mine, ref = out, g["out"].view(c.b, c.H, c.D)
    assert rel_err(mine, ref) <= OUT_REL_TOL
    # cache append: un-rotated key/value in slot `past`, earlier slots untouched, later slots untouched
    assert torch.equal(kc[:, :, c.past], g["cache_k_last"]) and torch.equal(vc[:, :, c.past], g["cache_v_last"])
    if c.past:
        assert torch.equal(kc[:, :, :c.past], d["past_k"])
    assert torch.isnan(kc[:, :, c.past + 1:]).all()


@pytest.mark.parametrize("b,H,H_kv,D,past", [
    (1, 32, 32, 128, 2048),      # BASELINE config 2 shape (one layer)
    (1, 32, 32, 128, 4095),
    (2, 40, 40, 128, 777),
    (8, 5, 5, 128, 3000),        # config-5-like: few heads per GPU, batch 8
    (1, 4, 4, 128, 1),
    (1, 1, 1, 128, 9),
    (3, 3, 3, 64, 130),
    (1, 12, 12, 64, 64),
    (1, 7, 7, 128, 15),
    (1, 32, 8, 128, 1000),       # GQA
    (8, 5, 5, 128, 16383),       # BASELINE config 5, one GPU's shard at full size
    (1, 40, 40, 128, 2047),      # 13B head count at the config-2 length
])
def test_decode_matches_oracle(b, H, H_kv, D, past):
    seed = 9000 + b * 131 + H * 17 + past
    q = det_normal((b, H, 1, D), seed, 1.0, torch.float16)
    k = det_normal((b, H_kv, 1, D), seed + 1, 1.0, torch.float16)
    v = det_normal((b, H_kv, 1, D), seed + 2, 1.0, torch.float16)
    pk = det_normal((b, H_kv, past, D), seed + 3, 1.0, torch.float16)
    pv = det_normal((b, H_kv, past, D), seed + 4, 1.0, torch.float16)
    pos = torch.full((b, 1), past, dtype=torch.int64)
    ref_out, ref_scores, _ = O.attn_forward_core(q, k, v, pk, pv, pos, None)
    out, logits, kc, vc, imp, himp = run_kernel(q, k, v, pk, pv, pos, None, want_importance=True)
    assert_logits_close(logits, ref_scores[:, :, 0, :])
    assert rel_err(out, ref_out[:, :, 0, :]) <= OUT_REL_TOL
    # cumulative importance increment == softmax probabilities of the exported logits (unpinned restatement)
    p = torch.softmax(logits.float(), dim=-1)
    assert torch.allclose(imp.cpu(), p, rtol=2e-3, atol=1e-6)
    hi = out.float().abs().sum(dim=(0, 2))
    assert torch.allclose(himp.cpu(), hi, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("b,H,D,past", [(1, 32, 128, 2047), (2, 16, 128, 2047), (8, 4, 128, 1500), (1, 64, 64, 777),
                                        (1, 128, 128, 600), (1, 30, 128, 2047)])
def test_cluster_and_ticket_variants_agree(b, H, D, past):
    """The DSMEM-merge (thread-block cluster) variant and the global-ticket variant are the same function."""
    from spatten_b200 import _lib
    seed = 7000 + H + past
    q = det_normal((b, H, 1, D), seed, 1.0, torch.float16)
    k = det_normal((b, H, 1, D), seed + 1, 1.0, torch.float16)
    v = det_normal((b, H, 1, D), seed + 2, 1.0, torch.float16)
    pk = det_normal((b, H, past, D), seed + 3, 1.0, torch.float16)
    pv = det_normal((b, H, past, D), seed + 4, 1.0, torch.float16)
    pos = torch.full((b, 1), past, dtype=torch.int64)
    ref_out, ref_scores, _ = O.attn_forward_core(q, k, v, pk, pv, pos, None)
    res = {}
    for name, fl in (("cluster", 0), ("ticket", _lib.DECODE_NO_CLUSTER), ("cluster_small", _lib.DECODE_SMALL_TILES),
                     ("ticket_small", _lib.DECODE_NO_CLUSTER | _lib.DECODE_SMALL_TILES)):
        out, logits, kc, vc, imp, himp = run_kernel(q, k, v, pk, pv, pos, None, want_importance=True, flags=fl)
        assert_logits_close(logits, ref_scores[:, :, 0, :], name)
        assert rel_err(out, ref_out[:, :, 0, :]) <= OUT_REL_TOL, name
        assert torch.equal(kc[:, :, past], k[:, :, 0]) and torch.equal(vc[:, :, past], v[:, :, 0])
        res[name] = (out, logits, imp.cpu(), himp.cpu())
    for other in ("ticket", "cluster_small", "ticket_small"):
        assert torch.equal(res["cluster"][1], res[other][1]), other                # logits: bit-identical
        assert rel_err(res["cluster"][0], res[other][0]) <= 1e-3, other
        assert torch.allclose(res["cluster"][2], res[other][2], rtol=1e-3, atol=1e-7), other
        assert torch.allclose(res["cluster"][3], res[other][3], rtol=1e-3, atol=1e-3), other


def test_decode_head_mask_and_noappend():
    from spatten_b200 import ops
    b, H, D, n = 1, 8, 128, 300
    q = det_normal((b, H, D), 1, 1.0, torch.float16).cuda()
    kc = det_normal((b, H, n, D), 2, 1.0, torch.float16).cuda()
    vc = det_normal((b, H, n, D), 3, 1.0, torch.float16).cuda()
    rope = ops.rope_table(D, n + 1, "cuda")
    hm = torch.tensor([1, 0, 1, 1, 0, 1, 1, 1], dtype=torch.uint8, device="cuda")
    pos = torch.full((b,), n - 1, dtype=torch.int64, device="cuda")
    out = ops.decode_attn(q, kc, vc, n, rope, q_pos=pos, head_mask=hm)
    # oracle: no append -> feed the last cached token as the "new" one
    ref, _, _ = O.attn_forward_core(q.cpu()[:, :, None], kc.cpu()[:, :, -1:], vc.cpu()[:, :, -1:],
                                    kc.cpu()[:, :, :-1], vc.cpu()[:, :, :-1], pos.cpu()[:, None], None)
    ref = ref[:, :, 0]
    live = hm.bool().cpu()
    assert rel_err(out.cpu()[:, live], ref[:, live]) <= OUT_REL_TOL
    assert (out.cpu()[:, ~live] == 0).all()


def test_decode_repeated_launches_reuse_workspace():
    """The split-KV tickets reset themselves: back-to-back launches on one workspace stay correct."""
    from spatten_b200 import ops
    b, H, D, n = 1, 32, 128, 2048
    q = det_normal((b, H, D), 11, 1.0, torch.float16).cuda()
    kc = det_normal((b, H, n, D), 12, 1.0, torch.float16).cuda()
    vc = det_normal((b, H, n, D), 13, 1.0, torch.float16).cuda()
    rope = ops.rope_table(D, n, "cuda")
    outs = [ops.decode_attn(q, kc, vc, n, rope).clone() for _ in range(5)]
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


def test_decode_bad_args_raise():
    from spatten_b200 import _lib, ops
    q = torch.zeros(1, 4, 96, dtype=torch.float16, device="cuda")
    kc = torch.zeros(1, 4, 16, 96, dtype=torch.float16, device="cuda")
    rope = torch.zeros(16, 96, dtype=torch.float16, device="cuda")
    with pytest.raises(_lib.SpattenError, match="UNSUPPORTED"):
        ops.decode_attn(q, kc, kc, 8, rope)
    q = torch.zeros(1, 4, 128, dtype=torch.float16, device="cuda")
    kc = torch.zeros(1, 4, 16, 128, dtype=torch.float16, device="cuda")
    rope = torch.zeros(4, 128, dtype=torch.float16, device="cuda")
    with pytest.raises(_lib.SpattenError, match="rope_len"):
        ops.decode_attn(q, kc, kc, 8, rope)
    with pytest.raises(RuntimeError, match="CUDA device"):
        ops.decode_attn(q.cpu(), kc, kc, 8, rope)
