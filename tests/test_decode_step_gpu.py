"""GPU parity of the persistent decode-step kernel (spatten_decode_step, one launch for all layers) against the
CPU oracle and against the per-layer kernel (spatten_decode_attn), which the reference goldens pin."""
import pytest
import torch

from detgen import det_normal
from helpers import OUT_REL_TOL, assert_logits_close, rel_err
from oracle import spatten_oracle as O

pytestmark = pytest.mark.gpu


def make_problem(L, b, H, H_kv, D, past, seed, append=True, cap_pad=5):
    dev = "cuda"
    q = det_normal((L, b, H, D), seed, 1.0, torch.float16)
    kn = det_normal((L, b, H_kv, D), seed + 1, 1.0, torch.float16)
    vn = det_normal((L, b, H_kv, D), seed + 2, 1.0, torch.float16)
    pk = det_normal((L, b, H_kv, past, D), seed + 3, 1.0, torch.float16)
    pv = det_normal((L, b, H_kv, past, D), seed + 4, 1.0, torch.float16)
    cap = past + 1 + cap_pad
    kc, vc = [], []
    for l in range(L):
        k = torch.full((b, H_kv, cap, D), float("nan"), dtype=torch.float16, device=dev)
        v = torch.full((b, H_kv, cap, D), float("nan"), dtype=torch.float16, device=dev)
        k[:, :, :past], v[:, :, :past] = pk[l].to(dev), pv[l].to(dev)
        kc.append(k)
        vc.append(v)
    return dict(q=q, kn=kn, vn=vn, pk=pk, pv=pv, kc=kc, vc=vc, cap=cap)


def oracle_layers(P, past, append=True):
    outs, scores = [], []
    L = P["q"].shape[0]
    b = P["q"].shape[1]
    for l in range(L):
        if append:
            pos = torch.full((b, 1), past, dtype=torch.int64)
            o, s, _ = O.attn_forward_core(P["q"][l][:, :, None], P["kn"][l][:, :, None], P["vn"][l][:, :, None],
                                          P["pk"][l], P["pv"][l], pos, None)
        else:
            pos = torch.full((b, 1), past - 1, dtype=torch.int64)
            o, s, _ = O.attn_forward_core(P["q"][l][:, :, None], P["pk"][l][:, :, -1:], P["pv"][l][:, :, -1:],
                                          P["pk"][l][:, :, :-1], P["pv"][l][:, :, :-1], pos, None)
        outs.append(o[:, :, 0])
        scores.append(s[:, :, 0])
    return torch.stack(outs), torch.stack(scores)


@pytest.mark.parametrize("L,b,H,H_kv,D,past", [
    (3, 1, 32, 32, 128, 2047),     # BASELINE config 2 shape, three layers
    (2, 2, 40, 40, 128, 777),      # 13B heads
    (2, 8, 5, 5, 128, 3000),       # config-5-like shard
    (4, 1, 4, 4, 128, 1),
    (2, 1, 1, 1, 128, 9),
    (3, 3, 3, 3, 64, 130),
    (2, 1, 12, 12, 64, 64),
    (2, 1, 7, 7, 128, 15),
    (2, 1, 32, 8, 128, 1000),      # GQA
    (1, 1, 2, 2, 128, 40000),      # more tiles per row than warps in the grid can hold as single tiles
    (1, 8, 5, 5, 128, 16383),      # BASELINE config 5, one GPU's shard at full size (13B: 40 heads / 8 GPUs)
    (1, 1, 40, 40, 128, 2047),     # 13B head count, config-2 length
])
@pytest.mark.parametrize("chain", [True, False])
def test_step_matches_oracle_and_per_layer_kernel(L, b, H, H_kv, D, past, chain):
    from spatten_b200 import ops
    if past >= 16000 and not chain:
        pytest.skip("large case runs once")
    seed = 12000 + L * 7 + b * 131 + H * 17 + past
    P = make_problem(L, b, H, H_kv, D, past, seed)
    dev = "cuda"
    rope = ops.rope_table(D, past + 1, dev)
    n = past + 1
    logits = torch.full((L, b, H, n + 3), float("nan"), dtype=torch.float16, device=dev)
    imp = torch.zeros((L, b, H, n), dtype=torch.float32, device=dev)
    himp = torch.zeros(H, dtype=torch.float32, device=dev)
    step = ops.DecodeStep(P["q"].to(dev), P["kc"], P["vc"], past, rope, k_new=P["kn"].to(dev), v_new=P["vn"].to(dev),
                          logits=logits, importance=imp, head_importance=himp, chain=chain)
    out = step.launch()
    torch.cuda.synchronize()
    assert not step.timed_out() and step.steps_completed() == 1
    if past <= 4096 or (b, H) == (8, 5):
        ref_out, ref_scores = oracle_layers(P, past)
        assert_logits_close(logits[..., :n].cpu(), ref_scores)
        assert rel_err(out.cpu(), ref_out) <= OUT_REL_TOL
    # the per-layer kernel on fresh copies of the same caches: bit-identical logits, same output
    for l in range(L):
        kc2 = P["kc"][l].clone()
        vc2 = P["vc"][l].clone()
        kc2[:, :, past:] = float("nan")
        vc2[:, :, past:] = float("nan")
        lg2 = torch.full((b, H, n), float("nan"), dtype=torch.float16, device=dev)
        out2 = ops.decode_attn(P["q"][l].to(dev), kc2, vc2, past, rope, k_new=P["kn"][l].to(dev), v_new=P["vn"][l].to(dev),
                               logits=lg2)
        assert torch.equal(lg2, logits[l][..., :n]), f"layer {l}: logits differ from spatten_decode_attn"
        assert rel_err(out[l], out2) <= OUT_REL_TOL
        # cache append: slot `past` holds the new un-rotated rows, nothing else was touched
        assert torch.equal(P["kc"][l][:, :, past].cpu(), P["kn"][l]) and torch.equal(P["vc"][l][:, :, past].cpu(), P["vn"][l])
        assert torch.equal(P["kc"][l][:, :, :past].cpu(), P["pk"][l])
        assert torch.isnan(P["kc"][l][:, :, past + 1:]).all() and torch.isnan(P["vc"][l][:, :, past + 1:]).all()
    assert torch.isnan(logits[..., n:]).all()
    p = torch.softmax(logits[..., :n].float(), dim=-1)
    assert torch.allclose(imp, p, rtol=2e-3, atol=1e-6)
    assert torch.allclose(himp.cpu(), out.float().abs().sum(dim=(0, 1, 3)).cpu(), rtol=1e-3, atol=1e-3)


def test_step_no_append_and_repeated_graph_replays():
    from spatten_b200 import ops
    L, b, H, D, n = 4, 1, 32, 128, 2048
    dev = "cuda"
    P = make_problem(L, b, H, H, D, n, 555, cap_pad=0)
    rope = ops.rope_table(D, n, dev)
    pos = torch.full((b,), n - 1, dtype=torch.int64, device=dev)
    kc = [k[:, :, :n].contiguous() for k in P["kc"]]
    vc = [v[:, :, :n].contiguous() for v in P["vc"]]
    step = ops.DecodeStep(P["q"].to(dev), kc, vc, n, rope, q_pos=pos, chain=True)
    ref_out, _ = oracle_layers(P, n, append=False)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        first = step.launch().clone()
        s.synchronize()
        assert rel_err(first.cpu(), ref_out) <= OUT_REL_TOL
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            step.launch()
        for _ in range(6):
            step.out.zero_()
            g.replay()
            s.synchronize()
            assert torch.equal(step.out, first)
    assert step.steps_completed() == 7 and not step.timed_out()      # 1 eager + 6 replays (capture does not run)


def test_step_head_mask_gqa_append_two_steps():
    """Pruned heads (also the first head of a GQA group, also a whole group) never leave holes in the KV cache:
    two decode steps with a per-layer head mask, every live head checked against the oracle on both steps."""
    from spatten_b200 import ops
    L, b, H, H_kv, D, past = 2, 2, 8, 4, 128, 300
    dev = "cuda"
    P = make_problem(L, b, H, H_kv, D, past, 4242, cap_pad=8)
    rope = ops.rope_table(D, past + 2, dev)
    hm = torch.tensor([[0, 1, 1, 1, 0, 0, 1, 0],      # group 0 loses its first head, group 2 is pruned entirely
                       [1, 1, 0, 0, 1, 0, 0, 1]], dtype=torch.uint8, device=dev)
    q2 = det_normal((L, b, H, D), 91, 1.0, torch.float16)
    kn2 = det_normal((L, b, H_kv, D), 92, 1.0, torch.float16)
    vn2 = det_normal((L, b, H_kv, D), 93, 1.0, torch.float16)
    out1 = ops.DecodeStep(P["q"].to(dev), P["kc"], P["vc"], past, rope, k_new=P["kn"].to(dev), v_new=P["vn"].to(dev),
                          head_mask=hm).launch().clone()
    out2 = ops.DecodeStep(q2.to(dev), P["kc"], P["vc"], past + 1, rope, k_new=kn2.to(dev), v_new=vn2.to(dev),
                          head_mask=hm).launch().clone()
    torch.cuda.synchronize()
    for l in range(L):
        # both appended rows are in the cache for EVERY kv head
        assert torch.equal(P["kc"][l][:, :, past].cpu(), P["kn"][l]) and torch.equal(P["vc"][l][:, :, past].cpu(), P["vn"][l])
        assert torch.equal(P["kc"][l][:, :, past + 1].cpu(), kn2[l]) and torch.equal(P["vc"][l][:, :, past + 1].cpu(), vn2[l])
        live = hm[l].bool().cpu()
        pos = torch.full((b, 1), past, dtype=torch.int64)
        r1, _, (k1, v1) = O.attn_forward_core(P["q"][l][:, :, None], P["kn"][l][:, :, None], P["vn"][l][:, :, None],
                                              P["pk"][l], P["pv"][l], pos, None)
        r2, _, _ = O.attn_forward_core(q2[l][:, :, None], kn2[l][:, :, None], vn2[l][:, :, None], k1, v1, pos + 1, None)
        assert rel_err(out1[l].cpu()[:, live], r1[:, live, 0]) <= OUT_REL_TOL
        assert rel_err(out2[l].cpu()[:, live], r2[:, live, 0]) <= OUT_REL_TOL
        assert (out1[l].cpu()[:, ~live] == 0).all() and (out2[l].cpu()[:, ~live] == 0).all()


def test_step_bad_args_raise():
    from spatten_b200 import _lib, ops
    dev = "cuda"
    q = torch.zeros(2, 1, 4, 128, dtype=torch.float16, device=dev)
    kc = [torch.zeros(1, 4, 16, 128, dtype=torch.float16, device=dev) for _ in range(2)]
    rope = torch.zeros(4, 128, dtype=torch.float16, device=dev)
    with pytest.raises(_lib.SpattenError, match="rope_len"):
        ops.DecodeStep(q, kc, kc, 8, rope).launch()
    with pytest.raises(ValueError, match="one K and one V cache per layer"):
        ops.DecodeStep(q, kc[:1], kc[:1], 8, rope)
