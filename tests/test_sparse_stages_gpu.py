"""GPU checks of the RTL-ordered stages (parity unpinned: compared with the self-consistent oracle restatements)."""
import pytest
import torch

from detgen import det_normal
from helpers import OUT_REL_TOL, assert_logits_close, rel_err
from oracle import spatten_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [True, False], ids=["one_launch", "three_launches"])
@pytest.mark.parametrize("b,H,D,n,keep", [(1, 8, 128, 600, 150), (2, 4, 128, 2049, 512), (1, 12, 64, 333, 100),
                                          (1, 4, 128, 64, 64), (1, 32, 128, 4096, 1024), (1, 32, 128, 4096, 2048),
                                          (1, 2, 128, 20000, 3000), (1, 160, 64, 1024, 200), (1, 40, 128, 17, 5)])
def test_local_v_attention_matches_oracle(b, H, D, n, keep, fused):
    from spatten_b200 import ops
    q = det_normal((b, H, 1, D), 3000 + n, 1.0, torch.float16)
    kc = det_normal((b, H, n, D), 3001 + n, 1.0, torch.float16)
    vc = det_normal((b, H, n, D), 3002 + n, 1.0, torch.float16)
    # oracle: rotate q and K with the reference's fp16 RoPE, then score / select / multiply survivors
    cos, sin = O.rotary_tables(D, n, torch.float16)
    pos = torch.full((b, 1), n - 1, dtype=torch.int64)
    q_rot = O.rope_single(q, cos, sin, pos)
    k_rot = O.rope_single(kc, cos, sin, torch.arange(n).unsqueeze(0))
    ref_out, ref_logits, ref_idx = O.local_value_attention(q_rot, k_rot, vc, keep)

    out, logits, idx = ops.local_v_attn(q[:, :, 0].cuda(), kc.cuda(), vc.cuda(), n, ops.rope_table(D, n, "cuda"), keep,
                                        q_pos=pos[:, 0].cuda(), fused=fused)
    torch.cuda.synchronize()
    assert_logits_close(logits.cpu(), ref_logits[:, :, 0, :])
    # the survivor set is chosen on each side's own logits: compare through the oracle selection on OUR logits
    mine_sel = O.local_value_attention(q_rot, k_rot, vc, keep)[2] if False else None
    order = torch.sort(logits.cpu().float(), dim=-1, descending=True, stable=True).indices
    exp_idx = order[..., :keep].sort(dim=-1).values if keep < n else torch.arange(n).expand(b, H, n)
    assert torch.equal(idx.cpu().long().view(b, H, -1), exp_idx)               # bit-exact top-k on identical scores
    same = (idx.cpu().long().view(b, H, -1) == ref_idx).float().mean().item()
    assert same > 0.97                                                          # and nearly always the oracle's set
    # output: recompute the oracle's multiply on OUR survivor set so only arithmetic is compared
    sel = torch.gather(logits.cpu().float(), 2, exp_idx)
    p = torch.softmax(sel, dim=-1)
    vsel = torch.gather(vc, 2, exp_idx[..., None].expand(b, H, exp_idx.shape[-1], D)).float()
    ref2 = torch.einsum("bhk,bhkd->bhd", p, vsel)
    assert rel_err(out.cpu(), ref2) <= OUT_REL_TOL
    if same == 1.0:
        assert rel_err(out.cpu(), ref_out[:, :, 0]) <= 2e-3


def test_head_select_and_masked_decode():
    from spatten_b200 import ops
    H = 16
    imp = torch.tensor([3., 1., 4., 1., 5., 9., 2., 6., 5., 3., 5., 8., 9., 7., 9., 3.])
    for keep in (0, 1, 5, 12, 16):
        mask = ops.head_select(imp.cuda(), keep).cpu()
        assert torch.equal(mask, O.head_prune_mask(imp, keep)), keep


def test_head_pruner_accumulates_and_masks():
    from spatten_b200.head_pruning import SpAttenHeadPruner
    from spatten_b200.modeling import AttentionStack
    from spatten_b200.enable_spatten_llm import enable_spatten_llm
    import contextlib, io
    H, hid = 8, 1024
    model = AttentionStack(2, hid, H, device="cuda", seed=3)
    with contextlib.redirect_stdout(io.StringIO()):
        enable_spatten_llm(model, 4, 16, 16)
    pruner = SpAttenHeadPruner(model, keep_heads=6)
    past = [None, None]
    outs = []
    with torch.no_grad():
        for step in range(4):
            hs = det_normal((1, 1, hid), 50 + step, 1.0, torch.float16).cuda()
            for l, m in enumerate(model.layers):
                o, _, past[l] = m(hs, past_key_value=past[l], use_cache=True)
            outs.append(o)
    imp0 = model.layers[0].spatten_head_importance.cpu()
    assert (imp0 > 0).all()
    pruner.update_masks()
    mask = model.layers[0].spatten_head_mask.cpu()
    assert int(mask.sum()) == 6 and torch.equal(mask, O.head_prune_mask(imp0, 6))
    # a masked head contributes nothing: its slice of the pre-o_proj output is zero
    from spatten_b200 import ops
    m = model.layers[0]
    k, v = past[0]
    q = det_normal((1, H, 128), 77, 1.0, torch.float16).cuda()
    rope = ops.rope_table(128, k.shape[2] + 1, "cuda")
    out = ops.decode_attn(q, k, v, k.shape[2], rope, head_mask=m.spatten_head_mask)
    assert (out.cpu()[0, ~mask.bool()] == 0).all() and (out.cpu()[0, mask.bool()].abs().sum(-1) > 0).all()
