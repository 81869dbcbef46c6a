"""Pin the oracle (oracle/spatten_oracle.py) to outputs of the unmodified reference (tests/golden)."""
import pytest
import torch

from cases import FWD_CASES, PRUNE_CASES, build_fwd_inputs, build_prune_inputs
from helpers import OUT_REL_TOL, assert_index_sets_match, assert_logits_close, load_golden, rel_err
from oracle import spatten_oracle as O


@pytest.mark.parametrize("c", FWD_CASES, ids=lambda c: c.name)
def test_forward_matches_reference(c):
    d, g = build_fwd_inputs(c), load_golden("fwd", c.name)
    if c.full:
        out, scores, cache = O.attn_forward(d["hidden"], d["wq"], d["wk"], d["wv"], d["wo"], c.H, c.H_kv,
                                            d["past_k"], d["past_v"], d["position_ids"], d["mask"])
        ref_out = g["out"]
    else:
        out, scores, cache = O.attn_forward_core(d["q"], d["k_new"], d["v_new"], d["past_k"], d["past_v"],
                                                 d["position_ids"], d["mask"])
        # reference returns [b, q, H*D] through an identity o_proj
        ref_out = g["out"].view(c.b, c.q_len, c.H, c.D).transpose(1, 2)
    assert scores.shape == g["attn_scores"].shape
    assert_logits_close(scores, g["attn_scores"], c.name)
    assert rel_err(out, ref_out) <= OUT_REL_TOL
    assert int(g["cache_len"]) == cache[0].shape[2] == c.past + c.q_len
    # cache holds UN-rotated keys (modify_llama.py:100): byte-exact
    assert torch.equal(cache[0][:, :, -1, :], g["cache_k_last"])
    assert torch.equal(cache[1][:, :, -1, :], g["cache_v_last"])


@pytest.mark.parametrize("c", PRUNE_CASES, ids=lambda c: c.name)
def test_prune_matches_reference(c):
    d, g = build_prune_inputs(c), load_golden("prune", c.name)
    new_past, keeps = O.token_prune(d["past"], c.num_coming, d["scores"], c.start, c.important, c.recent)
    if bool(g["identity"]):
        assert new_past is d["past"]
        return
    tail0 = c.S - c.recent + c.num_coming
    for l in range(c.L):
        k2, v2 = new_past[l]
        assert list(k2.shape) == g[f"shape_{l}"].tolist()
        assert k2.shape[2] == c.start + c.important + max(0, c.recent - c.num_coming)
        ref_idx = g[f"kept_idx_{l}"]                       # whole kept list incl. start / recent windows
        mid = ref_idx[:, c.start:c.start + c.important]
        score = O.importance_from_scores(d["scores"][l])
        assert_index_sets_match(keeps[l], mid, score, f"{c.name} layer {l}")
        # windows untouched
        assert torch.equal(ref_idx[:, :c.start].long(), torch.arange(c.start).expand(c.H, c.start))
        assert torch.equal(ref_idx[:, c.start + c.important:].long(),
                           torch.arange(tail0, c.S).expand(c.H, max(0, c.S - tail0)))
        if torch.equal(keeps[l].int(), mid.int()):        # byte-exact gather when no boundary tie
            if f"k_{l}" in g:
                assert torch.equal(k2, g[f"k_{l}"]) and torch.equal(v2, g[f"v_{l}"])
            w = torch.arange(1, k2.shape[2] + 1, dtype=torch.float64)[None, None, :, None]
            assert torch.equal((k2.double() * w).sum(dim=(0, 2)), g[f"k_wsum_{l}"])
            assert torch.equal((v2.double() * w).sum(dim=(0, 2)), g[f"v_wsum_{l}"])


def test_prune_edge_cases():
    assert O.token_prune(None, 3, [], 4, 8, 8) == (None, None)
    k = torch.zeros(2, 2, 40, 8)
    with pytest.raises(IndexError):          # reference: squeeze()[mask] with b=2 (kv_cache_token_pruning.py:68)
        O.token_prune([(k, k)], 0, [torch.zeros(2, 2, 1, 40)], 2, 4, 4)
    k1 = torch.zeros(1, 2, 40, 8)
    with pytest.raises(TypeError):           # important_size = 0 (:63)
        O.token_prune([(k1, k1)], 0, [torch.zeros(1, 2, 1, 40)], 2, 0, 4)
    with pytest.raises(RuntimeError):        # num_coming > recent clamps the window below `important` (torch.topk :60)
        O.token_prune([(k1, k1)], 20, [torch.zeros(1, 2, 1, 40)], 2, 45, 5)


def test_tie_rule_lowest_index_first():
    s = torch.tensor([[5., 1., 1., 1., 1., 0., 9., 9.]]).half()
    keep = O.select_keep_indices(s, seq_len=8, start=0, important=3, recent=2, num_coming=0)
    assert keep.tolist() == [[0, 1, 2]]       # 5 > kth(=1); then first two of the 1s by index


def test_rope_tables_are_hf433():
    cos, sin = O.rotary_tables(8, 5, torch.float32)
    inv = 1.0 / (10000 ** (torch.arange(0, 8, 2).float() / 8))
    assert torch.equal(cos[0, 0, 3, :4], torch.cos(3 * inv)) and torch.equal(cos[0, 0, :, :4], cos[0, 0, :, 4:])
    assert torch.equal(sin[0, 0, 2, 4:], torch.sin(2 * inv))


def test_quant_planes_roundtrip():
    from detgen import det_normal
    x = det_normal((2, 3, 17, 64), 7).half()
    msb, lsb, sc = O.quantize_kv_planes(x)
    full = O.dequant_planes(msb, lsb, sc)
    coarse = O.dequant_planes(msb, None, sc)
    assert (full - x.float()).abs().max() <= 0.51 * sc.float().max()
    assert (coarse - x.float()).abs().max() <= 8.51 * sc.float().max()
    assert msb.shape == (2, 3, 17, 32) and msb.dtype == torch.uint8


from cases import CHAIN_CASES, build_chain_inputs          # noqa: E402
from helpers import assert_chain_sets_match                # noqa: E402


@pytest.mark.parametrize("c", [c for c in CHAIN_CASES if c.past < 1000], ids=lambda c: c.name)
def test_oracle_decode_then_prune_chain_matches_reference(c):
    """The oracle's forward feeding the oracle's selection reproduces the reference's chained keep set."""
    d, g = build_chain_inputs(c), load_golden("chain", c.name)
    out, scores, (k, v) = O.attn_forward_core(d["q"], d["k_new"], d["v_new"], d["past_k"], d["past_v"],
                                              d["position_ids"], d["mask"])
    assert torch.equal(scores, g["attn_scores"])           # same library, same thread count: bit-identical logits
    keep = O.select_keep_indices(O.importance_from_scores(scores), c.past + 1, c.start, c.important, c.recent, c.num_coming)
    tail0 = c.past + 1 - c.recent + c.num_coming
    full = torch.cat([torch.arange(c.start).expand(c.H, -1), keep,
                      torch.arange(tail0, c.past + 1).expand(c.H, -1)], dim=1)
    rep = assert_chain_sets_match(full, g["kept_idx"], g["attn_scores"][0, :, 0], g["kth"], g["next"], c.name, slack_ulps=0)
    assert rep["identical"] + rep["boundary_flips_heads"] == c.H


def test_single_multiply_reproduces_the_fp16_logit_division():
    """csrc/prefill_fused.cu forms fp16(x / sqrt(128)) (modify_llama.py:111-113, fp32 opmath) as fp16(x * r) with
    r = nextafter(fl32(1/sqrt(128)), 1): exhaustive over every finite fp16 x.  The plain reciprocal misses 52 values."""
    import numpy as np
    h = np.arange(65536, dtype=np.uint32).astype(np.uint16).view(np.float16)
    h = h[np.isfinite(h)]
    c = np.float32(np.sqrt(np.float32(128.0)))
    ref = (h.astype(np.float32) / c).astype(np.float16)
    r0 = np.float32(1.0) / c
    r_up = np.nextafter(r0, np.float32(1.0), dtype=np.float32)
    assert ((h.astype(np.float32) * r_up).astype(np.float16).view(np.uint16) == ref.view(np.uint16)).all()
    assert ((h.astype(np.float32) * r0).astype(np.float16).view(np.uint16) != ref.view(np.uint16)).sum() == 52


def test_quant_plane_layout_chunks_hold_rope_partners():
    """The int4 plane row of a 128-dim token (spatten_kv_quantize, oracle.quantize_kv_planes): 16-byte chunk kq carries the
    MSB nibbles of dims [16kq, 16kq+16) followed by their RoPE partners [64+16kq, 64+16kq+16), 8 dims per little-endian u32
    in the nibble order [d0,d2,d4,d6,d1,d3,d5,d7] -- the layout one lane of the tensor-core kernel fetches with one load."""
    import numpy as np
    x = torch.zeros(1, 128, dtype=torch.float16)
    x[0] = torch.arange(128, dtype=torch.float16) - 64.0            # distinct codes: q = round(x / scale), scale = 64 / 127
    msb, lsb, sc = O.quantize_kv_planes(x)
    codes = O.dequant_ints(msb, lsb)[0].numpy()                     # int8 codes in dim order
    assert (np.diff(codes) > 0).all()                               # round trip through pack / unpack keeps the dim order
    u = (codes + 128).astype(np.int64)
    words = msb[0].numpy().view("<u4")                              # 16 words of 8 nibbles
    for kq in range(4):
        for j, base in enumerate((16 * kq, 16 * kq + 8, 64 + 16 * kq, 64 + 16 * kq + 8)):
            w = int(words[4 * kq + j])
            nib = [(w >> (4 * p)) & 15 for p in range(8)]
            dims = [base + d for d in (0, 2, 4, 6, 1, 3, 5, 7)]
            assert nib == [int(u[d]) >> 4 for d in dims], (kq, j)
