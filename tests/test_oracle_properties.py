"""Property tests (hypothesis) of the oracle's selection rule and of the unpinned restatements' self-consistency."""
import numpy as np
import pytest
import torch

pytest.importorskip("hypothesis")      # optional test dependency: skip this module instead of a collection error
from hypothesis import given, settings, strategies as st  # noqa: E402

from oracle import spatten_oracle as O


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 6), st.integers(20, 300), st.integers(0, 5), st.integers(1, 40), st.integers(0, 40), st.integers(0, 12),
       st.integers(0, 2 ** 31 - 1))
def test_selection_equals_reference_topk_on_tie_free_scores(H, S, start, important, recent, num_coming, seed):
    """On tie-free scores the oracle's rule is exactly the reference's topk(...).indices.sort() + start (:59-63)."""
    window_end = min(S, S - recent + num_coming)
    if S + num_coming <= start + important + recent or window_end - start < important:
        return
    rng = np.random.default_rng(seed)
    score = torch.from_numpy(np.stack([rng.permutation(S) for _ in range(H)]).astype(np.float32))   # all distinct
    mine = O.select_keep_indices(score, S, start, important, recent, num_coming)
    cand = score[:, start:S - recent + num_coming]
    ref = torch.topk(cand, important, dim=-1).indices.sort().values + start
    assert torch.equal(mine, ref)
    assert (mine[:, 1:] > mine[:, :-1]).all() and mine.min() >= start and mine.max() < window_end


@settings(max_examples=40, deadline=None)
@given(st.integers(2, 200), st.integers(1, 50), st.integers(0, 2 ** 31 - 1))
def test_tie_rule_keeps_score_multiset_and_lowest_indices(S, k, seed):
    if k >= S:
        return
    rng = np.random.default_rng(seed)
    score = torch.from_numpy(rng.integers(0, 4, size=(1, S)).astype(np.float32))     # heavy ties
    keep = O.select_keep_indices(score, S, 0, k, 0, 0)[0]
    ref = torch.topk(score[0], k).values.sort().values                               # same multiset of scores as torch.topk
    assert torch.equal(score[0, keep].sort().values, ref)
    kth = ref.min()
    eq_kept = [i for i in keep.tolist() if score[0, i] == kth]
    eq_all = [i for i in range(S) if score[0, i] == kth]
    assert eq_kept == eq_all[:len(eq_kept)]                                          # lowest-index-first among ties


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 3), st.integers(1, 4), st.integers(8, 80), st.integers(0, 2 ** 31 - 1))
def test_local_value_attention_with_full_keep_is_plain_attention(b, H, S, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn((b, H, 1, 16), generator=g)
    k = torch.randn((b, H, S, 16), generator=g)
    v = torch.randn((b, H, S, 16), generator=g)
    out, logits, idx = O.local_value_attention(q, k, v, S)
    p = torch.softmax(logits.float(), dim=-1)
    assert torch.allclose(out, torch.matmul(p, v), atol=1e-5)
    out2, _, idx2 = O.local_value_attention(q, k, v, max(1, S // 2))
    assert idx2.shape[-1] == max(1, S // 2) and (idx2[..., 1:] > idx2[..., :-1]).all()


def test_quant_planes_nibble_order_and_requant_monotonic():
    x = torch.arange(-64, 64, dtype=torch.float32).view(1, 1, 1, 128) / 8
    msb, lsb, sc = O.quantize_kv_planes(x.half())
    ints = O.dequant_ints(msb, lsb)
    assert torch.equal(ints.float() * sc.float()[..., None], O.dequant_planes(msb, lsb, sc))
    # nibble order [d0,d2,d4,d6,d1,d3,d5,d7]: byte 0 = (d0 | d2 << 4), byte 2 = (d1 | d3 << 4)
    u = (torch.clamp(torch.round(x / sc.float()), -127, 127) + 128).to(torch.int32)[0, 0, 0]
    assert int(msb[0, 0, 0, 0]) == int((u[0] >> 4) | ((u[2] >> 4) << 4))
    assert int(msb[0, 0, 0, 2]) == int((u[1] >> 4) | ((u[3] >> 4) << 4))
    assert (ints[0, 0, 0, 1:] >= ints[0, 0, 0, :-1]).all()                            # monotone code
    assert bool(O.requant_decision(torch.tensor(0.01), 0.05)) and not bool(O.requant_decision(torch.tensor(0.5), 0.05))
