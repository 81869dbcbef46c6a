"""GPU parity: token pruning (score reduce -> top-k select -> KV gather) vs reference goldens and oracle."""
import pytest
import torch

from cases import PRUNE_CASES, build_prune_inputs
from detgen import det_normal
from helpers import assert_index_sets_match, load_golden
from oracle import spatten_oracle as O

pytestmark = pytest.mark.gpu


def _cuda_past(d):
    return [(k.cuda(), v.cuda()) for k, v in d["past"]], [s.cuda() for s in d["scores"]]


@pytest.mark.parametrize("c", PRUNE_CASES, ids=lambda c: c.name)
def test_apply_token_pruning_matches_reference(c):
    from spatten_b200.kv_cache_token_pruning import SpAttenKVCache
    d, g = build_prune_inputs(c), load_golden("prune", c.name)
    if c.dtype == "fp32":
        pytest.skip("KV gather is dtype-agnostic but this case's cache is fp32; covered by test_select_fp32_scores")
    past, scores = _cuda_past(d)
    cache = SpAttenKVCache(start_size=c.start, recent_size=c.recent, important_size=c.important)
    new_past = cache.apply_token_pruning(past, c.num_coming, scores)
    torch.cuda.synchronize()
    if bool(g["identity"]):
        assert new_past is past
        return
    for l in range(c.L):
        k2, v2 = new_past[l]
        assert list(k2.shape) == g[f"shape_{l}"].tolist()
        ref_idx = g[f"kept_idx_{l}"]
        mine_idx = cache.keep_indices[l].cpu()
        score = O.importance_from_scores(d["scores"][l])
        assert_index_sets_match(mine_idx, ref_idx, score, f"{c.name} layer {l}")
        # the gathered bytes are exactly the rows named by the index list
        H, D = c.H, c.D
        gi = mine_idx.long()[None, :, :, None].expand(1, H, mine_idx.shape[1], D)
        assert torch.equal(k2.cpu(), torch.gather(d["past"][l][0], 2, gi))
        assert torch.equal(v2.cpu(), torch.gather(d["past"][l][1], 2, gi))
        if torch.equal(mine_idx.int(), ref_idx.int()):          # no boundary tie: byte-exact vs the reference
            if f"k_{l}" in g:
                assert torch.equal(k2.cpu(), g[f"k_{l}"]) and torch.equal(v2.cpu(), g[f"v_{l}"])
            w = torch.arange(1, k2.shape[2] + 1, dtype=torch.float64)[None, None, :, None]
            assert torch.equal((k2.cpu().double() * w).sum(dim=(0, 2)), g[f"k_wsum_{l}"])
            assert torch.equal((v2.cpu().double() * w).sum(dim=(0, 2)), g[f"v_wsum_{l}"])
        # importance_score attribute as the reference leaves it (:51)
        assert torch.equal(cache.importance_score[l].cpu(), score)


@pytest.mark.parametrize("c", [c for c in PRUNE_CASES if c.dtype == "fp32"], ids=lambda c: c.name)
def test_select_fp32_scores(c):
    from spatten_b200 import ops
    d, g = build_prune_inputs(c), load_golden("prune", c.name)
    for l in range(c.L):
        red = ops.score_reduce(d["scores"][l].cuda())
        ref_score = O.importance_from_scores(d["scores"][l])
        assert torch.allclose(red.cpu(), ref_score, rtol=1e-6, atol=1e-6)
        # select on the reference's own reduced score: index set must be bit-exact (tie-aware)
        keep = ops.topk_select(ref_score.cuda(), c.S, c.num_coming, c.start, c.important, c.recent)
        assert_index_sets_match(keep.cpu(), g[f"kept_idx_{l}"], ref_score, c.name)
        k, v = d["past"][l]
        k2, v2 = ops.kv_gather(k.cuda(), v.cuda(), keep)        # fp32 rows: 4*D bytes, dtype-agnostic gather
        gi = keep.cpu().long()[None, :, :, None].expand(1, c.H, keep.shape[1], c.D)
        assert torch.equal(k2.cpu(), torch.gather(k, 2, gi)) and torch.equal(v2.cpu(), torch.gather(v, 2, gi))


@pytest.mark.parametrize("S,start,imp,recent,coming,H", [
    (4096, 4, 1022, 1022, 0, 32),        # BASELINE config 2: 4K -> 2K
    (8192, 4, 2046, 2046, 0, 32),
    (32768, 4, 8190, 8190, 0, 5),        # config-5-like row length
    (513, 0, 1, 1, 0, 3),
    (300, 0, 297, 0, 0, 2),              # keep almost everything, no recent window
    (1000, 7, 400, 300, 123, 9),
])
def test_select_properties_large(S, start, imp, recent, coming, H):
    """Size-independent properties at full sizes: count, ordering, windows, threshold, tie rule."""
    from spatten_b200 import ops
    score = det_normal((H, S), 31 + S + H, 2.0, torch.float16)
    keep = ops.topk_select(score.cuda(), S, coming, start, imp, recent).cpu().long()
    ref = O.select_keep_indices(score, S, start, imp, recent, coming)
    e = min(S, S - recent + coming)
    assert keep.shape[1] == start + imp + (S - e)
    assert (keep[:, 1:] > keep[:, :-1]).all()                                   # strictly ascending
    assert torch.equal(keep[:, :start], torch.arange(start).expand(H, start))
    assert torch.equal(keep[:, start + imp:], torch.arange(e, S).expand(H, S - e))
    assert torch.equal(keep[:, start:start + imp], ref)                         # lowest-index tie rule == oracle
    for h in range(H):
        mid = keep[h, start:start + imp]
        dropped = torch.ones(S, dtype=torch.bool)
        dropped[mid] = False
        dropped[:start] = False
        dropped[e:] = False
        if dropped.any():
            assert score[h, mid].float().min() >= score[h, dropped].float().max()


def test_select_heavy_ties_lowest_index_first():
    from spatten_b200 import ops
    S, H = 700, 4
    score = torch.zeros((H, S), dtype=torch.float16)
    score[1] = torch.tensor([1.0, 2.0] * (S // 2), dtype=torch.float16)
    score[2, 100:400] = -0.0                                                     # -0 == +0
    score[3] = det_normal((S,), 5, 1.0, torch.float16).round()                   # few distinct values
    keep = ops.topk_select(score.cuda(), S, 0, 2, 100, 50).cpu().long()
    ref = O.select_keep_indices(score, S, 2, 100, 50, 0)
    assert torch.equal(keep[:, 2:102], ref)
    assert keep[0, 2:102].tolist() == list(range(2, 102))


def test_gather_in_place_equals_out_of_place():
    from spatten_b200 import ops
    b, H, S, D = 2, 6, 1500, 128
    k = det_normal((b, H, S, D), 71, 1.0, torch.float16).cuda()
    v = det_normal((b, H, S, D), 72, 1.0, torch.float16).cuda()
    score = det_normal((H, S), 73, 1.0, torch.float16).cuda()
    keep = ops.topk_select(score, S, 0, 4, 350, 350)
    k2, v2 = ops.kv_gather(k, v, keep)
    kin, vin = k.clone(), v.clone()
    ops.kv_gather(kin, vin, keep, k_out=kin, v_out=vin)                          # src is dst: in place
    n = keep.shape[1]
    assert torch.equal(kin[:, :, :n], k2) and torch.equal(vin[:, :, :n], v2)
    assert torch.equal(kin[:, :, n:], k[:, :, n:])                               # nothing beyond keep_total touched
    # batch extension: every sequence receives the same per-head index list
    gi = keep.long()[None, :, :, None].expand(b, H, n, D)
    assert torch.equal(k2, torch.gather(k, 2, gi))


def test_prune_then_decode_appends_in_place():
    """After a prune the returned views have spare capacity, so the next forward appends without realloc."""
    from spatten_b200.kv_cache_token_pruning import SpAttenKVCache
    from spatten_b200.cache_buffer import CacheBuffer
    H, S, D = 4, 400, 128
    k = det_normal((1, H, S, D), 81, 1.0, torch.float16).cuda()
    v = det_normal((1, H, S, D), 82, 1.0, torch.float16).cuda()
    s = det_normal((1, H, 1, S), 83, 1.0, torch.float16).cuda()
    cache = SpAttenKVCache(4, 50, 50)
    (k2, v2), = cache.apply_token_pruning([(k, v)], 10, [s])
    assert k2.shape[2] == 4 + 50 + 40
    cb = CacheBuffer.of(k2)
    assert cb is not None and cb.buf.shape[2] >= k2.shape[2] + 10 and CacheBuffer.of(v2) is not None


def test_prune_error_behaviour_matches_reference():
    from spatten_b200.kv_cache_token_pruning import SpAttenKVCache
    k = torch.zeros(1, 2, 40, 64, dtype=torch.float16, device="cuda")
    s = torch.zeros(1, 2, 1, 40, dtype=torch.float16, device="cuda")
    assert SpAttenKVCache(2, 4, 4).apply_token_pruning(None, 3, []) is None
    with pytest.raises(TypeError):
        SpAttenKVCache(2, 4, 0).apply_token_pruning([(k, k)], 0, [s])
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        SpAttenKVCache(2, 5, 45).apply_token_pruning([(k, k)], 20, [s])
    past = [(k, k)]
    assert SpAttenKVCache(2, 30, 30).apply_token_pruning(past, 0, [s]) is past
