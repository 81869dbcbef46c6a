"""Test helpers: golden loading, tolerances and tie-aware index-set comparison."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star: "fp16 attention output within 1e-3 rel".  Measured normwise: max|a-b| / max|b|.
OUT_REL_TOL = 1e-3
# exported logits are fp16( fp16(q.k) / sqrt(D) ).  A different fp32 summation order can move the inner
# fp16 rounding by one ulp (<= 2^-10 relative) and the outer rounding adds half an ulp more, so 2e-3
# relative is "one ulp of the reference's own intermediate"; tests also require >= 99 % bit-identical.
LOGIT_REL_TOL = 2e-3


def load_golden(kind: str, name: str):
    z = np.load(os.path.join(GOLDEN_DIR, f"{kind}_{name}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Normwise relative error max|a-b| / max|b| (b = reference)."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_logits_close(mine: torch.Tensor, ref: torch.Tensor, what=""):
    """Element-wise: |a-b| <= LOGIT_REL_TOL*|b| + 1e-4 (floor for values near 0)."""
    a, b = mine.double().cpu(), ref.double().cpu()
    finite = torch.isfinite(b)
    bad = ((a - b).abs() > LOGIT_REL_TOL * b.abs() + 1e-4) & finite
    assert not bad.any(), f"{what}: {int(bad.sum())} logits off; worst {float((a-b).abs().max())}"


def assert_index_sets_match(mine: torch.Tensor, ref: torch.Tensor, score: torch.Tensor, what=""):
    """Bit-exact index-set parity, tie-aware.

    mine/ref: [H, keep] ascending kept indices; score: [H, S] the score the selection ran on.
    Exact equality is required unless the reference's kept set and dropped set share a score value
    (a tie at the k-th boundary, where torch.topk's choice is implementation-defined);
    there the two sets must hold the same multiset of scores and differ only inside the tie group.
    """
    mine, ref, score = mine.cpu().long(), ref.cpu().long(), score.cpu().float()
    assert mine.shape == ref.shape, f"{what}: shape {tuple(mine.shape)} vs {tuple(ref.shape)}"
    for h in range(ref.shape[0]):
        if torch.equal(mine[h], ref[h]):
            continue
        sm, sr = set(mine[h].tolist()), set(ref[h].tolist())
        only_m, only_r = sorted(sm - sr), sorted(sr - sm)
        vals_m = sorted(score[h, only_m].tolist())
        vals_r = sorted(score[h, only_r].tolist())
        assert vals_m == vals_r and len(set(vals_m)) == 1, (
            f"{what}: head {h} index sets differ outside a boundary tie: mine-only {only_m[:8]} ref-only {only_r[:8]}")


def fp16_ulp(x: float) -> float:
    """Spacing of fp16 numbers at magnitude |x| (normal range)."""
    import math
    ax = max(abs(float(x)), 2.0 ** -14)
    return 2.0 ** (math.floor(math.log2(ax)) - 10)


def assert_chain_sets_match(mine: torch.Tensor, ref_kept: torch.Tensor, ref_scores: torch.Tensor, kth: torch.Tensor,
                            nxt: torch.Tensor, what="", slack_ulps: int = 2):
    """Index-set parity for a keep set derived from the path's OWN logits (decode -> prune chain).

    mine/ref_kept: [H, keep_total] ascending; ref_scores: [H, S] the REFERENCE's logits; kth/nxt: the reference's
    k-th and (k+1)-th candidate scores per head.  The CUDA logits differ from the CPU reference's by one fp16 ulp on
    ~0.1-1 % of entries (fp32 summation order), so: wherever the reference's boundary gap exceeds `slack_ulps` ulps
    the sets must be IDENTICAL; elsewhere they may differ only by indices whose reference score lies within
    `slack_ulps` ulps of the boundary (a tie or a one-ulp flip).  Returns a report dict.
    """
    mine, ref_kept, ref_scores = mine.cpu().long(), ref_kept.cpu().long(), ref_scores.cpu().float()
    assert mine.shape == ref_kept.shape, f"{what}: shape {tuple(mine.shape)} vs {tuple(ref_kept.shape)}"
    H = ref_kept.shape[0]
    identical = flipped_heads = flipped_idx = wide_margin_heads = 0
    for h in range(H):
        ulp = fp16_ulp(float(kth[h]))
        margin = float(kth[h]) - float(nxt[h])
        wide = margin > slack_ulps * ulp
        wide_margin_heads += int(wide)
        if torch.equal(mine[h], ref_kept[h]):
            identical += 1
            continue
        assert not wide, (f"{what}: head {h} differs although the reference's boundary gap {margin:.3e} is "
                          f"{margin / ulp:.1f} ulp wide")
        sm, sr = set(mine[h].tolist()), set(ref_kept[h].tolist())
        diff = sorted(sm ^ sr)
        off = (ref_scores[h, diff] - float(kth[h])).abs().max().item()
        assert off <= slack_ulps * ulp + 1e-12, (
            f"{what}: head {h}: index {diff[:6]} flipped but its reference score is {off / ulp:.1f} ulp from the boundary")
        flipped_heads += 1
        flipped_idx += len(diff) // 2
    return {"heads": H, "identical": identical, "boundary_flips_heads": flipped_heads, "boundary_flips_idx": flipped_idx,
            "heads_with_wide_margin": wide_margin_heads}
