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
