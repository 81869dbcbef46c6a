"""Repeat the chained persistent step (fresh caches every time) and check every repetition against the first one and the
per-layer kernel: a race in the segment hand-off would show up as a differing logit / output / appended row."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from spatten_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev, L, b, H, D, past = "cuda", 3, 1, 32, 128, 2047
n = past + 1
g = torch.Generator(device=dev).manual_seed(7)
q = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
kn = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
vn = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
pk = torch.randn((L, b, H, past, D), generator=g, device=dev, dtype=torch.float16)
pv = torch.randn((L, b, H, past, D), generator=g, device=dev, dtype=torch.float16)
rope = ops.rope_table(D, n, dev)
ref = None
bad = 0
for rep in range(reps):
    for chain in (True, False):
        kc = [torch.full((b, H, n + 5, D), float("nan"), dtype=torch.float16, device=dev) for _ in range(L)]
        vc = [torch.full((b, H, n + 5, D), float("nan"), dtype=torch.float16, device=dev) for _ in range(L)]
        for l in range(L):
            kc[l][:, :, :past], vc[l][:, :, :past] = pk[l], pv[l]
        logits = torch.full((L, b, H, n + 3), float("nan"), dtype=torch.float16, device=dev)
        imp = torch.zeros((L, b, H, n), dtype=torch.float32, device=dev)
        himp = torch.zeros(H, dtype=torch.float32, device=dev)
        step = ops.DecodeStep(q, kc, vc, past, rope, k_new=kn, v_new=vn, logits=logits, importance=imp, head_importance=himp, chain=chain)
        out = step.launch().clone()
        torch.cuda.synchronize()
        cur = (out, logits.clone(), imp.clone(), torch.stack([k[:, :, past] for k in kc]), torch.stack([v[:, :, past] for v in vc]))
        ok = (not step.timed_out()) and step.steps_completed() == 1 and not torch.isnan(out).any() and not torch.isnan(cur[1][..., :n]).any()
        ok = ok and torch.equal(cur[3], kn) and torch.equal(cur[4], vn) and torch.isnan(logits[..., n:]).all()
        if ref is None:
            ref = cur
        else:
            ok = ok and torch.equal(cur[1][..., :n], ref[1][..., :n]) and (cur[0].float() - ref[0].float()).abs().max() < 2e-3
            ok = ok and torch.allclose(cur[2], ref[2], rtol=2e-3, atol=1e-6)
        if not ok:
            bad += 1
            print("MISMATCH rep", rep, "chain", chain, "timed_out", step.timed_out(), "steps", step.steps_completed(),
                  "nan out", int(torch.isnan(out).sum()), "nan logits", int(torch.isnan(cur[1][..., :n]).sum()),
                  "logits differ", int((cur[1][..., :n] != ref[1][..., :n]).sum()), flush=True)
print("step stress:", reps, "repetitions x 2 modes,", bad, "mismatches")
