"""A few eager launches of the persistent decode-step kernel for ncu (tools/gpu_profile.sh)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatten_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="1,32,128,2048,8")
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--no-chain", action="store_true")
a = ap.parse_args()
b, H, D, n, L = map(int, a.shape.split(","))
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
kc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
vc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
q = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
kn = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
vn = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
rope = ops.rope_table(D, n + 8, dev)
logits = torch.empty((L, b, H, n + 8), device=dev, dtype=torch.float16)
step = ops.DecodeStep(q, kc, vc, n - 1, rope, k_new=kn, v_new=vn, logits=logits, chain=not a.no_chain)
for _ in range(a.launches):
    step.launch()
torch.cuda.synchronize()
print("ok", step.steps_completed(), step.timed_out())
