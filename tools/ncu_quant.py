"""A few eager config-4 quantised decode steps for ncu (launch list / full capture)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatten_b200 import ops  # noqa: E402
from spatten_b200.quant_kv import QuantKVCache  # noqa: E402

dev, b, H, D, n, L = "cuda", 1, 32, 128, 8192, 4
g = torch.Generator(device=dev).manual_seed(3)
rope = ops.rope_table(D, n + 8, dev)
caches = []
for _ in range(L):
    c = QuantKVCache(b, H, n)
    c.append(torch.randn((b, H, n, D), generator=g, device=dev, dtype=torch.float16),
             torch.randn((b, H, n, D), generator=g, device=dev, dtype=torch.float16))
    caches.append(c)
q = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
thr = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
for rep in range(3):
    for l in range(L):
        caches[l].decode(q[l], rope, threshold=thr, v_bits=4)
torch.cuda.synchronize()
print("ok")
