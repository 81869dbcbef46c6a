"""Time SpAttenKVCache.apply_token_pruning on a full model's cache (all layers in two launches)."""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatten_b200.kv_cache_token_pruning import SpAttenKVCache  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--heads", type=int, default=32)
ap.add_argument("--seq", type=int, default=4096)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
dev, L, H, S, D = "cuda", a.layers, a.heads, a.seq, 128
g = torch.Generator(device=dev).manual_seed(0)
past = [(torch.randn((1, H, S, D), generator=g, device=dev, dtype=torch.float16),
         torch.randn((1, H, S, D), generator=g, device=dev, dtype=torch.float16)) for _ in range(L)]
sc = [torch.randn((1, H, 1, S), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
keep = S // 2
with contextlib.redirect_stdout(io.StringIO()):
    cache = SpAttenKVCache(4, (keep - 4) // 2, (keep - 4) // 2)
for _ in range(2):
    cache.apply_token_pruning(past, 0, sc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(a.reps):
    cache.apply_token_pruning(past, 0, sc)
e1.record()
t_host = (time.perf_counter() - t0) / a.reps
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
by = 4 * L * H * keep * D * 2
print(json.dumps({"kind": "prune_layers", "layers": L, "heads": H, "seq": S, "keep": keep, "ms_per_call": round(ms, 4),
                  "us_per_layer": round(ms * 1e3 / L, 2), "host_ms_per_call": round(t_host * 1e3, 4),
                  "gather_GBps_incl_select": round(by / ms / 1e6, 1)}))
