"""Shared-memory hazard check (compute-sanitizer --tool racecheck) of the kernels added late in round 2: the fused prefill
(two query-tile pairs + a lone middle tile, mask tensor, importance) and the tensor-core quantised decode (both passes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatten_b200 import ops
from spatten_b200.quant_kv import QuantKVCache
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(s, generator=g, device=dev, dtype=torch.float16)
# fused prefill: two query-tile pairs + a lone middle tile, causal mask tensor, importance on
S = 384
q = rn(1, S, 2, 128); kc, vc = rn(1, 2, S, 128), rn(1, 2, S, 128); pos = torch.arange(S, device=dev).unsqueeze(0)
mask = torch.full((S, S), torch.finfo(torch.float16).min, dtype=torch.float16, device=dev).triu(1)[None].contiguous()
imp = torch.zeros((1, 2, S), dtype=torch.float32, device=dev)
ops.prefill_attn(q, kc, vc, S, ops.rope_table(128, S, dev), pos, mask, importance=imp)
# quantised decode, both passes
qc = QuantKVCache(1, 4, 200); qc.append(rn(1, 4, 200, 128), rn(1, 4, 200, 128))
qc.decode(rn(1, 4, 128), ops.rope_table(128, 200, dev), threshold=1.1, v_bits=4)
torch.cuda.synchronize()
print("race subset done")
