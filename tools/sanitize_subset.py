"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck): tiny shapes, all code paths."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatten_b200 import _lib, ops
from spatten_b200.kv_cache_token_pruning import SpAttenKVCache
from spatten_b200.quant_kv import QuantKVCache
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(s, generator=g, device=dev, dtype=torch.float16)
for (b, H, D, n, flags) in [(1, 32, 128, 600, 0), (1, 32, 128, 600, _lib.DECODE_NO_CLUSTER), (2, 3, 64, 77, 0), (1, 4, 128, 40, 0),
                            (1, 32, 128, 600, _lib.DECODE_SMALL_TILES), (1, 8, 128, 33, _lib.DECODE_PDL | _lib.DECODE_PDL_KV_STABLE)]:
    kc, vc = rn(b, H, n + 3, D), rn(b, H, n + 3, D)
    logits = torch.empty((b, H, n + 1), device=dev, dtype=torch.float16)
    imp = torch.zeros((b, H, n + 1), device=dev); himp = torch.zeros(H, device=dev)
    ops.decode_attn(rn(b, H, D), kc, vc, n, ops.rope_table(D, n + 1, dev), k_new=rn(b, H, D), v_new=rn(b, H, D), logits=logits,
                    importance=imp, head_importance=himp, flags=flags)
    if D == 128 and n >= 100:
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            c = SpAttenKVCache(4, 50, 50)
        c.apply_token_pruning([(kc[:1, :, :n + 1], vc[:1, :, :n + 1])], 3, [logits[:1].view(1, H, 1, n + 1)])
        ops.local_v_attn(rn(b, H, D), kc, vc, n, ops.rope_table(D, n + 1, dev), 64)
# prefill: row-wise and tcgen05
q = rn(1, 45, 4, 128); kc, vc = rn(1, 4, 45, 128), rn(1, 4, 45, 128); pos = torch.arange(45, device=dev).unsqueeze(0)
ops.decode_attn(q, kc, vc, 45, ops.rope_table(128, 45, dev), q_pos=pos, logits=torch.empty((1, 4, 45, 45), device=dev, dtype=torch.float16))
ops.prefill_attn(q, kc, vc, 45, ops.rope_table(128, 45, dev), pos)
q = rn(1, 200, 2, 128); kc, vc = rn(1, 2, 333, 128), rn(1, 2, 333, 128); pos = torch.arange(133, 333, device=dev).unsqueeze(0)
ops.prefill_attn(q, kc, vc, 333, ops.rope_table(128, 333, dev), pos)
# quantised path
qc = QuantKVCache(1, 4, 130); qc.append(rn(1, 4, 130, 128), rn(1, 4, 130, 128))
qc.decode(rn(1, 4, 128), ops.rope_table(128, 130, dev), threshold=1.1, v_bits=8, logits=torch.empty((1, 4, 130), device=dev, dtype=torch.float16))
ops.head_select(torch.rand(16, device=dev), 5)
torch.cuda.synchronize()
print("sanitize subset done")
