"""Kernel-level timing of spatten_decode_attn on synthetic Llama-shaped caches (CUDA graph + events)."""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatten_b200 import ops  # noqa: E402


def bench_shape(b, H, D, n, L, iters, logits_on=True, graph=True, flags=0):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    kc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    vc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    q = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
    kn = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
    vn = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
    rope = ops.rope_table(D, n + 8, dev)
    out = torch.empty((L, b, H, D), device=dev, dtype=torch.float16)
    logits = torch.empty((L, b, H, n + 8), device=dev, dtype=torch.float16) if logits_on else None

    def step():
        for l in range(L):
            ops.decode_attn(q[l], kc[l], vc[l], n - 1, rope, k_new=kn[l], v_new=vn[l],
                            logits=None if logits is None else logits[l], out=out[l], flags=flags)

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
        s.synchronize()
        if graph:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                step()
            run = gr.replay
        else:
            run = step
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.synchronize()
        e0.record(s)
        for _ in range(iters):
            run()
        e1.record(s)
        s.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bytes_layer = 2 * b * H * n * D * 2
    us_layer = ms * 1e3 / L
    return {"b": b, "H": H, "D": D, "n": n, "L": L, "logits": logits_on, "graph": graph, "flags": flags,
            "us_per_layer": round(us_layer, 3), "GBps": round(bytes_layer / us_layer / 1e3, 1),
            "frac_6445": round(bytes_layer / us_layer / 1e3 / 6445, 4)}


def bench_step(b, H, D, n, L, iters, chain=True, logits_on=True):
    """The persistent decode-step kernel: ONE launch for all L layers (spatten_decode_step)."""
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    kc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    vc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    q = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
    kn = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
    vn = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
    rope = ops.rope_table(D, n + 8, dev)
    logits = torch.empty((L, b, H, n + 8), device=dev, dtype=torch.float16) if logits_on else None
    step = ops.DecodeStep(q, kc, vc, n - 1, rope, k_new=kn, v_new=vn, logits=logits, chain=chain)
    ms = _time_graph(step.launch, iters)
    bytes_layer = 2 * b * H * n * D * 2
    us_layer = ms * 1e3 / L
    return {"kind": "decode_step", "b": b, "H": H, "D": D, "n": n, "L": L, "chain": chain, "logits": logits_on,
            "us_per_layer": round(us_layer, 3), "GBps": round(bytes_layer / us_layer / 1e3, 1),
            "frac_6445": round(bytes_layer / us_layer / 1e3 / 6445, 4), "timed_out": step.timed_out()}


def step_timeline(b, H, D, n, L, chain=True):
    """Per-layer phase timeline of the persistent step kernel (instrumented build), medians over CTAs, ns."""
    import ctypes as C
    import numpy as np
    from spatten_b200 import _lib
    lib = _lib.load()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    kc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    vc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    q = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
    rope = ops.rope_table(D, n + 8, dev)
    logits = torch.empty((L, b, H, n + 8), device=dev, dtype=torch.float16)
    step = ops.DecodeStep(q, kc, vc, n - 1, rope, k_new=q, v_new=q, logits=logits, chain=chain)
    for _ in range(3):
        step.launch()
    torch.cuda.synchronize()
    G, Wp = torch.cuda.get_device_properties(0).multi_processor_count, 14      # 12 compute warps + query warp + merge warp
    buf = torch.zeros(L * G * Wp * 8, dtype=torch.int64, device=dev)
    lib.spatten_debug_set_step_timeline(C.c_void_p(buf.data_ptr()))
    step.launch()
    torch.cuda.synchronize()
    lib.spatten_debug_set_step_timeline(None)
    t = buf.cpu().numpy().reshape(L, G, Wp, 8).astype(np.float64)
    t[t == 0] = np.nan
    t0 = np.nanmin(t)
    names = ["seg_reached", "prev_layer_seen", "first_tile+q", "tiles_done", "cta_last_arrived", "ticket_taken", "row_merged",
             "row_signalled"]
    print(f"step timeline b={b} H={H} n={n} L={L} chain={chain}: total {(np.nanmax(t) - t0) / 1e3:.1f} us")
    lay_end = np.array([np.nanmax(t[l, :, :, 7]) if np.isfinite(t[l, :, :, 7]).any() else np.nanmax(t[l]) for l in range(L)])
    print("  layer period (last row signalled), us: median %.2f" % np.median(np.diff(lay_end) / 1e3))
    for l in (L // 2, L // 2 + 1):
        base = lay_end[l - 1]
        print(f"  layer {l}: ns after layer {l-1}'s last row was signalled (min / median / max over warps)")
        for k, nm in enumerate(names):
            v = t[l, :, :, k].ravel()
            v = v[np.isfinite(v)] - base
            if v.size:
                print(f"    {nm:18s} {np.min(v):8.0f} {np.median(v):8.0f} {np.max(v):8.0f}   (n={v.size})")


def _time_graph(fn, iters):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            fn()
        for _ in range(3):
            gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.synchronize()
        e0.record(s)
        for _ in range(iters):
            gr.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_extra(iters):
    """Config-4 (int4 planes, 16K -> 8K tokens) and the RTL-ordered local-V path, one layer-step each, L layers cycled."""
    from spatten_b200.quant_kv import QuantKVCache
    dev, b, H, D, n, L = "cuda", 1, 32, 128, 8192, 16
    g = torch.Generator(device=dev).manual_seed(3)
    rope = ops.rope_table(D, n + 8, dev)
    caches = []
    for l in range(L):
        c = QuantKVCache(b, H, n)
        c.append(torch.randn((b, H, n, D), generator=g, device=dev, dtype=torch.float16),
                 torch.randn((b, H, n, D), generator=g, device=dev, dtype=torch.float16))
        caches.append(c)
    q = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
    logits = torch.empty((L, b, H, n), device=dev, dtype=torch.float16)
    for thr, vb in ((0.0, 4), (1e-4, 4), (1.1, 4), (0.0, 8)):
        needs = []

        def step():
            needs.clear()
            for l in range(L):
                needs.append(caches[l].decode(q[l], rope, threshold=thr, v_bits=vb, logits=logits[l])[1])
        ms = _time_graph(step, iters)
        r = float(torch.stack(needs).float().mean())
        kb = n * H * (D // 2) * (1 + r) + n * H * (D // 2) * (vb // 4) + 2 * n * H * 2      # K msb(+lsb on replay) + V + scales
        us = ms * 1e3 / L
        print(json.dumps({"kind": "quant_c4", "b": b, "H": H, "n": n, "thr": thr, "v_bits": vb, "requant_frac": round(r, 3),
                          "us_per_layer": round(us, 2), "alg_MiB": round(kb / 2**20, 2), "GBps": round(kb / us / 1e3, 1),
                          "frac_6445": round(kb / us / 1e3 / 6445, 4)}), flush=True)
    # local V: score all 4096 keys, multiply the top 2048 values (Mode B, 48 MiB/layer at config-2 scale)
    n, keep = 4096, 2048
    kc = [torch.randn((b, H, n, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    vc = [torch.randn((b, H, n, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    rope = ops.rope_table(D, n + 8, dev)

    def step_b():
        for l in range(L):
            ops.local_v_attn(q[l], kc[l], vc[l], n, rope, keep)
    ms = _time_graph(step_b, iters)
    kb = b * H * (n + keep) * D * 2
    us = ms * 1e3 / L
    print(json.dumps({"kind": "local_v_modeB", "b": b, "H": H, "n": n, "keep_v": keep, "us_per_layer": round(us, 2),
                      "alg_MiB": round(kb / 2**20, 2), "GBps": round(kb / us / 1e3, 1),
                      "frac_6445": round(kb / us / 1e3 / 6445, 4)}), flush=True)


def bench_prefill(iters, sizes=(1024, 4096, 8192)):
    """Prefill attention core (tcgen05 path) for one Llama-2-7B layer: S x S causal, logits exported."""
    dev, b, H, D = "cuda", 1, 32, 128
    g = torch.Generator(device=dev).manual_seed(5)
    for S in sizes:
        q = torch.randn((b, S, H, D), generator=g, device=dev, dtype=torch.float16)
        kc = torch.randn((b, H, S, D), generator=g, device=dev, dtype=torch.float16)
        vc = torch.randn((b, H, S, D), generator=g, device=dev, dtype=torch.float16)
        rope = ops.rope_table(D, S, dev)
        pos = torch.arange(S, device=dev).unsqueeze(0)
        mask = torch.full((S, S), torch.finfo(torch.float16).min, dtype=torch.float16, device=dev).triu(1)[None].contiguous()
        logits = torch.empty((b, H, S, S), dtype=torch.float16, device=dev)
        imp = torch.zeros((b, H, S), dtype=torch.float32, device=dev)
        for name, kw in (("fused", {}), ("fused+importance", {"importance": imp}), ("fused causal flag", {"causal": True}), ("fused no mask", {"nomask": True}),
                         ("three-launch", {"legacy": True}),
                         ("dbg: no logit stores", {"causal": True, "_debug": 1}), ("dbg: phase A only", {"causal": True, "_debug": 2}),
                         ("dbg: phase A only, no stores", {"causal": True, "_debug": 3}), ("dbg: phase A only, no stats", {"causal": True, "_debug": 6}),
                         ("dbg: phase A only, no stats no stores", {"causal": True, "_debug": 7}), ("dbg: no stats", {"causal": True, "_debug": 4})):
            if kw.get("_debug") and S < 8192:
                continue
            m = None if (kw.get("causal") or kw.get("nomask")) else mask
            kw = {k: v for k, v in kw.items() if k != "nomask"}
            for _ in range(2):
                ops.prefill_attn(q, kc, vc, S, rope, pos, m, logits, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_it = max(2, iters // 10)
            e0.record()
            for _ in range(n_it):
                ops.prefill_attn(q, kc, vc, S, rope, pos, m, logits, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n_it
            flops = 4.0 * S * S * H * D * b                  # both matmuls, dense (the reference computes the full square)
            print(json.dumps({"kind": "prefill", "path": name, "S": S, "H": H, "ms_per_layer": round(ms, 3),
                              "dense_equiv_TFLOPs": round(flops / ms / 1e9, 1), "logits_GB": round(b * H * S * S * 2 / 1e9, 2),
                              "logits_write_GBps": round(b * H * S * S * 2 / ms / 1e6, 1)}), flush=True)


def timeline(b, H, D, n, L, flags=0):
    """Phase timeline of the LAST of L back-to-back launches (warm caches), from %globaltimer stamps."""
    import ctypes as C
    import numpy as np
    from spatten_b200 import _lib
    lib = _lib.load()
    dev = "cuda"
    buf = torch.zeros(4096 * 8 * 8, dtype=torch.int64, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    kc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    vc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    q = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
    rope = ops.rope_table(D, n + 8, dev)
    logits = torch.empty((L, b, H, n + 8), device=dev, dtype=torch.float16)
    for rep in range(3):
        for l in range(L):
            if rep == 2 and l == L - 1:
                lib.spatten_debug_set_timeline(C.c_void_p(buf.data_ptr()))
            ops.decode_attn(q[l], kc[l], vc[l], n - 1, rope, k_new=q[l], v_new=q[l], logits=logits[l], flags=flags)
    torch.cuda.synchronize()
    lib.spatten_debug_set_timeline(None)
    t = buf.cpu().numpy().reshape(-1, 8, 8).astype(np.float64)
    used = t[:, :, 0].max(axis=1) > 0
    t = t[used]
    t0 = t[:, :, 0][t[:, :, 0] > 0].min()
    names = ["start", "q_rotated", "tile0_landed", "loop_done", "partial_written", "ticket", "merged", "ml_arrived"]
    print(f"timeline b={b} H={H} D={D} n={n}: {t.shape[0]} CTAs; ns relative to first warp start")
    for k, nm in enumerate(names):
        v = t[:, :, k]
        v = v[v > 0] - t0
        if k >= 6:   # merge-only stamps: report relative to the same warp's ticket stamp (merging warps only)
            sel = (t[:, :, k] > 0) & (t[:, :, 5] > 0)
            if k == 6 and (t[:, :, 7] > 0).sum() == 0:
                sel &= np.arange(8)[None, :] == 0          # cluster mode: warp 0 of every CTA merges
            dv = (t[:, :, k] - t[:, :, 5])[sel]
            if dv.size:
                print(f"  {nm:16s} after own ticket: min {dv.min():6.0f} median {np.median(dv):6.0f} max {dv.max():6.0f}")
        if v.size:
            print(f"  {nm:16s} min {v.min():8.0f}  median {np.median(v):8.0f}  max {v.max():8.0f}  (n={v.size})")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--shape", type=str, default=None, help="b,H,D,n,L: run only this shape (for ncu)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--timeline", action="store_true", help="dump per-phase globaltimer stamps for one launch")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--extra", action="store_true", help="config-4 quantised decode and Mode-B local-V timings")
    ap.add_argument("--prefill", action="store_true", help="tcgen05 prefill attention timings")
    ap.add_argument("--prefill-s", type=int, default=None, help="only this prefill length")
    ap.add_argument("--step", action="store_true", help="persistent decode-step kernel timings")
    ap.add_argument("--no-chain", action="store_true")
    ap.add_argument("--step-timeline", action="store_true")
    a = ap.parse_args()
    if a.step_timeline:
        step_timeline(*map(int, (a.shape or "1,32,128,2048,32").split(",")), chain=not a.no_chain)
        sys.exit(0)
    if a.step:
        if a.shape:
            shapes = [tuple(map(int, a.shape.split(",")))]
        else:
            shapes = [(1, 32, 128, 2048, 32), (1, 32, 128, 4096, 32), (1, 32, 128, 8192, 16), (8, 5, 128, 16384, 8),
                      (1, 24, 128, 4096, 32), (1, 4, 128, 2048, 32), (1, 32, 128, 512, 32), (1, 40, 128, 2048, 40)]
        for (b, H, D, n, L) in shapes:
            for chain in ((False,) if a.no_chain else (True, False)):
                print(json.dumps(bench_step(b, H, D, n, L, a.iters, chain=chain)), flush=True)
        sys.exit(0)
    if a.extra:
        bench_extra(a.iters)
        sys.exit(0)
    if a.prefill:
        bench_prefill(a.iters, (a.prefill_s,) if a.prefill_s else (1024, 4096, 8192))
        sys.exit(0)
    if a.timeline:
        timeline(*map(int, (a.shape or "1,32,128,2048,32").split(",")), flags=a.flags)
        sys.exit(0)
    if a.shape:
        b, H, D, n, L = map(int, a.shape.split(","))
        print(json.dumps(bench_shape(b, H, D, n, L, a.iters, True, graph=not a.no_graph)), flush=True)
        sys.exit(0)
    shapes = [(1, 32, 128, 2048, 32), (1, 32, 128, 4096, 32), (1, 32, 128, 8192, 16), (8, 5, 128, 16384, 8),
              (1, 4, 128, 2048, 32), (1, 32, 128, 512, 32)]
    for (b, H, D, n, L) in shapes:
        for fl in (0, 3, 7, 19, 23):
            print(json.dumps(bench_shape(b, H, D, n, L, a.iters, True, flags=fl)), flush=True)
    print(json.dumps(bench_shape(1, 32, 128, 2048, 32, a.iters, True, graph=False)), flush=True)
