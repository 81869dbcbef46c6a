#!/usr/bin/env python
"""Benchmark of the sparse-attention decode hot path (BASELINE.json metric) -- see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5] [--mode ...]

A STEP is one decode token through the attention hot path of every layer of the named model: RoPE(Q), KV append,
RoPE-on-K, Q.K^T/sqrt(D), logit export, softmax, P.V and the split-KV merge over the *pruned* KV cache, i.e. the attention
core of `llama_pos_shift_attention_forward` (reference modify_llama.py:86-144) x L.  Projections / MLP are outside the hot
path and outside the step, for both arms.  The per-step tensors of EVERY layer (q, k_new, v_new) are
inputs and are resident when the timed region starts.

  value  tok/s, device resident.  Three ways of running the same L layers are measured and all are reported under `modes`:
           independent  one persistent launch per step (`spatten_decode_step`), layers in order, each layer's query read
                        when its K/V tiles are consumed -- the headline, `value`.
           chained      the same launch with SPATTEN_STEP_CHAIN: layer l+1's query is read only after every row of layer
                        l has been stored (the ordering a decoder imposes); K/V prefetch still runs ahead.
           per_layer    L launches of `spatten_decode_attn` with programmatic dependent launch (round-1 path; a launch
                        waits for its predecessor before it reads its query).
  e2e    tok/s through the C-ABI call with HOST buffers (`spatten_decode_step_host`): every layer's q|k_new|v_new in one
         pinned buffer -> device, the launch, every layer's attention output -> host, stream sync; wall clock, max over
         ranks.  `e2e.per_layer_call` is the same through L x `spatten_decode_attn_host` (the reference's call granularity).
         The KV cache is device-resident state in both arms (the reference's past_key_values live on the device too).
  N > 1  weak scaling of the configuration (global batch N, heads sharded N ways: rank r owns H/N heads of every
         sequence), so every GPU streams the same number of KV bytes as the N=1 run; the exchange of the attention
         output is fused into the kernel (rows stored into every peer's gather buffer over NVLink + per-row flag).
         Before timing, one step's gathered output is compared bit-for-bit with an NCCL all-gather of the same launch
         without the exchange (`exchange_parity`).
  --impl reference   the UNMODIFIED reference forward (baseline/_ref/spatten_llm, bound to a transformers-4.33 shim
         `self`, baseline/ref_shim.py) on the box's host cores at the same shapes; each step is a bounded sample of layers
         scaled to L.  Falls back to the golden-pinned oracle port when baseline/_ref is absent.
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes as C
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _baseline_metric():
    """BASELINE.json's metric string, verbatim; the step definition lives in config.workload."""
    try:
        return json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    except Exception:
        return "decode tok/s Llama-2-7B seq=4K 50% KV-prune; attn HBM GB/s vs peak at 1/2/4/8 GPU"


METRIC = _baseline_metric()
UNIT = "tok/s"

# BASELINE.json configs[1..4] as shapes (SURVEY.md 8d "Configs as shapes")
CONFIGS = {
    # Llama-2-7B decode bs=1 seq=4K, 50 % KV-token prune, fp16 -- the configuration the metric is quoted on
    "c2": dict(name="llama-2-7b", L=32, H=32, D=128, batch=1, seq=4096, kept=2048, start=4, important=1022, recent=1022,
               live_heads=32, weak=True, quant=False),
    # Llama-2-7B seq=8K, cascade token (50 %) + head (75 %) prune: the decode step after the cascade
    "c3": dict(name="llama-2-7b", L=32, H=32, D=128, batch=1, seq=8192, kept=4096, start=4, important=2046, recent=2046,
               live_heads=24, weak=True, quant=False),
    # Llama-2-7B seq=16K, int4 MSB/LSB KV planes + 50 % token prune
    "c4": dict(name="llama-2-7b", L=32, H=32, D=128, batch=1, seq=16384, kept=8192, start=4, important=4094, recent=4094,
               live_heads=32, weak=True, quant=True),
    # Llama-2-13B bs=8 seq=32K, 50 % KV prune, heads sharded over the GPUs of the box (strong scaling: fixed problem)
    "c5": dict(name="llama-2-13b", L=40, H=40, D=128, batch=8, seq=32768, kept=16384, start=4, important=8190, recent=8190,
               live_heads=40, weak=False, quant=False),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def shard_shape(cfg, world):
    """(batch, heads per rank) of one rank."""
    if cfg["H"] % world:
        raise SystemExit(f"{cfg['H']} heads do not shard over {world} GPUs")
    b = cfg["batch"] * (world if cfg["weak"] else 1)
    return b, cfg["H"] // world


def workload_config(key, cfg, n_gpus, b, mode):
    scaling = "weak: batch x N, H/N heads per rank" if cfg["weak"] else "strong: fixed batch, H/N heads per rank"
    w = (f"{cfg['name']} decode, bs={b}, L={cfg['L']} layers, H={cfg['H']} heads x D={cfg['D']}, KV cache seq={cfg['seq']} "
         f"pruned 50% -> {cfg['kept']} tokens/head (start {cfg['start']} + important {cfg['important']} + recent "
         f"{cfg['recent']}), " + ("int4 MSB/LSB planes + fp16 scales" if cfg["quant"] else "fp16") +
         (f", {cfg['live_heads']} of {cfg['H']} heads live (head mask)" if cfg["live_heads"] != cfg["H"] else "") +
         f"; step = attention core of all {cfg['L']} layers (projections excluded), q/k_new/v_new of every layer resident")
    return {
        "workload": w, "baseline_config": key, "global_batch": b, "kv_len": cfg["kept"], "layers": cfg["L"],
        "mode": mode,
        "parallelism": "single GPU" if n_gpus == 1 else f"heads sharded x{n_gpus} ({scaling})",
        "l2": f"inputs larger than L2: {cfg['L']} layers of distinct KV per step vs ~126 MB L2",
    }


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference (or the oracle port) on host cores
# ------------------------------------------------------------------------------------------------------

class ReferenceAttentionCore:
    """The reference attention core at the step's shapes, on `device` ("cpu" for the baseline legs, "cuda" for the
    eager-GPU comparison).  kind "reference": the unmodified llama_pos_shift_attention_forward from baseline/_ref through
    the shim; kind "port": oracle/spatten_oracle.py::attn_forward_core (golden-pinned restatement)."""

    def __init__(self, b, H, D, past, n_caches, threads=None, device="cpu"):
        import torch
        self.torch = torch
        if threads:
            torch.set_num_threads(threads)
        g = torch.Generator().manual_seed(1234)
        mk = lambda *s: torch.randn(s, generator=g).half().to(device)
        self.caches = [(mk(b, H, past, D), mk(b, H, past, D)) for _ in range(n_caches)]
        self.q, self.k, self.v = mk(b, H, 1, D), mk(b, H, 1, D), mk(b, H, 1, D)
        self.pos = torch.full((b, 1), past, dtype=torch.int64, device=device)
        self.mask = torch.zeros((b, 1, 1, past + 1), dtype=torch.float16, device=device)
        self.hidden = torch.zeros((b, 1, H * D), dtype=torch.float16, device=device)
        self.i = 0
        try:
            from baseline import ref_shim
            self.mod = ref_shim.core_module(H, H, D, self.q, self.k, self.v, device=device)
            self.kind, self.where = "reference", ref_shim.import_reference()[2]
        except Exception as e:                                   # baseline/_ref absent: the pinned port
            from oracle import spatten_oracle as O
            self.O, self.mod, self.kind, self.where = O, None, "port", f"oracle port ({type(e).__name__})"

    def layer_step(self):
        pk, pv = self.caches[self.i % len(self.caches)]
        self.i += 1
        with self.torch.no_grad():
            if self.mod is not None:
                return self.mod(self.hidden, attention_mask=self.mask, position_ids=self.pos, past_key_value=(pk, pv),
                                use_cache=True)[0]
            return self.O.attn_forward_core(self.q, self.k, self.v, pk, pv, self.pos, self.mask)[0]

    def seconds_per_layer(self, n_calls):
        self.layer_step()
        t0 = time.perf_counter()
        for _ in range(n_calls):
            self.layer_step()
        return (time.perf_counter() - t0) / n_calls


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cfg = CONFIGS[args.config]
    world = max(1, args.gpus)
    b, _ = shard_shape(cfg, world)              # the whole job's batch; the CPU holds all H heads
    threads = os.cpu_count() or 1
    sample_layers = 4 if b * cfg["kept"] <= 4096 else 1
    core = ReferenceAttentionCore(b, cfg["H"], cfg["D"], cfg["kept"] - 1, min(4, sample_layers + 1), threads)
    per_step = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for _ in range(sample_layers):
            core.layer_step()
        if i >= args.warmup:
            per_step.append((time.perf_counter() - t0) * cfg["L"] / sample_layers)
    sec = statistics.mean(per_step)
    val = b / sec
    sample = (f"{sample_layers} of {cfg['L']} layers per step ({'llama_pos_shift_attention_forward, unmodified, ' + core.where if core.kind == 'reference' else 'oracle attn_forward_core'}"
              f", incl. the reference's cat / full-K RoPE / logit clone), scaled x{cfg['L'] / sample_layers:g}; "
              f"torch {torch.__version__} CPU fp16, bs={b}")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak" if cfg["weak"] else "strong",
        "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": workload_config(args.config, cfg, world, b, "reference"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": core.kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# clocks sampling
# ------------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        if not sm:
            return None
        mx = max(float(r[2]) for r in self.rows if len(r) >= 9)
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------

def run_ours(args):
    import faulthandler
    faulthandler.dump_traceback_later(1500, exit=True)     # a wedged exchange must not eat the GPU lease
    import torch
    import torch.distributed as dist
    from spatten_b200 import _lib, ops
    from spatten_b200.parallel import HeadShard, PeerStepExchange

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    key, cfg = args.config, CONFIGS[args.config]
    if cfg["quant"]:
        return run_quant_config(args, cfg, dev)
    L, H, D = cfg["L"], cfg["H"], cfg["D"]
    b, Hl = shard_shape(cfg, world)
    n = cfg["kept"]                                # tokens attended per head after the prune
    kv_len = n - 1                                 # cached; the step appends token n-1 (slot kv_len)
    cap = n + 64
    peak, peak_src = load_peaks()

    def timed_graph(fn, stream, steps, warmup, min_warm_s=0.0):
        """(ms per step, graphed?) of `fn` replayed from a CUDA graph: W warm-ups, K timed steps between barriers, CUDA
        events on the launching stream, max over ranks."""
        with torch.cuda.stream(stream):
            for _ in range(2):
                fn()
            stream.synchronize()
            graphed = True
            try:
                if args.no_graph:
                    raise RuntimeError("--no-graph")
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=stream):
                    fn()
                run = gr.replay
            except Exception as e:
                if rank == 0 and not args.no_graph:
                    print(f"[bench] graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
                torch.cuda.synchronize()
                run, graphed = fn, False
            t_w = time.perf_counter()
            for _ in range(warmup):
                run()
            stream.synchronize()
            est = torch.tensor([(time.perf_counter() - t_w) / max(1, warmup)], device=dev)
            if world > 1:
                dist.all_reduce(est, op=dist.ReduceOp.MAX)
            extra = int(min(20000, max(0.0, min_warm_s) / max(float(est.item()), 1e-6)))
            for i in range(extra):
                run()
                if i % 64 == 63:
                    stream.synchronize()
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                run()
            e1.record(stream)
            stream.synchronize()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, graphed

    # ---- device state: KV caches; every layer's q | k_new | v_new packed in one buffer (one H2D copy in the e2e path) ---
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    kc = [torch.randn((b, Hl, cap, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    vc = [torch.randn((b, Hl, cap, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    inp = torch.randn((3, L, b, Hl, D), generator=g, device=dev, dtype=torch.float16)
    q, kn, vn = inp[0], inp[1], inp[2]
    rope = ops.rope_table(D, cap, dev)
    logits = torch.empty((L, b, Hl, cap), device=dev, dtype=torch.float16)
    head_mask = None
    if cfg["live_heads"] != H:
        # head pruning keeps the same number of heads on every rank (the mask itself is data dependent; any mask with
        # this many live heads streams the same bytes)
        live_l = cfg["live_heads"] // world
        hm = torch.zeros((L, Hl), dtype=torch.uint8)
        for l in range(L):
            hm[l, torch.randperm(Hl, generator=torch.Generator().manual_seed(l))[:live_l]] = 1
        head_mask = hm.to(dev)
    live = cfg["live_heads"] // world
    shard = HeadShard(rank, world, H)
    stream = torch.cuda.Stream(device=dev)
    sptr = C.c_void_p(stream.cuda_stream)

    steps_by_mode, exch = {}, {}
    with torch.cuda.stream(stream):
        for mode, chain in (("independent", False), ("chained", True)):
            st = ops.DecodeStep(q, kc, vc, kv_len, rope, k_new=kn, v_new=vn, logits=logits, head_mask=head_mask, chain=chain)
            steps_by_mode[mode] = st
            if world > 1:
                exch[mode] = PeerStepExchange(st, shard)

    # ---- N > 1: the fused exchange must deliver exactly what every rank computed (bit-exact), before anything is timed --
    exchange_parity = None
    if world > 1:
        ok = True
        with torch.cuda.stream(stream):
            local_step = ops.DecodeStep(q, kc, vc, kv_len, rope, k_new=kn, v_new=vn, logits=logits, head_mask=head_mask,
                                        chain=False, sharded_layout=True)
            mine = local_step.launch(stream.cuda_stream).clone()
            stream.synchronize()
            every = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            for mode, px in exch.items():
                px.gathered.zero_()
                torch.cuda.synchronize()
                dist.barrier()
                steps_by_mode[mode].launch(stream.cuda_stream)
                stream.synchronize()
                dist.barrier()
                for r in range(world):
                    ok = ok and bool(torch.equal(px.gathered[:, r], every[r]))
                ok = ok and not steps_by_mode[mode].timed_out()
            del local_step, mine, every
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        exchange_parity = "ok" if int(t.item()) == 1 else "MISMATCH"
        if exchange_parity != "ok":
            raise SystemExit("[bench] fused exchange delivered different bytes than the all-gather of the local outputs")

    # ---- per-layer launches (round-1 path): PDL, device resident, N == 1 only ------------------------------------------
    layer_args, keep = [], []
    if world == 1:
        out_pl = torch.empty((L, b, Hl, D), device=dev, dtype=torch.float16)
        with torch.cuda.stream(stream):
            for l in range(L):
                a, _, ka = ops.make_decode_args(q[l], kc[l], vc[l], kv_len, rope, k_new=kn[l], v_new=vn[l], logits=logits[l],
                                                out=out_pl[l], head_mask=None if head_mask is None else head_mask[l],
                                                flags=0 if args.no_pdl else _lib.DECODE_PDL | _lib.DECODE_PDL_KV_STABLE)
                layer_args.append(a)
                keep.append(ka)

    def run_per_layer():
        for l in range(L):
            _lib.check(lib.spatten_decode_attn(C.byref(layer_args[l]), sptr))

    runners = {m: (lambda st=st: st.launch(stream.cuda_stream)) for m, st in steps_by_mode.items()}
    if world == 1:
        runners["per_layer"] = run_per_layer
    head_mode = args.mode

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ms_per_step, graphed = timed_graph(runners[head_mode], stream, args.steps, max(3, args.warmup), args.min_warm_s)
    clocks = sampler.stop() if sampler else None
    value = b / (ms_per_step * 1e-3)

    bytes_per_layer = 2 * b * live * n * D * 2                     # K rows + V rows attended, fp16
    modes = {}
    for m, fn in runners.items():
        ms = ms_per_step if m == head_mode else timed_graph(fn, stream, max(20, min(args.steps, 200)), 3)[0]
        us_layer = ms * 1e3 / L
        modes[m] = {"ms_per_step": ms, "tok_s": b / (ms * 1e-3), "us_per_layer": us_layer,
                    "achieved_GBps": bytes_per_layer / us_layer / 1e3, "frac": bytes_per_layer / us_layer / 1e3 / peak,
                    "launches_per_step": L if m == "per_layer" else 1}
    launches = modes[head_mode]["launches_per_step"]
    bytes_per_launch = bytes_per_layer * L // launches
    us_per_launch = ms_per_step * 1e3 / launches
    achieved = bytes_per_launch / (us_per_launch * 1e-6) / 1e9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tp) and key == "c2" and world == 1:
        tj = json.load(open(tp))
        traffic = tj.get(f"{head_mode}_dram_bytes_per_launch")
        traffic_src = tj.get("source")
    roofline = {"bound": "hbm",
                "kernel": "spatten::decode_attn_kernel<128>" if head_mode == "per_layer" else "spatten::decode_step_kernel<128>",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "us_per_launch": us_per_launch,
                "us_per_layer": ms_per_step * 1e3 / L, "frac_of_nominal_8TBps": achieved / 8000.0, "share_of_step": 1.0}

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region -------------------------------------
    st = steps_by_mode["independent" if head_mode == "per_layer" else head_mode]
    h_in = torch.randn(inp.shape, dtype=torch.float16).pin_memory()
    out_dev = exch[head_mode].gathered if world > 1 else st.out
    h_out = torch.empty(out_dev.shape, dtype=torch.float16).pin_memory()
    in_bytes, out_bytes = inp.numel() * 2, out_dev.numel() * 2

    def e2e_step():
        _lib.check(lib.spatten_decode_step_host(C.byref(st.args), h_in.data_ptr(), inp.data_ptr(), in_bytes,
                                                out_dev.data_ptr(), h_out.data_ptr(), out_bytes, sptr))

    def wall(fn, reps):
        for _ in range(3):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        s = (time.perf_counter() - t0) / reps
        if world > 1:
            t = torch.tensor([s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s = float(t.item())
        return s

    e2e_steps = max(3, min(args.steps, 50))
    e2e_s = wall(e2e_step, e2e_steps)
    e2e = {"value": b / e2e_s, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
           "ms_per_step": e2e_s * 1e3,
           "path": "spatten_decode_step_host: every layer's q|k_new|v_new in one pinned host buffer -> device, one persistent "
                   "launch, every layer's attention output" + (" of all ranks (gather buffer)" if world > 1 else "") +
                   " -> host, stream sync (wall clock, max over ranks)"}
    if world == 1:
        # the reference's call granularity: one C-ABI call with host buffers per layer, q|k_new|v_new packed per layer
        inp_l = torch.empty((L, 3, b, Hl, D), device=dev, dtype=torch.float16)
        h_in_l = torch.randn((L, 3, b, Hl, D), dtype=torch.float16).pin_memory()
        h_out_l = torch.empty((L, b, Hl, D), dtype=torch.float16).pin_memory()
        host_args = []
        for l in range(L):
            a = _lib.DecodeArgs.from_buffer_copy(layer_args[l])
            a.flags = 0
            a.q, a.k_new, a.v_new = inp_l[l, 0].data_ptr(), inp_l[l, 1].data_ptr(), inp_l[l, 2].data_ptr()
            host_args.append(a)

        def e2e_layers():
            for l in range(L):
                _lib.check(lib.spatten_decode_attn_host(C.byref(host_args[l]), h_in_l[l, 0].data_ptr(), h_in_l[l, 1].data_ptr(),
                                                        h_in_l[l, 2].data_ptr(), h_out_l[l].data_ptr(), sptr))
        s_l = wall(e2e_layers, max(3, min(args.steps, 20)))
        e2e["per_layer_call"] = {"value": b / s_l, "ms_per_step": s_l * 1e3, "calls_per_step": L,
                                 "path": "spatten_decode_attn_host per layer: packed q|k_new|v_new host -> device (one copy), "
                                         "kernel, output -> host, stream sync"}

    extras = {}
    if world == 1 and not args.no_extras:
        extras = run_extras(args, cfg, dev, stream, peak, ms_per_step, timed_graph)

    # ---- CPU baseline beside it (rank 0, N == 1) -----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        calls = 24 if b * n <= 4096 else 3
        core = ReferenceAttentionCore(b, H, D, kv_len, 4 if calls > 3 else 2, threads)
        sec_layer = core.seconds_per_layer(calls)
        cpu = {"value": b / (sec_layer * L), "unit": UNIT, "cores": threads, "kind": core.kind,
               "sample": f"{calls} layer-steps of " + ("the unmodified reference forward (baseline/_ref, shim self)" if core.kind == "reference"
                                                      else "oracle attn_forward_core") +
                         f" (incl. the reference's cat / full-K RoPE / logit clone) at the step's shapes, mean x {L} layers",
               "ms_per_layer": sec_layer * 1e3}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if cfg["weak"] else "strong",
            "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": workload_config(key, cfg, world, b, head_mode),
            "roofline": roofline, "modes": modes, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": launches * args.steps, "cuda_graph": graphed,
            "collective": None if world == 1 else (
                f"fused: spatten_decode_step writes every finished [D={D}] row into the local slot of a symmetric gather buffer; "
                f"the merge warp of one designated CTA per layer copies the finished layer ([b={b}, H/N={Hl}] rows) into the "
                f"{world - 1} peers' buffers over NVLink and releases one system-scope flag per peer; the launch ends when every "
                f"peer's flags of every layer have arrived (no NCCL call in the data path)"),
            "exchange_parity": exchange_parity,
            "peer_wait_timed_out": None if world == 1 else any(s.timed_out() for s in steps_by_mode.values()),
            "pdl": bool(head_mode == "per_layer" and not args.no_pdl),
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tearing the process group down while symmetric-memory mappings are live has hung on this pool; everything is
        # printed: leave hard.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_extras(args, cfg, dev, stream, peak, ms_per_step, timed_graph):
    """N == 1 side measurements reported next to the headline: prune cost, the reference ops on this GPU (eager), the
    full attention layer with projections, and the other BASELINE configurations' decode shapes."""
    import torch
    from spatten_b200 import ops
    L, H, D = cfg["L"], cfg["H"], cfg["D"]
    ex = {}
    g = torch.Generator(device=dev).manual_seed(99)

    # prune cost (once per chat round in the reference flow, run_spatten_llama.py:71-83)
    from spatten_b200.kv_cache_token_pruning import SpAttenKVCache
    S = cfg["seq"]
    with torch.cuda.stream(stream):
        pk = [(torch.randn((1, H, S, D), generator=g, device=dev, dtype=torch.float16),
               torch.randn((1, H, S, D), generator=g, device=dev, dtype=torch.float16)) for _ in range(L)]
        sc = [torch.randn((1, H, 1, S), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
        with contextlib.redirect_stdout(io.StringIO()):
            cache = SpAttenKVCache(cfg["start"], cfg["recent"], cfg["important"])
        cache.apply_token_pruning(pk, 0, sc)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(5):
            cache.apply_token_pruning(pk, 0, sc)
        ev1.record()
        host_ms = (time.perf_counter() - t0) * 1e3 / 5             # host time to enqueue one call (no synchronisation inside)
        torch.cuda.synchronize()
        per_layer_ms = ev0.elapsed_time(ev1) / (5 * L)             # the whole model's cache per call: two launches
    keep_total = cfg["start"] + cfg["important"] + cfg["recent"]
    ex["prune"] = {"ms_per_layer": per_layer_ms, "host_ms_per_call": host_ms,
                   "shape": f"{S} -> {keep_total} tokens, {H} heads, {L} layers per call (CUDA events; select + gather of the whole model in two launches)",
                   "gather_GBps": 4 * H * keep_total * D * 2 / (per_layer_ms * 1e-3) / 1e9,
                   "tok_s_incl_prune_every_64_tokens": 64.0 / (64 * ms_per_step * 1e-3 + per_layer_ms * L * 1e-3)}
    del pk, sc

    # the reference's own op sequence on THIS GPU (eager torch on cuBLAS): the "kernel to beat" of SURVEY.md 2.2
    try:
        n = cfg["kept"]
        core = ReferenceAttentionCore(1, H, D, n - 1, 4, device=dev)
        for _ in range(5):
            core.layer_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            core.layer_step()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 40
        ex["eager_gpu"] = {"us_per_layer": us, "kind": core.kind, "tok_s": 1.0 / (us * 1e-6 * L),
                           "what": "the reference attention core (llama_pos_shift_attention_forward between the projections) "
                                   "as eager torch ops on this GPU, same shapes, CUDA events over 40 layer-steps",
                           "speedup_of_headline": us / (ms_per_step * 1e3 / L)}
        del core
    except Exception as e:
        ex["eager_gpu"] = {"unavailable": f"{type(e).__name__}: {e}"}

    # full attention layers incl. the q/k/v/o projections through the patched forward (tok/s definition (2))
    try:
        from spatten_b200.enable_spatten_llm import enable_spatten_llm
        from spatten_b200.modeling import AttentionStack
        hid = H * D
        with torch.cuda.stream(stream), torch.no_grad():
            model = AttentionStack(L, hid, H, device=dev)
            with contextlib.redirect_stdout(io.StringIO()):
                enable_spatten_llm(model, cfg["start"], cfg["important"], cfg["recent"])
            n = cfg["kept"]
            hs = torch.randn((1, 1, hid), generator=g, device=dev, dtype=torch.float16)
            pos = torch.full((1, 1), n - 1, dtype=torch.int64, device=dev)
            pasts = []
            for lyr in model.layers:                     # one eager step allocates our capacity buffers
                k0 = torch.randn((1, H, n - 1, D), generator=g, device=dev, dtype=torch.float16)
                v0 = torch.randn((1, H, n - 1, D), generator=g, device=dev, dtype=torch.float16)
                _, _, present = lyr(hs, position_ids=pos, past_key_value=(k0, v0), use_cache=True)
                pasts.append(present)
            # re-arm every buffer at length n-1 so that each replay appends slot n-1 in place (steady-state decode step)
            from spatten_b200.cache_buffer import CacheBuffer
            pasts = [(CacheBuffer.of(k).view(n - 1), CacheBuffer.of(v).view(n - 1)) for k, v in pasts]

            def full_step():
                x = hs
                for lyr, past in zip(model.layers, pasts):
                    x = lyr(x, position_ids=pos, past_key_value=past, use_cache=False)[0]
                return x
        ms_full, graphed = timed_graph(full_step, stream, 50, 3)
        ex["full_decoder"] = {"tok_s": 1.0 / (ms_full * 1e-3), "ms_per_step": ms_full, "cuda_graph": graphed,
                              "what": f"{L} x patched LlamaAttention.forward (q/k/v/o projections on cuBLAS + spatten_decode_attn), "
                                      "no MLP / norms / embeddings: attention-block tok/s, fp16 weight streaming included"}
        del model, pasts
    except Exception as e:
        ex["full_decoder"] = {"unavailable": f"{type(e).__name__}: {e}"}

    # the other BASELINE configurations' decode shapes, one persistent launch over 8 cycled layers each
    others = {}

    def one_shape(name, b, Hs, n, live, Ls=8):
        with torch.cuda.stream(stream):
            kc = [torch.randn((b, Hs, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(Ls)]
            vc = [torch.randn((b, Hs, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(Ls)]
            q = torch.randn((Ls, b, Hs, D), generator=g, device=dev, dtype=torch.float16)
            kn = torch.randn((Ls, b, Hs, D), generator=g, device=dev, dtype=torch.float16)
            rope = ops.rope_table(D, n + 8, dev)
            lg = torch.empty((Ls, b, Hs, n + 8), device=dev, dtype=torch.float16)
            hm = None
            if live != Hs:
                hm = torch.zeros((Ls, Hs), dtype=torch.uint8)
                for l in range(Ls):
                    hm[l, torch.randperm(Hs, generator=torch.Generator().manual_seed(l))[:live]] = 1
                hm = hm.to(dev)
            st = ops.DecodeStep(q, kc, vc, n - 1, rope, k_new=kn, v_new=kn, logits=lg, head_mask=hm, chain=False)
        ms, _ = timed_graph(lambda: st.launch(stream.cuda_stream), stream, 30, 3)
        by = 2 * b * live * n * D * 2
        us = ms * 1e3 / Ls
        others[name] = {"us_per_layer": us, "achieved_GBps": by / us / 1e3, "frac": by / us / 1e3 / peak,
                        "algorithmic_MiB_per_layer": by / 2**20, "shape": f"b={b} H={Hs} (live {live}) n={n} D={D}",
                        "kernel": "decode_step_kernel (independent layers)"}
    try:
        one_shape("c3_decode_after_cascade", 1, 32, 4096, 24)
        one_shape("c5_per_gpu_shard", 8, 5, 16384, 5)
        others["c4_int4_planes"] = quant_shape(dev, stream, peak, timed_graph, Ls=8)
        others["c3_prefill_8k"] = prefill_shape(dev, stream)
    except Exception as e:
        others["error"] = f"{type(e).__name__}: {e}"
    ex["other_configs"] = others
    return ex


def prefill_shape(dev, stream, S=8192, H=32, D=128):
    """BASELINE config 3's prefill: one Llama-2-7B layer, S x S causal, every pre-mask logit exported (the reference's
    `attn_scores`), importance column sums on; the fused tcgen05 kernel next to the three-launch path it replaces."""
    import torch
    from spatten_b200 import ops
    g = torch.Generator(device=dev).manual_seed(11)
    out = {}
    with torch.cuda.stream(stream):
        q = torch.randn((1, S, H, D), generator=g, device=dev, dtype=torch.float16)
        kc = torch.randn((1, H, S, D), generator=g, device=dev, dtype=torch.float16)
        vc = torch.randn((1, H, S, D), generator=g, device=dev, dtype=torch.float16)
        rope = ops.rope_table(D, S, dev)
        pos = torch.arange(S, device=dev).unsqueeze(0)
        mask = torch.full((S, S), torch.finfo(torch.float16).min, dtype=torch.float16, device=dev).triu(1)[None].contiguous()
        logits = torch.empty((1, H, S, S), dtype=torch.float16, device=dev)
        imp = torch.zeros((1, H, S), dtype=torch.float32, device=dev)
        for name, kw in (("fused", {"importance": imp}), ("three_launch", {"importance": imp, "legacy": True})):
            for _ in range(2):
                ops.prefill_attn(q, kc, vc, S, rope, pos, mask, logits, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_it = 5 if name == "fused" else 2
            e0.record(stream)
            for _ in range(n_it):
                ops.prefill_attn(q, kc, vc, S, rope, pos, mask, logits, **kw)
            e1.record(stream)
            stream.synchronize()
            out[name + "_ms_per_layer"] = e0.elapsed_time(e1) / n_it
    ms = out["fused_ms_per_layer"]
    flops = 4.0 * S * S * H * D
    sustained = 1435.6
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        sustained = float(json.load(open(p)).get("bf16_tflops_sustained", sustained))
    out.update({"shape": f"b=1 H={H} S={S} D={D}, additive causal mask tensor (classified on the device), importance on",
                "dense_equivalent_TFLOPs": flops / ms / 1e9, "frac_of_sustained_tensor_peak": flops / ms / 1e9 / sustained,
                "logits_GB": H * S * S * 2 / 1e9, "logits_write_GBps": H * S * S * 2 / ms / 1e6,
                "kernel": "prefill_pack_kernel + mask_classify_kernel + prefill_fused_kernel (tcgen05, one launch for S and P.V)",
                "prefill_tok_s_attention_core_32_layers": S / (ms * 32 * 1e-3)})
    return out


def quant_shape(dev, stream, peak, timed_graph, Ls=8, b=1, H=32, n=8192, D=128, threshold=0.0):
    """BASELINE config 4: decode over int4 MSB/LSB planes (16K -> 8K tokens kept), one layer-step, Ls layers cycled."""
    import torch
    from spatten_b200 import ops
    from spatten_b200.quant_kv import QuantKVCache
    g = torch.Generator(device=dev).manual_seed(3)
    with torch.cuda.stream(stream):
        rope = ops.rope_table(D, n + 8, dev)
        caches = []
        for _ in range(Ls):
            c = QuantKVCache(b, H, n)
            c.append(torch.randn((b, H, n, D), generator=g, device=dev, dtype=torch.float16),
                     torch.randn((b, H, n, D), generator=g, device=dev, dtype=torch.float16))
            caches.append(c)
        q = torch.randn((Ls, b, H, D), generator=g, device=dev, dtype=torch.float16)
        needs = []

        def step():
            needs.clear()
            for l in range(Ls):
                needs.append(caches[l].decode(q[l], rope, threshold=threshold, v_bits=4)[1])
    ms, _ = timed_graph(step, stream, 30, 3)
    r = float(torch.stack(needs).float().mean())
    by = n * H * (D // 2) * (1 + r) + n * H * (D // 2) + 2 * n * H * 2      # K msb (+ lsb on replay) + V msb + scales
    us = ms * 1e3 / Ls
    return {"us_per_layer": us, "achieved_GBps": by / us / 1e3, "frac": by / us / 1e3 / peak,
            "algorithmic_MiB_per_layer": by / 2**20, "requant_fraction": r,
            "shape": f"b={b} H={H} n={n} D={D}, K int4 MSB (+LSB on replay) + V int4 + fp16 scales",
            "kernel": "decode_attn_q4_kernel (mma.sync tiles, ticket merge in the same launch)"}


def run_quant_config(args, cfg, dev):
    """--config c4 as the main line (single GPU)."""
    import torch
    if int(os.environ.get("WORLD_SIZE", "1")) != 1:
        raise SystemExit("--config c4 is a single-GPU configuration")
    peak, peak_src = load_peaks()
    stream = torch.cuda.Stream(device=dev)

    def timed_graph(fn, stream, steps, warmup, min_warm_s=0.0):
        with torch.cuda.stream(stream):
            for _ in range(2):
                fn()
            stream.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=stream):
                fn()
            for _ in range(warmup):
                gr.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            stream.synchronize()
            e0.record(stream)
            for _ in range(steps):
                gr.replay()
            e1.record(stream)
            stream.synchronize()
        return e0.elapsed_time(e1) / steps, True
    L = cfg["L"]
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    r = quant_shape(dev, stream, peak, timed_graph, Ls=L, n=cfg["kept"])
    clocks = sampler.stop()
    ms = r["us_per_layer"] * L / 1e3
    line = {"metric": METRIC, "value": 1.0 / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": 30, "warmup": 3,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int4+fp16",
            "data": "synthetic", "config": workload_config(args.config, cfg, 1, 1, "per_layer"),
            "roofline": {"bound": "hbm", "kernel": r["kernel"], "achieved": r["achieved_GBps"], "peak": peak, "unit": "GB/s",
                         "frac": r["frac"], "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": r["algorithmic_MiB_per_layer"] * 2**20,
                         "us_per_launch": r["us_per_layer"]},
            "cpu_baseline": None, "e2e": None, "clocks": clocks, "gpu_launches": 2 * L * 30, "cuda_graph": True}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c2",
                    help="BASELINE.json configuration (c2 = the one the metric is quoted on)")
    ap.add_argument("--mode", choices=["independent", "chained", "per_layer"], default="independent",
                    help="which way of running the L layers is the timed headline (all are reported under `modes`)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--min-warm-s", type=float, default=0.6,
                    help="keep warming up until this many seconds have run under load (clock sampling); 0 under ncu")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the N == 1 side measurements (prune, eager GPU, ...)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.mode == "per_layer" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            raise SystemExit("--mode per_layer is measured at N == 1 only")
        run_ours(args)


if __name__ == "__main__":
    main()
