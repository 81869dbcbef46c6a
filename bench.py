#!/usr/bin/env python
"""Benchmark of the sparse-attention decode hot path (BASELINE.json metric) -- see DESIGN.md "Measurement".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A STEP is one decode token through the attention hot path of every layer of the named model:
L launches of the fused kernel `spatten_decode_attn` (RoPE(Q), KV append, RoPE-on-K, Q.K^T/sqrt(D), logit
export, softmax, P.V, split-KV merge) over the *pruned* KV cache, i.e. the attention core of
`llama_pos_shift_attention_forward` (reference modify_llama.py:86-144) x L.  Projections / MLP are outside
the hot path and outside the step, for both arms.

  value  tok/s, device resident: the L launches replayed from a CUDA graph, inputs already in HBM.
  e2e    tok/s through the C-ABI call with HOST buffers (`spatten_decode_attn_host`): per layer, q/k_new/v_new
         come from pinned host memory (H2D inside the timed region) and the attention output is read back
         (D2H) before the call returns.  The KV cache is device-resident state in both arms (the reference's
         past_key_values live on the device too).
  N > 1  weak scaling: global batch N, heads sharded N ways (rank r owns H/N heads of every sequence), so
         every GPU streams the same number of KV bytes as the N=1 run; one NCCL collective per layer on the
         attention output (all-reduce of the zero-padded [N, b, H/N, D] buffer == the head all-gather).
  --impl reference   the reference algorithm on the box's host cores: the CPU oracle port
         (oracle/spatten_oracle.py::attn_forward_core, pinned to the unmodified reference by tests/golden)
         -- the Python reference itself cannot travel to the GPU box.  Each step is a bounded sample of
         layers scaled to L.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
# BASELINE.json configs[1]: Llama-2-7B decode bs=1 seq=4K, 50 % KV-token prune, fp16
MODEL = dict(name="llama-2-7b", L=32, H=32, D=128, seq=4096, start=4, important=1022, recent=1022)


def kept_len(m):
    return m["start"] + m["important"] + m["recent"]          # 2048


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------------

class CpuAttentionCore:
    """The reference attention core (oracle port) on host cores, at the step's shapes."""

    def __init__(self, b, H, D, past, n_caches, threads):
        import torch
        from oracle import spatten_oracle as O
        self.torch, self.O = torch, O
        torch.set_num_threads(threads)
        g = torch.Generator().manual_seed(1234)
        self.caches = [(torch.randn((b, H, past, D), generator=g).half(), torch.randn((b, H, past, D), generator=g).half())
                       for _ in range(n_caches)]
        self.q = torch.randn((b, H, 1, D), generator=g).half()
        self.k = torch.randn((b, H, 1, D), generator=g).half()
        self.v = torch.randn((b, H, 1, D), generator=g).half()
        self.pos = torch.full((b, 1), past, dtype=torch.int64)
        self.mask = torch.zeros((b, 1, 1, past + 1), dtype=torch.float16)
        self.i = 0

    def layer_step(self):
        pk, pv = self.caches[self.i % len(self.caches)]
        self.i += 1
        with self.torch.no_grad():
            self.O.attn_forward_core(self.q, self.k, self.v, pk, pv, self.pos, self.mask)

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
    m = MODEL
    threads = os.cpu_count() or 1
    sample_layers = 4
    core = CpuAttentionCore(1, m["H"], m["D"], kept_len(m) - 1, 4, threads)
    per_step = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        for _ in range(sample_layers):
            core.layer_step()
        if i >= args.warmup:
            per_step.append((time.perf_counter() - t0) * m["L"] / sample_layers)
    sec = statistics.mean(per_step)
    val = 1.0 / sec
    sample = (f"{sample_layers} of {m['L']} layers per step (attn_forward_core incl. the reference's cat / full-K RoPE / "
              f"logit clone), scaled x{m['L'] // sample_layers}; torch {torch.__version__} CPU fp16")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": workload_config(1, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, b):
    m = MODEL
    return {
        "workload": (f"{m['name']} decode, bs={b}, L={m['L']} layers, H={m['H']} heads x D={m['D']}, KV cache seq={m['seq']} "
                     f"pruned 50% -> {kept_len(m)} tokens/head (start {m['start']} + important {m['important']} + recent "
                     f"{m['recent']}), fp16; step = {m['L']} x spatten_decode_attn (attention core, projections excluded)"),
        "global_batch": b, "kv_len": kept_len(m), "layers": m["L"],
        "parallelism": "single GPU" if n_gpus == 1 else f"heads sharded x{n_gpus} (H/N heads per rank), batch {b}",
        "l2": "inputs larger than L2: 1 GiB of distinct KV per step vs ~126 MB L2",
    }


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
    faulthandler.dump_traceback_later(900, exit=True)      # a wedged collective must not eat the GPU lease
    import torch
    import torch.distributed as dist
    from spatten_b200 import _lib, ops

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
    m = MODEL
    L, H, D = m["L"], m["H"], m["D"]
    assert H % world == 0
    b, Hl = world, H // world                      # weak scaling: batch N, H/N heads per rank
    n = kept_len(m)                                # tokens attended per head after the 50 % prune
    kv_len = n - 1                                 # cached; the step appends token n-1 (slot kv_len)
    cap = n + 64

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    kc = [torch.randn((b, Hl, cap, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    vc = [torch.randn((b, Hl, cap, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
    q = torch.randn((L, b, Hl, D), generator=g, device=dev, dtype=torch.float16)
    kn = torch.randn((L, b, Hl, D), generator=g, device=dev, dtype=torch.float16)
    vn = torch.randn((L, b, Hl, D), generator=g, device=dev, dtype=torch.float16)
    rope = ops.rope_table(D, cap, dev)
    logits = torch.empty((L, b, Hl, cap), device=dev, dtype=torch.float16)
    # attention output of every layer: [L][N, b, H/N, D]; slot `rank` is written by the kernel, the collective fills the rest
    gathered = torch.zeros((L, world, b, Hl, D), device=dev, dtype=torch.float16)

    flags = _lib.DECODE_PDL | _lib.DECODE_PDL_KV_STABLE if (world == 1 and not args.no_pdl) else 0
    stream = torch.cuda.Stream(device=dev)
    sptr = C.c_void_p(stream.cuda_stream)
    layer_args, keep = [], []
    with torch.cuda.stream(stream):
        for l in range(L):
            a, _, ka = ops.make_decode_args(q[l], kc[l], vc[l], kv_len, rope, k_new=kn[l], v_new=vn[l], logits=logits[l],
                                            out=gathered[l, rank], flags=flags)
            layer_args.append(a)
            keep.append(ka)

    # the same launches without any exchange (local `out` only): the kernel-only region of the roofline at N > 1
    local_args = [_lib.DecodeArgs.from_buffer_copy(a) for a in layer_args]

    def attn_layer(l, exchange=True):
        _lib.check(lib.spatten_decode_attn(C.byref(layer_args[l] if exchange else local_args[l]), sptr))

    from spatten_b200.parallel import HeadShard, PeerExchange, exchange_heads
    shard = HeadShard(rank, world, H)
    px = None
    if world > 1 and args.collective == "fused":
        # attention output stored straight into every peer's gather buffer by the kernel; no collective kernel
        try:
            px = PeerExchange(L, shard, b, D, dev)
        except Exception as e:              # no peer-mappable memory on this box: use the NCCL formulation, and say so
            if rank == 0:
                print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); falling back to NCCL all-reduce",
                      file=sys.stderr)
            args.collective = "allreduce"
    local_out = gathered       # `local_args` (kernel-only roofline region) keep writing here: it must stay allocated
    if px is not None:
        gathered = px.gathered
        for l in range(L):
            # layer l waits, inside its own kernel, for layer l-1's exchange; one wait launch closes the step
            px.fill_args(layer_args[l], l, wait_on_layer=(l - 1) if l > 0 else None)
            layer_args[l].flags = _lib.DECODE_PDL | _lib.DECODE_PDL_KV_STABLE

    def collective(l):
        if px is not None:
            if l == L - 1:
                px.wait(l, stream.cuda_stream)      # last layer: all N producers have stored their heads
            return
        # zero-padded all-reduce (re-zeroes the peers' slots first) or all-gather of the per-rank head blocks
        exchange_heads(gathered[l], shard, args.collective)


    def step_eager(with_collective=True):
        for l in range(L):
            attn_layer(l, with_collective)
            if with_collective:
                collective(l)

    def make_runner(with_collective):
        """CUDA-graph the step when possible; fall back to eager launches (reported)."""
        if args.no_graph:
            return (lambda: step_eager(with_collective)), False
        try:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=stream):
                step_eager(with_collective)
            return gr.replay, True
        except Exception as e:                      # e.g. NCCL capture unsupported
            if rank == 0:
                print(f"[bench] graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            torch.cuda.synchronize()
            return (lambda: step_eager(with_collective)), False

    def timed(run, steps, warmup, min_warm_s=0.0):
        with torch.cuda.stream(stream):
            # at least `warmup` untimed steps, and long enough under load for the clock sampler to see it.
            # The count must be identical on every rank (each step contains collectives).
            t_w = time.perf_counter()
            for _ in range(warmup):
                run()
            stream.synchronize()
            est = torch.tensor([(time.perf_counter() - t_w) / warmup], device=dev)
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
        return ms

    with torch.cuda.stream(stream):
        for _ in range(2):
            step_eager(True)
        stream.synchronize()
        run_full, graphed = make_runner(True)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    total_ms = timed(run_full, args.steps, max(3, args.warmup), min_warm_s=args.min_warm_s)
    ms_per_step = total_ms / args.steps
    value = b / (ms_per_step * 1e-3)

    # kernel-only region for the roofline (identical to the step when N == 1)
    if world > 1:
        with torch.cuda.stream(stream):
            run_k, _ = make_runner(False)
        kernel_ms = timed(run_k, args.steps, 3) / args.steps
    else:
        kernel_ms = ms_per_step
    clocks = sampler.stop() if sampler else None

    bytes_per_launch = 2 * b * Hl * n * D * 2                      # K rows + V rows attended, fp16
    us_per_launch = kernel_ms * 1e3 / L
    peak, peak_src = load_peaks()
    achieved = bytes_per_launch / (us_per_launch * 1e-6) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("decode_attn_c2_dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "spatten::decode_attn_kernel<128>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "us_per_launch": us_per_launch,
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "share_of_step": kernel_ms / ms_per_step}

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region -------------------------
    hq = torch.randn((L, b, Hl, D), dtype=torch.float16).pin_memory()
    hk = torch.randn((L, b, Hl, D), dtype=torch.float16).pin_memory()
    hv = torch.randn((L, b, Hl, D), dtype=torch.float16).pin_memory()
    hout = torch.empty((L, world, b, Hl, D), dtype=torch.float16).pin_memory()
    e2e_args = []
    for l in range(L):
        a = _lib.DecodeArgs.from_buffer_copy(layer_args[l])
        a.flags = 0
        a.wait_n = 0                                # e2e waits for every layer's exchange explicitly (output goes to the host)
        e2e_args.append(a)

    def e2e_step():
        for l in range(L):
            if world == 1:
                _lib.check(lib.spatten_decode_attn_host(C.byref(e2e_args[l]), hq[l].data_ptr(), hk[l].data_ptr(),
                                                        hv[l].data_ptr(), hout[l, 0].data_ptr(), sptr))
            else:
                with torch.cuda.stream(stream):
                    q[l].copy_(hq[l], non_blocking=True)
                    kn[l].copy_(hk[l], non_blocking=True)
                    vn[l].copy_(hv[l], non_blocking=True)
                    _lib.check(lib.spatten_decode_attn(C.byref(e2e_args[l]), sptr))
                    if px is not None:
                        px.wait(l, stream.cuda_stream)
                    else:
                        collective(l)
                    hout[l].copy_(gathered[l], non_blocking=True)
                    stream.synchronize()

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = L * 3 * b * Hl * D * 2
    d2h = L * world * b * Hl * D * 2 if world > 1 else L * b * Hl * D * 2
    e2e = {"value": b / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": e2e_s * 1e3,
           "path": "spatten_decode_attn_host per layer: pinned host q/k_new/v_new -> device, kernel, attention output -> host, "
                   "stream sync (wall clock, max over ranks)"}

    # ---- prune cost (once per chat round in the reference flow, run_spatten_llama.py:71-83) ---------------
    extras = {}
    if world == 1:
        from spatten_b200.kv_cache_token_pruning import SpAttenKVCache
        S = m["seq"]
        pk = [(torch.randn((1, H, S, D), generator=g, device=dev, dtype=torch.float16),
               torch.randn((1, H, S, D), generator=g, device=dev, dtype=torch.float16)) for _ in range(4)]
        sc = [torch.randn((1, H, 1, S), generator=g, device=dev, dtype=torch.float16) for _ in range(4)]
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            cache = SpAttenKVCache(m["start"], m["recent"], m["important"])
        cache.apply_token_pruning(pk, 0, sc)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cache.apply_token_pruning(pk, 0, sc)
        torch.cuda.synchronize()
        per_layer_ms = (time.perf_counter() - t0) * 1e3 / 4
        extras["prune_ms_per_layer_4096_to_2048"] = per_layer_ms
        extras["tok_s_incl_prune_every_64_tokens"] = 64.0 / (64 * ms_per_step * 1e-3 + per_layer_ms * L * 1e-3)

    # ---- CPU baseline beside it (rank 0, N == 1) -----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        calls = 24
        sec_layer = CpuAttentionCore(1, H, D, kv_len, 4, threads).seconds_per_layer(calls)
        cpu = {"value": 1.0 / (sec_layer * L), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{calls} layer-steps of oracle attn_forward_core (reference op sequence incl. cat / full-K RoPE / clone) "
                         f"at the step's shapes, mean x {L} layers",
               "ms_per_layer": sec_layer * 1e3}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": workload_config(world, b),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": L * args.steps, "cuda_graph": graphed,
            "collective": None if world == 1 else (
                f"fused: spatten_decode_attn stores [b={b}, H/N={Hl}, D={D}] fp16 into all {world} peers over NVLink + flag wait, per layer"
                if px is not None else f"nccl {args.collective} of [N={world}, b={b}, H/N={Hl}, D={D}] fp16 per layer"),
            "peer_wait_timed_out": px.timed_out() if px is not None else None,
            "pdl": bool(flags),
            "local_out_bytes": local_out.numel() * 2,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tearing the process group down while captured NCCL work is still referenced by a CUDA graph hangs
        # (seen on this pool: destroy_process_group never returns).  Everything is printed; leave hard.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--collective", choices=["fused", "allreduce", "allgather"], default="fused",
                    help="exchange of the attention output at N>1: in-kernel peer stores + flag wait (fused), or NCCL")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--min-warm-s", type=float, default=0.6,
                    help="keep warming up until this many seconds have run under load (clock sampling); 0 under ncu")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
