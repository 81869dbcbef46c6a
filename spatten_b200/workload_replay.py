"""Replay the reference's attention workload traces.

The RTL simulator consumes per-(iteration, layer, head) CSV rows (header in
spatten_hardware/hardware/workloads/small.csv:1; reader sim/TestSpAtten.scala:54-100): how many keys are fetched
and scored, how many values survive the top-k (local V pruning), the K/V bit widths and whether the row refetches
LSBs.  This module parses the same format, applies the simulator's bit-width mapping (:75-87) and

  * `summarise(rows)` -- algorithmic DRAM bytes per request with the reference's own formula
    (spatten_hardware/simulator/src/bert.cpp:73-83: rows * D * bits / 8, in 32-byte transactions), per iteration/layer;
  * `replay(rows, device)` -- runs every (iteration, layer) group on the GPU with the RTL stage order
    (DotProduct over key_fetch_num keys -> TopK -> MultiplyValue over value_fetch_num survivors) on synthetic
    fp16 data of the trace's shapes (`ops.local_v_attn`), returning device time per group.  The traces carry no
    tensors (the simulator feeds random data, DRAMSimDPIDriverRamulator2.cpp:91-98), so there is nothing to compare
    numerically; the point is the like-for-like schedule.

CLI:  python -m spatten_b200.workload_replay trace.csv [--iteration N] [--gpu]
"""
from __future__ import annotations

import csv
import math
from dataclasses import dataclass
from typing import Dict, Iterable, List, Tuple

TRANSACTION_BYTES = 32          # MatrixFetcher.scala:54; bert.cpp:76-77


@dataclass(frozen=True)
class Request:
    iteration: int
    layer: int
    head: int
    D: int
    L: int
    key_fetch_num: int
    value_fetch_num: int
    key_bits: int               # after the simulator's mapping
    value_bits: int
    if_requant: bool
    if_topk: bool
    topk: int

    @property
    def key_bytes(self) -> int:
        """K fetch incl. the LSB refetch of a requantised row (4 extra bits, SpAttenController.scala:35-39)."""
        bits = self.key_bits + (4 if self.if_requant else 0)
        return _transactions(self.key_fetch_num, self.D, bits) * TRANSACTION_BYTES

    @property
    def value_bytes(self) -> int:
        return _transactions(self.value_fetch_num, self.D, self.value_bits) * TRANSACTION_BYTES


def _transactions(rows: int, d: int, bits: int) -> int:
    return rows * math.ceil(d * bits / 8 / TRANSACTION_BYTES)


def _map_bits(key_bits: int, value_bits: int, if_requant: bool) -> Tuple[int, int, bool]:
    """TestSpAtten.scala:75-87: -1 / 10 / 12 bit keys run as 8 bit + requant; -1 / 10 / 12 bit values as 8 bit."""
    if key_bits in (-1, 10, 12):
        key_bits, if_requant = 8, True
    if value_bits in (-1, 10, 12):
        value_bits = 8
    return key_bits, value_bits, if_requant


def parse(path: str, iteration: int | None = None) -> List[Request]:
    out: List[Request] = []
    with open(path, newline="") as f:
        for row in csv.DictReader(f):
            try:
                it = int(row["iteration_id"])
            except (TypeError, ValueError):
                continue                                   # trailing provenance line (TestSpAtten.scala:55-56 filter)
            if iteration is not None and it != iteration:
                continue
            kb, vb, rq = _map_bits(int(row["quant_key_bit"]), int(row["quant_value_bit"]),
                                   row["if_requant"].strip().lower() == "true")
            out.append(Request(it, int(row["layer_id"]), int(row["head_id"]), int(float(row["embedding_length_D"])),
                               int(row["sentence_length_L"]), int(row["key_fetch_num"]), int(row["value_fetch_num"]),
                               kb, vb, rq, row["if_topk"].strip().lower() == "true", int(row["topk"])))
    return out


def summarise(rows: Iterable[Request]) -> Dict[Tuple[int, int], Dict[str, float]]:
    """(iteration, layer) -> heads, keys, values, DRAM bytes (reference formula), dense fp16 bytes for comparison."""
    agg: Dict[Tuple[int, int], Dict[str, float]] = {}
    for r in rows:
        a = agg.setdefault((r.iteration, r.layer), dict(heads=0, keys=0, values=0, bytes=0, dense_fp16_bytes=0, requant=0))
        a["heads"] += 1
        a["keys"] = max(a["keys"], r.key_fetch_num)
        a["values"] = max(a["values"], r.value_fetch_num)
        a["bytes"] += r.key_bytes + r.value_bytes
        a["dense_fp16_bytes"] += 2 * r.L * r.D * 2
        a["requant"] += int(r.if_requant)
    return agg


GROUP_COLUMNS = ("iteration", "layer", "heads", "D", "L", "keys", "values", "key_bits", "value_bits", "requant_rows",
                 "topk_rows", "dram_bytes", "dense_fp16_bytes")


def group_requests(rows: Iterable[Request]):
    """Requests -> int64 array [groups, len(GROUP_COLUMNS)], one row per (iteration, layer) in trace order."""
    import numpy as np
    agg: Dict[Tuple[int, int], List[int]] = {}
    for r in rows:
        a = agg.setdefault((r.iteration, r.layer), [r.iteration, r.layer, 0, r.D, r.L, 0, 0, r.key_bits, r.value_bits, 0, 0, 0, 0])
        a[2] += 1
        a[5], a[6] = max(a[5], r.key_fetch_num), max(a[6], r.value_fetch_num)
        a[7], a[8] = max(a[7], r.key_bits), max(a[8], r.value_bits)
        a[9] += int(r.if_requant)
        a[10] += int(r.if_topk)
        a[11] += r.key_bytes + r.value_bytes
        a[12] += 2 * r.L * r.D * 2
    return np.array([agg[k] for k in sorted(agg)], dtype=np.int64).reshape(-1, len(GROUP_COLUMNS))


def replay_groups(groups, device="cuda", repeats: int = 3):
    """Run each (iteration, layer) group on the GPU in the RTL stage order -- DotProduct over `keys` keys -> TopK ->
    MultiplyValue over `values` survivors (`ops.local_v_attn`; plain decode attention when nothing is pruned) -- on
    synthetic fp16 data of the trace's shapes.  Returns microseconds per group (CUDA events, eager launches)."""
    import torch
    from . import ops
    g = torch.Generator(device=device).manual_seed(0)
    cache = {}
    times = []
    for row in groups:
        H, D, n, keep = int(row[2]), int(row[3]), int(row[5]), int(min(row[5], row[6]))
        if D not in (64, 128) or n < 1:
            raise ValueError(f"unsupported trace shape D={D}, keys={n}")
        key = (H, D, n)
        if key not in cache:
            cache[key] = (torch.randn((1, H, D), generator=g, device=device, dtype=torch.float16),
                          torch.randn((1, H, n, D), generator=g, device=device, dtype=torch.float16),
                          torch.randn((1, H, n, D), generator=g, device=device, dtype=torch.float16),
                          ops.rope_table(D, n, device))
        q, kc, vc, rope = cache[key]
        run = (lambda: ops.local_v_attn(q, kc, vc, n, rope, keep)) if keep < n else \
              (lambda: ops.decode_attn(q, kc, vc, n, rope))
        run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(repeats):
            run()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3 / repeats)
    return times


def replay(rows: Iterable[Request], device="cuda", repeats: int = 3):
    """Run each (iteration, layer) group on the GPU; returns {(iteration, layer): microseconds}."""
    groups = group_requests(rows)
    times = replay_groups(groups, device, repeats)
    return {(int(r[0]), int(r[1])): t for r, t in zip(groups, times)}


def main(argv=None):
    import argparse
    import json
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("csv")
    ap.add_argument("--iteration", type=int, default=None)
    ap.add_argument("--gpu", action="store_true", help="also replay on cuda:0")
    a = ap.parse_args(argv)
    rows = parse(a.csv, a.iteration)
    summ = summarise(rows)
    times = replay(rows) if a.gpu else {}
    for key in sorted(summ):
        rec = dict(iteration=key[0], layer=key[1], **summ[key])
        if key in times:
            rec["gpu_us"] = round(times[key], 2)
        print(json.dumps(rec))


if __name__ == "__main__":
    main()
