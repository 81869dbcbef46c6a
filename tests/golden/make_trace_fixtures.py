"""Reduce the reference's workload traces (spatten_hardware/hardware/workloads/*.csv, the RTL simulator's input format,
sim/TestSpAtten.scala:54-100) to compact per-(iteration, layer) schedules: tests/golden/trace_<name>.npz.

Run once in the build container (the GPU box has no /root/reference):   python tests/golden/make_trace_fixtures.py
Columns of `groups` (int64): iteration, layer, heads, D, L, keys (max over heads), values (max over heads), key_bits,
value_bits (after the simulator's bit mapping, :75-87), requant rows, topk rows, DRAM bytes (reference formula
bert.cpp:73-83, summed over heads), dense fp16 bytes.  Only numbers derived from the traces are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from spatten_b200 import workload_replay as W  # noqa: E402

SRC = "/root/reference/spatten_hardware/hardware/workloads"
TRACES = {
    "small": "small.csv",
    "gpt2_medium_ptb_per0": "summary-gpt2-medium-ptb-per0.csv",
    "gpt2_medium_wikitext2_per0_hp": "summary-gpt2-medium-wikitext2-per0-hp-step5+dedic+glob_p+local_p+hp+topk+quant.csv",
    "gpt2_small_wikitext2_per8_step1": "summary-gpt2-small-wikitext2-per8-step1+dedic.csv",
    "gpt2_small_wikitext2_per8": "summary-gpt2-small-wikitext2-per8.csv",
}


def main():
    for name, fn in TRACES.items():
        rows = W.parse(os.path.join(SRC, fn))
        groups = W.group_requests(rows)
        np.savez_compressed(os.path.join(HERE, f"trace_{name}.npz"), groups=groups, source=np.array(fn))
        print(name, groups.shape, "requests", len(rows))


if __name__ == "__main__":
    main()
