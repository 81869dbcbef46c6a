"""CPU test of the workload-trace parser and byte accounting (format: reference workloads/small.csv:1)."""
import os
import textwrap

from spatten_b200 import workload_replay as W

HEADER = ("iteration_id,layer_id,head_id,embedding_length_D,sentence_length_L,key_fetch_num,quant_key_bit,quant_query_bit,"
          "auto_requant_thres,if_requant,auto_requant_incre,value_fetch_num,quant_value_bit,if_accumulate_importance,"
          "if_rescale_previous_importance,if_topk,topk")


def _write(tmp_path):
    body = textwrap.dedent(f"""\
        {HEADER}
        0,0,0,64.0,1000,1000,-1,-1,-1,False,-1,1000,-1,True,False,False,-1
        0,0,1,64.0,1000,1000,6,12,0.05,False,4,250,6,True,False,True,900
        1,0,0,64.0,1001,900,6,12,0.05,True,4,200,8,True,False,False,-1
        1,1,0,64.0,1001,700,12,12,0.05,False,4,700,12,True,False,False,-1
        some/provenance/line.yml,,,,,,,,,,,,,,,,
        """)
    p = os.path.join(tmp_path, "trace.csv")
    open(p, "w").write(body)
    return p


def test_parse_and_bit_mapping(tmp_path):
    rows = W.parse(_write(str(tmp_path)))
    assert len(rows) == 4                                           # provenance line skipped (non-integer iteration)
    r0, r1, r2, r3 = rows
    assert (r0.key_bits, r0.value_bits, r0.if_requant) == (8, 8, True)      # -1 -> 8 bit + requant (TestSpAtten.scala:75-81)
    assert (r1.key_bits, r1.value_bits, r1.if_requant, r1.if_topk, r1.topk) == (6, 6, False, True, 900)
    assert (r3.key_bits, r3.value_bits, r3.if_requant) == (8, 8, True)      # 12 -> 8 bit + requant
    assert len(W.parse(_write(str(tmp_path)), iteration=1)) == 2


def test_byte_accounting_matches_reference_formula(tmp_path):
    rows = W.parse(_write(str(tmp_path)))
    r1 = rows[1]
    # 6-bit keys, D=64: 48 bytes -> 2 transactions of 32 B per row (bert.cpp:76-77)
    assert r1.key_bytes == 1000 * 2 * 32 and r1.value_bytes == 250 * 2 * 32
    r2 = rows[2]                                                    # requant: 6 + 4 bits = 80 B -> 3 transactions
    assert r2.key_bytes == 900 * 3 * 32 and r2.value_bytes == 200 * 2 * 32
    s = W.summarise(rows)
    assert set(s) == {(0, 0), (1, 0), (1, 1)}
    assert s[(0, 0)]["heads"] == 2 and s[(0, 0)]["values"] == 1000 and s[(0, 0)]["requant"] == 1
    assert s[(0, 0)]["dense_fp16_bytes"] == 2 * (2 * 1000 * 64 * 2)


import numpy as np      # noqa: E402
import pytest           # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRACES = ["small", "gpt2_medium_ptb_per0", "gpt2_medium_wikitext2_per0_hp", "gpt2_small_wikitext2_per8_step1",
          "gpt2_small_wikitext2_per8"]


def _groups(name):
    return np.load(os.path.join(GOLDEN, f"trace_{name}.npz"))["groups"]


def test_trace_fixtures_carry_the_schedules_the_survey_describes():
    """The reduced fixtures of the reference's four real traces (tests/golden/make_trace_fixtures.py): cascade key
    counts, local-V ratio and head pruning as SURVEY.md App. B.4 read them off the CSVs."""
    cols = {c: i for i, c in enumerate(W.GROUP_COLUMNS)}
    s = _groups("gpt2_small_wikitext2_per8")
    assert set(s[:, cols["D"]]) == {64} and set(s[:, cols["heads"]]) == {12} and s[:, cols["layer"]].max() == 11
    it1 = s[s[:, cols["iteration"]] == 1]
    assert list(it1[:, cols["keys"]][:3]) == [921, 716, 204]                      # cascade token pruning
    assert (it1[:, cols["values"]] < it1[:, cols["keys"]]).all()                   # local V pruning
    assert s[:, cols["requant_rows"]].sum() > 0 and set(s[:, cols["key_bits"]]) == {6}
    m = _groups("gpt2_medium_wikitext2_per0_hp")
    heads = m[m[:, cols["iteration"]] == 1][:, cols["heads"]]
    assert heads[0] == 16 and heads[-1] == 10 and sorted(set(heads)) == [10, 13, 16]      # cascade head pruning
    for name in TRACES:
        g = _groups(name)
        assert (g[:, cols["dram_bytes"]] > 0).all() and (g[:, cols["dram_bytes"]] <= g[:, cols["dense_fp16_bytes"]]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", TRACES)
def test_replay_real_traces_on_gpu(name):
    """Every (layer) group of the first generation iterations of each real trace runs through the RTL-ordered GPU path
    (DotProduct -> TopK -> MultiplyValue over survivors); the trace's own byte accounting rides along."""
    cols = {c: i for i, c in enumerate(W.GROUP_COLUMNS)}
    g = _groups(name)
    its = sorted(set(g[:, cols["iteration"]]))
    sel = g[np.isin(g[:, cols["iteration"]], its[:2] + its[-1:])]
    t = W.replay_groups(sel, repeats=2)
    assert len(t) == len(sel) and all(x > 0 for x in t)
    print(f"trace-replay {name}: {len(sel)} layer groups of iterations {its[:2] + its[-1:]}, "
          f"{sum(t) / 1e3:.2f} ms on the GPU, {sel[:, cols['dram_bytes']].sum() / 2**20:.1f} MiB by the reference's formula, "
          f"{sel[:, cols['dense_fp16_bytes']].sum() / 2**20:.1f} MiB dense fp16")
