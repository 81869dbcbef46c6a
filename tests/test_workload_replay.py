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
