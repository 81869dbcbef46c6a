"""Tiny on-hardware check of workload_replay.replay(): a synthetic GPT-2-small-shaped trace (12 heads, D=64)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatten_b200 import workload_replay as W
hdr = ("iteration_id,layer_id,head_id,embedding_length_D,sentence_length_L,key_fetch_num,quant_key_bit,quant_query_bit,"
       "auto_requant_thres,if_requant,auto_requant_incre,value_fetch_num,quant_value_bit,if_accumulate_importance,"
       "if_rescale_previous_importance,if_topk,topk")
lines = [hdr]
for layer, (keys, vals) in enumerate([(921, 184), (716, 143), (204, 204)]):
    for h in range(12):
        lines.append(f"1,{layer},{h},64.0,994,{keys},6,12,0.05,False,4,{vals},6,True,False,False,-1")
p = os.path.join(tempfile.mkdtemp(), "t.csv")
open(p, "w").write("\n".join(lines) + "\n")
rows = W.parse(p)
t = W.replay(rows)
s = W.summarise(rows)
for k in sorted(t):
    print(k, "keys", s[k]["keys"], "values", s[k]["values"], "gpu_us", round(t[k], 1))
