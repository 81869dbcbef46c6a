import sys, torch
sys.path.insert(0,'/root/repo')
from spatten_b200 import _lib
lib=_lib.load(); torch.zeros(1,device='cuda')
print("max clusters D=128:", lib.spatten_debug_max_clusters(128), "D=64:", lib.spatten_debug_max_clusters(64), lib.spatten_last_error())
