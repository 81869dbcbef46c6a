"""Stage-by-stage probe of the multi-GPU plumbing (prints progress; run under `timeout`)."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(70, exit=True)
import torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
def say(*a):
    print(f"[r{rank} {time.time()%1000:.1f}]", *a, flush=True)
torch.cuda.set_device(local); dev = torch.device("cuda", local)
say("init")
dist.init_process_group("nccl", device_id=dev)
say("inited")
x = torch.ones(4, device=dev); dist.all_reduce(x); torch.cuda.synchronize(); say("allreduce ok", x[0].item())
dist.barrier(); say("barrier ok")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    y = torch.ones(2, 4, device=dev)
    dist.all_reduce(y[0]); s.synchronize(); say("side-stream allreduce ok")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        dist.all_reduce(y[0])
    say("captured")
    g.replay(); s.synchronize(); say("replayed", y[0, 0].item())
dist.destroy_process_group(); say("done")
