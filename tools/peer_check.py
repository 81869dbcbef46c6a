"""2+ GPU check of the fused output exchange (run under torchrun): gathered == single-GPU output, bit for bit."""
import ctypes as C
import faulthandler
import os
import sys

faulthandler.dump_traceback_later(100, exit=True)
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spatten_b200 import _lib, ops
from spatten_b200.parallel import HeadShard, PeerExchange, gathered_to_heads

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
L, b, H, D, n = 3, world, 32, 128, 1500
sh = HeadShard(rank, world, H)
g = torch.Generator(device=dev).manual_seed(7)          # same seed everywhere: every rank can build the full problem
kc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
vc = [torch.randn((b, H, n + 8, D), generator=g, device=dev, dtype=torch.float16) for _ in range(L)]
q = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
kn = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
vn = torch.randn((L, b, H, D), generator=g, device=dev, dtype=torch.float16)
rope = ops.rope_table(D, n + 8, dev)
# reference: every shard's heads computed on THIS GPU with the shard's own launch shape (same split configuration,
# hence the same summation order => bit-identical), assembled in head order
full = []
for l in range(L):
    parts = []
    for r in range(world):
        sr = HeadShard(r, world, H).slice()
        parts.append(ops.decode_attn(q[l][:, sr].contiguous(), kc[l][:, sr].contiguous(), vc[l][:, sr].contiguous(), n, rope,
                                     k_new=kn[l][:, sr].contiguous(), v_new=vn[l][:, sr].contiguous()))
    full.append(torch.cat(parts, dim=1))
whole = [ops.decode_attn(q[l], kc[l].clone(), vc[l].clone(), n, rope, k_new=kn[l], v_new=vn[l]) for l in range(L)]
# sharded + fused exchange
s = sh.slice()
kcs = [kc[l][:, s].contiguous() for l in range(L)]
vcs = [vc[l][:, s].contiguous() for l in range(L)]
px = PeerExchange(L, sh, b, D, dev)
stream = torch.cuda.current_stream()
args = []
for l in range(L):
    a, _, keep = ops.make_decode_args(q[l][:, s].contiguous(), kcs[l], vcs[l], n, rope, k_new=kn[l][:, s].contiguous(),
                                      v_new=vn[l][:, s].contiguous(), out=px.gathered[l, rank])
    px.fill_args(a, l, wait_on_layer=(l - 1) if (l > 0 and os.environ.get("FOLD_WAIT") == "1") else None)
    args.append((a, keep))
ok = True
for step in range(4):
    px.gathered.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    for l in range(L):
        _lib.check(lib.spatten_decode_attn(C.byref(args[l][0]), C.c_void_p(stream.cuda_stream)))
        if os.environ.get("FOLD_WAIT") != "1" or l == L - 1:
            px.wait(l, stream.cuda_stream)
    torch.cuda.synchronize()
    for l in range(L):
        got = gathered_to_heads(px.gathered[l])
        rel = float((got.float() - whole[l].float()).abs().max() / whole[l].float().abs().max())
        if rel > 1e-3:                                   # vs the unsharded launch: same values up to summation order
            ok = False
            print(f"[r{rank}] step {step} layer {l}: differs from the unsharded launch by {rel}", flush=True)
        if not torch.equal(got, full[l]):
            ok = False
            print(f"[r{rank}] step {step} layer {l}: mismatch, max abs {float((got.float() - full[l].float()).abs().max())}", flush=True)
    dist.barrier()
print(f"[r{rank}] fused exchange {'OK' if ok and not px.timed_out() else 'FAILED'} timed_out={px.timed_out()}", flush=True)

# ---- the persistent decode-step kernel with the exchange fused in: chained and independent-layer modes ----------------
from spatten_b200.parallel import PeerStepExchange
ok_step = True
for chain in (True, False):
    kcs2 = [kc[l][:, s].contiguous() for l in range(L)]
    vcs2 = [vc[l][:, s].contiguous() for l in range(L)]
    step = ops.DecodeStep(q[:, :, s].contiguous(), kcs2, vcs2, n, rope, k_new=kn[:, :, s].contiguous(),
                          v_new=vn[:, :, s].contiguous(), chain=chain)
    pxs = PeerStepExchange(step, sh)
    # this rank's heads through the same step kernel without the exchange: the bit-exact reference for every slot
    local = ops.DecodeStep(q[:, :, s].contiguous(), [k.clone() for k in kcs2], [v.clone() for v in vcs2], n, rope,
                           k_new=kn[:, :, s].contiguous(), v_new=vn[:, :, s].contiguous(), chain=chain, sharded_layout=True)
    mine = local.launch().clone()
    torch.cuda.synchronize()
    every = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(every, mine)
    for it in range(3):
        pxs.gathered.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        step.launch()
        torch.cuda.synchronize()
        dist.barrier()
        for l in range(L):
            for r in range(world):
                if not torch.equal(pxs.gathered[l, r], every[r][l]):
                    ok_step = False
                    print(f"[r{rank}] step chain={chain} it {it} layer {l} slot {r}: mismatch", flush=True)
            rel = float((pxs.heads(l).float() - whole[l].float()).abs().max() / whole[l].float().abs().max())
            if rel > 1e-3:
                ok_step = False
                print(f"[r{rank}] step chain={chain} layer {l}: differs from the unsharded launch by {rel}", flush=True)
    if step.timed_out():
        ok_step = False
print(f"[r{rank}] step exchange {'OK' if ok_step else 'FAILED'}", flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0 if (ok and ok_step) else 1)
