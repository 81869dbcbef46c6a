"""world_size-2 gloo test (CPU) of the head-sharding plumbing: sharded attention + exchange == unsharded."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from detgen import det_normal
        from oracle import spatten_oracle as O
        from spatten_b200.parallel import HeadShard, exchange_heads, gathered_to_heads
        b, H, D, past = 2, 8, 64, 37
        q = det_normal((b, H, 1, D), 1, 1.0, torch.float32)
        k = det_normal((b, H, 1, D), 2, 1.0, torch.float32)
        v = det_normal((b, H, 1, D), 3, 1.0, torch.float32)
        pk = det_normal((b, H, past, D), 4, 1.0, torch.float32)
        pv = det_normal((b, H, past, D), 5, 1.0, torch.float32)
        pos = torch.full((b, 1), past, dtype=torch.int64)
        full, full_scores, _ = O.attn_forward_core(q, k, v, pk, pv, pos, None)
        sh = HeadShard(rank, world, H)
        s = sh.slice()
        loc, loc_scores, _ = O.attn_forward_core(q[:, s], k[:, s], v[:, s], pk[:, s], pv[:, s], pos, None)
        slots = torch.full((world, b, sh.count, D), 123.0)            # garbage in the peers' slots
        slots[rank] = loc[:, :, 0]
        exchange_heads(slots, sh, mode)
        got = gathered_to_heads(slots)
        ok = torch.equal(got, full[:, :, 0]) and torch.equal(loc_scores, full_scores[:, s])
        # pruning is per head: selecting on the shard's scores equals the shard of the global selection
        score = full_scores[0, :, 0]
        keep_all = O.select_keep_indices(score, past + 1, 2, 10, 10, 0)
        keep_loc = O.select_keep_indices(score[s], past + 1, 2, 10, 10, 0)
        ok = ok and torch.equal(keep_all[s], keep_loc)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["allreduce", "allgather"])
def test_head_sharding_exchange_world2(mode):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True


def test_head_shard_arithmetic():
    sys.path.insert(0, ROOT)
    from spatten_b200.parallel import HeadShard, gathered_to_heads
    sh = HeadShard(3, 8, 40)
    assert (sh.first, sh.count, sh.slice()) == (15, 5, slice(15, 20))
    with pytest.raises(ValueError):
        HeadShard(0, 3, 32)
    x = torch.arange(2 * 3 * 2 * 1).reshape(2, 3, 2, 1)          # [N=2, b=3, Hl=2, D=1]
    g = gathered_to_heads(x)
    assert g.shape == (3, 4, 1) and g[1, :, 0].tolist() == [2, 3, 8, 9]
