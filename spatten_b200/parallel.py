"""Head sharding across GPUs -- host-side plumbing only (torch.distributed; NCCL on GPUs, gloo in CPU tests).

Scores, top-k and the KV gather are per head in the reference (kv_cache_token_pruning.py:51,59-69), so a rank
owns a contiguous block of heads, their KV cache and their prune state; token pruning needs no communication.
The single exchange is on the attention output: every rank needs all heads' [b, H*D] before o_proj.
north_star words it as "one all-reduce on the attention output": each rank contributes its own heads in a
zero-padded [N, b, H/N, D] buffer, so the sum is bit-identical to the single-GPU output.  The same exchange
as an all-gather moves half the bytes; both are offered.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class HeadShard:
    """Contiguous head block [first, first + count) of rank `rank` out of `world`."""
    rank: int
    world: int
    num_heads: int

    def __post_init__(self):
        if self.num_heads % self.world != 0:
            raise ValueError(f"num_heads {self.num_heads} is not divisible by world size {self.world}")

    @property
    def count(self) -> int:
        return self.num_heads // self.world

    @property
    def first(self) -> int:
        return self.rank * self.count

    def slice(self) -> slice:
        return slice(self.first, self.first + self.count)


def exchange_heads(slots: torch.Tensor, shard: HeadShard, mode: str = "allreduce", group=None) -> torch.Tensor:
    """slots: [N, b, H/N, D]; slot `shard.rank` holds this rank's attention output.

    Returns `slots` with every rank's slot filled.  mode "allreduce": the other slots are zeroed and the buffer
    is summed across ranks (the literal north_star collective); "allgather": all_gather_into_tensor.
    Use `gathered_to_heads` to view the result as [b, H, D].
    """
    if slots.shape[0] != shard.world:
        raise ValueError("slots must have one leading slot per rank")
    if shard.world == 1:
        return slots
    if mode == "allgather":
        dist.all_gather(list(slots.unbind(0)), slots[shard.rank].clone(), group=group)
    elif mode == "allreduce":
        for r in range(shard.world):
            if r != shard.rank:
                slots[r].zero_()
        dist.all_reduce(slots, group=group)
    else:
        raise ValueError(f"unknown exchange mode {mode!r}")
    return slots


def gathered_to_heads(slots: torch.Tensor) -> torch.Tensor:
    """[N, b, H/N, D] -> [b, H, D] with heads in global order (rank-major head blocks)."""
    n, b, hl, d = slots.shape
    return slots.permute(1, 0, 2, 3).reshape(b, n * hl, d)


class PeerExchange:
    """Fused output exchange over NVLink peer memory (no collective kernel).

    One symmetric allocation per rank holds the gather buffers of all layers, fp16 [L, N, b, H/N, D], and one
    flag word per (layer, producing rank).  `spatten_decode_attn` stores its output rows straight into slot
    `rank` of every peer's buffer and, when its whole grid is done, bumps flag [layer, rank] on every peer;
    `wait(layer)` (spatten_wait_flags) blocks the stream until all N producers have reported for this step.
    Memory is mapped with torch.distributed._symmetric_memory (CUDA VMM / fabric handles under the hood).
    """

    def __init__(self, layers: int, shard: HeadShard, batch: int, head_dim: int, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self._C, self._lib = C, _lib
        self.shard, self.layers = shard, layers
        n, hl = shard.world, shard.count
        self.slot_elems = batch * hl * head_dim
        out_bytes = layers * n * self.slot_elems * 2
        self.flag_off = (out_bytes + 255) // 256 * 256
        total = self.flag_off + layers * n * 4
        self.buf = symm.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        group = group if group is not None else dist.group.WORLD
        self.hdl = symm.rendezvous(self.buf, group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.gathered = self.buf[:out_bytes].view(torch.float16).view(layers, n, batch, hl, head_dim)
        self.flags = self.buf[self.flag_off:self.flag_off + layers * n * 4].view(torch.int32).view(layers, n)
        self.done = torch.zeros(layers, dtype=torch.int32, device=device)
        self.expected = torch.zeros(layers, dtype=torch.int32, device=device)
        self.timeout = torch.zeros(1, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def fill_args(self, a, layer: int, wait_on_layer=None) -> None:
        """Point a SpattenDecodeArgs at the peers' slots for `layer`; with `wait_on_layer` the kernel itself waits
        for that (earlier) layer's exchange before it reads its query, so no wait launch is needed in between."""
        n, r = self.shard.world, self.shard.rank
        if wait_on_layer is not None:
            a.wait_n = n
            a.wait_flags = self.flags[wait_on_layer].data_ptr()
            a.wait_expected = self.expected[wait_on_layer:].data_ptr()
            a.wait_timeout = self.timeout.data_ptr()
        a.n_peers = n
        for p in range(n):
            a.peer_out[p] = self.ptrs[p] + ((layer * n + r) * self.slot_elems) * 2
            a.peer_flags[p] = self.ptrs[p] + self.flag_off + (layer * n + r) * 4
        a.done_counter = self.done[layer:].data_ptr()

    def wait(self, layer: int, stream_ptr: int) -> None:
        lib = self._lib.load()
        self._lib.check(lib.spatten_wait_flags(self.flags[layer].data_ptr(), self.shard.world,
                                               self.expected[layer:].data_ptr(), self.timeout.data_ptr(),
                                               self._C.c_void_p(stream_ptr)))

    def timed_out(self) -> bool:
        return bool(self.timeout.item())


class PeerStepExchange:
    """Fused output exchange for the persistent decode-step kernel (`spatten_decode_step`, ops.DecodeStep).

    Same symmetric-memory layout as `PeerExchange` -- gather buffers fp16 [L, N, b, H/N, D] and one u32 flag per
    (layer, producing rank) -- but the flags count ROWS: every finished (sequence, head) row is stored into all N
    peers' buffers by a whole warp and lanes 0..N-1 each add 1 to one peer's flag word (release at system scope), so
    there is no grid-wide completion counter on the critical path.  A consumer knows layer l of step s is complete
    when flags[l, r] == rows_per_rank * s for every r.  With `chain=True` the kernel itself waits for layer l's flags
    before it reads layer l+1's query; in every mode it waits for all layers before the launch completes.
    """

    def __init__(self, step, shard: HeadShard, group=None):
        import torch.distributed._symmetric_memory as symm
        a = step.args
        L, b, hl, D = a.num_layers, a.batch, a.num_heads, a.head_dim
        if hl != shard.count:
            raise ValueError("PeerStepExchange: the step must be built over this rank's heads only")
        n = shard.world
        self.shard, self.layers = shard, L
        self.slot_elems = b * hl * D
        out_bytes = L * n * self.slot_elems * 2
        self.flag_off = (out_bytes + 255) // 256 * 256
        total = self.flag_off + L * n * 4
        self.buf = symm.empty(total, dtype=torch.uint8, device=step.device)
        self.buf.zero_()
        group = group if group is not None else dist.group.WORLD
        self.hdl = symm.rendezvous(self.buf, group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.gathered = self.buf[:out_bytes].view(torch.float16).view(L, n, b, hl, D)
        self.flags = self.buf[self.flag_off:self.flag_off + L * n * 4].view(torch.int32).view(L, n)
        r = shard.rank
        a.n_peers, a.self_peer = n, r
        for p in range(n):
            a.peer_out[p] = self.ptrs[p] + r * self.slot_elems * 2            # layer 0, slot `rank` in rank p's buffer
            a.peer_flags[p] = self.ptrs[p] + self.flag_off + r * 4            # flag [0, rank] in rank p's memory
        a.peer_layer_stride = n * self.slot_elems
        a.peer_flag_layer_stride = n
        a.wait_flags = self.flags.data_ptr()
        self.step = step
        torch.cuda.synchronize(step.device)
        dist.barrier(group=group)

    def heads(self, layer: int) -> torch.Tensor:
        """[b, H, D] of `layer` in global head order (valid once the step's launch has completed)."""
        return gathered_to_heads(self.gathered[layer])
