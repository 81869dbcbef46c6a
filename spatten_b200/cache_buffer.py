"""Ownership record for KV-cache capacity buffers allocated by this package."""
from __future__ import annotations

import torch


class CacheBuffer:
    """A capacity buffer [b, H_kv, cap, D] this package allocated, plus how many rows of it are committed.

    The reference builds a fresh `present` with torch.cat every call and never mutates `past`
    (modify_llama.py:95-100).  The kernel appends in place instead, which is only safe when (1) the buffer is ours
    (never a caller's tensor: a truncated view, a fused kv layout or a shared storage would be overwritten) and
    (2) nobody has appended behind the view being extended yet (calling forward twice with the same `past` --
    prefix reuse, candidate scoring -- must not overwrite the first call's `present`).  Views handed out by this
    package carry their CacheBuffer in the `_spatten_cache` attribute; anything else, and any view that is no
    longer the newest one, is copied into a fresh buffer first.  `present` therefore aliases `past` exactly when
    `past` was the newest view of one of our buffers.
    """
    __slots__ = ("buf", "committed")

    def __init__(self, buf: torch.Tensor, committed: int):
        self.buf, self.committed = buf, committed

    def view(self, length: int) -> torch.Tensor:
        v = self.buf[:, :, :length]
        v._spatten_cache = self
        self.committed = length
        return v

    @staticmethod
    def of(t):
        """The CacheBuffer `t` is the newest view of, or None (caller-owned tensor / stale view)."""
        cb = getattr(t, "_spatten_cache", None) if t is not None else None
        if cb is None or cb.committed != t.shape[2] or t.data_ptr() != cb.buf.data_ptr():
            return None
        return cb
