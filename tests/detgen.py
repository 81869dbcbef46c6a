"""Bit-reproducible synthetic inputs (no dependence on any library's RNG stream).

`det_normal(shape, seed)` hashes the flat element index with splitmix64, sums four 16-bit
uniforms (Irwin-Hall, exact integer arithmetic) and scales to unit variance with one IEEE
multiply, so every platform and library version produces identical bits.  Golden outputs in
tests/golden/*.npz were produced by the reference on exactly these inputs.
"""
import numpy as np
import torch

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def det_normal(shape, seed: int, scale: float = 1.0, dtype=torch.float32) -> torch.Tensor:
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _splitmix64(idx ^ (np.uint64(seed) * np.uint64(0xD6E8FEB86659FD93) & _M64))
    s = ((h & np.uint64(0xFFFF)) + ((h >> np.uint64(16)) & np.uint64(0xFFFF))
         + ((h >> np.uint64(32)) & np.uint64(0xFFFF)) + ((h >> np.uint64(48)) & np.uint64(0xFFFF))).astype(np.int64)
    x = (s - 2 * 65535).astype(np.float64) / 65536.0          # var = 4/12
    x = x * 1.7320508075688772 * float(scale)                  # unit variance * scale
    return torch.from_numpy(x.astype(np.float32)).reshape(shape).to(dtype)
