"""CPU checks of the C-ABI boundary: the library loads, exports every symbol include/spatten_b200.h declares,
struct layouts match the ctypes mirrors, and argument validation returns error codes (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "spatten_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spatten_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from spatten_b200 import _lib
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/spatten_b200.h but not exported"


def test_struct_layouts_match_c():
    from spatten_b200 import _lib
    prog = r'''
    #include <stdio.h>
    #include "spatten_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(SpattenDecodeArgs), sizeof(SpattenSelectArgs), sizeof(SpattenGatherArgs),
             sizeof(SpattenStepLayer), sizeof(SpattenStepArgs), sizeof(SpattenPrefillArgs), sizeof(SpattenSparsePvArgs),
             sizeof(SpattenQuantDecodeArgs));
      printf("%zu %zu %zu\n", sizeof(SpattenPruneLayersArgs), sizeof(SpattenPruneLayer), sizeof(SpattenLocalVArgs));
      return 0;
    }'''
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "s.c"), os.path.join(td, "s")
        open(src, "w").write(prog)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = list(map(int, subprocess.check_output([exe]).split()))
    assert sizes == [C.sizeof(_lib.DecodeArgs), C.sizeof(_lib.SelectArgs), C.sizeof(_lib.GatherArgs),
                     C.sizeof(_lib.StepLayer), C.sizeof(_lib.StepArgs), C.sizeof(_lib.PrefillArgs),
                     C.sizeof(_lib.SparsePvArgs), C.sizeof(_lib.QuantDecodeArgs), C.sizeof(_lib.PruneLayersArgs), 40, C.sizeof(_lib.LocalVArgs)]
    assert C.sizeof(_lib.StepLayer) == 64


def test_argument_validation_without_gpu():
    from spatten_b200 import _lib
    lib = _lib.load()
    assert lib.spatten_abi_version() == 1
    a = _lib.DecodeArgs()
    assert lib.spatten_decode_attn(C.byref(a), None) == -1 and b"struct_size" in lib.spatten_last_error()
    a.struct_size = C.sizeof(a)
    a.batch, a.num_heads, a.num_kv_heads, a.head_dim, a.kv_len = 1, 4, 4, 80, 8
    assert lib.spatten_decode_attn(C.byref(a), None) == -2            # head_dim unsupported
    assert lib.spatten_decode_attn(None, None) == -1
    s = _lib.SelectArgs()
    assert lib.spatten_topk_select(C.byref(s), None) == -1
    g = _lib.GatherArgs()
    assert lib.spatten_kv_gather(C.byref(g), None) == -1
    kt, wb, we = C.c_int32(), C.c_int32(), C.c_int32()
    assert lib.spatten_prune_plan(128, 0, 4, 30, 30, C.byref(kt), C.byref(wb), C.byref(we)) == 0
    assert (kt.value, wb.value, we.value) == (64, 4, 98)
    assert lib.spatten_prune_plan(60, 0, 4, 30, 30, None, None, None) == 1
    assert lib.spatten_prune_plan(40, 0, 2, 0, 4, None, None, None) == -1


def test_missing_library_fails_loudly(monkeypatch):
    from spatten_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libspatten_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under spatten_b200/ may import, load or execute it."""
    pkg = os.path.join(ROOT, "spatten_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.\.?oracle)|importlib.*oracle|oracle\.spatten_oracle\.|"
                     r"#include\s+\".*oracle", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not pat.search(src), f"{f} imports / calls the oracle"


def test_cache_buffer_ownership_cpu():
    """In-place append is allowed only behind the NEWEST view of a buffer this package allocated."""
    import torch
    from spatten_b200.cache_buffer import CacheBuffer
    buf = torch.zeros(2, 3, 50, 8, dtype=torch.float16)
    cb = CacheBuffer(buf, 0)
    v20 = cb.view(20)
    assert CacheBuffer.of(v20) is cb and v20.shape == (2, 3, 20, 8)
    v21 = cb.view(21)
    assert CacheBuffer.of(v21) is cb
    assert CacheBuffer.of(v20) is None                     # stale: something was appended behind it already
    assert CacheBuffer.of(buf[:, :, :21]) is None          # a caller's own view of the same storage
    assert CacheBuffer.of(v21[:, :, :10]) is None and CacheBuffer.of(v21.clone()) is None
    assert CacheBuffer.of(None) is None


def test_prefill_and_quant_argument_validation_without_gpu():
    """Error codes of the entry points added in round 2 are decided before any CUDA call."""
    from spatten_b200 import _lib
    lib = _lib.load()
    a = _lib.PrefillArgs()
    assert lib.spatten_prefill_attn(C.byref(a), None) == -1 and b"struct_size" in lib.spatten_last_error()
    a.struct_size = C.sizeof(a)
    a.batch, a.num_heads, a.num_kv_heads, a.head_dim, a.q_len, a.kv_len, a.rope_len = 1, 4, 4, 64, 16, 16, 16
    assert lib.spatten_prefill_attn(C.byref(a), None) == -2            # head_dim unsupported
    a.head_dim, a.kv_len = 128, 8
    assert lib.spatten_prefill_attn(C.byref(a), None) == -1            # kv_len < q_len
    # scratch for the fused path: flag word + one 32 KiB image per (128-row tile) of Q, K and V
    assert lib.spatten_prefill_scratch_bytes(1, 32, 32, 8192, 8192) == 1024 + (32 * 64 + 2 * 32 * 64) * 32768
    assert lib.spatten_prefill_scratch_bytes(2, 8, 2, 129, 300) == 1024 + (2 * 8 * 2 + 2 * 2 * 2 * 3) * 32768
    q = _lib.QuantDecodeArgs()
    assert lib.spatten_decode_attn_quant(C.byref(q), None) == -1
    q.struct_size = C.sizeof(q)
    q.batch, q.num_heads, q.num_kv_heads, q.head_dim, q.kv_len, q.rope_len, q.v_bits = 1, 4, 4, 128, 16, 16, 5
    assert lib.spatten_decode_attn_quant(C.byref(q), None) == -1 and b"v_bits" in lib.spatten_last_error()
    assert lib.spatten_decode_quant_workspace_bytes(1, 32, 128) >= 8192 * (128 + 2) * 4 + 8192 * 4   # partials + tickets
