"""ctypes binding of the C-ABI library (include/spatten_b200.h).

The product path has no CPU fallback: if `lib/libspatten_b200.so` is missing or a call fails this
module raises.  Build with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C spatten_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libspatten_b200.so")

SPATTEN_OK = 0
DECODE_PDL = 1
DECODE_PDL_KV_STABLE = 2
DECODE_NO_CLUSTER = 4
DECODE_SCORES_ONLY = 8
DECODE_SMALL_TILES = 16
STEP_CHAIN = 1
STEP_SHARDED_LAYOUT = 2
ERR_NAMES = {-1: "SPATTEN_ERR_INVALID", -2: "SPATTEN_ERR_UNSUPPORTED", -3: "SPATTEN_ERR_CUDA",
             -4: "SPATTEN_ERR_WORKSPACE"}


class SpattenError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class DecodeArgs(C.Structure):
    """Mirror of `SpattenDecodeArgs` (include/spatten_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("batch", C.c_int32), ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32), ("head_dim", C.c_int32),
        ("kv_len", C.c_int32), ("rope_len", C.c_int32), ("append", C.c_int32), ("flags", C.c_uint32), ("q_len", C.c_int32),
        ("q", C.c_void_p), ("k_new", C.c_void_p), ("v_new", C.c_void_p),
        ("k_cache", C.c_void_p), ("v_cache", C.c_void_p),
        ("cache_stride_b", C.c_int64), ("cache_stride_h", C.c_int64),
        ("rope", C.c_void_p), ("q_pos", C.c_void_p), ("mask", C.c_void_p), ("head_mask", C.c_void_p),
        ("out", C.c_void_p), ("logits", C.c_void_p), ("logits_stride", C.c_int64),
        ("logits_stride_b", C.c_int64), ("logits_stride_q", C.c_int64),
        ("importance", C.c_void_p), ("importance_stride", C.c_int64), ("head_importance", C.c_void_p),
        ("n_peers", C.c_int32), ("peer_out", C.c_void_p * 8), ("peer_flags", C.c_void_p * 8),
        ("done_counter", C.c_void_p),
        ("wait_n", C.c_int32), ("wait_flags", C.c_void_p), ("wait_expected", C.c_void_p), ("wait_timeout", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class StepLayer(C.Structure):
    """Mirror of `SpattenStepLayer` (64 bytes; the array lives in device memory)."""
    _fields_ = [("q", C.c_void_p), ("k_new", C.c_void_p), ("v_new", C.c_void_p), ("k_cache", C.c_void_p),
                ("v_cache", C.c_void_p), ("out", C.c_void_p), ("logits", C.c_void_p), ("importance", C.c_void_p)]


class StepArgs(C.Structure):
    """Mirror of `SpattenStepArgs`."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("num_layers", C.c_int32), ("batch", C.c_int32), ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("kv_len", C.c_int32), ("rope_len", C.c_int32), ("append", C.c_int32),
        ("flags", C.c_uint32),
        ("layers", C.c_void_p),
        ("cache_stride_b", C.c_int64), ("cache_stride_h", C.c_int64),
        ("logits_stride", C.c_int64), ("importance_stride", C.c_int64),
        ("rope", C.c_void_p), ("q_pos", C.c_void_p), ("head_mask", C.c_void_p), ("head_importance", C.c_void_p),
        ("n_peers", C.c_int32), ("self_peer", C.c_int32), ("peer_out", C.c_void_p * 8), ("peer_layer_stride", C.c_int64),
        ("peer_flags", C.c_void_p * 8), ("peer_flag_layer_stride", C.c_int64), ("wait_flags", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class PrefillArgs(C.Structure):
    """Mirror of `SpattenPrefillArgs`."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("batch", C.c_int32), ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("q_len", C.c_int32), ("kv_len", C.c_int32), ("rope_len", C.c_int32),
        ("q", C.c_void_p), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p),
        ("cache_stride_b", C.c_int64), ("cache_stride_h", C.c_int64),
        ("rope", C.c_void_p), ("q_pos", C.c_void_p), ("mask", C.c_void_p),
        ("logits", C.c_void_p), ("logits_stride_b", C.c_int64), ("logits_stride_h", C.c_int64),
        ("logits_stride_q", C.c_int64), ("probs", C.c_void_p), ("out", C.c_void_p),
        ("importance", C.c_void_p), ("importance_stride", C.c_int64),
        ("flags", C.c_uint32), ("reserved0", C.c_uint32), ("scratch_bytes", C.c_int64),
    ]


PREFILL_CAUSAL = 1
PREFILL_LEGACY = 2


class SelectArgs(C.Structure):
    """Mirror of `SpattenSelectArgs`."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("num_rows", C.c_int32), ("seq_len", C.c_int32),
        ("num_coming", C.c_int32), ("start", C.c_int32), ("important", C.c_int32), ("recent", C.c_int32),
        ("score_dtype", C.c_int32),
        ("score", C.c_void_p), ("score_stride", C.c_int64),
        ("keep_idx", C.c_void_p), ("keep_stride", C.c_int64),
    ]


class GatherArgs(C.Structure):
    """Mirror of `SpattenGatherArgs`."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("batch", C.c_int32), ("num_heads", C.c_int32), ("keep_total", C.c_int32),
        ("row_bytes", C.c_int32),
        ("src_k", C.c_void_p), ("src_v", C.c_void_p), ("src_stride_b", C.c_int64), ("src_stride_h", C.c_int64),
        ("dst_k", C.c_void_p), ("dst_v", C.c_void_p), ("dst_stride_b", C.c_int64), ("dst_stride_h", C.c_int64),
        ("keep_idx", C.c_void_p), ("keep_stride", C.c_int64),
    ]


class PruneLayersArgs(C.Structure):
    """Mirror of `SpattenPruneLayersArgs` (the per-layer table is a device array of 5 pointers per layer)."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("num_layers", C.c_int32), ("batch", C.c_int32), ("num_heads", C.c_int32),
        ("seq_len", C.c_int32), ("num_coming", C.c_int32), ("start", C.c_int32), ("important", C.c_int32),
        ("recent", C.c_int32), ("score_dtype", C.c_int32), ("row_bytes", C.c_int32),
        ("layers", C.c_void_p), ("score_stride", C.c_int64),
        ("src_stride_b", C.c_int64), ("src_stride_h", C.c_int64), ("dst_stride_b", C.c_int64), ("dst_stride_h", C.c_int64),
        ("keep_idx", C.c_void_p), ("keep_stride", C.c_int64),
    ]


class SparsePvArgs(C.Structure):
    """Mirror of `SpattenSparsePvArgs`."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("batch", C.c_int32), ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("keep", C.c_int32),
        ("logits", C.c_void_p), ("logits_stride", C.c_int64),
        ("keep_idx", C.c_void_p), ("keep_stride", C.c_int64),
        ("v_cache", C.c_void_p), ("cache_stride_b", C.c_int64), ("cache_stride_h", C.c_int64),
        ("out", C.c_void_p),
    ]


class LocalVArgs(C.Structure):
    """Mirror of `SpattenLocalVArgs`."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("batch", C.c_int32), ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("kv_len", C.c_int32), ("rope_len", C.c_int32), ("keep", C.c_int32),
        ("q", C.c_void_p), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p),
        ("cache_stride_b", C.c_int64), ("cache_stride_h", C.c_int64),
        ("rope", C.c_void_p), ("q_pos", C.c_void_p), ("out", C.c_void_p), ("logits", C.c_void_p),
        ("logits_stride", C.c_int64), ("keep_idx", C.c_void_p), ("keep_stride", C.c_int64),
    ]


class QuantKV(C.Structure):
    """Mirror of `SpattenQuantKV`."""
    _fields_ = [
        ("k_msb", C.c_void_p), ("k_lsb", C.c_void_p), ("k_scale", C.c_void_p),
        ("v_msb", C.c_void_p), ("v_lsb", C.c_void_p), ("v_scale", C.c_void_p),
        ("plane_stride_b", C.c_int64), ("plane_stride_h", C.c_int64),
        ("scale_stride_b", C.c_int64), ("scale_stride_h", C.c_int64),
    ]


class QuantDecodeArgs(C.Structure):
    """Mirror of `SpattenQuantDecodeArgs`."""
    _fields_ = [
        ("struct_size", C.c_uint32), ("batch", C.c_int32), ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("kv_len", C.c_int32), ("rope_len", C.c_int32), ("v_bits", C.c_int32),
        ("requant_threshold", C.c_float),
        ("kv", QuantKV),
        ("q", C.c_void_p), ("rope", C.c_void_p), ("q_pos", C.c_void_p), ("out", C.c_void_p),
        ("logits", C.c_void_p), ("logits_stride", C.c_int64), ("need_requant", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


DTYPE_F16, DTYPE_F32 = 0, 1

_lib = None


def load() -> C.CDLL:
    """Load the shared library once; raise loudly when it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"spatten_b200: CUDA library not built ({LIB_PATH} missing). Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C spatten_b200/csrc`. "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.spatten_abi_version.restype = C.c_int
    lib.spatten_last_error.restype = C.c_char_p
    lib.spatten_device_info.argtypes = [C.POINTER(C.c_int)] * 3
    lib.spatten_decode_workspace_bytes.restype = C.c_size_t
    lib.spatten_decode_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.spatten_decode_attn.argtypes = [C.POINTER(DecodeArgs), C.c_void_p]
    lib.spatten_decode_attn_host.argtypes = [C.POINTER(DecodeArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
    _bind_optional(lib)
    if lib.spatten_abi_version() != 1:
        raise RuntimeError(f"spatten_b200: ABI version {lib.spatten_abi_version()} != 1")
    _lib = lib
    return lib


def _bind_optional(lib):
    lib.spatten_prune_plan.argtypes = [C.c_int32] * 5 + [C.POINTER(C.c_int32)] * 3
    lib.spatten_score_reduce.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    lib.spatten_topk_select.argtypes = [C.POINTER(SelectArgs), C.c_void_p]
    lib.spatten_kv_gather.argtypes = [C.POINTER(GatherArgs), C.c_void_p]
    lib.spatten_prune_layers.argtypes = [C.POINTER(PruneLayersArgs), C.c_void_p]
    lib.spatten_sparse_pv.argtypes = [C.POINTER(SparsePvArgs), C.c_void_p]
    lib.spatten_local_v_attn.argtypes = [C.POINTER(LocalVArgs), C.c_void_p]
    lib.spatten_prefill_attn.argtypes = [C.POINTER(PrefillArgs), C.c_void_p]
    lib.spatten_prefill_scratch_bytes.argtypes = [C.c_int32] * 5
    lib.spatten_prefill_scratch_bytes.restype = C.c_int64
    lib.spatten_wait_flags.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.spatten_kv_quantize.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.POINTER(QuantKV), C.c_int32, C.c_void_p]
    lib.spatten_decode_quant_workspace_bytes.restype = C.c_size_t
    lib.spatten_decode_quant_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.spatten_decode_attn_quant.argtypes = [C.POINTER(QuantDecodeArgs), C.c_void_p]
    lib.spatten_head_select.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.spatten_decode_step_workspace_bytes.restype = C.c_size_t
    lib.spatten_decode_step_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    lib.spatten_decode_step.argtypes = [C.POINTER(StepArgs), C.c_void_p]
    lib.spatten_decode_step_host.argtypes = [C.POINTER(StepArgs), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                             C.c_size_t, C.c_void_p]


def check(code: int):
    if code != SPATTEN_OK:
        raise SpattenError(code, load().spatten_last_error().decode("utf-8", "replace"))
