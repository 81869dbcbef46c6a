"""Public entry point, same call as the reference's (spatten_llm/enable_spatten_llm.py:5-23)."""
from .kv_cache_token_pruning import SpAttenKVCache
from .pos_shift.modify_llama import enable_llama_pos_shift_attention

__all__ = ["enable_spatten_llm"]


def enable_spatten_llm(model, start_size, important_size, recent_size):
    """Patch every LlamaAttention of `model` onto the CUDA path and return the pruning cache helper.

    Only llama-type models are supported, as in the reference (:6, :13-14): anything else raises
    ValueError(f"got {model_type}").  Llama caches are [batch, heads, seq, head_dim] -> seq dim 2 (:7).
    """
    model_type = model.config.model_type
    if "llama" not in model_type:
        raise ValueError(f"got {model_type}")
    enable_llama_pos_shift_attention(model)
    return SpAttenKVCache(start_size=start_size, important_size=important_size, recent_size=recent_size,
                          k_seq_dim=2, v_seq_dim=2)
