"""Qwen2 / Qwen2.5 decoder built from `QuantizedLinear` (reference: gbx_lm/models/qqwen2.py:15-191).
Identical block structure to Llama; q/k/v carry a bias (qqwen2.py:44-46), o_proj does not (:47)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Union

from . import qllama
from .qllama import BaseModelArgs


@dataclass
class ModelArgs(BaseModelArgs):  # gbx_lm/models/qqwen2.py:15-29
    model_type: str
    hidden_size: int
    num_hidden_layers: int
    intermediate_size: int
    num_attention_heads: int
    rms_norm_eps: float
    vocab_size: int
    num_key_value_heads: int
    max_position_embeddings: int = 32768
    rope_theta: float = 1000000
    rope_traditional: bool = False
    rope_scaling: Optional[Dict[str, Union[float, str]]] = None
    tie_word_embeddings: bool = True
    head_dim: Optional[int] = None


class Model(qllama.Model):
    QKV_BIAS = True

    def sanitize(self, weights):  # qqwen2.py:180-187
        if self.args.tie_word_embeddings:
            weights.pop("lm_head.weight", None)
        return {k: v for k, v in weights.items() if "self_attn.rotary_emb.inv_freq" not in k}
