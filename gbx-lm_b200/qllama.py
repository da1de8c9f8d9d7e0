"""Llama / Mistral decoder built from `QuantizedLinear` -- the CALLERS of the hot path
(reference: gbx_lm/models/qllama.py:39-209).  Same module tree and parameter names as the
reference so a gba2mlx checkpoint loads by key:

    model.embed_tokens.weight, model.layers.N.{input,post_attention}_layernorm.weight,
    model.layers.N.self_attn.{q,k,v,o}_proj.{qweight,scales,zeros[,bias]},
    model.layers.N.mlp.{gate,up,down}_proj.{qweight,scales,zeros}, model.norm.weight, lm_head.weight

Attention / RoPE / RMSNorm are glue written with PyTorch ops; at DECODE (one new token,
cache present) the glue of a block runs as four libgbxq launches instead (rope + cache write,
one-query attention, residual add + RMSNorm, silu * up).  The seven projections per block go through libgbxq.  Tensor parallelism
(new work): q/k/v/gate/up column-parallel, o/down row-parallel + sum all-reduce.
"""
from __future__ import annotations

import inspect
import math
import os
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Union

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .quantized_linear import QuantizedLinear
from .tp import TPContext


@dataclass
class BaseModelArgs:
    @classmethod
    def from_dict(cls, params):  # gbx_lm/models/base.py:15-24
        return cls(**{k: v for k, v in params.items() if k in inspect.signature(cls).parameters})


@dataclass
class ModelArgs(BaseModelArgs):  # gbx_lm/models/qllama.py:15-36
    model_type: str
    hidden_size: int
    num_hidden_layers: int
    intermediate_size: int
    num_attention_heads: int
    rms_norm_eps: float
    vocab_size: int
    head_dim: Optional[int] = None
    max_position_embeddings: Optional[int] = None
    num_key_value_heads: Optional[int] = None
    attention_bias: bool = False
    mlp_bias: bool = False
    rope_theta: float = 10000
    rope_traditional: bool = False
    rope_scaling: Optional[Dict[str, Union[float, str]]] = None
    tie_word_embeddings: bool = True

    def __post_init__(self):
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads


# ------------------------------------------------------------------------------------------ glue
def rope_inv_freq(dims: int, base: float, scaling: Optional[dict]) -> torch.Tensor:
    """Inverse frequencies as the reference's initialize_rope produces them
    (gbx_lm/models/rope_utils.py:197-257: default / linear / llama3)."""
    rope_type = "default"
    if scaling is not None:
        rope_type = scaling.get("type") or scaling.get("rope_type", "default")
    freqs = base ** (torch.arange(0, dims, 2, dtype=torch.float32) / dims)  # "wavelength-like", inv = 1/freqs
    if rope_type == "default":
        return 1.0 / freqs
    if rope_type == "linear":
        return (1.0 / freqs) / float(scaling["factor"])
    if rope_type == "llama3":  # rope_utils.py:70-108
        factor = scaling["factor"]
        low = scaling.get("low_freq_factor", 1.0)
        high = scaling.get("high_freq_factor", 4.0)
        old = scaling.get("original_max_position_embeddings", 8192)
        low_wl, high_wl = old / low, old / high
        wavelens = 2 * math.pi * freqs
        f = torch.where(wavelens > low_wl, freqs * factor, freqs)
        medium = (wavelens > high_wl) & (wavelens < low_wl)
        smooth = (old / wavelens - low) / (high - low)
        smooth_f = f / ((1 - smooth) / factor + smooth)
        return 1.0 / torch.where(medium, smooth_f, f)
    raise ValueError(f"Unsupported RoPE type {rope_type} (yarn/longrope are outside the configs)")


# Decode-step glue through libgbxq (rope_cache / decode_attention / add_rmsnorm / silu_mul) instead of ~45 framework
# launches per block; GBXQ_FUSED_DECODE=0 keeps the plain torch glue (the tests compare the two).
FUSED_DECODE = os.environ.get("GBXQ_FUSED_DECODE", "1") != "0"


def fused_decode_ok(x: torch.Tensor, L: int, cache) -> bool:
    return FUSED_DECODE and L == 1 and cache is not None and x.is_cuda and x.dtype == torch.bfloat16


class RoPE(nn.Module):
    """Non-traditional (half-split) rotary embedding, like mx.fast.rope(traditional=False)."""

    def __init__(self, dims: int, base: float, traditional: bool, scaling: Optional[dict]):
        super().__init__()
        if traditional:
            raise NotImplementedError("rope_traditional=True is not used by the configs")
        self.dims = dims
        self.register_buffer("inv_freq", rope_inv_freq(dims, base, scaling), persistent=False)

    def forward(self, x: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        # x [B, H, L, D]; positions int64 [L]
        ang = positions.to(torch.float32)[:, None] * self.inv_freq[None, :]  # [L, D/2]
        cos, sin = ang.cos()[None, None], ang.sin()[None, None]
        h = self.dims // 2
        x1, x2 = x[..., :h].float(), x[..., h : 2 * h].float()
        out = torch.cat([x1 * cos - x2 * sin, x1 * sin + x2 * cos], dim=-1)
        return out.to(x.dtype)


class KVCache:
    """Static pre-allocated KV cache [B, n_kv, max_len, D] (the reference grows in steps of 256,
    models/cache.py:215-277; a static buffer keeps the decode step CUDA-graph capturable)."""

    def __init__(self, batch: int, n_kv: int, max_len: int, head_dim: int, device, dtype=torch.bfloat16):
        self.keys = torch.zeros((batch, n_kv, max_len, head_dim), device=device, dtype=dtype)
        self.values = torch.zeros((batch, n_kv, max_len, head_dim), device=device, dtype=dtype)
        self.max_len = max_len
        self.offset = 0  # host-side count of valid positions (eager mode)

    def update(self, k: torch.Tensor, v: torch.Tensor, positions: torch.Tensor):
        self.keys.index_copy_(2, positions, k)
        self.values.index_copy_(2, positions, v)


def make_prompt_cache(model, batch: int = 1, max_len: int = 4096) -> List[KVCache]:
    a = model.args
    hd = a.head_dim or a.hidden_size // a.num_attention_heads
    dev = model.model.embed_tokens.weight.device
    n_kv = a.num_key_value_heads // model.tp.world
    return [KVCache(batch, n_kv, max_len, hd, dev) for _ in range(a.num_hidden_layers)]


class Attention(nn.Module):  # gbx_lm/models/qllama.py:39-96
    def __init__(self, args: ModelArgs, tp: TPContext, qkv_bias: Optional[bool] = None):
        super().__init__()
        dim = args.hidden_size
        self.tp = tp
        assert args.num_attention_heads % tp.world == 0 and args.num_key_value_heads % tp.world == 0, \
            "tensor parallel degree must divide the head counts (kv heads bound tp <= 8)"
        self.n_heads = args.num_attention_heads // tp.world
        self.n_kv_heads = args.num_key_value_heads // tp.world
        self.head_dim = head_dim = args.head_dim or args.hidden_size // args.num_attention_heads
        self.scale = head_dim ** -0.5
        attention_bias = getattr(args, "attention_bias", False) if qkv_bias is None else qkv_bias
        o_bias = getattr(args, "attention_bias", False) if qkv_bias is None else False
        self.q_proj = QuantizedLinear(dim, self.n_heads * head_dim, bias=attention_bias)
        self.k_proj = QuantizedLinear(dim, self.n_kv_heads * head_dim, bias=attention_bias)
        self.v_proj = QuantizedLinear(dim, self.n_kv_heads * head_dim, bias=attention_bias)
        self.o_proj = QuantizedLinear(self.n_heads * head_dim, dim, bias=o_bias)
        self.rope = RoPE(head_dim, args.rope_theta, args.rope_traditional, args.rope_scaling)

    def forward(self, x, positions, cache: Optional[KVCache], attend_len: Optional[int]):
        B, L, _ = x.shape
        # one grouped launch for the three projections of x (ref qllama.py:76 calls them back to back)
        q, k, v = ops.quantized_matmul_grouped(x, (self.q_proj, self.k_proj, self.v_proj))
        if fused_decode_ok(x, L, cache) and self.head_dim in (64, 128) and cache.max_len <= 49152:
            # decode step: RoPE + cache write in one launch, one-query attention in another
            qh = q.view(B, self.n_heads, self.head_dim)
            ops.rope_cache(qh, k.view(B, self.n_kv_heads, self.head_dim), v.view(B, self.n_kv_heads, self.head_dim),
                           positions, self.rope.inv_freq, cache.keys, cache.values)
            out = ops.decode_attention(qh, cache.keys, cache.values, positions, self.scale, attend_len)
            return self.tp.row_parallel(self.o_proj, out.view(B, 1, -1))
        q = q.view(B, L, self.n_heads, -1).transpose(1, 2)
        k = k.view(B, L, self.n_kv_heads, -1).transpose(1, 2)
        v = v.view(B, L, self.n_kv_heads, -1).transpose(1, 2)
        q = self.rope(q, positions)
        k = self.rope(k, positions)
        if cache is not None:
            cache.update(k, v, positions)
            T = attend_len if attend_len is not None else cache.max_len
            keys, values = cache.keys[:, :, :T], cache.values[:, :, :T]
        else:
            T = L
            keys, values = k, v
        # causal over absolute positions: key j visible to query at position p iff j <= p
        mask = torch.arange(T, device=x.device)[None, :] <= positions[:, None]
        out = F.scaled_dot_product_attention(q, keys, values, attn_mask=mask[None, None], scale=self.scale,
                                             enable_gqa=self.n_heads != self.n_kv_heads)
        out = out.transpose(1, 2).reshape(B, L, -1)
        return self.tp.row_parallel(self.o_proj, out)


class MLP(nn.Module):  # gbx_lm/models/qllama.py:99-115
    def __init__(self, args: ModelArgs, tp: TPContext):
        super().__init__()
        dim, hidden_dim = args.hidden_size, args.intermediate_size
        assert hidden_dim % tp.world == 0
        self.tp = tp
        mlp_bias = getattr(args, "mlp_bias", False)
        self.gate_proj = QuantizedLinear(dim, hidden_dim // tp.world, bias=mlp_bias)
        self.down_proj = QuantizedLinear(hidden_dim // tp.world, dim, bias=mlp_bias)
        self.up_proj = QuantizedLinear(dim, hidden_dim // tp.world, bias=mlp_bias)

    def forward(self, x):
        gate, up = ops.quantized_matmul_grouped(x, (self.gate_proj, self.up_proj))  # ref qllama.py:115
        if gate.is_cuda and gate.dtype == torch.bfloat16 and gate.numel() % 8 == 0 and FUSED_DECODE:
            act = ops.silu_mul(gate, up)
        else:
            act = F.silu(gate) * up
        return self.tp.row_parallel(self.down_proj, act)


class TransformerBlock(nn.Module):  # gbx_lm/models/qllama.py:118-141
    def __init__(self, args: ModelArgs, tp: TPContext, qkv_bias: Optional[bool] = None):
        super().__init__()
        self.self_attn = Attention(args, tp, qkv_bias)
        self.mlp = MLP(args, tp)
        self.input_layernorm = nn.RMSNorm(args.hidden_size, eps=args.rms_norm_eps)
        self.post_attention_layernorm = nn.RMSNorm(args.hidden_size, eps=args.rms_norm_eps)

    def forward(self, x, positions, cache, attend_len):
        h = x + self.self_attn(self.input_layernorm(x), positions, cache, attend_len)
        return h + self.mlp(self.post_attention_layernorm(h))


class LlamaModel(nn.Module):  # gbx_lm/models/qllama.py:144-174
    def __init__(self, args: ModelArgs, tp: TPContext, qkv_bias: Optional[bool] = None):
        super().__init__()
        assert args.vocab_size > 0
        self.embed_tokens = nn.Embedding(args.vocab_size, args.hidden_size)
        self.layers = nn.ModuleList([TransformerBlock(args, tp, qkv_bias) for _ in range(args.num_hidden_layers)])
        self.norm = nn.RMSNorm(args.hidden_size, eps=args.rms_norm_eps)

    def forward(self, inputs, positions, cache, attend_len):
        h = self.embed_tokens(inputs)
        if cache is None:
            cache = [None] * len(self.layers)
        if fused_decode_ok(h, inputs.shape[1], cache[0]) and h.shape[-1] % 8 == 0:
            # decode step: every residual add travels with the RMSNorm that follows it (one launch each); the add of a
            # block's MLP output happens in the next block's (or the final) norm launch
            resid, pend = h.contiguous(), None
            for layer, c in zip(self.layers, cache):
                resid, n = ops.add_rmsnorm(resid, pend, layer.input_layernorm.weight, layer.input_layernorm.eps)
                a = layer.self_attn(n, positions, c, attend_len)
                resid, n = ops.add_rmsnorm(resid, a, layer.post_attention_layernorm.weight, layer.post_attention_layernorm.eps)
                pend = layer.mlp(n)
            return ops.add_rmsnorm(resid, pend, self.norm.weight, self.norm.eps, want_h=False)[1]
        for layer, c in zip(self.layers, cache):
            h = layer(h, positions, c, attend_len)
        return self.norm(h)


class Model(nn.Module):  # gbx_lm/models/qllama.py:177-209
    QKV_BIAS: Optional[bool] = None  # None -> follow args.attention_bias (Llama); Qwen2 overrides

    def __init__(self, args: ModelArgs, tp: Optional[TPContext] = None):
        super().__init__()
        self.args = args
        self.model_type = args.model_type
        self.tp = tp or TPContext()
        self.model = LlamaModel(args, self.tp, self.QKV_BIAS)
        if not args.tie_word_embeddings:
            self.lm_head = nn.Linear(args.hidden_size, args.vocab_size, bias=False)

    def forward(self, inputs: torch.Tensor, cache: Optional[List[KVCache]] = None,
                positions: Optional[torch.Tensor] = None, attend_len: Optional[int] = None, hidden_states: bool = False):
        """inputs int64 [B, L].  positions int64 [L] (absolute); default = cache offset .. offset+L.
        attend_len: number of cache positions attended (None = whole static cache, for graph replay)."""
        L = inputs.shape[1]
        if positions is None:
            off = cache[0].offset if cache is not None else 0
            positions = torch.arange(off, off + L, device=inputs.device)
            if cache is not None:
                attend_len = off + L
                for c in cache:
                    c.offset = off + L
        out = self.model(inputs, positions, cache, attend_len)
        # vocabulary projection (qllama.py:194-198): decode-sized inputs go through the streaming bf16 GEMV
        # (gbxq_head_gemv), prefill through the dense matmul
        head = self.model.embed_tokens.weight if self.args.tie_word_embeddings else self.lm_head.weight
        logits = ops.head_linear(out, head) if FUSED_DECODE else F.linear(out, head)
        return (logits, out) if hidden_states else logits

    def sanitize(self, weights: Dict[str, Any]) -> Dict[str, Any]:  # qllama.py:201-205
        return {k: v for k, v in weights.items() if "self_attn.rotary_emb.inv_freq" not in k}

    @property
    def layers(self):
        return self.model.layers
