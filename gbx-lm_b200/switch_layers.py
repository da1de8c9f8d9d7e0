"""The two callers either side of the QuantizedLinear path, over `ops.gather_qmm` /
`ops.quantized_matmul(transpose=False)` (gbxq_gather_qmm, include/gbxq.h):

* `QuantizedSwitchLinear` / `SwitchGLU` -- the expert-indexed projections of the MoE blocks
  (gbx_lm/models/switch_layers.py:26-92,157-203; used by qqwen3_moe.py:111-140): same attribute names, shapes
  (`qweight [E, out, in*bits/32]`, `scales`/`zeros [E, out, in/group_size]`, `bias [E, out]`) and call signature.
* `quantized_scaled_dot_product_attention` -- attention on the group-quantized KV cache
  (gbx_lm/models/base.py:65-98): `scores = q @ K^T` with transpose=True and `out = softmax(scores) @ V` with
  transpose=False, K/V as (codes, scales, biases) triples quantized along head_dim.

CUDA only, like everything in this package."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import ops


class QuantizedSwitchLinear(nn.Module):
    """switch_layers.py:26-92.  Parameters are buffers (frozen, as the reference freezes them)."""

    def __init__(self, input_dims: int, output_dims: int, num_experts: int, bias: bool = True, group_size: int = 64,
                 bits: int = 4, dtype: torch.dtype = torch.bfloat16, device=None):
        super().__init__()
        if bits not in (2, 3, 4, 6, 8) or group_size not in (32, 64, 128):
            raise ValueError("bits must be one of 2, 3, 4, 6, 8 and group_size one of 32, 64, 128")
        if input_dims % 32 or input_dims % group_size:
            raise ValueError("input_dims must be a multiple of 32 and of group_size")
        self.scale = math.sqrt(1 / input_dims)
        self.input_dims, self.output_dims, self.num_experts = input_dims, output_dims, num_experts
        self.group_size, self.bits = group_size, bits
        self.register_buffer("qweight", torch.zeros((num_experts, output_dims, input_dims * bits // 32), dtype=torch.uint32, device=device))
        self.register_buffer("scales", torch.ones((num_experts, output_dims, input_dims // group_size), dtype=dtype, device=device))
        self.register_buffer("zeros", torch.zeros((num_experts, output_dims, input_dims // group_size), dtype=dtype, device=device))
        if bias:
            self.register_buffer("bias", torch.zeros((num_experts, output_dims), dtype=dtype, device=device))
        else:
            self.bias = None

    def set_bias_and_weight(self) -> None:
        """switch_layers.py:61-70: GBA stores w = q*scale - zero, MLX computes q*scale + biases."""
        self.zeros = -self.zeros
        if getattr(self, "q_perm", None) is not None:
            self.q_perm = self.q_perm.reshape(1, 1, -1)

    def forward(self, x: torch.Tensor, indices: torch.Tensor, sorted_indices: bool = False) -> torch.Tensor:
        # x + expand_dims(bias[indices], -2) (switch_layers.py:89-90) is fused as a separately rounded add
        return ops.gather_qmm(x, self.qweight, self.scales, self.zeros, rhs_indices=indices, transpose=True,
                              group_size=self.group_size, bits=self.bits, sorted_indices=sorted_indices, bias=self.bias)


def _gather_sort(x: torch.Tensor, indices: torch.Tensor):
    """switch_layers.py:11-16: order the (token, slot) pairs by expert."""
    m = indices.shape[-1]
    flat = indices.flatten()
    order = torch.argsort(flat, stable=True)
    inv_order = torch.argsort(order)
    return x.flatten(0, -3)[order // m], flat[order], inv_order


def _scatter_unsort(x: torch.Tensor, inv_order: torch.Tensor, shape=None) -> torch.Tensor:
    x = x[inv_order]
    if shape is not None:
        x = x.unflatten(0, tuple(shape))
    return x


class SwitchGLU(nn.Module):
    """switch_layers.py:157-203 with quant=True: down(silu(gate(x)) * up(x)) per (token, selected expert)."""

    def __init__(self, input_dims: int, hidden_dims: int, num_experts: int, bias: bool = False, group_size: int = 64,
                 bits: int = 4, dtype: torch.dtype = torch.bfloat16, device=None):
        super().__init__()
        kw = dict(bias=bias, group_size=group_size, bits=bits, dtype=dtype, device=device)
        self.gate_proj = QuantizedSwitchLinear(input_dims, hidden_dims, num_experts, **kw)
        self.up_proj = QuantizedSwitchLinear(input_dims, hidden_dims, num_experts, **kw)
        self.down_proj = QuantizedSwitchLinear(hidden_dims, input_dims, num_experts, **kw)

    def forward(self, x: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
        x = x.unsqueeze(-2).unsqueeze(-3)
        do_sort = indices.numel() >= 64
        idx, inv_order = indices, None
        if do_sort:
            x, idx, inv_order = _gather_sort(x, indices)
        x_up = self.up_proj(x, idx, sorted_indices=do_sort)
        x_gate = self.gate_proj(x, idx, sorted_indices=do_sort)
        x = self.down_proj(torch.nn.functional.silu(x_gate) * x_up, idx, sorted_indices=do_sort)
        if do_sort:
            x = _scatter_unsort(x, inv_order, indices.shape)
        return x.squeeze(-2)


QTensor = Tuple[torch.Tensor, torch.Tensor, torch.Tensor]


def quantized_scaled_dot_product_attention(queries: torch.Tensor, q_keys: QTensor, q_values: QTensor, scale: float,
                                           mask: Optional[torch.Tensor], group_size: int = 64, bits: int = 8) -> torch.Tensor:
    """base.py:65-98.  queries [B, Hq, L, D]; q_keys / q_values = (codes [B, Hkv, T, D*bits/32], scales, biases
    [B, Hkv, T, D/group_size]).  The softmax is evaluated in fp32 (`precise=True`)."""
    B, n_q_heads, L, D = queries.shape
    n_kv_heads = q_keys[0].shape[-3]
    n_repeats = n_q_heads // n_kv_heads
    queries = queries * scale
    if n_repeats > 1:
        queries = queries.reshape(B, n_kv_heads, n_repeats, L, D)
        q_keys = tuple(t.unsqueeze(-3) for t in q_keys)
        q_values = tuple(t.unsqueeze(-3) for t in q_values)
    scores = ops.quantized_matmul(queries, *q_keys, transpose=True, group_size=group_size, bits=bits)
    if mask is not None:
        scores = scores + mask
    scores = torch.softmax(scores.float(), dim=-1).to(queries.dtype)
    out = ops.quantized_matmul(scores, *q_values, transpose=False, group_size=group_size, bits=bits)
    if n_repeats > 1:
        out = out.reshape(B, n_q_heads, L, D)
    return out
