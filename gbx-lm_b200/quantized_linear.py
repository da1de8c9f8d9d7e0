"""`QuantizedLinear` -- host-side mirror of the reference module
`gbx_lm/models/quantized_linear_gba.py::QuantizedLinear` (:10-348) on top of libgbxq.

Same constructor, attribute names (`qweight/scales/zeros/bias/bits/group_size/input_dims/
output_dims`, `channel_scale`, optional `q_perm`), shapes and dtypes, and the same classmethods
(`reinit_module`, `prepare_scales_zeros`, `post_processing_and_release`, `from_linear`), so that a
checkpoint written by gba2mlx (`*.qweight`, `*.scales`, `*.zeros`; gbx_lm/utils.py:828-843,1089)
loads by key.  The forward is one call into the sm_100a library; there is no other code path.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch
from torch import nn

from . import ops

PROJ_KEYS = (  # match order of quantized_linear_gba.py:261 (first substring hit wins)
    "kv_a_proj_with_mqa", "kv_b_proj", "q_a_proj", "q_b_proj", "q_proj", "k_proj", "v_proj", "o_proj",
    "gate_proj", "up_proj", "down_proj", "qkv_proj", "gate_up_proj",
)


def strategy_params(name: str, strategy: Optional[dict], bits: int, group_size: int):
    """(bits, group_size) of the QuantizedLinear called `name` (`model.layers.<i>....<proj>`): its entry of
    `quant_strategy.json`'s "measurement" table, matched as `reinit_module` matches it (quantized_linear_gba.py:258-272),
    or the uniform defaults when there is no strategy (:278-281).  KeyError when the block has no entry for it."""
    if strategy is None:
        return bits, group_size
    block = strategy["model.layers.{}".format(name.split(".")[2])]
    for key in PROJ_KEYS:
        if key in name:
            if "shared_expert" in name:
                key = "moe_shared_expert_" + key
            if key in block:
                b = block[key]["bits"][0]
                return b, block[key]["group_size"][str(b)]
    raise KeyError(f"quant_strategy.json has no entry for {name}")


class QuantizedLinear(nn.Module):
    """y = x . dequant(qweight)^T (+ bias) with MLX group-affine packed weights.

    Args mirror quantized_linear_gba.py:36-43.  Parameters are frozen buffers (the reference
    freezes them, :57-58,162-166)."""

    def __init__(self, input_dims: int, output_dims: int, bias: bool = False, group_size: int = 64, bits: int = 4):
        super().__init__()
        self.group_size = group_size
        self.bits = bits
        self.output_dims = output_dims
        self.input_dims = input_dims
        self.double_group_size = 32
        if bias:
            self.register_buffer("bias", torch.zeros((output_dims,), dtype=torch.bfloat16))
        else:
            self.bias = None
        self.init_params(False, False)

    # ------------------------------------------------------------------ allocation (ref :60-117)
    def init_params(self, use_double_quantization: bool, use_q_perm: bool):
        if self.bits not in (2, 3, 4, 6, 8):
            raise ValueError(f"bits must be one of 2, 3, 4, 6, 8; got {self.bits}")
        dev = self.qweight.device if isinstance(getattr(self, "qweight", None), torch.Tensor) else None
        shape_w = (self.output_dims, self.input_dims // 32 * self.bits)
        shape_sz = (self.output_dims, self.input_dims // self.group_size)
        self._set("qweight", torch.zeros(shape_w, dtype=torch.uint32, device=dev))
        self._set("channel_scale", torch.ones((1, 1, self.input_dims), dtype=torch.float16, device=dev))
        if use_q_perm:
            self._set("q_perm", torch.zeros((1, 1, self.input_dims), dtype=torch.int16, device=dev))
        if use_double_quantization:
            g = math.ceil(self.input_dims / self.group_size)
            d = math.ceil(self.output_dims / self.double_group_size)
            self._set("qstatistic", torch.zeros((g, d, self.double_group_size), dtype=torch.uint8, device=dev))
            for name, fill in (("qzeros_zeros", 0.0), ("qzeros_scales", 1.0), ("qscales_zeros", 0.0), ("qscales_scales", 1.0)):
                self._set(name, torch.full((g, d, 1), fill, dtype=torch.float16, device=dev))
            for name in ("scales", "zeros"):
                if name in self._buffers:
                    del self._buffers[name]
        else:
            self._set("scales", torch.ones(shape_sz, dtype=torch.float16, device=dev))
            self._set("zeros", torch.zeros(shape_sz, dtype=torch.float16, device=dev))

    def _set(self, name: str, t: Optional[torch.Tensor]):
        if name in self._buffers:
            self._buffers[name] = t
        else:
            if hasattr(self, name):
                delattr(self, name)
            self.register_buffer(name, t)

    # MLX-style aliases named in the task statement (read-only views of the same storage)
    @property
    def weight(self) -> torch.Tensor:
        return self.qweight

    @property
    def biases(self) -> torch.Tensor:
        return self.zeros

    # ------------------------------------------------------------------ conversion helpers
    def create_scales_zeros(self):
        """Expand GBA double-quantised statistics into fp16 scales/zeros (ref :119-149)."""
        g = math.ceil(self.input_dims / self.group_size)
        qs = self.qstatistic
        qscales = (qs & 0xF0) >> 4
        qzeros = qs & 0x0F
        zeros = ((qzeros.to(torch.float16) - self.qzeros_zeros) * self.qzeros_scales).reshape(g, self.output_dims)
        scales = ((qscales.to(torch.float16) - self.qscales_zeros) * self.qscales_scales).reshape(g, self.output_dims)
        self._set("scales", scales.t().contiguous())
        self._set("zeros", zeros.t().contiguous())
        for name in ("qstatistic", "qzeros_zeros", "qzeros_scales", "qscales_zeros", "qscales_scales"):
            del self._buffers[name]

    def set_bias_and_weight(self):
        """GBA stores a subtractive zero-point, MLX adds its `biases`: negate (ref :151-160)."""
        self._set("zeros", -self.zeros)
        if "q_perm" in self._buffers:
            self._set("q_perm", self.q_perm.reshape(1, 1, -1))

    def extra_repr(self) -> str:
        # the reference derives input_dims as shape[1] * (32 // bits), wrong for 3/6-bit (ref :169-172;
        # SURVEY 8a quirks); the exact formula is shape[1] * 32 // bits
        out_dims, words = self.qweight.shape
        in_dims = words * 32 // self.bits
        return (
            f"input_dims={in_dims}, output_dims={out_dims}, bias={self.bias is not None}, "
            f"group_size={self.group_size}, bits={self.bits}"
        )

    # ------------------------------------------------------------------ forward (ref :179-207)
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.quantized_matmul(
            x, self.qweight, self.scales, self.zeros, transpose=True, group_size=self.group_size, bits=self.bits,
            bias=self.bias,
        )

    # ------------------------------------------------------------------ per-layer bits / group size (ref :210-283)
    @classmethod
    def reinit_module(
        cls,
        model: nn.Module,
        group_size: int = 64,
        bits: int = 4,
        strategy: Optional[dict] = None,
        use_double_quantization: bool = False,
        use_q_perm: bool = False,
        gba_linear_class_predicate: Callable[[nn.Module], bool] = lambda m: isinstance(m, QuantizedLinear),
    ):
        if strategy is None:
            for m in model.modules():
                if gba_linear_class_predicate(m):
                    m.group_size = group_size
                    m.bits = bits
                    m.init_params(use_double_quantization, use_q_perm)
            return
        for name, child in model.named_modules():
            if not isinstance(child, QuantizedLinear):
                continue
            layer_number = name.split(".")[2]
            strategy_per_block = strategy["model.layers.{}".format(layer_number)]
            strg = None
            for key in PROJ_KEYS:
                if key in name:
                    if "shared_expert" in name:
                        key = "moe_shared_expert_" + key
                    if key in strategy_per_block:
                        strg = strategy_per_block[key]
                        break
            if strg is None:
                # the reference silently reuses the previous layer's entry here; refuse instead
                raise KeyError(f"quant_strategy.json has no entry for {name}")
            child.bits = strg["bits"][0]
            child.group_size = strg["group_size"][str(child.bits)]
            assert child.group_size in [32, 64, 128], f"The group size value ({child.group_size}) must be 32, 64 or 128."
            child.init_params(use_double_quantization, use_q_perm)

    @classmethod
    def prepare_scales_zeros(cls, model: nn.Module, gba_linear_class_predicate=lambda m: isinstance(m, QuantizedLinear)):
        for m in model.modules():
            if gba_linear_class_predicate(m):
                m.create_scales_zeros()

    @classmethod
    def post_processing_and_release(cls, model: nn.Module, gba_linear_class_predicate=lambda m: isinstance(m, QuantizedLinear)):
        for m in model.modules():
            if gba_linear_class_predicate(m):
                m.set_bias_and_weight()

    @classmethod
    def from_linear(cls, linear_layer: nn.Module, group_size: int = 64, bits: int = 4, q_perm=None, channel_scale=None):
        """Quantise a dense layer (ref :327-348, whose `weight` is unbound unless q_perm is given --
        SURVEY 8a quirks; here the un-permuted case works).  Uses the affine recipe of mx.quantize."""
        from .packing import quantize_affine

        weight = linear_layer.weight.detach()
        output_dims, input_dims = weight.shape
        ql = cls(input_dims, output_dims, False, group_size, bits)
        if q_perm is not None:
            weight = weight[:, q_perm.reshape(-1).long()]
            ql._set("q_perm", q_perm.reshape(1, 1, -1))
        if channel_scale is not None:
            ql._set("channel_scale", channel_scale)
        qw, s, z = quantize_affine(weight.to(torch.bfloat16), group_size, bits)
        ql._set("qweight", qw)
        ql._set("scales", s)
        ql._set("zeros", z)
        if getattr(linear_layer, "bias", None) is not None:
            ql._set("bias", linear_layer.bias.detach().to(torch.bfloat16))
        return ql
