"""GBA checkpoint -> the MLX layout the hot path reads (host logic, no kernels).

What the reference does when it loads an ORIGINAL GreenBitAI "GBA" checkpoint with `is_conversion=True`
(gbx_lm/utils.py:828-843,869-877 and gbx_lm/models/quantized_linear_gba.py:119-160), restated on torch tensors:

  * `qweight` is stored K-major, `[K*bits/32, N]`: transpose to the MLX `[N, K*bits/32]` (the packing along K inside
    a word is unchanged), dtype uint32                                                     utils.py:831-832
  * plain checkpoints: `scales`, `zeros` `[K/gs, N]` -> transposed `[N, K/gs]`, bf16          utils.py:834-835
  * double-quantised checkpoints: `qstatistic` uint8 `[G, D, 32]` holds a 4-bit code of the scale (high nibble) and of
    the zero (low nibble) of every group, with fp16 second-level `q{scales,zeros}_{scales,zeros}` `[G, D, 1]`;
        zeros  = (q_zero  - qzeros_zeros)  * qzeros_scales      in fp16
        scales = (q_scale - qscales_zeros) * qscales_scales     in fp16
    viewed as `[G, N]`, transposed to `[N, G]`                      quantized_linear_gba.py:119-149
  * MLX adds its `biases`, GBA subtracts its zeros: `zeros <- -zeros`                        quantized_linear_gba.py:151-155
  * norms, biases, router gates, lm_head, embed_tokens, channel_scale -> bf16; scales / zeros -> bf16 at every load
    (the fp16 values of the double-quant expansion are truncated to bf16 here, as in the reference)   utils.py:837-843

`convert_gba_weights` returns a state dict in exactly the layout `gbx_lm_b200.utils.load_model` (and gba2mlx's output)
uses -- `load_model(..., is_conversion=True)` calls it right after reading the files; `q_perm` / `channel_scale` are carried along untouched (unused by the forward: quantized_linear_gba.py:187-192).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

DQ_LEAVES = ("qstatistic", "qzeros_zeros", "qzeros_scales", "qscales_zeros", "qscales_scales")


def parameter_usage(weights: Dict[str, torch.Tensor]) -> Tuple[bool, bool]:
    """(uses double quantisation, uses q_perm) -- gbx_lm/utils.py:696-719."""
    dq = any(any(leaf in k for leaf in DQ_LEAVES) for k in weights)
    return dq, any("q_perm" in k for k in weights)


def expand_double_quant(qstatistic: torch.Tensor, qzeros_zeros: torch.Tensor, qzeros_scales: torch.Tensor,
                        qscales_zeros: torch.Tensor, qscales_scales: torch.Tensor, input_dims: int, output_dims: int,
                        group_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp16 (scales, zeros) `[N, G]` from the second-level codes (quantized_linear_gba.py:119-149); zeros still with
    GBA's sign."""
    g = math.ceil(input_dims / group_size)
    qs = ((qstatistic.to(torch.int32) & 0xF0) >> 4).to(torch.float16)
    qz = (qstatistic.to(torch.int32) & 0x0F).to(torch.float16)
    zeros = ((qz - qzeros_zeros.to(torch.float16)) * qzeros_scales.to(torch.float16)).reshape(g, output_dims)
    scales = ((qs - qscales_zeros.to(torch.float16)) * qscales_scales.to(torch.float16)).reshape(g, output_dims)
    return scales.t().contiguous(), zeros.t().contiguous()


def convert_gba_weights(weights: Dict[str, torch.Tensor], group_size_of=None) -> Dict[str, torch.Tensor]:
    """GBA state dict -> MLX-layout state dict (see module docstring).  `group_size_of(module_name) -> group_size` is
    needed only for double-quantised checkpoints (to size the expanded statistics); default 64."""
    dq, _ = parameter_usage(weights)
    out: Dict[str, torch.Tensor] = {}
    mods = sorted({k.rsplit(".", 1)[0] for k in weights if k.endswith(".qweight")})
    handled = set()
    for m in mods:
        qw = weights[m + ".qweight"]
        if qw.dtype not in (torch.int32, torch.uint32):
            qw = qw.to(torch.int64).to(torch.int32)
        qw_t = qw.view(torch.int32).t().contiguous().view(torch.uint32)   # [N, K*bits/32]
        out[m + ".qweight"] = qw_t
        handled.add(m + ".qweight")
        n = qw_t.shape[0]
        if dq and m + ".qstatistic" in weights:
            gs = group_size_of(m) if group_size_of is not None else 64
            st = weights[m + ".qstatistic"]
            g = st.shape[0]
            scales, zeros = expand_double_quant(st, weights[m + ".qzeros_zeros"], weights[m + ".qzeros_scales"],
                                                weights[m + ".qscales_zeros"], weights[m + ".qscales_scales"],
                                                g * gs, n, gs)
            handled.update(m + "." + leaf for leaf in DQ_LEAVES)
        else:
            scales = weights[m + ".scales"].t().contiguous()
            zeros = weights[m + ".zeros"].t().contiguous()
            handled.update((m + ".scales", m + ".zeros"))
        out[m + ".scales"] = scales.to(torch.bfloat16)
        out[m + ".zeros"] = (-zeros.float()).to(zeros.dtype).to(torch.bfloat16)  # exact sign flip, then the loader's bf16 cast
    for k, v in weights.items():
        if k in handled:
            continue
        if any(s in k for s in ("norm.weight", "bias", "gate.weight", "lm_head", "embed_tokens", "channel_scale")):
            v = v.to(torch.bfloat16)
        out[k] = v
    return out


def expand_statistics(weights: Dict[str, torch.Tensor], group_size_of=None) -> Dict[str, torch.Tensor]:
    """A checkpoint ALREADY in the MLX orientation that still carries double-quantised statistics (the reference's
    `use_double_quantization and not is_conversion` case: `prepare_scales_zeros` without `post_processing_and_release`,
    gbx_lm/utils.py:864-868): only the second-level codes are expanded to `scales` / `zeros`; no transpose, no sign flip."""
    out = {k: v for k, v in weights.items() if not any(k.endswith("." + leaf) for leaf in DQ_LEAVES)}
    for k, st in weights.items():
        if not k.endswith(".qstatistic"):
            continue
        m = k.rsplit(".", 1)[0]
        gs = group_size_of(m) if group_size_of is not None else 64
        n = weights[m + ".qweight"].shape[0]
        scales, zeros = expand_double_quant(st, weights[m + ".qzeros_zeros"], weights[m + ".qzeros_scales"],
                                            weights[m + ".qscales_zeros"], weights[m + ".qscales_scales"],
                                            st.shape[0] * gs, n, gs)
        out[m + ".scales"] = scales.to(torch.bfloat16)
        out[m + ".zeros"] = zeros.to(torch.bfloat16)
    return out
