"""Host-side (CPU torch) helpers for the MLX packed-uint32 layout: pack/unpack codes, the affine
quantisation recipe of `mx.quantize`, and synthetic layer fabrication.

These exist to WRITE checkpoints in the layout gba2mlx emits (gbx_lm/gba2mlx.py:47-65,
gbx_lm/utils.py:1055-1104) and for `QuantizedLinear.from_linear` (quantized_linear_gba.py:342);
they are load-time utilities, not on the matmul path, and never used to compute a forward.

Layout: row n of qweight is the LSB-first bitstream of the row's K codes, cut into little-endian
uint32 words: code k occupies stream bits [k*bits, (k+1)*bits)."""
from __future__ import annotations

import math

import torch

BITS = (2, 3, 4, 6, 8)
GROUP_SIZES = (32, 64, 128)


def pack_codes(q: torch.Tensor, bits: int) -> torch.Tensor:
    """uint8 codes [N, K] -> uint32 [N, K*bits/32]."""
    assert bits in BITS and q.dim() == 2 and q.shape[1] % 32 == 0
    n, k = q.shape
    q = q.to(torch.int64).cpu()
    nblk = k // 32  # 32 codes <-> `bits` words
    qb = q.reshape(n, nblk, 32)
    out = torch.zeros((n, nblk, bits), dtype=torch.int64)
    for i in range(32):
        bit = i * bits
        wi, off = bit // 32, bit % 32
        out[:, :, wi] |= (qb[:, :, i] << off) & 0xFFFFFFFF
        if off + bits > 32:
            out[:, :, wi + 1] |= qb[:, :, i] >> (32 - off)
    out = out.reshape(n, nblk * bits)
    # int64 -> uint32 bit pattern
    return (out & 0xFFFFFFFF).to(torch.uint32)


def unpack_codes(w: torch.Tensor, bits: int) -> torch.Tensor:
    """uint32 [N, K*bits/32] -> uint8 codes [N, K]."""
    assert bits in BITS and w.dim() == 2
    n, nw = w.shape
    wi64 = w.cpu().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    nblk = nw // bits
    wb = wi64.reshape(n, nblk, bits)
    out = torch.zeros((n, nblk, 32), dtype=torch.int64)
    mask = (1 << bits) - 1
    for i in range(32):
        bit = i * bits
        wi, off = bit // 32, bit % 32
        v = wb[:, :, wi] >> off
        if off + bits > 32:
            v = v | (wb[:, :, wi + 1] << (32 - off))
        out[:, :, i] = v & mask
    return out.reshape(n, nblk * 32).to(torch.uint8)


def quantize_affine(w: torch.Tensor, group_size: int = 64, bits: int = 4):
    """Affine group quantisation following mx.quantize's recipe: returns (qweight uint32,
    scales, biases) with scales/biases in w.dtype.  W ~= scales * q + biases."""
    assert bits in BITS and group_size in GROUP_SIZES
    n, k = w.shape
    assert k % group_size == 0
    dt = w.dtype
    g = w.float().cpu().reshape(n, k // group_size, group_size)
    n_bins = float((1 << bits) - 1)
    w_max = g.amax(-1)
    w_min = g.amin(-1)
    mask = w_min.abs() > w_max.abs()
    scales = ((w_max - w_min) / n_bins).clamp_min(1e-7)
    scales = torch.where(mask, scales, -scales)
    edge = torch.where(mask, w_min, w_max)
    q0 = torch.round(edge / scales)
    scales = torch.where(q0 != 0, edge / torch.where(q0 != 0, q0, torch.ones_like(q0)), scales)
    biases = torch.where(q0 == 0, torch.zeros_like(edge), edge)
    # the codes come from the UNROUNDED fp32 scale / bias; only the stored statistics are cast to w.dtype (the order of
    # MLX's CPU quantize: rint((w - bias) / scale) before the static_cast of scales / biases)
    q = torch.round((g - biases[..., None]) / scales[..., None]).clamp(0, n_bins).to(torch.uint8)
    return pack_codes(q.reshape(n, k), bits), scales.to(dt), biases.to(dt)


def synth_layer(n: int, k: int, bits: int, group_size: int, seed: int = 0, with_bias: bool = False,
                dtype: torch.dtype = torch.bfloat16):
    """Synthetic QuantizedLinear tensors: codes ~ U{0..2^b-1};
    scales ~ U(.5,1.5) * (2/sqrt(K)) / (2^b-1); zeros = -scales*(2^b-1)/2*(1+eps), eps ~ U(-.05,.05).
    Dequantised weights are ~U(-1/sqrt(K), 1/sqrt(K))."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    nb = (1 << bits) - 1
    q = torch.randint(0, nb + 1, (n, k), generator=gen, dtype=torch.uint8)
    s = (torch.rand((n, k // group_size), generator=gen) + 0.5) * (2.0 / math.sqrt(k) / nb)
    eps = torch.rand((n, k // group_size), generator=gen) * 0.1 - 0.05
    s = s.to(dtype)
    z = (-s.float() * (nb / 2.0) * (1.0 + eps)).to(dtype)
    out = {"qweight": pack_codes(q, bits), "scales": s, "zeros": z, "codes": q}
    if with_bias:
        out["bias"] = (torch.randn((n,), generator=gen) * 0.02).to(dtype)
    return out
