"""CPU tests of the GBA -> MLX-layout converter (gbx_lm_b200/gba_convert.py; reference: gbx_lm/utils.py:828-843,
quantized_linear_gba.py:119-155): a synthetic GBA checkpoint, converted, must dequantise (oracle) to the weights GBA's
own convention defines -- w[k, n] = scale[g, n] * q[k, n] - zero[g, n] -- for plain and double-quantised statistics."""
import numpy as np
import torch

from gbx_lm_b200 import gba_convert as G
from oracle import mlx_affine as A


def _gba_layer(n, k, bits, gs, seed, double_quant):
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 1 << bits, size=(n, k), dtype=np.uint8)
    packed = A.pack_codes(q, bits)                                  # [N, K*bits/32]  (MLX orientation)
    qweight_gba = torch.from_numpy(np.ascontiguousarray(packed.T).view(np.int32))   # GBA stores it K-major
    g = k // gs
    out = {"qweight": qweight_gba}
    if double_quant:
        d = n // 32
        stat = rng.integers(0, 256, size=(g, d, 32), dtype=np.uint8)
        zz = rng.integers(0, 8, size=(g, d, 1)).astype(np.float16)
        zs = (rng.random((g, d, 1)) * 0.01 + 0.005).astype(np.float16)
        sz = rng.integers(0, 4, size=(g, d, 1)).astype(np.float16) - 6
        ss = (rng.random((g, d, 1)) * 0.001 + 0.0005).astype(np.float16)
        out.update(qstatistic=torch.from_numpy(stat), qzeros_zeros=torch.from_numpy(zz), qzeros_scales=torch.from_numpy(zs),
                   qscales_zeros=torch.from_numpy(sz), qscales_scales=torch.from_numpy(ss))
        qs = ((stat.astype(np.int32) & 0xF0) >> 4).astype(np.float16)
        qz = (stat.astype(np.int32) & 0x0F).astype(np.float16)
        scales = ((qs - sz) * ss).reshape(g, n)       # fp16 arithmetic, as the reference
        zeros = ((qz - zz) * zs).reshape(g, n)
    else:
        scales = (rng.random((g, n)) * 0.01 + 0.002).astype(np.float16)
        zeros = (rng.random((g, n)) * 0.08).astype(np.float16)
        out.update(scales=torch.from_numpy(scales), zeros=torch.from_numpy(zeros))
    # GBA's definition of the dense weight, with the bf16 truncation the loader applies to scales / zeros
    s_bf = torch.from_numpy(scales.astype(np.float32)).to(torch.bfloat16).float().numpy()
    z_bf = torch.from_numpy(zeros.astype(np.float32)).to(torch.bfloat16).float().numpy()
    return out, q, s_bf, z_bf


def _check(bits, gs, double_quant):
    n, k = 64, 256
    layer, q, s_bf, z_bf = _gba_layer(n, k, bits, gs, seed=bits + gs, double_quant=double_quant)
    w = {f"model.layers.0.mlp.up_proj.{leaf}": t for leaf, t in layer.items()}
    w["model.norm.weight"] = torch.ones(8, dtype=torch.float16)
    w["model.layers.0.mlp.up_proj.q_perm"] = torch.arange(k, dtype=torch.int16)
    out = G.convert_gba_weights(w, group_size_of=lambda m: gs)
    p = "model.layers.0.mlp.up_proj."
    assert out[p + "qweight"].shape == (n, k * bits // 32) and out[p + "qweight"].dtype == torch.uint32
    assert out[p + "scales"].shape == (n, k // gs) and out[p + "scales"].dtype == torch.bfloat16
    assert out["model.norm.weight"].dtype == torch.bfloat16 and p + "q_perm" in out
    assert not any(leaf in key for key in out for leaf in G.DQ_LEAVES)
    qw = out[p + "qweight"].view(torch.int32).numpy().view(np.uint32)
    assert (A.unpack_codes(qw, bits) == q).all()
    sb = out[p + "scales"].view(torch.int16).numpy().view(np.uint16)
    zb = out[p + "zeros"].view(torch.int16).numpy().view(np.uint16)
    deq = A.dequantize(qw, sb, zb, gs, bits, "bf16")
    # GBA convention: w = scale * q - zero  (per group along K), evaluated as the path does (two bf16 roundings)
    want = A._round_to(A._round_to(np.repeat(s_bf.T, gs, 1) * q.astype(np.float32), "bf16") - np.repeat(z_bf.T, gs, 1), "bf16")
    assert (deq == want).all()


def test_plain_gba_checkpoint():
    for bits, gs in ((4, 64), (2, 128), (3, 32), (8, 64)):
        _check(bits, gs, False)


def test_double_quantised_gba_checkpoint():
    for bits, gs in ((4, 64), (2, 32)):
        _check(bits, gs, True)
    assert G.parameter_usage({"a.qstatistic": 0, "a.q_perm": 0}) == (True, True)
    assert G.parameter_usage({"a.scales": 0}) == (False, False)
