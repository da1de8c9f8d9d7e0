"""GPU parity of the device-side quantiser (gbxq_quantize / gbxq_quantize_rows) and of the KV
caches built on it, against the restatement of mx.quantize (oracle/mlx_affine.py::quantize, MLX_SPEC R6).
Bar: codes, scales and biases BIT-EXACT (integer / byte work plus IEEE fp32 divisions); the quantized attention on top
within 2e-2 of the fp32 evaluation on the oracle-dequantised cache (same bar as tests/test_gpu_gather.py)."""
import numpy as np
import pytest
import torch

from oracle import mlx_affine as A

pytestmark = pytest.mark.gpu

DT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}


def _u32(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().view(torch.int32).numpy().view(np.uint32)


def _check(w: torch.Tensor, got, gs: int, bits: int, dtype: str, what=""):
    oq, os_, ob = A.quantize(w.detach().float().cpu().numpy(), gs, bits, dtype)
    assert got[0].dtype == torch.uint32 and got[1].dtype == w.dtype and got[2].dtype == w.dtype
    assert tuple(got[0].shape) == oq.shape and tuple(got[1].shape) == os_.shape == tuple(got[2].shape), what
    assert np.array_equal(got[1].float().cpu().numpy(), os_), f"{what}: scales differ"
    assert np.array_equal(got[2].float().cpu().numpy(), ob), f"{what}: biases differ"
    assert np.array_equal(_u32(got[0]), oq), f"{what}: codes differ"


@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("bits", A.SUPPORTED_BITS)
def test_quantize_bit_exact(cuda_device, bits, dtype):
    from gbx_lm_b200 import ops

    gen = torch.Generator().manual_seed(100 * bits)
    for gs in (32, 64, 128):
        # decode-sized (8 heads x head_dim 128), ragged (less than one 256-element chunk; rows that straddle chunks),
        # many chunks per CTA, and leading batch dimensions
        for shape in ((8, 128), (1, gs), (3, 3 * gs), (257, 512), (2, 3, 5, 256)):
            w = (torch.randn(shape, generator=gen) * 3.0).to(DT[dtype]).to(cuda_device)
            _check(w, ops.quantize(w, gs, bits), gs, bits, dtype, f"b{bits} g{gs} {dtype} {shape}")


import json  # noqa: E402
import os  # noqa: E402

QG = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "quantize_golden.json")))


@pytest.mark.parametrize("case", QG["cases"], ids=lambda c: f"b{c['bits']}-g{c['group_size']}")
def test_quantize_golden_fixture(cuda_device, case):
    """The committed known answers (written by the C restatement, tests/golden/make_quantize_golden.py)."""
    from gbx_lm_b200 import ops

    rows, K, bits, gs = case["rows"], case["K"], case["bits"], case["group_size"]
    wb = np.array(case["w_bf16"], dtype=np.uint16).reshape(rows, K)
    w = torch.from_numpy(wb.view(np.int16)).view(torch.bfloat16).to(cuda_device)
    q, s, b = ops.quantize(w, gs, bits)
    assert np.array_equal(_u32(q).reshape(-1), np.array(case["codes"], dtype=np.uint32))
    bits16 = lambda t: t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16).reshape(-1)  # noqa: E731
    assert np.array_equal(bits16(s), np.array(case["scales_bf16"], dtype=np.uint16))
    assert np.array_equal(bits16(b), np.array(case["biases_bf16"], dtype=np.uint16))


@pytest.mark.parametrize("bits", A.SUPPORTED_BITS)
def test_quantize_special_groups(cuda_device, bits):
    """Zero and constant groups, one-sided groups (all negative / all positive), an outlier, magnitudes from 1e-30 to
    1e30, signed zeros: the same codes and statistics as the restatement."""
    from gbx_lm_b200 import ops

    gs = 64
    gen = torch.Generator().manual_seed(bits)
    w = torch.randn((12, 2 * gs), generator=gen)
    w[0, :gs] = 0.0
    w[1, :gs] = 0.37
    w[2, :gs] = -w[2, :gs].abs()
    w[3, :gs] = w[3, :gs].abs()
    w[4, 5] = 3000.0
    w[5] *= 1e-30
    w[6] *= 1e30
    w[7, :gs] = 0.0
    w[7, 3] = -0.0
    w[8, :gs] = -2.5
    w[9] *= 1e-6
    for dtype in ("bf16", "f32"):
        x = w.to(DT[dtype]).to(cuda_device)
        _check(x, ops.quantize(x, gs, bits), gs, bits, dtype, f"special b{bits} {dtype}")
    x = w.clamp(-6e4, 6e4).to(torch.float16).to(cuda_device)
    _check(x, ops.quantize(x, gs, bits), gs, bits, "f16", f"special b{bits} f16")


@pytest.mark.parametrize("bits", [8, 4, 3])
def test_quantize_into_cache_rows(cuda_device, bits):
    """gbxq_quantize_rows: the new positions of every (batch, head) land at the offset of a [.., capacity, ..] buffer;
    every other row keeps its content."""
    from gbx_lm_b200 import ops

    B, H, cap, D, gs = 2, 3, 16, 128, 64
    gen = torch.Generator().manual_seed(bits)
    for steps, offset in ((1, 0), (1, 15), (5, 7), (16, 0)):
        x = torch.randn((B, H, steps, D), generator=gen).to(torch.bfloat16).to(cuda_device)
        oq = torch.full((B, H, cap, D * bits // 32), 0x5A5A5A5A, dtype=torch.int32, device=cuda_device).view(torch.uint32)
        os_ = torch.full((B, H, cap, D // gs), 7.0, dtype=torch.bfloat16, device=cuda_device)
        ob = torch.full((B, H, cap, D // gs), -7.0, dtype=torch.bfloat16, device=cuda_device)
        ops.quantize_into(x, (oq, os_, ob), offset, gs, bits)
        dq, ds, db = ops.quantize(x, gs, bits)
        sl = slice(offset, offset + steps)
        assert torch.equal(oq.view(torch.int32)[:, :, sl], dq.view(torch.int32))
        assert torch.equal(os_[:, :, sl], ds) and torch.equal(ob[:, :, sl], db)
        keep = torch.ones(cap, dtype=torch.bool, device=cuda_device)
        keep[sl] = False
        assert (oq.view(torch.int32)[:, :, keep] == 0x5A5A5A5A).all()
        assert (os_[:, :, keep] == 7.0).all() and (ob[:, :, keep] == -7.0).all()
    # a transposed (non-contiguous) source, as attention layers produce keys: [B, L, H, D] -> [B, H, L, D]
    xt = torch.randn((B, 4, H, D), generator=gen).to(torch.bfloat16).to(cuda_device).transpose(1, 2)
    assert not xt.is_contiguous()
    oq.view(torch.int32).zero_()
    ops.quantize_into(xt, (oq, os_, ob), 2, gs, bits)
    assert torch.equal(oq.view(torch.int32)[:, :, 2:6], ops.quantize(xt, gs, bits)[0].view(torch.int32))
    with pytest.raises(ValueError):
        ops.quantize_into(x, (oq, os_, ob), 1, gs, bits)           # 16 rows at offset 1 do not fit
    with pytest.raises(ValueError):
        ops.quantize_into(x, (oq, os_.float(), ob), 0, gs, bits)   # statistics must be in w's dtype
    with pytest.raises(ValueError):
        ops.quantize_into(x, (oq[:, :, ::2], os_[:, :, ::2], ob[:, :, ::2]), 0, gs, bits)  # non-contiguous buffers


def test_quantize_full_size_round_trip(cuda_device):
    """Size-independent properties at full sizes: a whole prefill of keys (8 heads x 2048 positions x 128) and an
    8B-sized weight matrix go encode -> decode on the device (gbxq_quantize -> gbxq_dequantize); every element is
    within one step of its group, every group's anchored edge comes back exactly (up to the bf16 rounding of the bias),
    and the result is deterministic."""
    from gbx_lm_b200 import ops

    gen = torch.Generator(device=cuda_device).manual_seed(5)
    for shape, gs, bits in (((8 * 2048, 128), 64, 8), ((4096, 14336), 64, 4), ((4096, 4096), 128, 3), ((1024, 4096), 32, 2)):
        w = torch.randn(shape, generator=gen, device=cuda_device).to(torch.bfloat16)
        q, s, b = ops.quantize(w, gs, bits)
        q2, s2, b2 = ops.quantize(w, gs, bits)
        assert torch.equal(q.view(torch.int32), q2.view(torch.int32)) and torch.equal(s, s2) and torch.equal(b, b2)
        d = ops.dequantize(q, s, b, gs, bits).float().reshape(shape[0], -1, gs)
        g = w.float().reshape(shape[0], -1, gs)
        step = s.float().abs().unsqueeze(-1)
        # one step of quantisation error + the bf16 roundings of scale (2^-9 relative over up to 2^bits steps), bias and result
        bound = step * (1.0 + (1 << bits) * 2.0 ** -8) + g.abs().amax(-1, keepdim=True) * 2.0 ** -7
        assert ((d - g).abs() <= bound).all()
        lo, hi = g.amin(-1), g.amax(-1)
        edge = torch.where(lo.abs() > hi.abs(), lo, hi)
        assert torch.equal(b.float().reshape(edge.shape), edge)  # bias = the edge (bf16 values in, bf16 out: exact)
        codes_used = ops.dequantize(q, torch.ones_like(s), torch.zeros_like(b), gs, bits)
        assert codes_used.min() == 0 and codes_used.max() == (1 << bits) - 1


def test_quantized_kv_cache_and_attention_on_device(cuda_device):
    """QuantizedKVCache.update_and_fetch (prompt, decode steps, growth across the 256-position block border) holds
    exactly the oracle's quantisation of the concatenated keys / values, and the quantized attention over it
    (gbx_lm/models/base.py:65-98) matches the fp32 evaluation on the oracle-dequantised cache."""
    from gbx_lm_b200 import cache, switch_layers as SL

    dev = cuda_device
    B, Hq, Hkv, D, gs, bits = 1, 8, 2, 128, 64, 8
    gen = torch.Generator().manual_seed(11)
    c = cache.QuantizedKVCache(group_size=gs, bits=bits)
    ks, vs = [], []
    qk = qv = None
    for n in (250, 1, 1, 1, 1, 1, 1, 1, 3):
        k = torch.randn((B, Hkv, n, D), generator=gen).to(torch.bfloat16).to(dev)
        v = torch.randn((B, Hkv, n, D), generator=gen).to(torch.bfloat16).to(dev)
        ks.append(k), vs.append(v)
        qk, qv = c.update_and_fetch(k, v)
    T = 260
    assert c.offset == T and c.keys[0].shape[-2] == 512 and qk[0].shape == (B, Hkv, T, D * bits // 32)
    kcat, vcat = torch.cat(ks, 2), torch.cat(vs, 2)
    _check(kcat, tuple(t.contiguous() for t in qk), gs, bits, "bf16", "cache keys")
    _check(vcat, tuple(t.contiguous() for t in qv), gs, bits, "bf16", "cache values")

    q = torch.randn((B, Hq, 1, D), generator=gen).to(torch.bfloat16).to(dev)
    out = SL.quantized_scaled_dot_product_attention(q, qk, qv, D ** -0.5, None, group_size=gs, bits=bits)
    assert out.shape == (B, Hq, 1, D)

    def deq(t3):
        w = A.dequantize(_u32(t3[0]).reshape(-1, t3[0].shape[-1]), t3[1].float().cpu().numpy().reshape(-1, t3[1].shape[-1]),
                         t3[2].float().cpu().numpy().reshape(-1, t3[2].shape[-1]), gs, bits, "f32")
        return torch.from_numpy(w).reshape(B, Hkv, T, D).to(dev)

    kd, vd = deq(qk).repeat_interleave(Hq // Hkv, 1), deq(qv).repeat_interleave(Hq // Hkv, 1)
    ref = torch.softmax((q.float() * D ** -0.5) @ kd.transpose(-1, -2), -1) @ vd
    assert (out.float() - ref).abs().max() <= 2e-2 * ref.abs().max()
    # and the quantisation itself is a faithful cache: close to the attention on the unquantised keys / values
    dense = torch.softmax((q.float() * D ** -0.5) @ kcat.float().repeat_interleave(Hq // Hkv, 1).transpose(-1, -2), -1) \
        @ vcat.float().repeat_interleave(Hq // Hkv, 1)
    assert (out.float() - dense).abs().max() <= 5e-2 * dense.abs().max()

    # KVCache -> to_quantized (gbx_lm/utils.py:204-214) gives the same cache content for the same positions
    dcache = [cache.KVCache()]
    dcache[0].update_and_fetch(kcat[:, :, :250], vcat[:, :, :250])
    cache.maybe_quantize_kv_cache(dcache, quantized_kv_start=0, kv_group_size=gs, kv_bits=bits)
    assert isinstance(dcache[0], cache.QuantizedKVCache) and dcache[0].offset == 250
    assert torch.equal(dcache[0].keys[0].view(torch.int32)[:, :, :250], c.keys[0].view(torch.int32)[:, :, :250])
    assert torch.equal(dcache[0].values[1][:, :, :250], c.values[1][:, :, :250])
