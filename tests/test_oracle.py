"""CPU tests of the checker itself: the numpy + C restatements of MLX affine quantisation
(oracle/) against the committed golden vectors, hand-derived constants and torch's own bf16
arithmetic.  MLX is not available, so this pins our reading of the format (PARITY UNPINNED)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as C
from oracle import mlx_affine as A

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "affine_golden.json")))

# hand-derived in SURVEY.md 8c from the byte rules, independent of any code in this repo
SURVEY_RAMPS = {
    2: [0xE4E4E4E4, 0xE4E4E4E4],
    3: [0x88FAC688, 0xC688FAC6, 0xFAC688FA],
    4: [0x76543210, 0xFEDCBA98, 0x76543210, 0xFEDCBA98],
    6: [0x440C2040, 0xA2481C61, 0x3CE34C2C, 0x544D2450, 0xA6585D65, 0x7DE75C6D],
    8: [0x03020100, 0x07060504, 0x0B0A0908, 0x0F0E0D0C, 0x13121110, 0x17161514, 0x1B1A1918, 0x1F1E1D1C],
}


@pytest.mark.parametrize("bits", A.SUPPORTED_BITS)
def test_ramp_words_match_survey_and_golden(bits):
    q = (np.arange(32) % (1 << bits)).astype(np.uint8)[None]
    w = A.pack_codes(q, bits)[0].tolist()
    assert w == SURVEY_RAMPS[bits]
    assert w == GOLD["ramps"][str(bits)]
    assert C.pack(q, bits)[0].tolist() == w


@pytest.mark.parametrize("bits", A.SUPPORTED_BITS)
def test_single_hot_straddlers(bits):
    # 3-bit codes 10, 21 and 6-bit codes 5, 10, 21, 26 straddle a 32-bit word border
    for k in range(32):
        q = np.zeros((1, 32), dtype=np.uint8)
        q[0, k] = (1 << bits) - 1
        w = A.pack_codes(q, bits)
        assert w[0].tolist() == GOLD["single_hot"][str(bits)][str(k)]
        # total set bits == bits, located at stream bits [k*bits, (k+1)*bits)
        stream = np.unpackbits(w.astype("<u4").view(np.uint8), bitorder="little")
        assert stream.sum() == bits and stream[k * bits : (k + 1) * bits].all()
        assert (A.unpack_codes(w, bits) == q).all()
        assert (A.unpack_codes_byte_rules(w, bits) == q).all()
        assert (C.unpack(w, bits) == q).all()
    straddle = [k for k in range(32) if (k * bits) // 32 != ((k + 1) * bits - 1) // 32]
    assert straddle == {2: [], 3: [10, 21], 4: [], 6: [5, 10, 21, 26], 8: []}[bits]


@pytest.mark.parametrize("bits", A.SUPPORTED_BITS)
def test_pack_unpack_random_roundtrip(bits):
    rng = np.random.default_rng(bits)
    q = rng.integers(0, 1 << bits, size=(7, 224), dtype=np.uint8)
    w = A.pack_codes(q, bits)
    assert w.shape == (7, 224 * bits // 32)
    assert (A.unpack_codes(w, bits) == q).all()
    assert (A.unpack_codes_byte_rules(w, bits) == q).all()
    assert (C.pack(q, bits) == w).all() and (C.unpack(w, bits) == q).all()


def test_empty_and_invalid():
    assert A.pack_codes(np.zeros((0, 64), np.uint8), 4).shape == (0, 8)
    with pytest.raises(ValueError):
        A.pack_codes(np.zeros((1, 48), np.uint8), 4)  # K % 32
    with pytest.raises(ValueError):
        A.pack_codes(np.full((1, 32), 4, np.uint8), 2)  # code out of range
    with pytest.raises(ValueError):
        A.dequantize(np.zeros((2, 8), np.uint32), np.zeros((2, 1), np.float32), np.zeros((2, 1), np.float32), 64, 5, "f32")
    with pytest.raises(ValueError):
        A.dequantize(np.zeros((2, 8), np.uint32), np.zeros((2, 1), np.float32), np.zeros((2, 1), np.float32), 48, 4, "f32")
    with pytest.raises(ValueError):  # K mismatch between qweight and scales
        A.dequantize(np.zeros((2, 8), np.uint32), np.zeros((2, 2), np.float32), np.zeros((2, 2), np.float32), 64, 4, "f32")


def test_bf16_rounding_matches_torch():
    rng = np.random.default_rng(0)
    a = rng.standard_normal(10000).astype(np.float32) * np.float32(3.7)
    a[:4] = [0.0, -0.0, np.inf, 1.0 + 2**-8]  # tie -> even
    ours = A.f32_to_bf16_bits(a)
    ref = torch.from_numpy(a).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert (ours == ref).all()


@pytest.mark.parametrize("case", GOLD["dequant"], ids=lambda c: f"b{c['bits']}-{c['dtype']}")
def test_dequant_golden_and_torch(case):
    bits, dtype = case["bits"], case["dtype"]
    w = np.array([case["qweight"]], dtype=np.uint32)
    tdt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[dtype]
    s_t = torch.tensor([[0.0123]], dtype=tdt)
    b_t = torch.tensor([[-0.771]], dtype=tdt)
    raw = (lambda t: t.view(torch.int16).numpy().view(np.uint16)) if dtype != "f32" else (lambda t: t.numpy())
    s, b = raw(s_t), raw(b_t)
    if dtype == "f16":
        s, b = s.view(np.float16), b.view(np.float16)
    d = A.dequantize(w, s, b, 32, bits, dtype)
    assert d.astype(np.float32).view(np.uint32)[0].tolist() == case["out_f32_hex"]
    # independent: torch's T-typed multiply then add (two roundings) on the codes
    q = torch.tensor(case["codes"], dtype=tdt)
    ref = (s_t[0, 0] * q) + b_t[0, 0]
    assert torch.equal(ref.float(), torch.from_numpy(d[0].astype(np.float32)))
    # C oracle, bit for bit
    cs = s if dtype == "f32" else np.ascontiguousarray(s).view(np.uint16)
    cb = b if dtype == "f32" else np.ascontiguousarray(b).view(np.uint16)
    dc = C.dequantize(w, cs, cb, 32, bits, dtype)
    if dtype == "f32":
        assert (dc.view(np.uint32) == d.view(np.uint32)).all()
    elif dtype == "bf16":
        assert (dc == A.f32_to_bf16_bits(d)).all()
    else:
        assert (dc == d.astype(np.float16).view(np.uint16)).all()


def test_dequant_two_roundings_differ_from_fma():
    """The vector exercises both roundings: a fused multiply-add would give a different bf16."""
    found = False
    s = A.bf16_bits_to_f32(A.f32_to_bf16_bits(np.array([0.0123], np.float32)))[0]
    b = A.bf16_bits_to_f32(A.f32_to_bf16_bits(np.array([-0.771], np.float32)))[0]
    for q in range(256):
        two = A._round_to(A._round_to(np.float32(s * q), "bf16") + np.float32(b), "bf16")
        one = A._round_to(np.float32(np.float64(s) * q + np.float64(b)), "bf16")
        found |= bool(two != one)
    assert found


@pytest.mark.parametrize("case", GOLD["qmm"], ids=lambda c: f"b{c['bits']}-g{c['group_size']}")
def test_qmm_golden(case):
    w = np.array(case["qweight"], dtype=np.uint32)
    s = np.array(case["scales_bf16"], dtype=np.uint16)
    z = np.array(case["zeros_bf16"], dtype=np.uint16)
    x = np.array(case["x_bf16"], dtype=np.uint16)
    gold = np.array(case["y_bf16"], dtype=np.uint16)
    y = A.quantized_matmul(x, w, s, z, case["group_size"], case["bits"], "bf16", "f64")
    assert (A.f32_to_bf16_bits(y) == gold).all()
    yc = C.qmm(x, w, s, z, case["group_size"], case["bits"], "bf16", "f64")
    assert (yc == gold).all()
    # the matmul against the explicitly dequantised (twice-rounded) matrix stays within 1 bf16 ulp-ish
    yd = A.quantized_matmul(x, w, s, z, case["group_size"], case["bits"], "bf16", "deq")
    assert np.abs(yd - y).max() <= 1e-2 * np.abs(y).max()
    # fp32 k-ordered (MLX SIMD-style) variant agrees with the truth to bf16 resolution
    y32 = C.qmm(x, w, s, z, case["group_size"], case["bits"], "bf16", "f32")
    assert np.abs(A.bf16_bits_to_f32(y32) - y).max() <= 2.0 ** -7 * np.abs(y).max()


def test_qmm_linearity_and_bias():
    L = A.synth_layer(16, 128, 4, 64, seed=3, with_bias=True)
    x1 = A.bf16_bits_to_f32(A.synth_x(1, 128, seed=4))
    x2 = A.bf16_bits_to_f32(A.synth_x(1, 128, seed=5))
    sf, zf = A.bf16_bits_to_f32(L["scales"]), A.bf16_bits_to_f32(L["zeros"])
    f = lambda x: A.quantized_matmul(x, L["qweight"], sf, zf, 64, 4, "f32", "f64")
    np.testing.assert_allclose(f(x1 + x2), f(x1) + f(x2), rtol=0, atol=1e-5)
    yb = A.quantized_matmul(A.f32_to_bf16_bits(x1), L["qweight"], L["scales"], L["zeros"], 64, 4, "bf16", "f64", bias=L["bias"])
    y0 = A.quantized_matmul(A.f32_to_bf16_bits(x1), L["qweight"], L["scales"], L["zeros"], 64, 4, "bf16", "f64")
    ref = A._round_to(y0 + A.bf16_bits_to_f32(L["bias"]), "bf16")
    assert (yb == ref).all()
    yc = C.qmm(A.f32_to_bf16_bits(x1), L["qweight"], L["scales"], L["zeros"], 64, 4, "bf16", "f64", bias=L["bias"])
    assert (yc == A.f32_to_bf16_bits(yb)).all()


def test_mlx_x86_scalar_accumulation_is_coarse():
    """Documents why parity is stated against the fp64 truth: the T-accumulate variant (MLX's x86
    scalar `_qmm_t`, accumulating in bf16) is itself percent-level away from the truth."""
    L = A.synth_layer(32, 1024, 4, 64, seed=9)
    x = A.synth_x(1, 1024, seed=10)
    y = A.quantized_matmul(x, L["qweight"], L["scales"], L["zeros"], 64, 4, "bf16", "f64")
    yt = A.bf16_bits_to_f32(C.qmm(x, L["qweight"], L["scales"], L["zeros"], 64, 4, "bf16", "t"))
    rel = np.abs(yt - y).max() / np.abs(y).max()
    assert 1e-3 < rel < 0.5


def test_quantize_roundtrip_error_bound():
    rng = np.random.default_rng(1)
    w = rng.standard_normal((8, 256)).astype(np.float32)
    for bits in A.SUPPORTED_BITS:
        qw, s, b = A.quantize(w, 64, bits, "f32")
        d = A.dequantize(qw, s, b, 64, bits, "f32")
        step = np.abs(np.repeat(s, 64, -1))
        assert (np.abs(d - w) <= 1.01 * step + 1e-6).all()  # edge-anchored scale: at most one step at the far end
