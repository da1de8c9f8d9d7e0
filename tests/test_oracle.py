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


# ---------------------------------------------------------------------------------------------------------------
# Pins that do not come from the code under test (round 2):
#   * tests/golden/hf_metal_golden.json -- written by HuggingFace transformers' Metal/MLX affine-quantisation host code
#     (tests/golden/make_hf_golden.py), a foreign producer of the layout mx.quantized_matmul reads (2/4/8-bit)
#   * oracle/indep_check.py -- bit-by-bit, exact-rational evaluation of the documentation formulas (oracle/MLX_SPEC.md)
# ---------------------------------------------------------------------------------------------------------------
from fractions import Fraction  # noqa: E402

from oracle import indep_check as I  # noqa: E402

HF = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hf_metal_golden.json")))


@pytest.mark.parametrize("case", HF["cases"], ids=lambda c: f"hf-b{c['bits']}-g{c['group_size']}")
def test_foreign_vectors_pin_packing_and_affine_map(case):
    bits, gs, n, k = case["bits"], case["group_size"], case["N"], case["K"]
    w = np.array(case["qweight"], dtype=np.uint32)
    s = np.array(case["scales_f32_hex"], dtype=np.uint32).view(np.float32)
    b = np.array(case["biases_f32_hex"], dtype=np.uint32).view(np.float32)
    want = np.array(case["dequant_f32_hex"], dtype=np.uint32)
    assert w.shape == (n, k * bits // 32)
    # numpy oracle and C oracle reproduce the foreign dequantised matrix bit for bit (fp32)
    d = A.dequantize(w, s, b, gs, bits, "f32")
    assert (d.view(np.uint32) == want).all()
    dc = C.dequantize(w, s, b, gs, bits, "f32")
    assert (dc.view(np.uint32) == want).all()
    # the ramp group of row 0 exercises every code: unpacked codes are monotone and span 0 .. 2^bits-1
    q = A.unpack_codes(w, bits)
    assert q[0, :gs].min() == 0 and q[0, :gs].max() == (1 << bits) - 1
    assert (np.diff(q[0, :gs].astype(np.int32)) >= 0).all()
    # the foreign fp64 matmul against the oracle's fp64 truth
    x = np.array(case["x_f32_hex"], dtype=np.uint32).view(np.float32)
    y = A.quantized_matmul(x, w, s, b, gs, bits, "f32", "f64")
    np.testing.assert_allclose(y, np.array(case["y_f64"]), rtol=1e-5, atol=1e-6)
    # independent bit-by-bit checker on two rows
    for r in (0, n - 1):
        sc = [Fraction(float(v)) for v in s[r]]
        bi = [Fraction(float(v)) for v in b[r]]
        row = I.dequantize_row(w[r], sc, bi, gs, bits, "f32", k)
        assert [I.to_f32_bits(v) for v in row] == want[r].tolist()


@pytest.mark.parametrize("bits", A.SUPPORTED_BITS)
@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
def test_three_independent_checkers_agree(bits, dtype):
    """numpy oracle == C oracle == bit-by-bit rational checker: packing (incl. the straddling 3-/6-bit codes), the
    two-rounding dequantize in every dtype, and the exact matmul."""
    rng = np.random.default_rng(100 + bits)
    gs, n, k = 32, 3, 96
    q = rng.integers(0, 1 << bits, size=(n, k), dtype=np.uint8)
    w = A.pack_codes(q, bits)
    # independent packer: bit by bit
    for r in range(n):
        words = [0] * (k * bits // 32)
        for kk in range(k):
            I.put_code(words, kk, bits, int(q[r, kk]))
        assert words == w[r].tolist()
        assert [I.code_at(w[r], kk, bits) for kk in range(k)] == q[r].tolist()
    sf = (rng.random((n, k // gs)).astype(np.float32) + 0.5) * np.float32(0.013)
    bf = -(rng.random((n, k // gs)).astype(np.float32) + 0.5) * np.float32(0.4)
    if dtype == "bf16":
        s, b = A.f32_to_bf16_bits(sf), A.f32_to_bf16_bits(bf)
        sbits, bbits = s, b
    elif dtype == "f16":
        s, b = sf.astype(np.float16), bf.astype(np.float16)
        sbits, bbits = s.view(np.uint16), b.view(np.uint16)
    else:
        s, b = sf, bf
        sbits, bbits = s.view(np.uint32), b.view(np.uint32)
    d = A.dequantize(w, s, b, gs, bits, dtype).astype(np.float32)
    cs = s if dtype == "f32" else np.ascontiguousarray(sbits).view(np.uint16)
    cb = b if dtype == "f32" else np.ascontiguousarray(bbits).view(np.uint16)
    dc = C.dequantize(w, cs, cb, gs, bits, dtype)
    for r in range(n):
        sc = [I.from_bits(int(v), dtype) for v in sbits[r]]
        bi = [I.from_bits(int(v), dtype) for v in bbits[r]]
        row = I.dequantize_row(w[r], sc, bi, gs, bits, dtype, k)
        assert [I.to_f32_bits(v) for v in row] == d[r].view(np.uint32).tolist()
        if dtype == "bf16":
            assert [I.to_f32_bits(v) >> 16 for v in row] == dc[r].tolist()
        # exact rational matmul of one x row against the oracle's fp64 truth
        x = rng.standard_normal(k).astype(np.float32)
        xq = A._round_to(x, dtype)
        yt = A.quantized_matmul(xq if dtype != "bf16" else A.f32_to_bf16_bits(x), w[r:r + 1], s[r:r + 1], b[r:r + 1], gs, bits, dtype, "f64")
        exact = I.qmm_row([Fraction(float(v)) for v in xq], w[r], sc, bi, gs, bits, k)
        # the oracle returns the fp64 sum rounded once to the output dtype: equal to the exact sum rounded (or its
        # neighbour when the fp64 sum sits within 1e-12 of a tie)
        ulp = {"bf16": 2.0 ** -7, "f16": 2.0 ** -10, "f32": 2.0 ** -23}[dtype] * max(abs(float(exact)), 1e-3)
        got = float(np.asarray(yt, dtype=np.float64).ravel()[0])
        assert got == float(I.rne(exact, dtype)) or abs(got - float(exact)) <= ulp


def test_indep_rne_matches_hardware_formats():
    rng = np.random.default_rng(5)
    v = rng.standard_normal(300).astype(np.float64) * 7.3
    v[:3] = [1.0 + 2.0 ** -8, 1.0 + 3 * 2.0 ** -8, 2.0 ** -130]  # ties and a bf16/f32 subnormal
    for val in v:
        fr = Fraction(float(val))
        assert float(I.rne(fr, "f32")) == float(np.float32(val))
        assert float(I.rne(fr, "f16")) == float(np.float16(val)) or abs(val) < 6e-5
        want = A.bf16_bits_to_f32(A.f32_to_bf16_bits(np.array([np.float32(val)])))[0]
        # bf16 from the fp64 value directly vs via fp32: identical unless the fp32 step itself lands on a tie
        assert float(I.rne(Fraction(float(np.float32(val))), "bf16")) == float(want)


# ---------------------------------------------------------------------------------- transpose=False / gather_qmm (8f-3, 8f-4)
def _stack(E, N, K, bits, gs, seed):
    Ls = [A.synth_layer(N, K, bits, gs, seed=seed + e) for e in range(E)]
    return {k: np.stack([L[k] for L in Ls]) for k in ("qweight", "scales", "zeros")}


@pytest.mark.parametrize("bits", A.SUPPORTED_BITS)
def test_qmm_transpose_false_element_by_element(bits):
    """x @ W with W quantized along its LAST axis: a pure-Python triple loop over the unpacked codes (no shared code
    with the vectorised restatement beyond unpack_codes, itself checked above), bf16 scales."""
    gs, K, N, M = 32, 5, 64, 3
    L = A.synth_layer(K, N, bits, gs, seed=bits)           # rows = k, quantized axis = n
    x = A.synth_x(M, K, seed=9)
    y = A.quantized_matmul_nt(x, L["qweight"], L["scales"], L["zeros"], gs, bits, "bf16")
    q = A.unpack_codes(L["qweight"], bits)
    s, b, xf = A.bf16_bits_to_f32(L["scales"]), A.bf16_bits_to_f32(L["zeros"]), A.bf16_bits_to_f32(x)
    for m in range(M):
        for n in range(N):
            acc = 0.0
            for k in range(K):
                acc += float(xf[m, k]) * (float(s[k, n // gs]) * int(q[k, n]) + float(b[k, n // gs]))
            want = A.bf16_bits_to_f32(A.f32_to_bf16_bits(np.float32(acc)))
            assert y[m, n] == want or abs(y[m, n] - want) <= 2.0 ** -7 * abs(want)  # fp64 sum order may move a tie


def test_qmm_transpose_false_broadcasts_like_gqa_attention():
    """scores [B, kv, rep, L, T] @ V codes [B, kv, 1, T, D]: the expanded-heads call of models/base.py:80-92."""
    B, kv, rep, Lq, T, D, bits, gs = 2, 2, 3, 2, 7, 64, 8, 64
    Ls = [A.synth_layer(T, D, bits, gs, seed=40 + i) for i in range(B * kv)]
    pk = {k: np.stack([L[k] for L in Ls]).reshape(B, kv, 1, T, -1) for k in ("qweight", "scales", "zeros")}
    x = A.synth_x(B * kv * rep * Lq, T, seed=3).reshape(B, kv, rep, Lq, T)
    y = A.quantized_matmul_nt(x, pk["qweight"], pk["scales"], pk["zeros"], gs, bits, "bf16")
    assert y.shape == (B, kv, rep, Lq, D)
    for b in range(B):
        for h in range(kv):
            L = Ls[b * kv + h]
            Wd = A.dequantize(L["qweight"], L["scales"], L["zeros"], gs, bits, "bf16")  # bf16-rounded weights: loose bound
            ref = A.bf16_bits_to_f32(x[b, h]).astype(np.float64) @ Wd.astype(np.float64)
            assert np.abs(y[b, h] - ref).max() <= 2.0 ** -6 * np.abs(ref).max()


@pytest.mark.parametrize("bits", (2, 3, 4, 6, 8))
def test_gather_qmm_is_indexed_quantized_matmul(bits):
    """gather_qmm == quantized_matmul of the selected (x row block, expert) pairs; indices broadcast; bias per expert."""
    E, N, K, gs = 5, 24, 64, 32
    st = _stack(E, N, K, bits, gs, seed=bits)
    rng = np.random.default_rng(bits)
    bias = A.f32_to_bf16_bits(rng.standard_normal((E, N)).astype(np.float32))
    # the SwitchGLU call shape: x [T, 1, 1, K] against indices [T, topk] (switch_layers.py:181-196)
    T, topk = 6, 3
    x = A.synth_x(T, K, seed=1).reshape(T, 1, 1, K)
    idx = rng.integers(0, E, size=(T, topk))
    lhs = np.arange(T).reshape(T, 1)
    y = A.gather_qmm(x, st["qweight"], st["scales"], st["zeros"], lhs, idx, True, gs, bits, "bf16", bias=bias)
    assert y.shape == (T, topk, 1, N)
    for t in range(T):
        for j in range(topk):
            e = idx[t, j]
            ref = A.quantized_matmul(x[t, 0], st["qweight"][e], st["scales"][e], st["zeros"][e], gs, bits, "bf16", "f64", bias=bias[e])
            assert np.array_equal(y[t, j], ref)
    # default indices: batch item i of x against matrix i of w
    x2 = A.synth_x(E * 2, K, seed=2).reshape(E, 2, K)
    y2 = A.gather_qmm(x2, st["qweight"], st["scales"], st["zeros"], None, None, True, gs, bits, "bf16")
    for e in range(E):
        assert np.array_equal(y2[e], A.quantized_matmul(x2[e], st["qweight"][e], st["scales"][e], st["zeros"][e], gs, bits, "bf16", "f64"))
