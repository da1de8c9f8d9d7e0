"""Third, independently written checker of the MLX affine format (TEST INFRASTRUCTURE ONLY).

Derived directly from the documentation passages quoted in oracle/MLX_SPEC.md (D1, D2, D3) and deliberately sharing
no code, helper or numeric trick with oracle/mlx_affine.py or oracle/gbxq_oracle.c: codes are read one BIT at a time
from the word stream, arithmetic is exact (`fractions.Fraction`), and rounding to bf16 / fp16 / fp32 is a from-scratch
round-to-nearest-even on the exact rational.  Pure-Python loops: small cases only.  tests/test_oracle.py compares the
numpy and C oracles and the golden fixtures against it."""
from fractions import Fraction
import struct

FORMATS = {"bf16": (8, 8), "f16": (5, 11), "f32": (8, 24)}  # (exponent bits, significand bits incl. hidden one)


def code_at(words, k, bits):
    """Code k of a row: stream bits [k*bits, (k+1)*bits), LSB first; stream bit j lives in bit j%32 of word j//32 (D1)."""
    v = 0
    for i in range(bits):
        j = k * bits + i
        v |= ((int(words[j // 32]) >> (j % 32)) & 1) << i
    return v


def put_code(words, k, bits, q):
    for i in range(bits):
        j = k * bits + i
        if (q >> i) & 1:
            words[j // 32] |= 1 << (j % 32)


def rne(v: Fraction, fmt: str) -> Fraction:
    """Round an exact rational to the nearest value of the binary format `fmt` (ties to even); finite, normal or
    subnormal results only (the tests stay far from overflow)."""
    if v == 0:
        return Fraction(0)
    ebits, p = FORMATS[fmt]
    emin = 2 - (1 << (ebits - 1))  # exponent of the smallest normal
    sign = -1 if v < 0 else 1
    a = -v if v < 0 else v
    e = a.numerator.bit_length() - a.denominator.bit_length()  # floor(log2 a) or one above
    if Fraction(2) ** e > a:
        e -= 1
    e = max(e, emin)
    ulp = Fraction(2) ** (e - (p - 1))
    n = a / ulp
    f = n.numerator // n.denominator
    r = n - f
    if r > Fraction(1, 2) or (r == Fraction(1, 2) and f % 2 == 1):
        f += 1
    return sign * f * ulp


def from_bits(u: int, fmt: str) -> Fraction:
    if fmt == "bf16":
        return Fraction(struct.unpack("<f", struct.pack("<I", (u & 0xFFFF) << 16))[0])
    if fmt == "f16":
        return Fraction(struct.unpack("<e", struct.pack("<H", u & 0xFFFF))[0])
    return Fraction(struct.unpack("<f", struct.pack("<I", u & 0xFFFFFFFF))[0])


def to_f32_bits(v: Fraction) -> int:
    """bit pattern of v as fp32 (v must be exactly representable: every bf16/f16/f32 value is)"""
    return struct.unpack("<I", struct.pack("<f", float(v)))[0]


def dequantize_row(words, scales, biases, group_size, bits, fmt, K):
    """D2 in the dtype of `scales`: multiply rounded to T, then add rounded to T (rule R3).  `scales` / `biases` are
    exact Fractions of T values."""
    out = []
    for k in range(K):
        q = code_at(words, k, bits)
        g = k // group_size
        out.append(rne(rne(scales[g] * q, fmt) + biases[g], fmt))
    return out


def qmm_row(x, words, scales, biases, group_size, bits, K):
    """D3, exact: sum_k x[k] * (s * q + beta) as a rational."""
    acc = Fraction(0)
    for k in range(K):
        g = k // group_size
        acc += x[k] * (scales[g] * code_at(words, k, bits) + biases[g])
    return acc


def _f32(v: float) -> float:
    """one fp32 rounding of a double: for +, -, / of two fp32 operands the double result rounded to fp32 equals the
    correctly rounded fp32 operation (53 >= 2 * 24 + 2 significand bits)"""
    return struct.unpack("<f", struct.pack("<f", v))[0]


def _rint(v: float) -> float:
    f = float(int(v))  # toward zero
    d = v - f
    if abs(d) > 0.5 or (abs(d) == 0.5 and int(f) % 2 != 0):
        f += 1.0 if v > 0 else -1.0
    return f


def quantize_group(ws, bits):
    """R6 for ONE group: `ws` are the fp32 values of the group's elements (python floats holding T values).  Returns
    (codes, scale, bias) with scale / bias still in fp32 -- the caller rounds them to T for storage.  Written from the
    step list of R6 with scalar python arithmetic, one explicit fp32 rounding per operation."""
    n_bins = float((1 << bits) - 1)
    w_max, w_min = max(ws), min(ws)
    scale = max(_f32(_f32(w_max - w_min) / n_bins), _f32(1e-7))
    if abs(w_min) > abs(w_max):
        edge = w_min
    else:
        edge, scale = w_max, -scale
    q0 = _rint(_f32(edge / scale))
    bias = 0.0
    if q0 != 0:
        scale, bias = _f32(edge / q0), edge
    codes = [int(min(max(_rint(_f32(_f32(w - bias) / scale)), 0.0), n_bins)) for w in ws]
    return codes, scale, bias
