"""CPU oracle for the gbx-lm quantized-matmul path (TEST INFRASTRUCTURE ONLY).

This module restates, in numpy, the arithmetic that the reference reaches through

    mx.quantized_matmul(x, qweight, scales=scales, biases=zeros, transpose=True,
                        group_size=gs, bits=b)      gbx_lm/models/quantized_linear_gba.py:195-203
    mx.dequantize(w, scales, biases, gs, bits)      gbx_lm/tuner/lora.py:62-68,
                                                    gbx_lm/tuner/utils.py:214-220,
                                                    gbx_lm/tuner/dora.py:101

The algorithm itself lives in the third-party dependency `mlx` (ml-explore/mlx; the
reference requires `mlx>=0.25.0`, requirements.txt:2 / setup.py:12 -- a floor, not an
exact pin).  MLX is neither under /root/reference nor installable here (no wheel, no
network), and the reference's own tests hold no golden vector for this path
(SURVEY.md section 8c).  Therefore:

    ****  PARITY UNPINNED  ****
    The semantics below are MLX's published affine-quantisation format restated from
    its documentation and source as remembered (mlx/ops.cpp `dequantize`,
    mlx/backend/cpu/quantized.cpp `_qmm_t` / `extract_bits`); they are pinned against
    hand-derived known-answer vectors (tests/golden/), NOT against a live MLX run.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (gbx_lm_b200) never does.

Format (MLX "affine" quantisation):
  * Row n of `w` (uint32 [N, K*bits/32]) is an LSB-first bitstream of the K codes of
    output feature n: code k occupies stream bits [k*bits, (k+1)*bits); the stream is cut
    into little-endian uint32 words.  For bits in {3, 6} codes straddle byte/word borders.
  * dequantize:  W[n,k] = T( T(scales[n,k//gs] * T(q[n,k])) + biases[n,k//gs] ), T = dtype of
    `scales`; two roundings (multiply, then add), each exact-in-fp32 then round-to-nearest-even.
  * quantized_matmul(transpose=True):  y[m,n] = sum_k x[m,k] * W[n,k]; the summation order and
    accumulator precision are NOT bit-specified by MLX (Metal: fp32; CPU SIMD: fp32; CPU scalar
    x86: T).  The oracle offers an fp64 "truth", an fp32 k-ordered variant and the T-accumulate
    x86-scalar variant.
"""
from __future__ import annotations

import numpy as np

SUPPORTED_BITS = (2, 3, 4, 6, 8)
SUPPORTED_GROUP_SIZES = (32, 64, 128)

# ----------------------------------------------------------------------------------------
# dtype helpers.  bf16 has no numpy dtype: a bf16 array is carried as uint16 bit patterns
# wrapped in `BF16` so that shapes/indexing stay numpy-like.
# ----------------------------------------------------------------------------------------


def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even), returning uint16 bit patterns."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32)
    nan = np.isnan(a)
    rounding_bias = ((u >> 16) & 1) + np.uint32(0x7FFF)
    r = ((u + rounding_bias) >> 16).astype(np.uint16)
    r = np.where(nan, np.uint16(0x7FC0), r).astype(np.uint16)
    return r


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


class BF16:
    """A bf16 tensor as uint16 bits (`.bits`) with a float32 view (`.f32`)."""

    def __init__(self, bits: np.ndarray):
        self.bits = np.ascontiguousarray(bits, dtype=np.uint16)

    @classmethod
    def from_f32(cls, a) -> "BF16":
        return cls(f32_to_bf16_bits(np.asarray(a, dtype=np.float32)))

    @property
    def f32(self) -> np.ndarray:
        return bf16_bits_to_f32(self.bits)

    @property
    def shape(self):
        return self.bits.shape


def _round_to(a: np.ndarray, dtype: str) -> np.ndarray:
    """Round an fp32 (or fp64) array to `dtype` and return it widened back to fp32."""
    if dtype == "bf16":
        return bf16_bits_to_f32(f32_to_bf16_bits(np.asarray(a, dtype=np.float32)))
    if dtype == "f16":
        return np.asarray(a, dtype=np.float32).astype(np.float16).astype(np.float32)
    if dtype == "f32":
        return np.asarray(a, dtype=np.float32)
    raise ValueError(f"dtype must be bf16/f16/f32, got {dtype}")


def _as_f32(a, dtype: str) -> np.ndarray:
    """Accept BF16 / float16 / float32 containers and return exact fp32 values."""
    if isinstance(a, BF16):
        if dtype != "bf16":
            raise ValueError("BF16 container passed for a non-bf16 dtype")
        return a.f32
    a = np.asarray(a)
    if a.dtype == np.uint16 and dtype == "bf16":
        return bf16_bits_to_f32(a)
    return _round_to(a.astype(np.float32), dtype)


# ----------------------------------------------------------------------------------------
# packing  (SURVEY.md section 8c rule 1; MLX: mlx/backend/cpu/quantized.cpp extract_bits,
# allocation shape at quantized_linear_gba.py:61 -> (out, in // 32 * bits))
# ----------------------------------------------------------------------------------------


def _check_bits_gs(bits: int, group_size: int | None = None):
    if bits not in SUPPORTED_BITS:
        raise ValueError(f"bits must be one of {SUPPORTED_BITS}, got {bits}")
    if group_size is not None and group_size not in SUPPORTED_GROUP_SIZES:
        raise ValueError(f"group_size must be one of {SUPPORTED_GROUP_SIZES}, got {group_size}")


def pack_codes(q: np.ndarray, bits: int) -> np.ndarray:
    """codes uint8 [..., K] (values < 2**bits) -> uint32 [..., K*bits/32], LSB-first bitstream."""
    _check_bits_gs(bits)
    q = np.asarray(q)
    K = q.shape[-1]
    if K % 32:
        raise ValueError("K must be a multiple of 32")
    if q.size and int(q.max()) >= (1 << bits):
        raise ValueError("code out of range")
    lead = q.shape[:-1]
    q2 = q.reshape(-1, K).astype(np.uint8)
    # bit matrix [R, K, bits] (LSB first) -> flat stream -> bytes (little bit order) -> u32 LE
    bitsel = ((q2[:, :, None] >> np.arange(bits, dtype=np.uint8)) & 1).astype(np.uint8)
    stream = bitsel.reshape(q2.shape[0], K * bits)
    by = np.packbits(stream, axis=-1, bitorder="little")
    words = by.view("<u4") if by.size else by.reshape(q2.shape[0], 0).view("<u4")
    return np.ascontiguousarray(words.reshape(*lead, K * bits // 32)).astype(np.uint32)


def unpack_codes(w: np.ndarray, bits: int) -> np.ndarray:
    """uint32 [..., K*bits/32] -> codes uint8 [..., K]  (inverse of pack_codes)."""
    _check_bits_gs(bits)
    w = np.ascontiguousarray(np.asarray(w, dtype=np.uint32))
    lead = w.shape[:-1]
    nw = w.shape[-1]
    if (nw * 32) % bits:
        raise ValueError("word count incompatible with bits")
    K = nw * 32 // bits
    by = w.reshape(-1, nw).astype("<u4").view(np.uint8)
    stream = np.unpackbits(by, axis=-1, bitorder="little")
    stream = stream.reshape(-1, K, bits)
    q = (stream.astype(np.uint16) << np.arange(bits, dtype=np.uint16)).sum(axis=-1)
    return q.astype(np.uint8).reshape(*lead, K)


def unpack_codes_byte_rules(w: np.ndarray, bits: int) -> np.ndarray:
    """Independent restatement for the straddling widths, written from MLX's byte rules
    (3-bit: 8 codes per 3 bytes; 6-bit: 4 codes per 3 bytes).  Used only to
    cross-check `unpack_codes`."""
    w = np.ascontiguousarray(np.asarray(w, dtype=np.uint32))
    lead = w.shape[:-1]
    by = w.reshape(-1, w.shape[-1]).astype("<u4").view(np.uint8).astype(np.uint16)
    R = by.shape[0]
    if bits == 3:
        b = by.reshape(R, -1, 3)
        b0, b1, b2 = b[..., 0], b[..., 1], b[..., 2]
        v = np.stack(
            [
                b0 & 7,
                (b0 >> 3) & 7,
                (b0 >> 6) | ((b1 & 1) << 2),
                (b1 >> 1) & 7,
                (b1 >> 4) & 7,
                (b1 >> 7) | ((b2 & 3) << 1),
                (b2 >> 2) & 7,
                b2 >> 5,
            ],
            axis=-1,
        )
    elif bits == 6:
        b = by.reshape(R, -1, 3)
        b0, b1, b2 = b[..., 0], b[..., 1], b[..., 2]
        v = np.stack(
            [b0 & 63, (b0 >> 6) | ((b1 & 15) << 2), (b1 >> 4) | ((b2 & 3) << 4), b2 >> 2],
            axis=-1,
        )
    elif bits in (2, 4, 8):
        per = 8 // bits
        v = np.stack([(by >> (bits * i)) & ((1 << bits) - 1) for i in range(per)], axis=-1)
    else:
        raise ValueError(bits)
    return v.reshape(R, -1).astype(np.uint8).reshape(*lead, -1)


# ----------------------------------------------------------------------------------------
# validation (reference asserts at quantized_linear_gba.py:250,272, utils.py:819-821)
# ----------------------------------------------------------------------------------------


def check_shapes(w_shape, s_shape, b_shape, group_size: int, bits: int, x_last: int | None = None):
    _check_bits_gs(bits, group_size)
    if tuple(s_shape) != tuple(b_shape):
        raise ValueError("scales and biases must have the same shape")
    if len(w_shape) < 2 or tuple(w_shape[:-1]) != tuple(s_shape[:-1]):
        raise ValueError("qweight / scales leading dims disagree")
    K = w_shape[-1] * 32 // bits
    if (w_shape[-1] * 32) % bits or K != s_shape[-1] * group_size:
        raise ValueError(
            f"qweight last dim {w_shape[-1]} (bits={bits}) does not match scales last dim "
            f"{s_shape[-1]} (group_size={group_size})"
        )
    if x_last is not None and x_last != K:
        raise ValueError(f"x last dim {x_last} != K {K}")
    return K


# ----------------------------------------------------------------------------------------
# dequantize  (mx.dequantize)
# ----------------------------------------------------------------------------------------


def dequantize(w, scales, biases, group_size: int = 64, bits: int = 4, dtype: str = "bf16") -> np.ndarray:
    """Returns the dequantised matrix as fp32 values that are exactly representable in `dtype`.

    W = T( T(scale * T(q)) + bias ): q <= 255 is exact in every T; the product and the sum are
    each computed exactly enough in fp32/fp64 and rounded to T (two roundings, no FMA)."""
    w = np.asarray(w, dtype=np.uint32)
    s = _as_f32(scales, dtype)
    b = _as_f32(biases, dtype)
    K = check_shapes(w.shape, s.shape, b.shape, group_size, bits)
    q = unpack_codes(w, bits).astype(np.float32)
    s_full = np.repeat(s, group_size, axis=-1)
    b_full = np.repeat(b, group_size, axis=-1)
    # MLX's bf16/f16 scalar types compute every binary op as float(a) op float(b) in fp32 and
    # round the fp32 result to T; mirror exactly that (fp32 op, then one rounding to T) per op.
    prod = _round_to((s_full * q).astype(np.float32), dtype)
    out = _round_to((prod + b_full).astype(np.float32), dtype)
    assert out.shape[-1] == K
    return out


# ----------------------------------------------------------------------------------------
# quantized_matmul  (mx.quantized_matmul transpose=True)
# ----------------------------------------------------------------------------------------


def quantized_matmul(
    x,
    w,
    scales,
    biases,
    group_size: int = 64,
    bits: int = 4,
    dtype: str = "bf16",
    accum: str = "f64",
    bias=None,
) -> np.ndarray:
    """y[..., n] = sum_k x[..., k] * (scales[n,k//gs]*q[n,k] + biases[n,k//gs])  (+ bias[n]).

    accum:
      "f64"   exact affine weights (no intermediate rounding), fp64 accumulate -> the "truth";
      "f32"   MLX CPU SIMD-path order: fp32 weights (scale*q+bias in fp32), fp32 accumulate in k order;
      "deq"   matmul against the T-rounded `dequantize` matrix, fp64 accumulate.
    Result is rounded once to `dtype` and returned as fp32.  The optional `bias` follows
    QuantizedLinear.__call__ (quantized_linear_gba.py:204-205): a second, separately rounded add."""
    xf = _as_f32(x, dtype)
    w = np.asarray(w, dtype=np.uint32)
    s = _as_f32(scales, dtype)
    b = _as_f32(biases, dtype)
    K = check_shapes(w.shape, s.shape, b.shape, group_size, bits, xf.shape[-1])
    lead = xf.shape[:-1]
    x2 = xf.reshape(-1, K)
    q = unpack_codes(w, bits)
    if accum == "f64":
        W = np.repeat(s, group_size, -1).astype(np.float64) * q.astype(np.float64) + np.repeat(
            b, group_size, -1
        ).astype(np.float64)
        y = x2.astype(np.float64) @ W.T
    elif accum == "deq":
        W = dequantize(w, s, b, group_size, bits, dtype).astype(np.float64)
        y = x2.astype(np.float64) @ W.T
    elif accum == "f32":
        W = (np.repeat(s, group_size, -1) * q.astype(np.float32) + np.repeat(b, group_size, -1)).astype(
            np.float32
        )
        y = np.zeros((x2.shape[0], W.shape[0]), dtype=np.float32)
        for k in range(K):  # strict k order, fp32 accumulate
            y += x2[:, k : k + 1] * W[None, :, k]
    else:
        raise ValueError(accum)
    y = _round_to(y, dtype)
    if bias is not None:
        y = _round_to(y.astype(np.float64) + _as_f32(bias, dtype).astype(np.float64), dtype)
    return y.reshape(*lead, -1)


# ----------------------------------------------------------------------------------------
# quantized_matmul(transpose=False) and gather_qmm  (oracle/MLX_SPEC.md section 5)
# ----------------------------------------------------------------------------------------


def _affine_f64(w, s, b, group_size: int, bits: int) -> np.ndarray:
    """Exact affine matrix scale*q+bias in fp64 along the LAST (quantized) axis; any leading dimensions."""
    q = unpack_codes(np.asarray(w, dtype=np.uint32), bits).astype(np.float64)
    return np.repeat(s, group_size, -1).astype(np.float64) * q + np.repeat(b, group_size, -1).astype(np.float64)


def quantized_matmul_nt(x, w, scales, biases, group_size: int = 64, bits: int = 4, dtype: str = "bf16") -> np.ndarray:
    """mx.quantized_matmul(x, w, scales, biases, transpose=False): y[..., m, n] = sum_k x[..., m, k] * W[..., k, n] with
    W[..., k, :] = scales[..., k, n//gs] * q[..., k, n] + biases[..., k, n//gs] -- the matrix is quantized along its last
    axis N (as `scores @ V` on the quantized cache uses it, gbx_lm/models/base.py:90-92).  Batch dimensions of x and w
    broadcast like a matmul's.  Exact weights, fp64 accumulation, one rounding to `dtype` (the "truth")."""
    xf = _as_f32(x, dtype).astype(np.float64)
    s = _as_f32(scales, dtype)
    b = _as_f32(biases, dtype)
    W = _affine_f64(w, s, b, group_size, bits)  # [..., K, N]
    if xf.shape[-1] != W.shape[-2]:
        raise ValueError(f"[quantized_matmul] x last dim {xf.shape[-1]} != K {W.shape[-2]}")
    return _round_to(np.matmul(xf, W), dtype)


def gather_qmm(x, w, scales, biases, lhs_indices=None, rhs_indices=None, transpose: bool = True, group_size: int = 64,
               bits: int = 4, dtype: str = "bf16", bias=None) -> np.ndarray:
    """mx.gather_qmm as QuantizedSwitchLinear.__call__ uses it (gbx_lm/models/switch_layers.py:78-92): x [..., M, K],
    w [..., N, K*bits/32] (transpose) a stack of matrices; output batch item i multiplies x.reshape(-1, M, K)[lhs[i]]
    with the dequantised matrix w.reshape(-1, ...)[rhs[i]]; lhs / rhs broadcast against each other and default to the
    operands' own flattened batches.  `bias` [E, N]: the layer's separately rounded `+ bias[indices]` (:89-90)."""
    xf = _as_f32(x, dtype).astype(np.float64)
    s = _as_f32(scales, dtype)
    b = _as_f32(biases, dtype)
    W = _affine_f64(w, s, b, group_size, bits)
    x3 = xf.reshape(-1, xf.shape[-2], xf.shape[-1])
    W3 = W.reshape(-1, W.shape[-2], W.shape[-1])
    lhs = np.arange(x3.shape[0]).reshape(xf.shape[:-2]) if lhs_indices is None else np.asarray(lhs_indices)
    rhs = np.arange(W3.shape[0]).reshape(W.shape[:-2]) if rhs_indices is None else np.asarray(rhs_indices)
    lhs, rhs = np.broadcast_arrays(lhs, rhs)
    A = x3[lhs.reshape(-1)]
    B = W3[rhs.reshape(-1)]
    y = np.matmul(A, np.swapaxes(B, -1, -2) if transpose else B)
    y = _round_to(y, dtype)
    if bias is not None:
        bf = _as_f32(bias, dtype).astype(np.float64).reshape(-1, y.shape[-1])
        y = _round_to(y.astype(np.float64) + bf[rhs.reshape(-1)][:, None, :], dtype)
    return y.reshape(*lhs.shape, y.shape[-2], y.shape[-1])


# ----------------------------------------------------------------------------------------
# quantize  (mx.quantize -- MLX_SPEC R6).  Used to fabricate test weights and as the checker of the device-side
# quantiser (gbxq_quantize: the quantized KV cache, gbx_lm/models/cache.py:176-177,251-263).
# ----------------------------------------------------------------------------------------


def quantize(wf: np.ndarray, group_size: int = 64, bits: int = 4, dtype: str = "bf16"):
    """Affine quantisation per MLX's recipe (R6): returns (packed uint32, scales, biases), the statistics as fp32 arrays
    holding T values.  Everything is evaluated in fp32, one rounding per operation; the codes come from the UNROUNDED
    fp32 scale / bias and only the stored statistics are cast to T.  `wf` is any [..., K] array of T values."""
    _check_bits_gs(bits, group_size)
    wf = _round_to(np.asarray(wf, dtype=np.float32), dtype)
    if wf.ndim < 2:
        raise ValueError("[quantize] the matrix to be quantized must have at least 2 dimensions")
    lead, K = wf.shape[:-1], wf.shape[-1]
    if K % group_size:
        raise ValueError("[quantize] the last dimension must be divisible by the group size")
    g = wf.reshape(-1, K // group_size, group_size)
    N = g.shape[0]
    n_bins = np.float32((1 << bits) - 1)
    w_max = g.max(-1)
    w_min = g.min(-1)
    mask = np.abs(w_min) > np.abs(w_max)
    scales = np.maximum((w_max - w_min) / n_bins, np.float32(1e-7))
    scales = np.where(mask, scales, -scales)
    edge = np.where(mask, w_min, w_max)
    q0 = np.rint(edge / scales)
    scales = np.where(q0 != 0, edge / np.where(q0 != 0, q0, np.float32(1)), scales).astype(np.float32)
    biases = np.where(q0 == 0, np.float32(0), edge).astype(np.float32)
    q = np.clip(np.rint((g - biases[..., None]) / scales[..., None]), 0, n_bins).astype(np.uint8)
    packed = pack_codes(q.reshape(N, K), bits)
    return (packed.reshape(*lead, packed.shape[-1]), _round_to(scales, dtype).reshape(*lead, K // group_size),
            _round_to(biases, dtype).reshape(*lead, K // group_size))


# ----------------------------------------------------------------------------------------
# synthetic inputs: identical recipe for tests, bench and the CUDA side
# ----------------------------------------------------------------------------------------


def synth_layer(N: int, K: int, bits: int, group_size: int, seed: int = 0, with_bias: bool = False):
    """codes ~ U{0..2^b-1}; scales ~ U(.5,1.5)*(2/sqrt(K))/(2^b-1) (bf16);
    zeros = -scales*(2^b-1)/2*(1+eps), eps~U(-.05,.05) (bf16).  Returns dict of numpy arrays
    (scales/zeros as bf16 bit patterns in uint16)."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 1 << bits, size=(N, K), dtype=np.uint8)
    nb = (1 << bits) - 1
    s = rng.uniform(0.5, 1.5, size=(N, K // group_size)).astype(np.float32) * np.float32(2.0 / np.sqrt(K) / nb)
    eps = rng.uniform(-0.05, 0.05, size=s.shape).astype(np.float32)
    s_bits = f32_to_bf16_bits(s)
    z = -bf16_bits_to_f32(s_bits) * np.float32(nb / 2.0) * (1.0 + eps)
    out = {
        "codes": q,
        "qweight": pack_codes(q, bits),
        "scales": s_bits,
        "zeros": f32_to_bf16_bits(z),
    }
    if with_bias:
        out["bias"] = f32_to_bf16_bits(rng.normal(0, 0.02, size=(N,)).astype(np.float32))
    return out


def synth_x(M: int, K: int, seed: int = 1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return f32_to_bf16_bits(rng.standard_normal((M, K)).astype(np.float32))
