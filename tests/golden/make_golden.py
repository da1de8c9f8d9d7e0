"""Generates tests/golden/affine_golden.json -- known-answer vectors for the MLX affine format.

MLX is not importable in this container (nor on the GPU box), so these vectors pin OUR READING of
the format: they were produced by the numpy oracle (oracle/mlx_affine.py) and are
cross-checked in tests/test_oracle.py against (a) the hex constants derived by hand in SURVEY.md
8c, (b) MLX's byte rules for 3/6-bit restated independently, (c) torch's own bf16 arithmetic.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mlx_affine as A  # noqa: E402


def main():
    out = {"ramps": {}, "single_hot": {}, "dequant": [], "qmm": []}
    for bits in A.SUPPORTED_BITS:
        q = (np.arange(32) % (1 << bits)).astype(np.uint8)[None]
        out["ramps"][str(bits)] = [int(v) for v in A.pack_codes(q, bits)[0]]
        hots = {}
        for k in range(32):
            q = np.zeros((1, 32), dtype=np.uint8)
            q[0, k] = (1 << bits) - 1
            hots[str(k)] = [int(v) for v in A.pack_codes(q, bits)[0]]
        out["single_hot"][str(bits)] = hots
    # dequant known answers: scale=bf16(0.0123), bias=bf16(-0.771), all codes of each width
    s = A.f32_to_bf16_bits(np.array([[0.0123]], dtype=np.float32))
    b = A.f32_to_bf16_bits(np.array([[-0.771]], dtype=np.float32))
    for bits in A.SUPPORTED_BITS:
        q = (np.arange(32) * 37 % (1 << bits)).astype(np.uint8)[None]
        w = A.pack_codes(q, bits)
        for dtype in ("bf16", "f16", "f32"):
            if dtype == "bf16":
                sc, bi = s, b
            elif dtype == "f16":
                sc = np.array([[0.0123]], dtype=np.float16)
                bi = np.array([[-0.771]], dtype=np.float16)
            else:
                sc = np.array([[0.0123]], dtype=np.float32)
                bi = np.array([[-0.771]], dtype=np.float32)
            d = A.dequantize(w, sc, bi, 32, bits, dtype)
            out["dequant"].append(
                {
                    "bits": bits,
                    "dtype": dtype,
                    "group_size": 32,
                    "codes": [int(v) for v in q[0]],
                    "qweight": [int(v) for v in w[0]],
                    "out_f32_hex": [int(v) for v in d[0].astype(np.float32).view(np.uint32)],
                }
            )
    # small matmul cases (seeded), fp64-truth outputs rounded to bf16
    case = 0
    for bits in A.SUPPORTED_BITS:
        for gs in (32, 64, 128):
            N, K, M = 8, 256, 2
            L = A.synth_layer(N, K, bits, gs, seed=1000 + case)
            x = A.synth_x(M, K, seed=2000 + case)
            y = A.quantized_matmul(x, L["qweight"], L["scales"], L["zeros"], gs, bits, "bf16", "f64")
            out["qmm"].append(
                {
                    "bits": bits,
                    "group_size": gs,
                    "N": N,
                    "K": K,
                    "M": M,
                    "qweight": L["qweight"].tolist(),
                    "scales_bf16": L["scales"].tolist(),
                    "zeros_bf16": L["zeros"].tolist(),
                    "x_bf16": x.tolist(),
                    "y_bf16": A.f32_to_bf16_bits(y).tolist(),
                }
            )
            case += 1
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "affine_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
