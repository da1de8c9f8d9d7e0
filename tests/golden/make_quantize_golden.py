"""Generates tests/golden/quantize_golden.json -- known-answer vectors for the device-side quantiser (gbxq_quantize).

MLX is not importable here, so these vectors pin OUR READING of mx.quantize (oracle/MLX_SPEC.md R6).  They
are produced by the C restatement (oracle/gbxq_oracle.c::gbxq_oracle_quantize), NOT by the numpy one the GPU tests
compare against; tests/test_quantize_cpu.py checks the numpy oracle, the scalar checker and two hand-derived cases
against them, tests/test_gpu_quantize.py checks the kernel.

    python tests/golden/make_quantize_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle as C  # noqa: E402
from oracle import mlx_affine as A  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    cases = []
    for bits in A.SUPPORTED_BITS:
        for gs in (32, 64, 128):
            rows, K = 4, 2 * gs
            w = (rng.standard_normal((rows, K)) * rng.choice([0.02, 1.0, 40.0], size=(rows, 1))).astype(np.float32)
            w[1, :gs] = 0.0                      # a zero group
            w[2, :gs] = -np.abs(w[2, :gs])       # a one-sided group
            wb = A.f32_to_bf16_bits(w)
            q, s, b = C.quantize(wb, gs, bits, "bf16")
            cases.append({"bits": bits, "group_size": gs, "rows": rows, "K": K, "dtype": "bf16",
                          "w_bf16": [int(v) for v in wb.reshape(-1)], "codes": [int(v) for v in q.reshape(-1)],
                          "scales_bf16": [int(v) for v in s.reshape(-1)], "biases_bf16": [int(v) for v in b.reshape(-1)]})
    out = {"generator": "oracle/gbxq_oracle.c::gbxq_oracle_quantize via tests/golden/make_quantize_golden.py", "cases": cases}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quantize_golden.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print(path, os.path.getsize(path), "bytes,", len(cases), "cases")


if __name__ == "__main__":
    main()
