"""Generates tests/golden/hf_metal_golden.json -- vectors for MLX's affine format produced by code that is NOT ours.

Source of the vectors: `transformers.integrations.metal_quantization` (HuggingFace transformers, the copy
installed in this image; functions `_affine_quantize_tensor` / `_affine_dequantize_tensor`).  That module is the
host side of HF's "Metal affine quantization": the tensors it packs are consumed by the MLX Metal kernels
(`kernels-community/mlx-quantization-metal-kernels`, entry `affine_qmm_t(x, weight, scales, qbiases, group_size,
bits)`), i.e. it is an independently written producer of exactly the layout `mx.quantized_matmul(transpose=True)`
reads (the reference's call: gbx_lm/models/quantized_linear_gba.py:195-203).  It covers the power-of-two widths
(2, 4, 8 bits: `elems_per_int = 32 // bits`); 3- and 6-bit stay pinned only by the byte rules in oracle/MLX_SPEC.md.

What the fixture pins (checked in tests/test_oracle.py against oracle/mlx_affine.py, oracle/gbxq_oracle.c and
oracle/indep_check.py, and on the GPU in tests/test_gpu_parity.py):
  * packing: code k of a row sits at bits [bits*(k % (32/bits)) ...) of word k // (32/bits)   (LSB first)
  * the affine map  w = scale * q + bias  with `biases` ADDITIVE (the sign the reference fixes up at
    quantized_linear_gba.py:151-155), per group of `group_size` consecutive k
  * fp32 dequantised values (HF's helper evaluates in fp32 with one rounding; MLX's own two-rounding rule for
    16-bit scales is NOT exercised by this source and is stated as such in oracle/MLX_SPEC.md)

    python tests/golden/make_hf_golden.py          (needs `transformers`; run in the build container)
"""
import json
import os

import numpy as np
import torch
from transformers.integrations import metal_quantization as HF

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hf_metal_golden.json")


def main():
    import transformers

    cases = []
    g = torch.Generator().manual_seed(20261018)
    for bits in (2, 4, 8):
        for gs in (32, 64, 128):
            n, k = 6, 2 * gs if gs >= 64 else 4 * gs
            w = torch.randn((n, k), generator=g) * 0.05
            w[0, :gs] = torch.linspace(-0.3, 0.4, gs)  # a ramp group: every code appears
            wq, scales, biases = HF._affine_quantize_tensor(w, gs, bits)
            deq = HF._affine_dequantize_tensor(wq, scales, biases, gs, bits)
            x = torch.randn((3, k), generator=g)
            y = x.double() @ deq.double().T
            cases.append({
                "bits": bits, "group_size": gs, "N": n, "K": k,
                "qweight": wq.view(torch.int32).numpy().view(np.uint32).tolist(),
                "scales_f32_hex": scales.float().numpy().view(np.uint32).tolist(),
                "biases_f32_hex": biases.float().numpy().view(np.uint32).tolist(),
                "dequant_f32_hex": deq.float().numpy().view(np.uint32).tolist(),
                "x_f32_hex": x.float().numpy().view(np.uint32).tolist(),
                "y_f64": y.numpy().tolist(),
            })
    json.dump({"source": "transformers.integrations.metal_quantization", "transformers_version": transformers.__version__,
               "cases": cases}, open(OUT, "w"))
    print(OUT, len(cases), "cases")


if __name__ == "__main__":
    main()
