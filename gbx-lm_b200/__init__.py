"""gbx_lm_b200 -- B200 (sm_100a) drop-in for the quantized-matmul hot path of GreenBitAI/gbx-lm.

Scope: `QuantizedLinear.__call__` -> `mx.quantized_matmul` / `mx.dequantize`
(gbx_lm/models/quantized_linear_gba.py:179-207), the layer-mix weight loader
(gbx_lm/utils.py:752-879) and tensor-parallel sharding of the Llama / Qwen2 projections.
Everything numerical runs in libgbxq.so (hand-written CUDA behind the C ABI of include/gbxq.h);
there is no CPU, Triton or PyTorch fallback."""

__version__ = "0.1.0"

from .ops import dequantize, quantized_matmul, quantized_matmul_grouped  # noqa: F401
from .quantized_linear import QuantizedLinear  # noqa: F401
