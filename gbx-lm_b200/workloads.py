"""Synthetic workloads of BASELINE.json (`configs`), shared by bench.py, the tests and the checkpoint
writer: model dimensions and the deterministic layer-mix strategies in
the reference's `quant_strategy.json` schema (read at gbx_lm/utils.py:787-792, consumed at
gbx_lm/models/quantized_linear_gba.py:256-276)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

PROJS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


@dataclass(frozen=True)
class ModelDims:
    name: str
    model_type: str
    hidden: int
    inter: int
    layers: int
    heads: int
    kv_heads: int
    head_dim: int
    vocab: int
    tie: bool
    rope_theta: float = 500000.0
    rope_scaling: Optional[dict] = None
    qkv_bias: bool = False
    rms_eps: float = 1e-5
    max_pos: int = 131072

    def proj_shapes(self) -> Dict[str, Tuple[int, int]]:
        """(N, K) of the seven projections of one block (qllama.py:55-58,110-112)."""
        q = self.heads * self.head_dim
        kv = self.kv_heads * self.head_dim
        return {
            "q_proj": (q, self.hidden), "k_proj": (kv, self.hidden), "v_proj": (kv, self.hidden),
            "o_proj": (self.hidden, q), "gate_proj": (self.inter, self.hidden),
            "up_proj": (self.inter, self.hidden), "down_proj": (self.hidden, self.inter),
        }


_L3 = {"factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
       "original_max_position_embeddings": 8192, "rope_type": "llama3"}
_L31 = dict(_L3, factor=8.0)

MODELS = {
    "llama-3.2-1b": ModelDims("llama-3.2-1b", "llama", 2048, 8192, 16, 32, 8, 64, 128256, True, rope_scaling=_L3),
    "llama-3.2-3b": ModelDims("llama-3.2-3b", "llama", 3072, 8192, 28, 24, 8, 128, 128256, True, rope_scaling=_L3),
    "llama-3-8b": ModelDims("llama-3-8b", "llama", 4096, 14336, 32, 32, 8, 128, 128256, False, max_pos=8192),
    "qwen2.5-32b": ModelDims("qwen2.5-32b", "qwen2", 5120, 27648, 64, 40, 8, 128, 152064, False,
                             rope_theta=1000000.0, qkv_bias=True, rms_eps=1e-6, max_pos=32768),
    "llama-3-70b": ModelDims("llama-3-70b", "llama", 8192, 28672, 80, 64, 8, 128, 128256, False, max_pos=8192),
    # tiny architecture-complete models for tests
    "tiny-llama": ModelDims("tiny-llama", "llama", 256, 512, 2, 4, 2, 64, 512, True, rope_scaling=_L3, max_pos=2048),
    "tiny-qwen2": ModelDims("tiny-qwen2", "qwen2", 256, 512, 2, 4, 2, 64, 512, False, rope_theta=1000000.0,
                            qkv_bias=True, rms_eps=1e-6, max_pos=2048),
}


def strategy_bpw40(layers: int) -> dict:
    """layer-mix bpw-4.0 ("2/4-bit layers, group_size 64"): everything 4-bit gs64 except up_proj in
    even layers and gate_proj in odd layers at 2-bit gs64."""
    meas = {}
    for i in range(layers):
        blk = {}
        for p in PROJS:
            two = (p == "up_proj" and i % 2 == 0) or (p == "gate_proj" and i % 2 == 1)
            b = 2 if two else 4
            blk[p] = {"bits": [b], "group_size": {str(b): 64}}
        meas[f"model.layers.{i}"] = blk
    return {"measurement": meas}


def strategy_bpw22(layers: int) -> dict:
    """layer-mix bpw-2.2 ("2-bit-heavy, with 3/6-bit packed layers"): 2-bit gs128 everywhere except,
    in every 4th layer, o_proj and down_proj at 3-bit gs64 and v_proj at 6-bit gs64."""
    meas = {}
    for i in range(layers):
        blk = {}
        for p in PROJS:
            b, g = 2, 128
            if i % 4 == 0 and p in ("o_proj", "down_proj"):
                b, g = 3, 64
            if i % 4 == 0 and p == "v_proj":
                b, g = 6, 64
            blk[p] = {"bits": [b], "group_size": {str(b): g}}
        meas[f"model.layers.{i}"] = blk
    return {"measurement": meas}


def strategy_uniform(layers: int, bits: int, group_size: int) -> dict:
    blk = {p: {"bits": [bits], "group_size": {str(bits): group_size}} for p in PROJS}
    return {"measurement": {f"model.layers.{i}": dict(blk) for i in range(layers)}}


STRATEGIES = {"bpw-4.0": strategy_bpw40, "bpw-2.2": strategy_bpw22}


def layer_plan(dims: ModelDims, strategy: Optional[dict], default_bits: int = 4, default_gs: int = 64):
    """[(layer, proj, N, K, bits, group_size)] for every QuantizedLinear of the model, resolved the way
    reinit_module does (bits = strg["bits"][0]; group_size = strg["group_size"][str(bits)])."""
    shapes = dims.proj_shapes()
    plan = []
    for i in range(dims.layers):
        for p in PROJS:
            if strategy is None:
                b, g = default_bits, default_gs
            else:
                st = strategy["measurement"][f"model.layers.{i}"][p]
                b = st["bits"][0]
                g = st["group_size"][str(b)]
            n, k = shapes[p]
            plan.append((i, p, n, k, b, g))
    return plan


def qmm_bytes(m: int, n: int, k: int, bits: int, gs: int, with_bias: bool = False, esz: int = 2) -> int:
    """ALGORITHMIC bytes of one quantized matmul launch (SURVEY.md 8d / BASELINE.md section 3):
    packed weights + scales + biases + x + y (+ bias)."""
    return n * k * bits // 8 + 2 * n * (k // gs) * esz + m * k * esz + m * n * esz + (n * esz if with_bias else 0)


def stored_bpw(plan) -> float:
    wbits = sum(n * k * b + 2 * n * (k // g) * 16 for (_, _, n, k, b, g) in plan)
    return wbits / sum(n * k for (_, _, n, k, _, _) in plan)
