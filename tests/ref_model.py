"""Independent fp32 PyTorch reference of the Llama / Qwen2 decoder (reference: gbx_lm/models/qllama.py,
qqwen2.py) on DENSE weights obtained from the CPU oracle's dequantize -- used only by tests to check the
callers of the hot path (logits / greedy tokens).  Shares no code with gbx_lm_b200."""
import math

import numpy as np
import torch

from oracle import mlx_affine as A


def dense_weights_c(ckpt: dict, plan_bits, device="cpu"):
    """As dense_weights, through the C/OpenMP oracle (bit-identical to the numpy one, tests/test_oracle.py) and moved to
    `device` layer by layer: for full-size architectures (Llama-3.2-1B: 1.2e9 weights)."""
    from oracle import c_oracle as C

    out = {}
    for m in sorted({k.rsplit(".", 1)[0] for k in ckpt if k.endswith(".qweight")}):
        bits, gs = plan_bits(m)
        qw = ckpt[m + ".qweight"].view(torch.int32).numpy().view(np.uint32)
        s = ckpt[m + ".scales"].view(torch.int16).numpy().view(np.uint16)
        z = ckpt[m + ".zeros"].view(torch.int16).numpy().view(np.uint16)
        d = C.dequantize(qw, s, z, gs, bits, "bf16")  # uint16 bf16 bit patterns
        out[m + ".weight"] = torch.from_numpy(d.view(np.int16)).view(torch.bfloat16).to(device).float()
        if m + ".bias" in ckpt:
            out[m + ".bias"] = ckpt[m + ".bias"].to(device).float()
    for k, v in ckpt.items():
        if k.rsplit(".", 1)[1] not in ("qweight", "scales", "zeros", "bias") or k.endswith("norm.bias"):
            out[k] = v.to(device).float()
    return out


def dense_weights(ckpt: dict, plan_bits):
    """ckpt: name -> tensor (cpu). plan_bits(module_name) -> (bits, gs). Returns fp32 dense dict."""
    out = {}
    mods = sorted({k.rsplit(".", 1)[0] for k in ckpt if k.endswith(".qweight")})
    for m in mods:
        bits, gs = plan_bits(m)
        qw = ckpt[m + ".qweight"].view(torch.int32).numpy().view(np.uint32)
        s = ckpt[m + ".scales"].view(torch.int16).numpy().view(np.uint16)
        z = ckpt[m + ".zeros"].view(torch.int16).numpy().view(np.uint16)
        out[m + ".weight"] = torch.from_numpy(A.dequantize(qw, s, z, gs, bits, "bf16").astype(np.float32))
        if m + ".bias" in ckpt:
            out[m + ".bias"] = ckpt[m + ".bias"].float()
    for k, v in ckpt.items():
        if k.rsplit(".", 1)[1] not in ("qweight", "scales", "zeros", "bias") or k.endswith("norm.bias"):
            out[k] = v.float()
    return out


def _inv_freq(dims, base, scaling):
    freqs = base ** (torch.arange(0, dims, 2, dtype=torch.float64) / dims)
    if scaling and (scaling.get("rope_type") or scaling.get("type")) == "llama3":
        factor, low, high = scaling["factor"], scaling.get("low_freq_factor", 1.0), scaling.get("high_freq_factor", 4.0)
        old = scaling.get("original_max_position_embeddings", 8192)
        wl = 2 * math.pi * freqs
        f = torch.where(wl > old / low, freqs * factor, freqs)
        smooth = (old / wl - low) / (high - low)
        sf = f / ((1 - smooth) / factor + smooth)
        freqs = torch.where((wl > old / high) & (wl < old / low), sf, f)
    return (1.0 / freqs).float()


def _rope(x, pos, inv):
    ang = pos.float()[:, None] * inv[None, :]
    c, s = ang.cos()[None, None], ang.sin()[None, None]
    h = x.shape[-1] // 2
    x1, x2 = x[..., :h], x[..., h:]
    return torch.cat([x1 * c - x2 * s, x1 * s + x2 * c], -1)


def _rms(x, w, eps):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def forward(W, cfg, tokens):
    """tokens int64 [B, L] -> logits fp32 [B, L, V] (full causal attention, no cache)."""
    H, nh, nkv = cfg["hidden_size"], cfg["num_attention_heads"], cfg["num_key_value_heads"]
    hd = cfg.get("head_dim") or H // nh
    eps = cfg["rms_norm_eps"]
    inv = _inv_freq(hd, cfg["rope_theta"], cfg.get("rope_scaling"))
    B, L = tokens.shape
    dev = tokens.device
    inv = inv.to(dev)
    pos = torch.arange(L, device=dev)
    h = W["model.embed_tokens.weight"][tokens]
    lin = lambda x, name: torch.nn.functional.linear(x, W[name + ".weight"], W.get(name + ".bias"))
    for i in range(cfg["num_hidden_layers"]):
        p = f"model.layers.{i}."
        x = _rms(h, W[p + "input_layernorm.weight"], eps)
        q = lin(x, p + "self_attn.q_proj").view(B, L, nh, hd).transpose(1, 2)
        k = lin(x, p + "self_attn.k_proj").view(B, L, nkv, hd).transpose(1, 2)
        v = lin(x, p + "self_attn.v_proj").view(B, L, nkv, hd).transpose(1, 2)
        q, k = _rope(q, pos, inv), _rope(k, pos, inv)
        k = k.repeat_interleave(nh // nkv, 1)
        v = v.repeat_interleave(nh // nkv, 1)
        att = (q @ k.transpose(-1, -2)) * hd ** -0.5
        att = att.masked_fill(torch.triu(torch.ones(L, L, dtype=torch.bool, device=dev), 1), float("-inf")).softmax(-1)
        o = (att @ v).transpose(1, 2).reshape(B, L, -1)
        h = h + lin(o, p + "self_attn.o_proj")
        x = _rms(h, W[p + "post_attention_layernorm.weight"], eps)
        h = h + lin(torch.nn.functional.silu(lin(x, p + "mlp.gate_proj")) * lin(x, p + "mlp.up_proj"), p + "mlp.down_proj")
    h = _rms(h, W["model.norm.weight"], eps)
    head = W["model.embed_tokens.weight"] if cfg.get("tie_word_embeddings", True) else W["lm_head.weight"]
    return h @ head.t()
