"""Layer-mix weight loader, checkpoint writer and decode loop -- host-side mirror of the parts of
`gbx_lm/utils.py` that sit on either side of the quantized-matmul path:

    load_model / load        gbx_lm/utils.py:752-920   (strategy -> reinit_module -> load_weights)
    make_shards/save_weights gbx_lm/utils.py:967-988, 1055-1104
    save_config              gbx_lm/utils.py:1107-1127
    generate_step            gbx_lm/utils.py:217-338   (batch-1 prefill in 2048-token chunks + greedy decode)

The on-disk layout is exactly what gba2mlx emits (gbx_lm/gba2mlx.py:47-65): `model*.safetensors` with
`<module>.{qweight,scales,zeros[,bias]}`, metadata {"format": "mlx"}, `config.json` carrying
`"quantization": {"group_size", "bits"}` and an optional `quant_strategy.json`.
Tensor-parallel loading (new work) slices each projection for this rank before it reaches the GPU.
"""
from __future__ import annotations

import glob
import json
import logging
import time
from pathlib import Path
from typing import Any, Callable, Dict, Generator, List, Optional, Tuple, Type, Union

import torch
from torch import nn

from . import qllama, qqwen2
from .quantized_linear import QuantizedLinear
from .tp import TPContext, proj_kind, shard_tensor

MODEL_MAPPING = {  # subset of gbx_lm/utils.py:52-61 that BASELINE.json's configs use
    "llama": qllama,
    "mistral": qllama,  # same remapping as the reference (utils.py:141-159)
    "qwen2": qqwen2,
}
MAX_FILE_SIZE_GB = 5


def _get_classes(config: dict):
    model_type = config["model_type"]
    if model_type not in MODEL_MAPPING:
        msg = f"Model type {model_type} not supported."
        logging.error(msg)
        raise ValueError(msg)
    arch = MODEL_MAPPING[model_type]
    return arch.Model, arch.ModelArgs


def get_parameter_usage_info(weights: Dict[str, Any]) -> Tuple[bool, bool]:
    """(double-quantised statistics present, q_perm present) -- gbx_lm/utils.py:696-719; one implementation, shared
    with the converter."""
    from .gba_convert import parameter_usage

    return parameter_usage(weights)


def _load_safetensors(path: str) -> Dict[str, torch.Tensor]:
    from safetensors.torch import load_file

    return load_file(path, device="cpu")


def load_weights(model: nn.Module, weights: Dict[str, torch.Tensor], strict: bool = False, device=None) -> List[str]:
    """`model.load_weights(list(weights.items()), strict=False)` (utils.py:862): assign by dotted key.
    Unlike MLX (which does not check shapes with strict=False -- SURVEY 3.1) a shape mismatch raises."""
    targets = dict(model.named_buffers())
    targets.update(dict(model.named_parameters()))
    missing = []
    for key, t in weights.items():
        if key not in targets:
            if strict:
                raise KeyError(f"unexpected weight {key}")
            missing.append(key)
            continue
        mod_name, _, leaf = key.rpartition(".")
        mod = model.get_submodule(mod_name) if mod_name else model
        cur = targets[key]
        if cur is not None and tuple(cur.shape) != tuple(t.shape):
            raise ValueError(f"{key}: checkpoint shape {tuple(t.shape)} != module shape {tuple(cur.shape)}")
        t = t.to(device) if device is not None else t
        if leaf in mod._parameters:
            mod._parameters[leaf] = nn.Parameter(t, requires_grad=False)
        else:
            mod._buffers[leaf] = t
    return missing


def load_model(
    model_path: Union[str, Path],
    lazy: bool = False,
    model_config: dict = {},
    bits: int = 4,
    group_size: int = 64,
    is_conversion: bool = False,
    get_model_classes: Callable[[dict], Tuple[Type[nn.Module], Type]] = _get_classes,
    device: Union[str, torch.device] = "cuda",
    tp: Optional[TPContext] = None,
) -> Tuple[nn.Module, dict]:
    """Mirror of gbx_lm/utils.py:752-879 (same steps, same order, same error behaviour)."""
    model_path = Path(model_path)
    tp = tp or TPContext()
    # ======== load strategy.json file ========= #  (utils.py:787-792)
    strategy = None
    try:
        with open(model_path / "quant_strategy.json", "r") as f:
            strategy = json.load(f)["measurement"]
    except FileNotFoundError:
        logging.info(f"[WARNING] Strategy config file not found in {model_path}")
    # ===== load quantization config file ====== #  (utils.py:795-803)
    try:
        with open(model_path / "config.json", "r") as f:
            config = json.load(f)
            quantization = config.get("quantization", None)
            if quantization is None:
                quantization = {"group_size": group_size, "bits": bits}
    except FileNotFoundError:
        logging.info(f"[WARNING] Quantization config file not found in {model_path}")
        raise
    config.update(model_config)

    weight_files = glob.glob(str(model_path / "model*.safetensors"))
    if not weight_files:
        logging.error(f"No safetensors found in {model_path}")
        raise FileNotFoundError(f"No safetensors found in {model_path}")
    weights: Dict[str, torch.Tensor] = {}
    for wf in weight_files:
        weights.update(_load_safetensors(wf))

    use_double_quantization, use_q_perm = get_parameter_usage_info(weights)
    if not use_q_perm:
        assert quantization["group_size"] in [32, 64, 128], \
            f"The group size value ({group_size}) must be 32, 64 or 128."
    if is_conversion or use_double_quantization:
        # an original GBA checkpoint (K-major qweight / statistics, subtractive zeros, optionally double-quantised
        # statistics): brought to the layout of the path on the host, before any module sees it -- the reference's
        # transposes (utils.py:828-838), `prepare_scales_zeros` and `post_processing_and_release` (:864-873)
        from .gba_convert import convert_gba_weights, expand_statistics
        from .quantized_linear import strategy_params

        def gs_of(mod_name: str) -> int:
            return strategy_params(mod_name, strategy, quantization["bits"], quantization["group_size"])[1]

        weights = convert_gba_weights(weights, gs_of) if is_conversion else expand_statistics(weights, gs_of)
        for k in [k for k in weights if k.endswith(".q_perm")]:
            weights[k] = weights[k].reshape(1, 1, -1)  # quantized_linear_gba.py:157-158
        use_double_quantization = False  # plain scales / zeros from here on
    # scales / zeros -> bf16 (utils.py:841-843)
    for k, v in weights.items():
        if "scale" in k or "zeros" in k:
            weights[k] = v.to(torch.bfloat16)

    model_class, model_args_class = get_model_classes(config=config)
    model_args = model_args_class.from_dict(config)
    model = model_class(model_args, tp) if tp.world > 1 else model_class(model_args)
    if hasattr(model, "sanitize"):
        weights = model.sanitize(weights)

    # per-layer (bits, group_size) from the strategy (utils.py:854-860 -> quantized_linear_gba.py:210-283)
    QuantizedLinear.reinit_module(
        model, **quantization, strategy=strategy, use_double_quantization=use_double_quantization, use_q_perm=use_q_perm
    )
    if tp.world > 1:
        mods = dict(model.named_modules())
        sharded = {}
        for k, v in weights.items():
            mod_name = k.rpartition(".")[0]
            m = mods.get(mod_name)
            if isinstance(m, QuantizedLinear) and proj_kind(mod_name) is not None:
                sharded[k] = shard_tensor(k, v, m.bits, m.group_size, tp.rank, tp.world)
            else:
                sharded[k] = v
        weights = sharded
    # embeddings / norms / head follow the activation dtype of the path
    for k, v in list(weights.items()):
        if v.is_floating_point() and v.dtype != torch.bfloat16 and "channel_scale" not in k:
            weights[k] = v.to(torch.bfloat16)

    dev = torch.device(device)
    load_weights(model, weights, strict=False, device=dev)
    # anything the checkpoint did not provide (RoPE tables, absent optional buffers) moves too
    model.to(dev)
    for m in model.modules():
        if isinstance(m, (nn.RMSNorm, nn.Embedding, nn.Linear)):
            m.to(torch.bfloat16)
    if not lazy and dev.type == "cuda":
        torch.cuda.synchronize(dev)
    model.eval()
    for prm in model.parameters():
        prm.requires_grad_(False)
    return model, config


def load(path: Union[str, Path], model_config: dict = {}, lazy: bool = False, device="cuda", tp: Optional[TPContext] = None):
    """gbx_lm.load without the tokenizer / hub download (no network; token ids in, token ids out)."""
    model, config = load_model(Path(path), lazy, model_config=model_config, device=device, tp=tp)
    return model, config


# ------------------------------------------------------------------------------------------ writer
def _nbytes(t: torch.Tensor) -> int:
    return t.numel() * t.element_size()


def make_shards(weights: dict, max_file_size_gb: int = MAX_FILE_SIZE_GB) -> list:
    """The on-disk split of gbx_lm/utils.py:967-988: tensors in insertion order, a tensor that would take the current
    file over the cap opens the next one.  (The reference also emits an EMPTY first shard when the very first tensor is
    over the cap; no empty file is written here.)"""
    cap = max_file_size_gb << 30
    shards, used = [{}], 0
    for name, t in weights.items():
        if shards[-1] and used + _nbytes(t) > cap:
            shards.append({})
            used = 0
        shards[-1][name] = t
        used += _nbytes(t)
    return shards


def save_weights(save_path: Union[str, Path], weights: Dict[str, torch.Tensor], *, donate_weights: bool = False,
                 max_file_size_gb: int = MAX_FILE_SIZE_GB) -> None:
    """The layout `load_model` (and gbx_lm/utils.py:1055-1104) expects: `model.safetensors`, or
    `model-0000i-of-0000n.safetensors` when the split has n > 1 files, each with metadata {"format": "mlx"}, plus
    `model.safetensors.index.json` = {"metadata": {"total_size"}, "weight_map": name -> file, sorted by name}."""
    from safetensors.torch import save_file

    out = Path(save_path)
    out.mkdir(parents=True, exist_ok=True)
    shards = make_shards(weights, max_file_size_gb)
    n = len(shards)
    index = {"metadata": {"total_size": sum(_nbytes(t) for t in weights.values())}, "weight_map": {}}
    if donate_weights:
        weights.clear()  # the shards hold the only references from here on
    for i in range(1, n + 1):
        shard = shards.pop(0)  # dropped after the write, so that a donated checkpoint is freed file by file
        fname = f"model-{i:05d}-of-{n:05d}.safetensors" if n > 1 else "model.safetensors"
        save_file({k: v.contiguous().cpu() for k, v in shard.items()}, str(out / fname), metadata={"format": "mlx"})
        index["weight_map"].update(dict.fromkeys(shard, fname))
    index["weight_map"] = dict(sorted(index["weight_map"].items()))
    (out / "model.safetensors.index.json").write_text(json.dumps(index, indent=4))


def save_config(config: dict, config_path: Union[str, Path]) -> None:
    """gbx_lm/utils.py:1107-1127: keys sorted, `_name_or_path` dropped (from the written copy only)."""
    Path(config_path).write_text(json.dumps({k: config[k] for k in sorted(config) if k != "_name_or_path"}, indent=4))


def write_synthetic_checkpoint(path: Union[str, Path], dims, strategy: Optional[dict], seed: int = 0,
                               default_bits: int = 4, default_gs: int = 64, embed_scale: float = 1.0) -> dict:
    """Random-init checkpoint of `dims` (workloads.ModelDims) in the gba2mlx layout, with the synthetic
    recipe of SURVEY.md 8d.  `strategy` is a quant_strategy.json dict or None (uniform bits/gs from
    config["quantization"], the `strategy is None` branch of reinit_module)."""
    from . import packing
    from .workloads import PROJS, layer_plan

    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    gen = torch.Generator().manual_seed(seed)
    weights: Dict[str, torch.Tensor] = {}
    h = dims.hidden
    weights["model.embed_tokens.weight"] = (torch.randn((dims.vocab, h), generator=gen) * embed_scale).to(torch.bfloat16)
    weights["model.norm.weight"] = (1.0 + 0.05 * torch.randn((h,), generator=gen)).to(torch.bfloat16)
    if not dims.tie:
        weights["lm_head.weight"] = (torch.randn((dims.vocab, h), generator=gen) / h ** 0.5).to(torch.bfloat16)
    for (i, p, n, k, b, g) in layer_plan(dims, strategy, default_bits, default_gs):
        sub = "self_attn" if p in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
        L = packing.synth_layer(n, k, b, g, seed=seed * 100003 + i * 7 + PROJS.index(p),
                                with_bias=dims.qkv_bias and p in ("q_proj", "k_proj", "v_proj"))
        base = f"model.layers.{i}.{sub}.{p}"
        weights[base + ".qweight"] = L["qweight"]
        weights[base + ".scales"] = L["scales"]
        weights[base + ".zeros"] = L["zeros"]
        if "bias" in L:
            weights[base + ".bias"] = L["bias"]
    for i in range(dims.layers):
        for nm in ("input_layernorm", "post_attention_layernorm"):
            weights[f"model.layers.{i}.{nm}.weight"] = (1.0 + 0.05 * torch.randn((h,), generator=gen)).to(torch.bfloat16)
    config = {
        "model_type": dims.model_type, "hidden_size": h, "num_hidden_layers": dims.layers,
        "intermediate_size": dims.inter, "num_attention_heads": dims.heads, "num_key_value_heads": dims.kv_heads,
        "rms_norm_eps": dims.rms_eps, "vocab_size": dims.vocab, "rope_theta": dims.rope_theta,
        "rope_scaling": dims.rope_scaling, "tie_word_embeddings": dims.tie, "max_position_embeddings": dims.max_pos,
        "head_dim": dims.head_dim, "quantization": {"group_size": default_gs, "bits": default_bits},
    }
    if dims.model_type == "qwen2":
        config.pop("head_dim")
    save_weights(path, weights)
    save_config(config, path / "config.json")
    if strategy is not None:
        with open(path / "quant_strategy.json", "w") as f:
            json.dump(strategy, f)
    return config


# ------------------------------------------------------------------------------------------ decode loop
def fast_argmax(x: torch.Tensor) -> torch.Tensor:
    """argmax over the last (vocabulary) dimension, first occurrence on ties like torch.argmax, as two row-parallel
    reductions: a [B, V] single-row argmax is one slow block on the GPU (~0.5 ms at V = 128256, profiles/README.md), a
    [B*a, V/a] max followed by an argmax over the a segment maxima is not."""
    v = x.shape[-1]
    a = next((c for c in (256, 192, 128, 96, 64, 48, 32) if v % c == 0 and v // c >= 64), 0)
    if not a:
        return torch.argmax(x, dim=-1)
    seg = x.reshape(*x.shape[:-1], a, v // a)
    m, i = seg.max(dim=-1)              # first maximal index inside every segment
    j = m.argmax(dim=-1, keepdim=True)  # first segment holding the global maximum
    return (j * (v // a) + i.gather(-1, j)).squeeze(-1)



class DecodeGraph:
    """One decode step (token ids [B,1] -> logits) captured in a CUDA graph: the whole step --
    7*L QuantizedLinear launches plus the glue -- replays as one submission."""

    def __init__(self, model, cache, batch: int = 1):
        self.model, self.cache = model, cache
        dev = cache[0].keys.device
        self.tok = torch.zeros((batch, 1), dtype=torch.long, device=dev)
        self.pos = torch.zeros((1,), dtype=torch.long, device=dev)
        self.logits = None
        self.graph = None

    def capture(self, tokens: torch.Tensor, position: int, device_loop: int = 0):
        """Warm-up + capture at (tokens, position) = the step that will be replayed first, so the KV
        entries written during warm-up/capture are exactly the ones that step writes anyway.
        device_loop = n > 0: the greedy sampler is part of the graph -- argmax of the logits (= argmax of the
        log-probabilities, utils.py:285,305-306), fed back as the next step's token, position advanced and the token
        appended to `self.out` [n] on the device -- so that n steps replay back to back without a host round trip."""
        self.tok.copy_(tokens.reshape(self.tok.shape))
        self.pos.fill_(position)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(2):
                self.model(self.tok, self.cache, positions=self.pos, attend_len=None)
        torch.cuda.current_stream().wait_stream(s)
        if device_loop:
            self.out = torch.zeros((device_loop, self.tok.shape[0]), dtype=torch.long, device=self.tok.device)
            self.idx = torch.zeros((1,), dtype=torch.long, device=self.tok.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.logits = self.model(self.tok, self.cache, positions=self.pos, attend_len=None)
            if device_loop:
                nxt = fast_argmax(self.logits[:, -1, :])  # [B]; bf16 -> fp32 is order preserving, so no cast
                self.out.index_copy_(0, self.idx, nxt[None])
                self.tok.copy_(nxt[:, None])
                self.pos.add_(1)
                self.idx.add_(1)
        return self

    def step(self, tokens: torch.Tensor, position: int) -> torch.Tensor:
        self.tok.copy_(tokens.reshape(self.tok.shape))
        self.pos.fill_(position)
        self.graph.replay()
        for c in self.cache:
            c.offset = position + 1
        return self.logits[:, -1, :]


def generate_step(
    prompt: torch.Tensor,
    model: nn.Module,
    *,
    max_tokens: int = 256,
    sampler: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
    prompt_cache: Optional[Any] = None,
    prefill_step_size: int = 2048,
    max_kv_size: Optional[int] = None,
    use_cuda_graph: bool = True,
) -> Generator[Tuple[int, torch.Tensor], None, None]:
    """Mirror of gbx_lm/utils.py:217-338 for the plain-KV, greedy path: batch-1 prompt of token ids,
    prefill in `prefill_step_size` chunks, then one token per step; yields (token, logprobs).
    The default sampler is the argmax of the log-probabilities (utils.py:285,305-306)."""
    y = prompt.reshape(-1).to(next(model.parameters()).device)
    sampler = sampler or (lambda lp: torch.argmax(lp, dim=-1))
    max_len = max_kv_size or (y.numel() + max_tokens + 1)
    cache = prompt_cache if prompt_cache is not None else qllama.make_prompt_cache(model, 1, max_len)

    def _logprobs(logits):
        logits = logits.float()
        return logits - torch.logsumexp(logits, dim=-1, keepdim=True)

    with torch.no_grad():
        while y.numel() > prefill_step_size:
            model(y[:prefill_step_size][None], cache)
            y = y[prefill_step_size:]
        logits = model(y[None], cache)[:, -1, :]
        lp = _logprobs(logits)
        tok = sampler(lp)
        dg = DecodeGraph(model, cache).capture(tok, cache[0].offset) if (use_cuda_graph and tok.is_cuda) else None
        n = 0
        while True:
            if n == max_tokens:
                break
            yield int(tok.item()), lp.squeeze(0)
            n += 1
            pos = cache[0].offset
            if dg is not None:
                logits = dg.step(tok, pos)
            else:
                logits = model(tok.reshape(1, 1), cache)[:, -1, :]
            lp = _logprobs(logits)
            tok = sampler(lp)


def generate_tokens_device(model, prompt: torch.Tensor, max_tokens: int = 128, prefill_step_size: int = 2048) -> Tuple[List[int], dict]:
    """Greedy generation with the sampler inside the captured decode step (DecodeGraph device_loop): the same tokens as
    `generate_tokens` (argmax of the log-probabilities = argmax of the logits), but the host only enqueues replays and
    reads the tokens once at the end -- the per-token host round trip of utils.py:299-337 (~0.3-0.6 ms on a B200 box)
    is gone.  Same timing definitions as generate_tokens."""
    y = prompt.reshape(-1).to(next(model.parameters()).device)
    n_prompt = y.numel()
    cache = qllama.make_prompt_cache(model, 1, n_prompt + max_tokens + 1)
    t0 = time.perf_counter()
    with torch.no_grad():
        while y.numel() > prefill_step_size:
            model(y[:prefill_step_size][None], cache)
            y = y[prefill_step_size:]
        tok = torch.argmax(model(y[None], cache)[:, -1, :].float(), dim=-1)
    pos0 = cache[0].offset
    dg = DecodeGraph(model, cache).capture(tok, pos0, device_loop=max(max_tokens - 1, 1))
    first = int(tok.item())
    t_first = time.perf_counter()
    for _ in range(max_tokens - 1):
        dg.graph.replay()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    toks = [first] + [int(t) for t in dg.out[: max_tokens - 1, 0].tolist()]
    for c in cache:
        c.offset = pos0 + max_tokens - 1
    stats = {
        "prompt_tokens": int(n_prompt), "prompt_tps": n_prompt / max(t_first - t0, 1e-9),
        "generation_tokens": len(toks), "generation_tps": max(len(toks) - 1, 1) / max(t1 - t_first, 1e-9),
    }
    return toks, stats


def generate_tokens(model, prompt: torch.Tensor, max_tokens: int = 128, **kw) -> Tuple[List[int], dict]:
    """Greedy generation + the reference's timing definitions (utils.py:590-607):
    prompt_tps = prompt_len / time_to_first_token, generation_tps = n / (t_now - t_first)."""
    toks = []
    t0 = time.perf_counter()
    t_first = None
    for tok, _ in generate_step(prompt, model, max_tokens=max_tokens, **kw):
        if t_first is None:
            t_first = time.perf_counter()
        toks.append(tok)
    t1 = time.perf_counter()
    stats = {
        "prompt_tokens": int(prompt.numel()), "prompt_tps": prompt.numel() / max(t_first - t0, 1e-9),
        "generation_tokens": len(toks), "generation_tps": max(len(toks) - 1, 1) / max(t1 - t_first, 1e-9),
    }
    return toks, stats
