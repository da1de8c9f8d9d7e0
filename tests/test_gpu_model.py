"""GPU tests of the CALLERS of the hot path: checkpoint -> load_model ->
Llama / Qwen2 forward through libgbxq, against an independent fp32 reference on oracle-dequantised
weights; greedy decode (CUDA-graph vs eager, teacher-forced top-1 vs the reference)."""
import json

import numpy as np
import pytest
import torch

from gbx_lm_b200 import QuantizedLinear, utils, workloads as W
from tests import ref_model

pytestmark = pytest.mark.gpu

CASES = [
    ("tiny-llama", "bpw-4.0", 4, 64),
    ("tiny-llama", "bpw-2.2", 2, 128),
    ("tiny-qwen2", None, 4, 128),
]


def _make(tmp_path, name, strat_name, bits, gs, seed=7):
    dims = W.MODELS[name]
    strat = W.STRATEGIES[strat_name](dims.layers) if strat_name else None
    cfg = utils.write_synthetic_checkpoint(tmp_path, dims, strat, seed=seed, default_bits=bits, default_gs=gs, embed_scale=1.0)
    from safetensors.torch import load_file

    ckpt = load_file(str(tmp_path / "model.safetensors"))
    plan = {}
    for (i, p, n, k, b, g) in W.layer_plan(dims, strat, bits, gs):
        sub = "self_attn" if p in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
        plan[f"model.layers.{i}.{sub}.{p}"] = (b, g)
    dense = ref_model.dense_weights(ckpt, lambda m: plan[m])
    return dims, cfg, dense


@pytest.mark.parametrize("name,strat,bits,gs", CASES)
def test_forward_matches_reference(cuda_device, tmp_path, name, strat, bits, gs):
    dims, cfg, dense = _make(tmp_path, name, strat, bits, gs)
    model, _ = utils.load_model(tmp_path, device=cuda_device)
    assert all(m.qweight.is_cuda for m in model.modules() if isinstance(m, QuantizedLinear))
    toks = torch.randint(0, dims.vocab, (2, 19), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        logits = model(toks.to(cuda_device)).float().cpu()
        ref = ref_model.forward(dense, cfg, toks)
    assert logits.shape == ref.shape == (2, 19, dims.vocab)
    err = (logits - ref).abs().max() / ref.abs().max()
    assert err < 3e-2, err
    # top-1 agrees wherever the reference's margin is not a rounding tie
    top2 = ref.topk(2, -1).values
    clear = (top2[..., 0] - top2[..., 1]) > 0.05 * ref.abs().max()
    assert (logits.argmax(-1)[clear] == ref.argmax(-1)[clear]).all()


def test_chunked_prefill_and_cache_consistency(cuda_device, tmp_path):
    dims, cfg, dense = _make(tmp_path, "tiny-llama", "bpw-4.0", 4, 64)
    model, _ = utils.load_model(tmp_path, device=cuda_device)
    from gbx_lm_b200 import qllama

    toks = torch.randint(0, dims.vocab, (1, 24), generator=torch.Generator().manual_seed(2)).to(cuda_device)
    with torch.no_grad():
        full = model(toks).float()
        cache = qllama.make_prompt_cache(model, 1, 64)
        a = model(toks[:, :9], cache).float()
        b = model(toks[:, 9:20], cache).float()
        outs = [a, b]
        for i in range(20, 24):
            outs.append(model(toks[:, i:i + 1], cache).float())
        inc = torch.cat(outs, 1)
    assert cache[0].offset == 24
    assert (inc - full).abs().max() <= 2e-2 * full.abs().max()


@pytest.mark.parametrize("name,strat,bits,gs", CASES[:1] + CASES[2:])
def test_greedy_decode_graph_equals_eager_and_reference(cuda_device, tmp_path, name, strat, bits, gs):
    dims, cfg, dense = _make(tmp_path, name, strat, bits, gs, seed=11)
    model, _ = utils.load_model(tmp_path, device=cuda_device)
    prompt = torch.randint(0, dims.vocab, (12,), generator=torch.Generator().manual_seed(3))
    g_toks, stats = utils.generate_tokens(model, prompt, max_tokens=48, use_cuda_graph=True)
    e_toks, _ = utils.generate_tokens(model, prompt, max_tokens=48, use_cuda_graph=False)
    assert len(g_toks) == 48 and stats["generation_tps"] > 0
    assert g_toks == e_toks  # the graph replays exactly the eager computation
    # the sampler inside the captured step (no host round trip per token) yields the same tokens
    d_toks, d_stats = utils.generate_tokens_device(model, prompt, max_tokens=48)
    assert d_toks == g_toks and d_stats["generation_tps"] > 0
    # chunked prefill (utils.py:312-319 semantics) does not change the tokens
    c_toks, _ = utils.generate_tokens(model, prompt, max_tokens=16, prefill_step_size=5)
    assert c_toks == g_toks[:16]
    # teacher-forced check against the fp32 reference: our token is the reference's top-1 wherever the
    # reference's top-1 margin exceeds bf16 noise; report how many steps had a clear margin
    seq = torch.cat([prompt, torch.tensor(g_toks)])[None]
    with torch.no_grad():
        ref = ref_model.forward(dense, cfg, seq)[0, len(prompt) - 1:-1]
    top2 = ref.topk(2, -1).values
    clear = (top2[:, 0] - top2[:, 1]) > 0.05 * ref.abs().max()
    assert clear.sum() >= 24, f"only {int(clear.sum())} decisive steps"
    assert (ref.argmax(-1)[clear] == torch.tensor(g_toks)[clear]).all()


def test_fast_argmax_is_torch_argmax(cuda_device):
    """Two-stage vocabulary argmax of the device-side sampler: first occurrence on ties, any vocabulary size."""
    g = torch.Generator(device=cuda_device).manual_seed(0)
    for v in (128256, 152064, 512, 1000, 32000):
        x = torch.randn((3, v), generator=g, device=cuda_device).to(torch.bfloat16)  # bf16: ties for the maximum do occur
        x[1, 7] = x[1, v - 5] = x[1].max() + 1  # forced tie far apart
        x[2, v - 1] = 100.0
        assert torch.equal(utils.fast_argmax(x), torch.argmax(x, dim=-1)), v
        assert torch.equal(utils.fast_argmax(x.float()), torch.argmax(x.float(), dim=-1)), v


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[0]: Llama-3.2-1B architecture, random-init, layer-mix bpw-4.0, batch-1 greedy generate of 128
# tokens (the reference's loop: gbx_lm/utils.py:217-338, argmax of the log-probabilities :285-306).  north_star asks for
# identical greedy tokens over 128 steps against the reference's CPU path; MLX is not installable (oracle/MLX_SPEC.md), so
# the comparison model is the fp32 decoder of tests/ref_model.py on ORACLE-dequantised weights, teacher-forced on the
# tokens the CUDA model produced.  A step is "decisive" when the reference's top-1 margin exceeds
# twice the largest logit difference observed between the two models: there the tokens MUST agree.
# The 128 tokens and the per-step margins are committed as tests/golden/greedy128_llama32_1b.json.
# ---------------------------------------------------------------------------------------------------------------
GREEDY_FIXTURE = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "greedy128_llama32_1b.json")


def test_greedy_128_tokens_llama32_1b_architecture(cuda_device, tmp_path):
    import os

    from safetensors.torch import load_file

    dims = W.MODELS["llama-3.2-1b"]
    strat = W.STRATEGIES["bpw-4.0"](dims.layers)
    # embed_scale 1/sqrt(h) on the tied head keeps logits O(1) so that margins are comparable with bf16 resolution
    cfg = utils.write_synthetic_checkpoint(tmp_path, dims, strat, seed=20261018, default_bits=4, default_gs=64, embed_scale=1.0)
    model, _ = utils.load_model(tmp_path, device=cuda_device)
    prompt = torch.randint(0, dims.vocab, (16,), generator=torch.Generator().manual_seed(5))
    toks, stats = utils.generate_tokens_device(model, prompt, max_tokens=128)
    toks_host, _ = utils.generate_tokens(model, prompt, max_tokens=128)
    assert len(toks) == 128 and toks == toks_host  # device-side and host-side sampler: the same 128 tokens
    ckpt = {}
    for f in sorted(os.listdir(tmp_path)):
        if f.endswith(".safetensors"):
            ckpt.update(load_file(str(tmp_path / f)))
    plan = {}
    for (i, p, n, k, b, g) in W.layer_plan(dims, strat, 4, 64):
        sub = "self_attn" if p in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
        plan[f"model.layers.{i}.{sub}.{p}"] = (b, g)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        dense = ref_model.dense_weights_c(ckpt, lambda m: plan[m], device=cuda_device)
        seq = torch.cat([prompt, torch.tensor(toks)])[None].to(cuda_device)
        with torch.no_grad():
            ref = ref_model.forward(dense, cfg, seq)[0, len(prompt) - 1:-1].float()   # logits that chose token t, fp32
            ours = model(seq).float()[0, len(prompt) - 1:-1]                          # the CUDA model, teacher-forced
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    noise_t = (ours - ref).abs().max(-1).values.cpu()  # per step: largest logit difference between the two models
    noise = float(noise_t.max())
    top2 = ref.topk(2, -1).values
    margin = (top2[:, 0] - top2[:, 1]).cpu()
    decisive = margin > 2.0 * noise_t
    ref_top1 = ref.argmax(-1).cpu()
    got = torch.tensor(toks)
    match = ref_top1 == got
    report = {"tokens": toks, "reference_top1": ref_top1.tolist(), "margin": [round(float(v), 5) for v in margin],
              "logit_noise_max_abs": noise, "logit_scale_max_abs": float(ref.abs().max()), "decisive_steps": int(decisive.sum()),
              "all_128_match": bool(match.all()), "matching_steps": int(match.sum()), "generation_tps": stats["generation_tps"]}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(report, open(os.path.join("gpurun_out", "greedy128_llama32_1b.json"), "w"))
    print("greedy-128:", {k: v for k, v in report.items() if k not in ("tokens", "reference_top1", "margin")})
    assert noise <= 3e-2 * float(ref.abs().max()), noise
    assert bool(match[decisive].all()), "a decisive step disagrees with the fp32 reference"
    assert int(decisive.sum()) >= 64, report["decisive_steps"]
    assert int(match.sum()) >= 110, report["matching_steps"]  # the other steps are within rounding of a tie; mismatches must be rare
    if os.path.exists(GREEDY_FIXTURE):
        fx = json.load(open(GREEDY_FIXTURE))
        fx_dec = torch.tensor(fx["margin"]) > 2.0 * max(fx["logit_noise_max_abs"], noise)  # conservative: the global noise
        same = torch.tensor(fx["tokens"]) == got
        # the committed run: tokens are reproduced at least up to the first non-decisive step
        first_tie = int((~fx_dec).nonzero()[0]) if (~fx_dec).any() else 128
        assert bool(same[:first_tie].all()), (first_tie, fx["tokens"][:first_tie], toks[:first_tie])
