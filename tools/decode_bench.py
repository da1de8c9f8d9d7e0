#!/usr/bin/env python
"""End-to-end decode through the CALLERS of the hot path (BASELINE.json configs[0]): synthetic
layer-mix checkpoint in the gba2mlx layout -> load_model (quant_strategy.json, safetensors) -> greedy generate of N
tokens at batch 1 (CUDA-graph decode step: 7 L QuantizedLinear forwards through libgbxq + torch glue + bf16 head),
timed with the reference's own definitions (gbx_lm/utils.py:590-607).  Prints one JSON line.

    python tools/decode_bench.py [--model llama-3.2-1b] [--strategy bpw-4.0] [--tokens 128] [--prompt 32]"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gbx_lm_b200 import ops, utils  # noqa: E402
from gbx_lm_b200 import workloads as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3.2-1b")
    ap.add_argument("--strategy", default="bpw-4.0", choices=["bpw-4.0", "bpw-2.2", "uniform"])
    ap.add_argument("--tokens", type=int, default=128)
    ap.add_argument("--prompt", type=int, default=32)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    dims = W.MODELS[args.model]
    strat = None if args.strategy == "uniform" else W.STRATEGIES[args.strategy](dims.layers)
    d = Path(tempfile.mkdtemp(prefix="gbxq_ckpt_"))
    try:
        t0 = time.perf_counter()
        utils.write_synthetic_checkpoint(d, dims, strat, seed=0, default_bits=4, default_gs=64, embed_scale=1.0)
        t_write = time.perf_counter() - t0
        t0 = time.perf_counter()
        model, cfg = utils.load_model(d, device="cuda")
        torch.cuda.synchronize()
        t_load = time.perf_counter() - t0
    finally:
        shutil.rmtree(d, ignore_errors=True)
    prompt = torch.randint(0, dims.vocab, (args.prompt,), generator=torch.Generator().manual_seed(1))
    best, toks0 = None, None
    for r in range(args.repeat):
        n0 = ops.launch_count()
        toks, stats = utils.generate_tokens(model, prompt, max_tokens=args.tokens, use_cuda_graph=True)
        stats["gbxq_launches"] = ops.launch_count() - n0
        if toks0 is None:
            toks0 = toks
        assert toks == toks0, "greedy decode is not reproducible run to run"
        if best is None or stats["generation_tps"] > best["generation_tps"]:
            best = stats
    d_toks, d_stats = utils.generate_tokens_device(model, prompt, max_tokens=args.tokens)
    d_toks, d_stats = utils.generate_tokens_device(model, prompt, max_tokens=args.tokens)
    assert d_toks == toks0, "device-side greedy sampler diverged from the host-side one"
    # device time of the captured decode step alone (no sampling, no host sync between steps)
    from gbx_lm_b200 import qllama

    cache = qllama.make_prompt_cache(model, 1, args.prompt + args.tokens + 1)
    with torch.no_grad():
        model(prompt[None].cuda(), cache)
    dg = utils.DecodeGraph(model, cache).capture(torch.tensor([1], device="cuda"), args.prompt)
    for _ in range(5):
        dg.graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        dg.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / 50
    plan = W.layer_plan(dims, strat, 4, 64)
    body = sum(W.qmm_bytes(1, n, k, b, g) for (_, _, n, k, b, g) in plan)
    head = dims.vocab * dims.hidden * 2
    print(json.dumps({
        "tool": "decode_bench", "model": args.model, "strategy": args.strategy, "batch": 1, "prompt_tokens": args.prompt,
        "generation_tokens": best["generation_tokens"], "generation_tps": round(best["generation_tps"], 1),
        "prompt_tps": round(best["prompt_tps"], 1), "ms_per_token": round(1e3 / best["generation_tps"], 4),
        "generation_tps_device_sampler": round(d_stats["generation_tps"], 1), "graph_ms_per_step": round(graph_ms, 4), "gbxq_launches_per_generate": best["gbxq_launches"],
        "qmm_bytes_per_token": body, "head_bytes_per_token": head,
        "hbm_gbs_body_plus_head": round((body + head) * best["generation_tps"] / 1e9, 1),
        "checkpoint_write_s": round(t_write, 1), "load_s": round(t_load, 1), "data": "synthetic (random-init, seed 0)",
        "note": "whole decode step: quantized forwards + decode glue through libgbxq (GBXQ_FUSED_DECODE=0: torch glue), bf16 head via cuBLAS; generation_tps syncs with the host every token, generation_tps_device_sampler does not",
    }), flush=True)


if __name__ == "__main__":
    main()
