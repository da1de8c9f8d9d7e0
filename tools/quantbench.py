#!/usr/bin/env python
"""Device-side mx.quantize (gbxq_quantize / gbxq_quantize_rows): microseconds per call, algorithmic GB/s
(sizeof(T) read + bits/8 written per element + 2 * sizeof(T) per group) and fraction of the measured HBM peak, for the
shapes the quantized KV cache produces (gbx_lm/models/cache.py:176-177,251-263).   python tools/quantbench.py [case substring]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbx_lm_b200 import ops  # noqa: E402

p = os.path.join(ROOT, "MEASURED_PEAKS.json")
pk = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0
dev = torch.device("cuda:0")
print(f"{'case':44s} {'bits':>4s} {'gs':>3s} {'us/call':>9s} {'GB/s':>8s} {'frac':>6s}")
CASES = (
    ("decode step: 8 kv heads x 1 x 128 (into cache)", (1, 8, 1, 128), True),
    ("decode batch 64: 64 x 8 x 1 x 128 (into cache)", (64, 8, 1, 128), True),
    ("prefill chunk: 8 heads x 2048 x 128 (into cache)", (1, 8, 2048, 128), True),
    ("to_quantized: 8 heads x 32768 x 128", (1, 8, 32768, 128), False),
    ("weight matrix 4096 x 14336", (4096, 14336), False),
)
ONLY = sys.argv[1] if len(sys.argv) > 1 else ""  # substring filter on the case name
for name, shape, into in CASES:
    if ONLY not in name:
        continue
    for bits, gs in ((8, 64), (4, 64), (3, 64)):
        w = torch.randn(shape, device=dev).to(torch.bfloat16)
        if into:
            cap = max(256, shape[-2])
            k = shape[-1]
            out = (torch.zeros((*shape[:-2], cap, k * bits // 32), dtype=torch.int32, device=dev).view(torch.uint32),
                   torch.zeros((*shape[:-2], cap, k // gs), dtype=torch.bfloat16, device=dev),
                   torch.zeros((*shape[:-2], cap, k // gs), dtype=torch.bfloat16, device=dev))
            fn = lambda: ops.quantize_into(w, out, 0, gs, bits)  # noqa: E731
        else:
            fn = lambda: ops.quantize(w, gs, bits)  # noqa: E731
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        reps = 20
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        n = w.numel()
        by = n * 2 + n * bits // 8 + 2 * 2 * n // gs
        print(f"{name:44s} {bits:4d} {gs:3d} {us:9.2f} {by / us / 1e3:8.1f} {by / us / 1e3 / pk:6.3f}", flush=True)
    del w
