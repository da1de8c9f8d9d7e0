#!/usr/bin/env python
"""bf16 vocabulary projection at decode sizes: gbxq_head_gemv against the framework's dense matmul (cuBLAS), GB/s and
fraction of the measured HBM peak.   python tools/headbench.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbx_lm_b200 import ops  # noqa: E402

pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = torch.device("cuda:0")
print(f"{'V':>7s} {'K':>5s} {'M':>2s} {'gbxq us':>9s} {'GB/s':>8s} {'frac':>6s} {'cuBLAS us':>10s} {'GB/s':>8s}")
for (v, k) in ((128256, 2048), (128256, 3072), (128256, 4096), (152064, 5120), (128256, 8192)):
    w = (torch.randn((v, k), device=dev) / k ** 0.5).to(torch.bfloat16)
    for m in (1, 4, 8):
        x = torch.randn((m, k), device=dev).to(torch.bfloat16)
        res = []
        for fn in (lambda: ops.head_linear(x, w), lambda: torch.nn.functional.linear(x, w)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(10):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) * 100.0)  # us per call
        by = v * k * 2 + m * k * 2 + m * v * 2
        print(f"{v:7d} {k:5d} {m:2d} {res[0]:9.1f} {by / res[0] / 1e3:8.1f} {by / res[0] / 1e3 / pk:6.3f} {res[1]:10.1f} {by / res[1] / 1e3:8.1f}", flush=True)
    del w
