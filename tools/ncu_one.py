#!/usr/bin/env python
"""Tiny driver for ncu captures: runs ONE shape of the quantized matmul a few times.
    ncu ... python tools/ncu_one.py N K bits gs M [kernel] [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbx_lm_b200 import ops  # noqa: E402

n, k, bits, gs, m = (int(v) for v in sys.argv[1:6])
kernel = sys.argv[6] if len(sys.argv) > 6 else "auto"
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 6
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
nb = (1 << bits) - 1
copies = 4
ws = [torch.randint(-(2 ** 31), 2 ** 31 - 1, (n, k * bits // 32), generator=gen, device=dev, dtype=torch.int64).to(torch.int32).view(torch.uint32) for _ in range(copies)]
s = ((torch.rand((n, k // gs), generator=gen, device=dev) + 0.5) * (2.0 / (k ** 0.5) / nb)).to(torch.bfloat16)
z = (-s.float() * (nb / 2.0)).to(torch.bfloat16)
x = torch.randn((m, k), generator=gen, device=dev).to(torch.bfloat16)
kid = {"auto": 0, "generic": 1, "gemv": 2, "gemm": 3, "skinny": 4, "mmv": 5, "mmv8": 6, "gemm_ts": 7}[kernel]
for i in range(iters):
    y = ops._qmm_impl(x, ws[i % copies], s, z, None, gs, bits, kid)
torch.cuda.synchronize()
print("done", float(y.float().abs().mean()))
