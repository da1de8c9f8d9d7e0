#!/usr/bin/env python
"""In-kernel timeline of the TMEM-operand GEMM (gbxq_gemm_ts_sm100.cu), CTA 0: when each stage's packed words landed,
when its A stage was stored, when its MMAs were issued and when its second x tile landed (ns from the CTA's entry).
    python tools/ts_timeline.py N K bits gs M"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbx_lm_b200 import _lib, ops  # noqa: E402


def main():
    n, k, bits, gs, m = (int(v) for v in sys.argv[1:6])
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(0)
    w = torch.randint(-(2 ** 31), 2 ** 31 - 1, (n, k * bits // 32), generator=gen, device=dev, dtype=torch.int64).to(torch.int32).view(torch.uint32)
    s = (torch.rand((n, k // gs), generator=gen, device=dev) * 0.01).to(torch.bfloat16)
    x = torch.randn((m, k), generator=gen, device=dev).to(torch.bfloat16)
    for _ in range(3):
        ops._qmm_impl(x, w, s, s, None, gs, bits, 7)
    # evict the weights from L2
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    junk.fill_(1)
    buf = torch.zeros(8 + 4 * 64, dtype=torch.int64, device=dev)
    lib = _lib.get()
    lib.gbxq_debug_ts_timeline.argtypes = [ctypes.c_void_p]
    lib.gbxq_debug_ts_timeline(buf.data_ptr())
    ops._qmm_impl(x, w, s, s, None, gs, bits, 7)
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    t0 = t[0]
    print(f"N={n} K={k} bits={bits} gs={gs} M={m}: setup {t[1]-t0} ns, epilogue starts {t[2]-t0}, exit {t[3]-t0}")
    print("stage  W-landed  A-stored  MMA-issued  x-landed   (dW = W-landed - previous)")
    prev = None
    for st in range(64):
        a = t[8 + 4 * st: 12 + 4 * st]
        if a[0] == 0:
            break
        d = "" if prev is None else f"{a[0]-prev:6d}"
        print(f"{st:5d} {a[0]-t0:9d} {a[1]-t0:9d} {a[2]-t0:11d} {a[3]-t0:9d}   {d}")
        prev = a[0]


if __name__ == "__main__":
    main()
