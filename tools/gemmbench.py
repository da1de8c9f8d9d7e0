#!/usr/bin/env python
"""Prefill / batched GEMM sweep: TFLOP/s (2*M*N*K) and fraction of the measured bf16 tensor peak,
plus algorithmic GB/s for the skinny regime.   python tools/gemmbench.py [--ms 17,32,64,128,512,2048]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbx_lm_b200 import ops, workloads as W  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d["bf16_tflops"])
    return 6650.0, 1590.0


def bench(n, k, bits, gs, m, kernel, dev, iters=10):
    gen = torch.Generator(device=dev).manual_seed(0)
    nb = (1 << bits) - 1
    copies = max(2, min(16, (300 << 20) // (n * k * bits // 8)))
    ws = [torch.randint(-(2 ** 31), 2 ** 31 - 1, (n, k * bits // 32), generator=gen, device=dev, dtype=torch.int64).to(torch.int32).view(torch.uint32) for _ in range(copies)]
    s = ((torch.rand((n, k // gs), generator=gen, device=dev) + 0.5) * (2.0 / (k ** 0.5) / nb)).to(torch.bfloat16)
    z = (-s.float() * (nb / 2.0)).to(torch.bfloat16)
    x = torch.randn((m, k), generator=gen, device=dev).to(torch.bfloat16)
    kid = {"auto": 0, "generic": 1, "gemv": 2, "gemm": 3, "skinny": 4, "mmv": 5, "mmv8": 6, "gemm_ts": 7}[kernel]
    for i in range(3):
        ops._qmm_impl(x, ws[i % copies], s, z, None, gs, bits, kid)
    torch.cuda.synchronize()
    # one CUDA graph of `iters` launches: host-side launch cost (tensor-map encodes, Python) stays out of the timing
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(iters):
            ops._qmm_impl(x, ws[i % copies], s, z, None, gs, bits, kid)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    return us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ms", default="17,32,64,128,256,512,2048")
    ap.add_argument("--kernel", default="gemm")
    ap.add_argument("--bits", default="4")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    hbm, tf = peaks()
    shapes = [("3b q/o", 3072, 3072), ("3b gate", 8192, 3072), ("3b down", 3072, 8192), ("8b gate", 14336, 4096), ("8b down", 4096, 14336)]
    rows = []
    print(f"{'shape':8s} {'N':>6s} {'K':>6s} bits   M {'us':>9s} {'TFLOP/s':>8s} {'fracT':>6s} {'GB/s':>8s} {'fracH':>6s}")
    for bits in [int(b) for b in args.bits.split(",")]:
        for (name, n, k) in shapes:
            for m in [int(v) for v in args.ms.split(",")]:
                us = bench(n, k, bits, 64, m, args.kernel, dev)
                tfl = 2.0 * m * n * k / us / 1e6
                gbs = W.qmm_bytes(m, n, k, bits, 64) / us / 1e3
                rows.append({"shape": name, "N": n, "K": k, "bits": bits, "M": m, "us": us, "tflops": tfl, "gbs": gbs})
                print(f"{name:8s} {n:6d} {k:6d} {bits:4d} {m:4d} {us:9.2f} {tfl:8.1f} {tfl / tf:6.3f} {gbs:8.1f} {gbs / hbm:6.3f}", flush=True)
    if args.json:
        json.dump({"hbm_gbs": hbm, "bf16_tflops": tf, "kernel": args.kernel, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
