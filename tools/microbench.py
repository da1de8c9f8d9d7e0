#!/usr/bin/env python
"""Per-shape roofline sweep of the quantized matmul kernels: times each
(N, K, bits, group_size, M) over a rotation of distinct weight copies larger than 2x L2 (so every
launch streams from HBM), CUDA events on the launching stream, prints GB/s and fraction of the
measured HBM peak.   python tools/microbench.py [--quick] [--kernel auto|gemv|gemm|generic] [--json out]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbx_lm_b200 import ops, workloads as W  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def bench_shape(n, k, bits, gs, m, kernel, dev, iters=30, min_bytes=320 << 20, l2=False):
    wbytes = n * k * bits // 8
    copies = 1 if l2 else max(2, min(64, (min_bytes + wbytes - 1) // wbytes))
    gen = torch.Generator(device=dev).manual_seed(0)
    nb = (1 << bits) - 1
    ws, ss, zs = [], [], []
    for _ in range(copies):
        ws.append(torch.randint(-(2 ** 31), 2 ** 31 - 1, (n, k * bits // 32), generator=gen, device=dev, dtype=torch.int64).to(torch.int32).view(torch.uint32))
        s = ((torch.rand((n, k // gs), generator=gen, device=dev) + 0.5) * (2.0 / (k ** 0.5) / nb)).to(torch.bfloat16)
        ss.append(s)
        zs.append((-s.float() * (nb / 2.0)).to(torch.bfloat16))
    x = torch.randn((m, k), generator=gen, device=dev).to(torch.bfloat16)
    kid = {"auto": 0, "generic": 1, "gemv": 2, "gemm": 3, "skinny": 4, "mmv": 5, "mmv8": 6, "gemm_ts": 7}[kernel]

    def run(i):
        j = i % copies
        return ops._qmm_impl(x, ws[j], ss[j], zs[j], None, gs, bits, kid)

    for i in range(copies + 3):
        run(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            run(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    by = W.qmm_bytes(m, n, k, bits, gs)
    return us, by / us / 1e3  # GB/s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--json", default=None)
    ap.add_argument("--ms", default="1")
    ap.add_argument("--shapes", default="8b")
    ap.add_argument("--l2", action="store_true", help="one weight copy: launches after the first stream from L2 (compute/L2 ceiling, not a roofline number)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    pk = peak()
    shapes8b = [("q/o", 4096, 4096), ("k/v", 1024, 4096), ("gate/up", 14336, 4096), ("down", 4096, 14336)]
    shapes70b = [("q/o", 8192, 8192), ("k/v", 1024, 8192), ("gate/up", 28672, 8192), ("down", 8192, 28672)]
    shapes = shapes70b if args.shapes == "70b" else shapes8b
    if args.shapes == "big":  # asymptotic streaming rate: ~8x gate_proj rows
        shapes = [("big", 114688, 4096), ("bigk", 32768, 14336)]
    combos = [(4, 64), (2, 64)] if args.quick else [(4, 64), (4, 128), (2, 64), (2, 128), (3, 64), (6, 64), (8, 64), (4, 32), (2, 32)]
    ms = [int(v) for v in args.ms.split(",")]
    rows = []
    print(f"{'shape':8s} {'N':>6s} {'K':>6s} bits gs   M {'us':>8s} {'GB/s':>8s} {'frac':>6s}")
    for (name, n, k) in shapes:
        for (bits, gs) in combos:
            for m in ms:
                try:
                    us, gbs = bench_shape(n, k, bits, gs, m, args.kernel, dev, l2=args.l2)
                except Exception as e:  # noqa: BLE001
                    print(name, n, k, bits, gs, m, "ERR", e)
                    continue
                rows.append({"shape": name, "N": n, "K": k, "bits": bits, "gs": gs, "M": m, "us": us, "gbs": gbs, "frac": gbs / pk})
                print(f"{name:8s} {n:6d} {k:6d} {bits:4d} {gs:3d} {m:3d} {us:8.2f} {gbs:8.1f} {gbs / pk:6.3f}", flush=True)
    if args.json:
        json.dump({"peak_gbs": pk, "kernel": args.kernel, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
