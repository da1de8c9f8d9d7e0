#!/usr/bin/env python
"""Timeline of a chain of dependent mmv8 launches (CTA 0 globaltimer stamps): where the per-launch fixed cost goes.
    python tools/timeline.py N K bits gs [launches] [pdl]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gbx_lm_b200 import _lib, ops  # noqa: E402

n, k, bits, gs = (int(v) for v in sys.argv[1:5])
L = int(sys.argv[5]) if len(sys.argv) > 5 else 8
pdl = int(sys.argv[6]) if len(sys.argv) > 6 else 2
ops.set_pdl_mode(pdl)
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
nb = (1 << bits) - 1
ws = [torch.randint(-(2 ** 31), 2 ** 31 - 1, (n, k * bits // 32), generator=gen, device=dev, dtype=torch.int64).to(torch.int32).view(torch.uint32) for _ in range(L)]
s = ((torch.rand((n, k // gs), generator=gen, device=dev) + 0.5) * (2.0 / (k ** 0.5) / nb)).to(torch.bfloat16)
z = (-s.float() * (nb / 2.0)).to(torch.bfloat16)
x = torch.randn((1, k), generator=gen, device=dev).to(torch.bfloat16)
lib = _lib.get()
lib.gbxq_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.gbxq_debug_timeline.restype = None
buf = torch.zeros((L, 8), dtype=torch.int64, device=dev)


def chain():
    for i in range(L):
        ops._qmm_impl(x, ws[i], s, z, None, gs, bits, 6)


chain()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
lib.gbxq_debug_timeline(buf.data_ptr(), L)
with torch.cuda.graph(g):
    chain()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
t = buf.cpu().numpy()
t0 = t[0, 0]
names = ["entry", "waited", "prologue", "stage0", "loopend", "ctabar", "exit"]
print(f"N={n} K={k} bits={bits} gs={gs} pdl={pdl}: ns relative to launch 0 entry; per launch: " + " ".join(names))
for i in range(L):
    print(i, " ".join(f"{int(v - t0):7d}" for v in t[i, :7]), "  | dur", int(t[i, 6] - t[i, 0]), " wait->x", int(t[i, 7] - t[i, 1]),
          " exit->next waited", int(t[i + 1, 1] - t[i, 6]) if i + 1 < L else "")

# ---- all-CTA timeline: where the hand-over between dependent launches goes (entry / waited / loop end / exit of EVERY
# CTA of every launch; per launch the spread of each stamp over the grid and the gap to the next launch)
if os.environ.get("TIMELINE_ALL", "1") != "0":
    import numpy as np

    lib.gbxq_debug_timeline_all.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.gbxq_debug_timeline_all.restype = None
    STRIDE = 2048
    big = torch.zeros((L, STRIDE, 8), dtype=torch.int64, device=dev)
    g2 = torch.cuda.CUDAGraph()
    lib.gbxq_debug_timeline_all(big.data_ptr(), L, STRIDE)
    with torch.cuda.graph(g2):
        chain()
    lib.gbxq_debug_timeline_all(0, 0, 0)
    for _ in range(3):
        g2.replay()
    torch.cuda.synchronize()
    a = big.cpu().numpy()
    live = a[:, :, 0] != 0
    base = a[0, :, 0][live[0]].min()
    print("all CTAs: per launch, ns relative to the first entry of launch 0: min / median / max over the grid")
    cols = {"entry": 0, "waited": 1, "prologue": 2, "stage0": 3, "loopend": 4, "exit": 6}
    for i in range(L):
        m = live[i]
        n_cta = int(m.sum())
        row = [f"L{i} ctas={n_cta}"]
        for nm, c in cols.items():
            v = a[i, m, c] - base
            row.append(f"{nm} {int(v.min())}/{int(np.median(v))}/{int(v.max())}")
        print("  ".join(row))
        if i + 1 < L and live[i + 1].any():
            nxt = a[i + 1, live[i + 1], 1] - base
            ex = a[i, m, 6] - base
            print(f"     last exit -> next waited(min) {int(nxt.min() - ex.max())} ns;  exit spread {int(ex.max() - ex.min())} ns;"
                  f"  loop time min/med/max {int((a[i, m, 4] - a[i, m, 3]).min())}/{int(np.median(a[i, m, 4] - a[i, m, 3]))}/{int((a[i, m, 4] - a[i, m, 3]).max())}")
    # exit time versus SM for one middle launch: are the stragglers whole SMs?
    i = min(3, L - 1)
    m = live[i]
    sm = a[i, m, 5]
    ex = a[i, m, 6] - base
    en = a[i, m, 0] - base
    order = np.argsort(ex)
    print(f"launch {i}: 10 earliest exits (exit, entry, smid):", [(int(ex[j]), int(en[j]), int(sm[j])) for j in order[:10]])
    print(f"launch {i}: 10 latest exits   (exit, entry, smid):", [(int(ex[j]), int(en[j]), int(sm[j])) for j in order[-10:]])
