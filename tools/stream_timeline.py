#!/usr/bin/env python
"""Where the time of a chain launch (gbxq_qmm_stream) goes: per-call %globaltimer stamps of CTA 0 over the decode step of
bench.py (llama-3-8b bpw-4.0, q|k|v and gate|up grouped).  Development aid; run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gbx_lm_b200 import QuantizedLinear, ops  # noqa: E402
from gbx_lm_b200 import workloads as W  # noqa: E402


def main():
    layers_n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda", 0)
    dims = W.MODELS["llama-3-8b"]
    plan = [e for e in W.layer_plan(dims, W.STRATEGIES["bpw-4.0"](dims.layers), 4, 64) if e[0] < layers_n]
    gen = torch.Generator(device=dev).manual_seed(1)
    mods = []
    for (i, p, n, k, b, g) in plan:
        m = QuantizedLinear(k, n, bias=False, group_size=g, bits=b)
        m._set("qweight", torch.randint(-(2 ** 31), 2 ** 31 - 1, (n, k * b // 32), generator=gen, device=dev, dtype=torch.int64).to(torch.int32).view(torch.uint32))
        s = ((torch.rand((n, k // g), generator=gen, device=dev) + 0.5) * (2.0 / (k ** 0.5) / ((1 << b) - 1))).to(torch.bfloat16)
        m._set("scales", s)
        m._set("zeros", (-s.float() * ((1 << b) - 1) / 2).to(torch.bfloat16))
        m._set("channel_scale", None)
        mods.append((p, m))
    xbuf = {k: torch.randn((1, k), generator=gen, device=dev).to(torch.bfloat16) for k in {e[3] for e in plan}}
    chain = ops.StreamChain(1)
    names, nbytes, i = [], [], 0
    while i < len(mods):
        p = mods[i][0]
        take = 3 if p == "q_proj" else (2 if p == "gate_proj" else 1)
        ms = [m for _, m in mods[i:i + take]]
        chain.add(xbuf[ms[0].input_dims], ms)
        names.append({"q_proj": "qkv", "gate_proj": "gate_up"}.get(p, p))
        nbytes.append(sum(W.qmm_bytes(1, m.output_dims, m.input_dims, m.bits, m.group_size) for m in ms))
        i += take
    chain.finalize(debug_timeline=True)
    for _ in range(5):
        chain.run()
    torch.cuda.synchronize()
    t = chain.timeline.cpu().numpy().astype("int64")
    t0 = t[0, 0]
    print(f"{'call':<8}{'MB':>7}{'hbm_us':>8} | {'enter':>8}{'ordered':>8}{'x_here':>8}{'frags':>8}{'stage0':>8}{'loopend':>8}{'stored':>8} | "
          f"{'wait':>6}{'xload':>6}{'build':>6}{'loop':>6}{'epi':>6}{'total':>7}  (ns)")
    agg = {}
    for c in range(len(names)):
        r = t[c]
        enter, ordered, xh, fr, st0, le, stored = (r[0] - t0, r[5] - t0, r[7] - t0, r[2] - t0, r[3] - t0, r[4] - t0, r[6] - t0)
        nxt = (t[c + 1, 0] - t0) if c + 1 < len(names) else stored
        d = (ordered - enter, xh - ordered, fr - xh, le - fr, stored - le, nxt - enter)
        agg.setdefault(names[c], []).append(d)
        if c < 12:
            print(f"{names[c]:<8}{nbytes[c] / 1e6:>7.1f}{nbytes[c] / 6.5e3:>8.0f} | {enter:>8}{ordered:>8}{xh:>8}{fr:>8}{st0:>8}{le:>8}{stored:>8} | "
                  + "".join(f"{v:>6}" for v in d[:5]) + f"{d[5]:>7}")
    print("-- mean per call type over the chain (ns): wait xload build loop epi total")
    for k, v in agg.items():
        n = len(v)
        print(f"{k:<8}" + "".join(f"{sum(x[j] for x in v) / n:>8.0f}" for j in range(6)))
    print("chain total us:", (t[-1, 6] - t0) / 1e3, "calls", len(names), "timed_out", chain.timed_out())


if __name__ == "__main__":
    main()
