import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import gbx_lm_b200 as g
from oracle import mlx_affine as A
from tests.gpu_util import bf16_from_bits, layer_to_cuda
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
for (M, N, K, bits, gs) in ((256, 1024, 3072, 8, 64), (256, 1024, 3072, 4, 64)):
    L = A.synth_layer(N, K, bits, gs, seed=M + N, with_bias=True)
    d = layer_to_cuda(L, dev)
    xb = A.synth_x(M, K, seed=K)
    x = bf16_from_bits(xb, dev)
    ref = A.quantized_matmul(xb, L["qweight"], L["scales"], L["zeros"], gs, bits, "bf16", "f64")   # no bias
    y = g.quantized_matmul(x, d["qweight"], d["scales"], d["zeros"], True, gs, bits, kernel="gemm").float().cpu().numpy()
    ysk = g.quantized_matmul(x[:8], d["qweight"], d["scales"], d["zeros"], True, gs, bits, kernel="skinny").float().cpu().numpy()
    Wd = g.dequantize(d["qweight"], d["scales"], d["zeros"], gs, bits).float()
    yt = (x.float() @ Wd.t()).cpu().numpy()
    # single-rounded weights in fp32 (what the GEMM feeds the tensor core)
    q = torch.from_numpy(A.unpack_codes(L["qweight"], bits).astype(np.float32)).to(dev)
    s = bf16_from_bits(L["scales"], dev).float().repeat_interleave(gs, 1)
    b = bf16_from_bits(L["zeros"], dev).float().repeat_interleave(gs, 1)
    W1 = torch.addcmul(b, s, q).to(torch.bfloat16).float()
    y1 = (x.float() @ W1.t()).cpu().numpy()
    W0 = torch.addcmul(b.double(), s.double(), q.double())
    y0 = (x.double() @ W0.t()).cpu().numpy()
    sc = np.abs(ref).max()
    e = np.abs(y - ref); i = np.unravel_index(e.argmax(), e.shape)
    print(f"b{bits}: gemm-vs-oracle {e.max()/sc:.2e} at {i}: gemm {y[i]:.5f} oracle {ref[i]:.5f} torch(deq 2-round) {yt[i]:.5f} torch(1-round bf16 W) {y1[i]:.5f} torch f64 exact W {y0[i]:.5f}")
    print(f"     oracle-vs-f64torch {np.abs(ref - y0).max()/sc:.2e}  (gemm - y1 fp32) max {np.abs(y - y1).max()/sc:.2e}  skinny rows0-7 vs oracle {np.abs(ysk - ref[:8]).max()/sc:.2e}")
    print("     errors per column n of worst:", np.sort(e.max(0))[-5:], "median col max", np.median(e.max(0)))
