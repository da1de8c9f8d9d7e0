"""GPU parity tests of the index-batched quantized matmul (gbxq_gather_qmm): mx.gather_qmm as QuantizedSwitchLinear calls
it (gbx_lm/models/switch_layers.py:78-92) and mx.quantized_matmul(transpose=False) / batched transpose=True as the
quantized-KV attention calls them (gbx_lm/models/base.py:85-93), against the oracle's fp64 restatements."""
import numpy as np
import pytest
import torch

from oracle import mlx_affine as A
from tests.gpu_util import assert_close_to_truth, bf16_from_bits, bits_from_bf16, u32_to_torch

pytestmark = pytest.mark.gpu

BITS = (2, 3, 4, 6, 8)


def _stack(E, N, K, bits, gs, seed):
    Ls = [A.synth_layer(N, K, bits, gs, seed=seed + e) for e in range(E)]
    return {k: np.stack([L[k] for L in Ls]) for k in ("qweight", "scales", "zeros")}


def _cuda(st, dev):
    return u32_to_torch(st["qweight"], dev), bf16_from_bits(st["scales"], dev), bf16_from_bits(st["zeros"], dev)


@pytest.mark.parametrize("bits", BITS)
@pytest.mark.parametrize("gs", (32, 64, 128))
def test_gather_qmm_switch_linear_call_shape(cuda_device, bits, gs):
    """x [T, 1, 1, K] against indices [T, topk] (the SwitchGLU call, switch_layers.py:181-196), per-expert bias."""
    from gbx_lm_b200 import ops

    E, N, K, T, topk = 6, 72, 256, 5, 3
    st = _stack(E, N, K, bits, gs, seed=bits + gs)
    rng = np.random.default_rng(bits * gs)
    bias = A.f32_to_bf16_bits(rng.standard_normal((E, N)).astype(np.float32) * 0.05)
    x = A.synth_x(T, K, seed=4).reshape(T, 1, 1, K)
    idx = rng.integers(0, E, size=(T, topk))
    qw, s, z = _cuda(st, cuda_device)
    y = ops.gather_qmm(bf16_from_bits(x, cuda_device), qw, s, z, rhs_indices=torch.from_numpy(idx).to(cuda_device),
                       transpose=True, group_size=gs, bits=bits, bias=bf16_from_bits(bias, cuda_device))
    assert y.shape == (T, topk, 1, N) and y.dtype == torch.bfloat16
    ref = A.gather_qmm(x, st["qweight"], st["scales"], st["zeros"], np.arange(T).reshape(T, 1), idx, True, gs, bits, "bf16", bias=bias)
    assert_close_to_truth(y, ref, f"gather_qmm b{bits} g{gs}", 1e-2)
    # deterministic, and independent of the scheduling hint
    y2 = ops.gather_qmm(bf16_from_bits(x, cuda_device), qw, s, z, rhs_indices=torch.from_numpy(idx).to(cuda_device),
                        transpose=True, group_size=gs, bits=bits, bias=bf16_from_bits(bias, cuda_device), sorted_indices=True)
    assert torch.equal(y, y2)


def test_gather_qmm_rows_per_item_lhs_indices_and_dtypes(cuda_device):
    """Several rows of x per batch item (M = 6 spans two row tiles), explicit lhs_indices, f16 / f32 operands, and
    agreement with the QuantizedLinear kernel on the same (x, expert) pairs."""
    from gbx_lm_b200 import ops
    import gbx_lm_b200 as g

    E, N, K, bits, gs = 4, 130, 512, 4, 64
    st = _stack(E, N, K, bits, gs, seed=11)
    qw, s, z = _cuda(st, cuda_device)
    x = A.synth_x(3 * 6, K, seed=5).reshape(3, 6, K)
    lhs = np.array([2, 0, 0, 1, 2])
    rhs = np.array([3, 3, 0, 1, 2])
    xt = bf16_from_bits(x, cuda_device)
    y = ops.gather_qmm(xt, qw, s, z, torch.from_numpy(lhs).to(cuda_device), torch.from_numpy(rhs).to(cuda_device), True, gs, bits)
    assert y.shape == (5, 6, N)
    ref = A.gather_qmm(x, st["qweight"], st["scales"], st["zeros"], lhs, rhs, True, gs, bits, "bf16")
    assert_close_to_truth(y, ref, "gather_qmm lhs/rhs", 1e-2)
    for i in range(5):
        single = g.quantized_matmul(xt[lhs[i]], qw[rhs[i]], s[rhs[i]], z[rhs[i]], True, gs, bits, kernel="generic")
        assert torch.equal(y[i], single)  # same arithmetic as the shape-agnostic QuantizedLinear kernel
    for dt, name in ((torch.float16, "f16"), (torch.float32, "f32")):
        xs, ss, zs = xt.to(dt), s.to(dt), z.to(dt)
        yd = ops.gather_qmm(xs, qw, ss, zs, torch.from_numpy(lhs).to(cuda_device), torch.from_numpy(rhs).to(cuda_device), True, gs, bits)
        assert yd.dtype == dt
        assert (yd.float() - torch.from_numpy(ref).to(cuda_device)).abs().max() <= 1e-2 * np.abs(ref).max()


@pytest.mark.parametrize("bits", BITS)
def test_qmm_transpose_false_gqa_broadcast(cuda_device, bits):
    """scores [B, kv, rep, L, T] @ V codes [B, kv, 1, T, D] (base.py:80-92): V quantized along head_dim."""
    import gbx_lm_b200 as g

    B, kv, rep, Lq, T, D, gs = 2, 2, 4, 3, 37, 128, 64
    Ls = [A.synth_layer(T, D, bits, gs, seed=60 + i) for i in range(B * kv)]
    pk = {k: np.stack([L[k] for L in Ls]).reshape(B, kv, 1, T, -1) for k in ("qweight", "scales", "zeros")}
    x = A.synth_x(B * kv * rep * Lq, T, seed=3).reshape(B, kv, rep, Lq, T)
    qw, s, z = _cuda(pk, cuda_device)
    y = g.quantized_matmul(bf16_from_bits(x, cuda_device), qw, s, z, False, gs, bits)
    assert y.shape == (B, kv, rep, Lq, D)
    ref = A.quantized_matmul_nt(x, pk["qweight"], pk["scales"], pk["zeros"], gs, bits, "bf16")
    assert_close_to_truth(y, ref, f"qmm transpose=False b{bits}", 1e-2)


def test_qmm_transpose_false_wide_and_long(cuda_device):
    """N wider than one block of octets (N = 640) and a long contraction (K = 3000 cached tokens), 2-D operands."""
    import gbx_lm_b200 as g

    K, N, bits, gs, M = 3000, 640, 4, 32, 5
    L = A.synth_layer(K, N, bits, gs, seed=8)
    x = A.synth_x(M, K, seed=2)
    y = g.quantized_matmul(bf16_from_bits(x, cuda_device), u32_to_torch(L["qweight"], cuda_device),
                           bf16_from_bits(L["scales"], cuda_device), bf16_from_bits(L["zeros"], cuda_device), False, gs, bits)
    ref = A.quantized_matmul_nt(x, L["qweight"], L["scales"], L["zeros"], gs, bits, "bf16")
    assert y.shape == (M, N)
    assert_close_to_truth(y, ref, "qmm transpose=False wide", 1e-2)


def test_switch_glu_and_quantized_attention_modules(cuda_device):
    """The two callers (gbx_lm_b200.switch_layers) against fp32 torch models on oracle-dequantised weights: SwitchGLU
    below and above the 64-index sort threshold, and attention on a group-quantized KV cache with GQA and a mask."""
    from gbx_lm_b200 import packing, switch_layers as SL

    dev = cuda_device
    E, H, I, bits, gs = 8, 256, 384, 4, 64
    glu = SL.SwitchGLU(H, I, E, bias=False, group_size=gs, bits=bits, device=dev)
    deq = {}
    for name, (n, k) in (("gate_proj", (I, H)), ("up_proj", (I, H)), ("down_proj", (H, I))):
        st = _stack(E, n, k, bits, gs, seed=hash(name) % 1000)
        mod = getattr(glu, name)
        mod.qweight, mod.scales, mod.zeros = _cuda(st, dev)
        deq[name] = torch.from_numpy(np.stack([A.dequantize(st["qweight"][e], st["scales"][e], st["zeros"][e], gs, bits, "bf16") for e in range(E)])).to(dev)
    for T, topk in ((4, 2), (40, 2)):  # 8 and 80 indices: unsorted and sorted paths
        x = torch.randn((1, T, H), generator=torch.Generator().manual_seed(T)).to(torch.bfloat16).to(dev)
        idx = torch.randint(0, E, (1, T, topk), generator=torch.Generator().manual_seed(T + 1)).to(dev)
        y = glu(x, idx)
        assert y.shape == (1, T, topk, H)
        xf = x.float()
        ref = torch.empty((1, T, topk, H), device=dev)
        for t in range(T):
            for j in range(topk):
                e = int(idx[0, t, j])
                up = (xf[0, t] @ deq["up_proj"][e].t()).to(torch.bfloat16).float()
                gate = (xf[0, t] @ deq["gate_proj"][e].t()).to(torch.bfloat16).float()
                h = (torch.nn.functional.silu(gate.to(torch.bfloat16)) * up.to(torch.bfloat16)).float()
                ref[0, t, j] = h @ deq["down_proj"][e].t()
        assert (y.float() - ref).abs().max() <= 2e-2 * ref.abs().max()

    B, Hq, Hkv, Lq, T, D, kb = 1, 8, 2, 2, 50, 128, 8
    gen = torch.Generator().manual_seed(3)
    q = torch.randn((B, Hq, Lq, D), generator=gen).to(torch.bfloat16).to(dev)
    k = torch.randn((B, Hkv, T, D), generator=gen).to(torch.bfloat16)
    v = torch.randn((B, Hkv, T, D), generator=gen).to(torch.bfloat16)

    def quant(t):
        qw, s, z = packing.quantize_affine(t.reshape(-1, D), gs, kb)
        return tuple(a.reshape(B, Hkv, T, -1).to(dev) for a in (qw, s, z))

    qk, qv = quant(k), quant(v)
    mask = torch.zeros((Lq, T), dtype=torch.bfloat16, device=dev)
    mask[0, -1] = float("-inf")  # causal: the first query does not see the last key
    out = SL.quantized_scaled_dot_product_attention(q, qk, qv, D ** -0.5, mask, group_size=gs, bits=kb)
    assert out.shape == (B, Hq, Lq, D)

    def deq_kv(t3):
        w = A.dequantize(t3[0].cpu().view(torch.int32).numpy().view(np.uint32).reshape(-1, t3[0].shape[-1]),
                         bits_from_bf16(t3[1]).reshape(-1, t3[1].shape[-1]), bits_from_bf16(t3[2]).reshape(-1, t3[2].shape[-1]), gs, kb, "bf16")
        return torch.from_numpy(w).reshape(B, Hkv, T, D).to(dev)

    kd, vd = deq_kv(qk).repeat_interleave(Hq // Hkv, 1), deq_kv(qv).repeat_interleave(Hq // Hkv, 1)
    sc = (q.float() * D ** -0.5) @ kd.transpose(-1, -2) + mask.float()
    ref = torch.softmax(sc, -1) @ vd
    assert (out.float() - ref).abs().max() <= 2e-2 * ref.abs().max()


def test_gather_qmm_validation(cuda_device):
    from gbx_lm_b200 import ops

    dev = cuda_device
    st = _stack(2, 8, 64, 4, 64, seed=1)
    qw, s, z = _cuda(st, dev)
    x = torch.zeros((3, 1, 64), dtype=torch.bfloat16, device=dev)
    idx = torch.tensor([0, 1, 1], device=dev)
    assert ops.gather_qmm(x, qw, s, z, rhs_indices=idx).shape == (3, 1, 8)
    with pytest.raises(ValueError):
        ops.gather_qmm(x, qw[0], s[0], z[0], rhs_indices=idx)            # not a stack
    with pytest.raises(ValueError):
        ops.gather_qmm(x, qw, s, z, rhs_indices=idx, bits=5)
    with pytest.raises(ValueError):
        ops.gather_qmm(x[..., :32], qw, s, z, rhs_indices=idx)           # K mismatch
    with pytest.raises(ValueError):
        ops.gather_qmm(x.float(), qw, s, z, rhs_indices=idx)             # dtype mismatch with the scales
    with pytest.raises(RuntimeError):
        ops.gather_qmm(x.cpu(), qw, s, z, rhs_indices=idx)               # no CPU path
    # out-of-range expert indices are clamped, not dereferenced
    y = ops.gather_qmm(x + 1, qw, s, z, rhs_indices=torch.tensor([5, -3, 1], device=dev))
    ref = ops.gather_qmm(x + 1, qw, s, z, rhs_indices=torch.tensor([1, 0, 1], device=dev))
    assert torch.equal(y, ref)
