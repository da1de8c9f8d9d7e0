"""Helpers shared by the GPU parity tests: numpy oracle arrays <-> CUDA tensors."""
import numpy as np
import torch

from oracle import mlx_affine as A


def bf16_from_bits(bits: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(bits).view(np.int16)).view(torch.bfloat16).to(device)


def bits_from_bf16(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)


def u32_to_torch(w: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(w).view(np.int32)).view(torch.uint32).to(device)


def layer_to_cuda(L: dict, device):
    out = {
        "qweight": u32_to_torch(L["qweight"], device),
        "scales": bf16_from_bits(L["scales"], device),
        "zeros": bf16_from_bits(L["zeros"], device),
    }
    if "bias" in L:
        out["bias"] = bf16_from_bits(L["bias"], device)
    return out


def assert_close_to_truth(y_gpu: torch.Tensor, y_ref: np.ndarray, what: str = "", tol: float = 2.0 ** -7):
    """Tolerance of the path (BASELINE.json north_star: max relative error <= 1e-2 in bf16),
    stated two ways: normalised by the output scale, and element-wise with an rms floor."""
    y = y_gpu.detach().float().cpu().numpy().reshape(y_ref.shape)
    err = np.abs(y - y_ref)
    scale = np.abs(y_ref).max() + 1e-30
    rms = np.sqrt((y_ref.astype(np.float64) ** 2).mean()) + 1e-30
    assert np.isfinite(y).all(), what
    assert err.max() <= tol * scale, f"{what}: max err {err.max():.3e} vs scale {scale:.3e}"
    assert (err <= 1e-2 * np.abs(y_ref) + 1e-2 * rms).all(), f"{what}: element-wise 1e-2 bound violated"
    return err.max() / scale
