"""ctypes access to oracle/libgbxq_oracle.so (TEST INFRASTRUCTURE ONLY; see gbxq_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgbxq_oracle.so")
DTYPES = {"bf16": 0, "f16": 1, "f32": 2}
ACCUM = {"f32": 0, "t": 1, "f64": 2}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gbxq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libgbxq_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        i64, vp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
        _lib.gbxq_oracle_unpack.argtypes = [vp, vp, i64, i64, ci]
        _lib.gbxq_oracle_pack.argtypes = [vp, vp, i64, i64, ci]
        _lib.gbxq_oracle_dequantize.argtypes = [vp, vp, vp, vp, i64, i64, ci, ci, ci]
        _lib.gbxq_oracle_qmm.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i64, ci, ci, ci, ci, ci]
        _lib.gbxq_oracle_qmm_fast.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, ci, ci, ci, ci]
        _lib.gbxq_oracle_quantize.argtypes = [vp, vp, vp, vp, i64, i64, ci, ci, ci]
        _lib.gbxq_oracle_max_threads.restype = ci
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _np_t(dtype: str):
    return np.float32 if dtype == "f32" else np.uint16  # bf16/f16 travel as raw 16-bit patterns


def pack(codes: np.ndarray, bits: int) -> np.ndarray:
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    N, K = codes.shape
    w = np.zeros((N, K * bits // 32), dtype=np.uint32)
    assert lib().gbxq_oracle_pack(_p(codes), _p(w), N, K, bits) == 0
    return w


def unpack(w: np.ndarray, bits: int) -> np.ndarray:
    w = np.ascontiguousarray(w, dtype=np.uint32)
    N = w.shape[0]
    K = w.shape[1] * 32 // bits
    codes = np.zeros((N, K), dtype=np.uint8)
    assert lib().gbxq_oracle_unpack(_p(w), _p(codes), N, K, bits) == 0
    return codes


def dequantize(w, scales, biases, group_size, bits, dtype="bf16") -> np.ndarray:
    """scales/biases/out are raw arrays in `dtype` (uint16 bit patterns for bf16/f16)."""
    w = np.ascontiguousarray(w, dtype=np.uint32)
    s = np.ascontiguousarray(scales, dtype=_np_t(dtype))
    b = np.ascontiguousarray(biases, dtype=_np_t(dtype))
    N = w.shape[0]
    K = w.shape[1] * 32 // bits
    out = np.zeros((N, K), dtype=_np_t(dtype))
    rc = lib().gbxq_oracle_dequantize(_p(w), _p(s), _p(b), _p(out), N, K, bits, group_size, DTYPES[dtype])
    if rc:
        raise ValueError("oracle rejected the arguments")
    return out


def qmm(x, w, scales, biases, group_size, bits, dtype="bf16", accum="f64", bias=None, nthreads=0):
    w = np.ascontiguousarray(w, dtype=np.uint32)
    x = np.ascontiguousarray(x, dtype=_np_t(dtype))
    s = np.ascontiguousarray(scales, dtype=_np_t(dtype))
    b = np.ascontiguousarray(biases, dtype=_np_t(dtype))
    M, K = x.shape
    N = w.shape[0]
    y = np.zeros((M, N), dtype=_np_t(dtype))
    bp = None
    if bias is not None:
        bias = np.ascontiguousarray(bias, dtype=_np_t(dtype))
        bp = _p(bias)
    rc = lib().gbxq_oracle_qmm(
        _p(x), _p(w), _p(s), _p(b), bp, _p(y), M, N, K, bits, group_size, DTYPES[dtype], ACCUM[accum], nthreads
    )
    if rc:
        raise ValueError("oracle rejected the arguments")
    return y


def qmm_fast(x, w, scales, biases, group_size, bits, nthreads=0):
    """bf16-only timing variant (bench.py CPU arm)."""
    w = np.ascontiguousarray(w, dtype=np.uint32)
    x = np.ascontiguousarray(x, dtype=np.uint16)
    s = np.ascontiguousarray(scales, dtype=np.uint16)
    b = np.ascontiguousarray(biases, dtype=np.uint16)
    M, K = x.shape
    N = w.shape[0]
    y = np.zeros((M, N), dtype=np.uint16)
    rc = lib().gbxq_oracle_qmm_fast(_p(x), _p(w), _p(s), _p(b), _p(y), M, N, K, bits, group_size, 0, nthreads)
    if rc:
        raise ValueError("oracle rejected the arguments")
    return y


def quantize(w, group_size, bits, dtype="bf16"):
    """w: [rows, K] raw array in `dtype` (uint16 bit patterns for bf16/f16).  Returns (codes uint32, scales, biases raw)."""
    w = np.ascontiguousarray(w, dtype=_np_t(dtype))
    rows, K = w.shape
    q = np.zeros((rows, K * bits // 32), dtype=np.uint32)
    s = np.zeros((rows, K // group_size), dtype=_np_t(dtype))
    b = np.zeros((rows, K // group_size), dtype=_np_t(dtype))
    rc = lib().gbxq_oracle_quantize(_p(w), _p(q), _p(s), _p(b), rows, K, bits, group_size, DTYPES[dtype])
    if rc:
        raise ValueError("oracle rejected the arguments")
    return q, s, b


def max_threads() -> int:
    return lib().gbxq_oracle_max_threads()
