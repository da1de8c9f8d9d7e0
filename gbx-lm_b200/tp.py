"""Tensor-parallel sharding of the QuantizedLinear projections (the reference itself
has no TP -- section 2.2 -- this is the new work BASELINE.json's north_star defines).

  column-parallel  q_proj k_proj v_proj gate_proj up_proj : split N (rows of qweight/scales/zeros, bias)
  row-parallel     o_proj down_proj                       : split K (packed columns at (K/tp)*bits/32
                                                            words, scale columns at K/(tp*gs)) and sum
                                                            the partial [M, hidden] outputs across ranks

Pure slicing -- no re-packing -- legal iff (K/tp) % group_size == 0 and (K/tp)*bits % 32 == 0.
One process per GPU; the collective is NCCL over NVLink (`torch.distributed`), or the library's
one-shot peer-memory all-reduce (`gbxq_allreduce_oneshot`) for the latency-bound decode messages.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

COLUMN_PARALLEL = ("q_proj", "k_proj", "v_proj", "gate_proj", "up_proj")
ROW_PARALLEL = ("o_proj", "down_proj")


class TPContext:
    """Rank/world of the tensor-parallel group + the all-reduce used after row-parallel layers."""

    def __init__(self, rank: int = 0, world: int = 1, group=None, oneshot: "Optional[OneShotAllReduce]" = None,
                 fused: "Optional[FusedRowParallel]" = None):
        self.rank = rank
        self.world = world
        self.group = group
        self.oneshot = oneshot
        self.fused = fused

    def row_parallel(self, layer, x: torch.Tensor) -> torch.Tensor:
        """y = all-reduce(x_shard . W_shard^T) of a row-parallel QuantizedLinear (o_proj / down_proj): ONE kernel when the
        decode kernel serves the call (gbxq_qmm_rowpar_allreduce), otherwise the matmul followed by the all-reduce."""
        if self.world > 1 and self.fused is not None:
            y = self.fused(layer, x)
            if y is not None:
                return y
        return self.all_reduce(layer(x))

    def all_reduce(self, y: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return y
        if self.oneshot is not None and y.is_cuda and self.oneshot.fits(y):
            return self.oneshot(y)
        import torch.distributed as dist

        dist.all_reduce(y, group=self.group)
        return y


def proj_kind(name: str) -> Optional[str]:
    for p in COLUMN_PARALLEL:
        if f".{p}." in name or name.endswith("." + p):
            return "column"
    for p in ROW_PARALLEL:
        if f".{p}." in name or name.endswith("." + p):
            return "row"
    return None


def check_row_split(k: int, bits: int, group_size: int, world: int):
    if k % world or (k // world) % group_size or ((k // world) * bits) % 32:
        raise ValueError(
            f"row-parallel split of K={k} over {world} ranks needs (K/tp) % group_size == 0 and "
            f"(K/tp)*bits % 32 == 0 (bits={bits}, group_size={group_size})"
        )


def shard_tensor(name: str, t: torch.Tensor, bits: int, group_size: int, rank: int, world: int) -> torch.Tensor:
    """Slice one checkpoint tensor of a QuantizedLinear (`<module>.{qweight,scales,zeros,bias}`) for `rank`."""
    if world == 1:
        return t
    kind = proj_kind(name.rsplit(".", 1)[0])
    leaf = name.rsplit(".", 1)[1]
    if kind is None or leaf not in ("qweight", "scales", "zeros", "bias"):
        return t
    if kind == "column":
        n = t.shape[0]
        if n % world:
            raise ValueError(f"{name}: N={n} not divisible by tp={world}")
        per = n // world
        return t[rank * per : (rank + 1) * per].contiguous()
    # row-parallel
    if leaf == "bias":
        # added once: rank 0 keeps it, the others add zero
        return t if rank == 0 else torch.zeros_like(t)
    if leaf == "qweight":
        k = t.shape[1] * 32 // bits
        check_row_split(k, bits, group_size, world)
        per = (k // world) * bits // 32
    else:
        k = t.shape[1] * group_size
        check_row_split(k, bits, group_size, world)
        per = (k // world) // group_size
    return t[:, rank * per : (rank + 1) * per].contiguous()


def shard_state_dict(weights: Dict[str, torch.Tensor], bits_of, rank: int, world: int) -> Dict[str, torch.Tensor]:
    """bits_of(module_name) -> (bits, group_size) for every QuantizedLinear module name."""
    if world == 1:
        return weights
    out = {}
    for k, v in weights.items():
        mod = k.rsplit(".", 1)[0]
        if proj_kind(mod) is not None:
            b, g = bits_of(mod)
            out[k] = shard_tensor(k, v, b, g, rank, world)
        else:
            out[k] = v
    return out


class OneShotAllReduce:
    """Latency-optimised sum all-reduce over NVLink peer memory (gbxq_allreduce_oneshot): staging
    buffers and flags live in torch symmetric memory so every rank can address every peer."""

    def __init__(self, group, device: torch.device, capacity_elems: int = 1 << 20, dtype=torch.bfloat16):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.dtype = dtype
        self.capacity = capacity_elems
        self.buf = symm.empty((capacity_elems,), dtype=dtype, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group.group_name)
        # control words behind the flags (include/gbxq.h: world * GBXQ_AR_MAX_CTAS + 3 entries are needed): +0 the
        # device-resident sequence number (seq = 0 mode: CUDA-graph replayable), +1 the CTA counter, +2 the time-out word
        self.flags = symm.empty((self.world * _lib.AR_MAX_CTAS + 4,), dtype=torch.int32, device=device)
        self.flags.zero_()
        self.fhdl = symm.rendezvous(self.flags, self.group.group_name)
        bufs = [self.hdl.buffer_ptrs[r] for r in range(self.world)]
        flgs = [self.fhdl.buffer_ptrs[r] for r in range(self.world)]
        self.bufs_dev = torch.tensor(bufs, dtype=torch.int64, device=device)
        self.flags_dev = torch.tensor(flgs, dtype=torch.int64, device=device)
        self.seq = 0
        torch.cuda.synchronize(device)
        dist.barrier(self.group)

    # The one-shot kernel is for the latency-bound decode messages (M*hidden*2 bytes: 16 KB .. 1 MB); a prefill chunk's
    # 40 MB partial goes to NCCL's bandwidth-optimal rings (r03e: Qwen2.5-32B prefill of 4096 tokens at tp4 took 1165 ms
    # with every message on the one-shot kernel, 5x the single-GPU time).
    MAX_BYTES = 1 << 20

    def timed_out(self) -> bool:
        """True once a wait for a peer gave up (4 s): results of that call are undefined (synchronises)."""
        from . import _lib

        return bool(self.flags[self.world * _lib.AR_MAX_CTAS + 2].item())

    def fits(self, y: torch.Tensor) -> bool:
        nbytes = y.numel() * y.element_size()
        return y.dtype == self.dtype and y.numel() * 2 <= self.capacity and nbytes % 16 == 0 and nbytes <= self.MAX_BYTES

    def __call__(self, y: torch.Tensor) -> torch.Tensor:
        from . import _lib

        if not y.is_contiguous():
            raise ValueError("OneShotAllReduce reduces in place: y must be contiguous")
        dt = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}[y.dtype]
        st = torch.cuda.current_stream().cuda_stream
        rc = _lib.get().gbxq_allreduce_oneshot(
            y.data_ptr(), y.data_ptr(), y.numel(), dt, self.bufs_dev.data_ptr(), self.flags_dev.data_ptr(),
            self.capacity, self.rank, self.world, 0, st,
        )
        _lib.check(rc, "gbxq_allreduce_oneshot")
        return y


class FusedRowParallel:
    """Row-parallel QuantizedLinear with the all-reduce inside the matmul kernel (gbxq_qmm_rowpar_allreduce): fp32
    staging buffers and epoch flags in torch symmetric memory, addressed by every peer over NVLink."""

    def __init__(self, group, device: torch.device, max_elems: int = 4 * 8192):
        import ctypes

        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.stage_elems = 2 * self.world * max_elems  # two halves x world partials of [M, N], 8-byte {fp32, epoch} words
        self.stage = symm.empty((self.stage_elems,), dtype=torch.int64, device=device)
        self.stage.zero_()
        self.shdl = symm.rendezvous(self.stage, self.group.group_name)
        self.flags = symm.empty((self.world * _lib.RP_MAX_CTAS + 4,), dtype=torch.int32, device=device)
        self.flags.zero_()
        self.fhdl = symm.rendezvous(self.flags, self.group.group_name)
        self._stage_ptrs = (ctypes.c_void_p * self.world)(*[self.shdl.buffer_ptrs[r] for r in range(self.world)])
        self._flag_ptrs = (ctypes.c_void_p * self.world)(*[self.fhdl.buffer_ptrs[r] for r in range(self.world)])
        self.comm = _lib.Comm(self._stage_ptrs, self._flag_ptrs, self.stage_elems, self.rank, self.world)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)

    def timed_out(self) -> bool:
        from . import _lib

        return bool(self.flags[self.world * _lib.RP_MAX_CTAS + 2].item())

    def __call__(self, layer, x: torch.Tensor) -> Optional[torch.Tensor]:
        """Returns the reduced y, or None when the fused kernel does not serve this call (nothing was enqueued)."""
        import ctypes

        from . import _lib

        if x.dtype != torch.bfloat16 or layer.scales.dtype != torch.bfloat16:
            return None
        k = layer.input_dims
        x2 = x.reshape(-1, k)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        m, n = x2.shape[0], layer.output_dims
        if m > 4 or 2 * self.world * m * n > self.stage_elems:
            return None
        y = torch.empty((m, n), dtype=torch.bfloat16, device=x.device)
        bias = getattr(layer, "bias", None)
        rc = _lib.get().gbxq_qmm_rowpar_allreduce(
            x2.data_ptr(), layer.qweight.data_ptr(), layer.scales.data_ptr(), layer.zeros.data_ptr(),
            bias.data_ptr() if bias is not None else None, y.data_ptr(), m, n, k, layer.bits, layer.group_size, 0,
            ctypes.byref(self.comm), torch.cuda.current_stream().cuda_stream)
        if rc == -9:
            return None
        _lib.check(rc, "gbxq_qmm_rowpar_allreduce")
        return y.reshape(*x.shape[:-1], n)
