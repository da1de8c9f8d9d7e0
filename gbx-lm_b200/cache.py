"""The two KV caches either side of the quantized-KV attention path, with the interface of the
reference's `KVCache` (gbx_lm/models/cache.py:216-266) and `QuantizedKVCache` (cache.py:133-213): attributes `keys`,
`values`, `offset`, `step`, (`group_size`, `bits`), methods `update_and_fetch`, `state`, `meta_state`, `trim`,
`is_trimmable`, `to_quantized`, and the same growth rule, so that a caller written against the reference
(`maybe_quantize_kv_cache`, gbx_lm/utils.py:204-214; the attention layers) runs unchanged.

Built differently underneath.  Both caches are one block-grown store (`_BlockStore`): tuples of position-major CUDA
buffers `[B, H, capacity, width]` that share `offset`; the dense cache holds one buffer per side, the quantized one
three (codes, scales, biases).  What the reference does with `mx.quantize` followed by three slice assignments per
side (cache.py:176-180) is ONE `gbxq_quantize_rows` launch per side that writes codes and statistics at the offset
(`ops.quantize_into`); `to_quantized` (cache.py:251-263) is `gbxq_quantize` over the buffers.  The consumer is
`switch_layers.quantized_scaled_dot_product_attention` (gbx_lm/models/base.py:65-98).  The decode path of `qllama.py`
keeps its own static bf16 cache.  CUDA only, like everything in this package."""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch

from . import ops

QTensor = Tuple[torch.Tensor, torch.Tensor, torch.Tensor]
Bufs = Tuple[torch.Tensor, ...]


def _zeros(shape, dtype: torch.dtype, device) -> torch.Tensor:
    # uint32 has few torch kernels: codes are allocated / concatenated as int32 and viewed
    if dtype == torch.uint32:
        return torch.zeros(shape, dtype=torch.int32, device=device).view(torch.uint32)
    return torch.zeros(shape, dtype=dtype, device=device)


def _cat_positions(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    if a.dtype == torch.uint32:
        return torch.cat([a.view(torch.int32), b.view(torch.int32)], dim=-2).view(torch.uint32)
    return torch.cat([a, b], dim=-2)


class _BlockStore:
    """`offset` valid positions in buffers that grow by whole blocks of `step` positions.

    Growth rule (the reference's, cache.py:145-171 / 224-239, kept because `state` exposes the capacity): when the new
    positions do not fit, ceil(new / step) blocks are appended -- after cutting the buffers back to `offset` if that is
    not a block border."""

    def __init__(self):
        self.offset = 0
        self.step = 256

    @staticmethod
    def _capacity(bufs: Optional[Bufs]) -> int:
        return 0 if bufs is None else bufs[0].shape[-2]

    def _make_room(self, sides: Sequence[Optional[Bufs]], new: int, fresh: Callable[[int], Sequence[Bufs]]):
        """The sides (tuples of buffers) with room for `new` more positions; `fresh(n)` allocates n zeroed positions
        of every buffer of every side."""
        if sides[0] is not None and self.offset + new <= self._capacity(sides[0]):
            return list(sides)
        blocks = -(-new // self.step) * self.step
        grown = []
        for old, add in zip(sides, fresh(blocks)):
            if old is None:
                grown.append(tuple(add))
                continue
            cut = self.offset if self.offset % self.step else self._capacity(old)
            grown.append(tuple(_cat_positions(o[..., :cut, :], a) for o, a in zip(old, add)))
        return grown

    def _valid(self, bufs: Bufs) -> Bufs:
        return tuple(t[..., : self.offset, :] for t in bufs)

    def is_trimmable(self) -> bool:
        return True

    def trim(self, n: int) -> int:
        n = min(self.offset, n)
        self.offset -= n
        return n


class QuantizedKVCache(_BlockStore):
    """cache.py:133-213.  `keys` / `values` are (codes uint32 [B, H, capacity, D*bits/32], scales, biases [B, H,
    capacity, D/group_size]) triples; `update_and_fetch` returns views of the first `offset` positions."""

    def __init__(self, group_size: int = 64, bits: int = 8):
        super().__init__()
        self.keys: Optional[QTensor] = None
        self.values: Optional[QTensor] = None
        self.group_size = group_size
        self.bits = bits

    # the device operation, overridable so that growth / trim / state can be exercised without a GPU
    @staticmethod
    def _quantize_into(x: torch.Tensor, out: QTensor, offset: int, group_size: int, bits: int) -> None:
        ops.quantize_into(x, out, offset, group_size, bits)

    def _fresh(self, lead, dims, like: torch.Tensor):
        def alloc(n: int):
            # D * bits / 32 words per position: the reference's D // (32 // bits) (cache.py:150,156) is the same for
            # 2/4/8-bit and mis-sized for the straddling 3-/6-bit packings (32 // 3 = 10)
            return [(_zeros((*lead, n, d * self.bits // 32), torch.uint32, like.device),
                     _zeros((*lead, n, d // self.group_size), like.dtype, like.device),
                     _zeros((*lead, n, d // self.group_size), like.dtype, like.device)) for d in dims]

        return alloc

    def update_and_fetch(self, keys: torch.Tensor, values: torch.Tensor) -> Tuple[QTensor, QTensor]:
        new = keys.shape[-2]
        k_side, v_side = self._make_room((self.keys, self.values), new,
                                         self._fresh(keys.shape[:-2], (keys.shape[-1], values.shape[-1]), keys))
        # the kernel writes into the buffers in place: they must be dense (a `state` set from views is not)
        self.keys = tuple(t if t.is_contiguous() else t.contiguous() for t in k_side)
        self.values = tuple(t if t.is_contiguous() else t.contiguous() for t in v_side)
        at = self.offset
        self.offset = at + new
        self._quantize_into(keys, self.keys, at, self.group_size, self.bits)
        self._quantize_into(values, self.values, at, self.group_size, self.bits)
        return self._valid(self.keys), self._valid(self.values)

    @property
    def state(self):
        if self.offset == self._capacity(self.keys):
            return self.keys, self.values
        return self._valid(self.keys), self._valid(self.values)

    @state.setter
    def state(self, v):
        self.keys, self.values = v

    @property
    def meta_state(self):
        return tuple(str(v) for v in (self.step, self.offset, self.group_size, self.bits))

    @meta_state.setter
    def meta_state(self, v):
        self.step, self.offset, self.group_size, self.bits = (int(s) for s in v)


class KVCache(_BlockStore):
    """cache.py:216-266: the dense cache a prompt is processed into before `to_quantized`."""

    def __init__(self):
        super().__init__()
        self.keys: Optional[torch.Tensor] = None
        self.values: Optional[torch.Tensor] = None

    @staticmethod
    def _quantize(x: torch.Tensor, group_size: int, bits: int) -> QTensor:
        return ops.quantize(x, group_size, bits)

    def update_and_fetch(self, keys: torch.Tensor, values: torch.Tensor):
        new = keys.shape[-2]

        def fresh(n: int):
            return [(_zeros((*t.shape[:-2], n, t.shape[-1]), t.dtype, t.device),) for t in (keys, values)]

        sides = self._make_room([None if t is None else (t,) for t in (self.keys, self.values)], new, fresh)
        self.keys, self.values = sides[0][0], sides[1][0]
        at = self.offset
        self.offset = at + new
        self.keys[..., at: self.offset, :] = keys
        self.values[..., at: self.offset, :] = values
        return self._valid((self.keys, self.values))

    @property
    def state(self):
        if self.offset == self.keys.shape[-2]:
            return self.keys, self.values
        return self._valid((self.keys, self.values))

    @state.setter
    def state(self, v):
        self.keys, self.values = v
        self.offset = self.keys.shape[-2]

    def to_quantized(self, group_size: int = 64, bits: int = 4) -> QuantizedKVCache:
        """The whole buffers are quantized (unused positions are zero groups), `offset` carries over."""
        q = QuantizedKVCache(group_size=group_size, bits=bits)
        q.offset = self.offset
        if self.keys is not None:
            q.keys = self._quantize(self.keys, group_size, bits)
            q.values = self._quantize(self.values, group_size, bits)
        return q


def maybe_quantize_kv_cache(prompt_cache: list, quantized_kv_start: int, kv_group_size: int, kv_bits: Optional[int]) -> None:
    """gbx_lm/utils.py:204-214: once the prompt is longer than `quantized_kv_start`, every dense cache of the list is
    replaced by its quantized form (the list is edited in place; a list that already starts quantized is left alone)."""
    if kv_bits is None or isinstance(prompt_cache[0], QuantizedKVCache) or prompt_cache[0].offset <= quantized_kv_start:
        return
    prompt_cache[:] = [c.to_quantized(group_size=kv_group_size, bits=kv_bits) if isinstance(c, KVCache) else c
                       for c in prompt_cache]
