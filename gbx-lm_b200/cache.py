"""The two KV caches either side of the quantized-KV attention path: `KVCache`
(gbx_lm/models/cache.py:216-266) and `QuantizedKVCache` (cache.py:133-213), same attributes, growth rule (blocks of
`step` = 256 positions), `state` / `meta_state` / `trim` semantics and `update_and_fetch` signature.

What is new underneath: `mx.quantize` of the fresh keys / values (cache.py:176-177) plus the three slice assignments
(cache.py:178-180) are ONE `gbxq_quantize_rows` launch per tensor that writes codes, scales and biases at the current
offset of the cache buffers (ops.quantize_into); `KVCache.to_quantized` (cache.py:251-263) is `gbxq_quantize` on the
whole prefix.  The consumer is `switch_layers.quantized_scaled_dot_product_attention` (gbx_lm/models/base.py:65-98).
The decode path of `qllama.py` keeps its own static bf16 cache; these classes serve `maybe_quantize_kv_cache`-style
callers (gbx_lm/utils.py:204-214).  CUDA only, like everything in this package."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops

QTensor = Tuple[torch.Tensor, torch.Tensor, torch.Tensor]


class QuantizedKVCache:
    """cache.py:133-213.  `keys` / `values` are (codes uint32 [B, H, cap, D*bits/32], scales, biases [B, H, cap,
    D/group_size]) triples; `update_and_fetch` returns views of the first `offset` positions."""

    def __init__(self, group_size: int = 64, bits: int = 8):
        self.keys: Optional[QTensor] = None
        self.values: Optional[QTensor] = None
        self.offset = 0
        self.step = 256
        self.group_size = group_size
        self.bits = bits

    # the two device operations, overridable so that the growth / trim logic can be exercised without a GPU
    @staticmethod
    def _quantize_into(x: torch.Tensor, out: QTensor, offset: int, group_size: int, bits: int) -> None:
        ops.quantize_into(x, out, offset, group_size, bits)

    def _init_quant(self, shape, dim: int, dtype: torch.dtype, device) -> QTensor:
        # the reference allocates dim // (32 // bits) words (cache.py:150,156), which is dim * bits / 32 for 2/4/8-bit and
        # mis-sized for the straddling 3-/6-bit packings (32 // 3 = 10): the stream length is used here
        words = dim * self.bits // 32
        # uint32 has few torch kernels: allocate / concatenate as int32 and view
        return (torch.zeros((*shape, words), dtype=torch.int32, device=device).view(torch.uint32),
                torch.zeros((*shape, dim // self.group_size), dtype=dtype, device=device),
                torch.zeros((*shape, dim // self.group_size), dtype=dtype, device=device))

    def update_and_fetch(self, keys: torch.Tensor, values: torch.Tensor) -> Tuple[QTensor, QTensor]:
        B, n_kv_heads, num_steps, k_head_dim = keys.shape
        v_head_dim = values.shape[-1]
        prev = self.offset
        if self.keys is None or (prev + num_steps) > self.keys[0].shape[-2]:
            new_steps = (self.step + num_steps - 1) // self.step * self.step
            shape = (B, n_kv_heads, new_steps)
            if self.keys is not None:
                def expand(x: torch.Tensor) -> torch.Tensor:
                    if prev % self.step != 0:
                        x = x[..., :prev, :]
                    codes = x.dtype == torch.uint32
                    if codes:
                        x = x.view(torch.int32)
                    x = torch.cat([x, torch.zeros((*shape, x.shape[-1]), dtype=x.dtype, device=x.device)], dim=-2)
                    return x.view(torch.uint32) if codes else x

                self.keys = tuple(expand(x) for x in self.keys)
                self.values = tuple(expand(x) for x in self.values)
            else:
                self.keys = self._init_quant(shape, k_head_dim, keys.dtype, keys.device)
                self.values = self._init_quant(shape, v_head_dim, values.dtype, values.device)
        if not all(x.is_contiguous() for x in (*self.keys, *self.values)):  # a `state` set from views
            self.keys = tuple(x.contiguous() for x in self.keys)
            self.values = tuple(x.contiguous() for x in self.values)
        self.offset += num_steps
        self._quantize_into(keys, self.keys, prev, self.group_size, self.bits)
        self._quantize_into(values, self.values, prev, self.group_size, self.bits)
        return (tuple(x[..., : self.offset, :] for x in self.keys), tuple(x[..., : self.offset, :] for x in self.values))

    @property
    def state(self):
        if self.offset == self.keys[0].shape[2]:
            return self.keys, self.values
        return (tuple(x[..., : self.offset, :] for x in self.keys), tuple(x[..., : self.offset, :] for x in self.values))

    @state.setter
    def state(self, v):
        self.keys, self.values = v

    @property
    def meta_state(self):
        return tuple(map(str, (self.step, self.offset, self.group_size, self.bits)))

    @meta_state.setter
    def meta_state(self, v):
        self.step, self.offset, self.group_size, self.bits = map(int, v)

    def is_trimmable(self) -> bool:
        return True

    def trim(self, n: int) -> int:
        n = min(self.offset, n)
        self.offset -= n
        return n


class KVCache:
    """cache.py:216-266: the dense cache a prompt is processed into before `to_quantized`."""

    def __init__(self):
        self.keys: Optional[torch.Tensor] = None
        self.values: Optional[torch.Tensor] = None
        self.offset = 0
        self.step = 256

    @staticmethod
    def _quantize(x: torch.Tensor, group_size: int, bits: int) -> QTensor:
        return ops.quantize(x, group_size, bits)

    def update_and_fetch(self, keys: torch.Tensor, values: torch.Tensor):
        prev = self.offset
        if self.keys is None or (prev + keys.shape[2]) > self.keys.shape[2]:
            B, n_kv_heads, _, k_head_dim = keys.shape
            v_head_dim = values.shape[3]
            n_steps = (self.step + keys.shape[2] - 1) // self.step
            new_k = torch.zeros((B, n_kv_heads, n_steps * self.step, k_head_dim), dtype=keys.dtype, device=keys.device)
            new_v = torch.zeros((B, n_kv_heads, n_steps * self.step, v_head_dim), dtype=values.dtype, device=values.device)
            if self.keys is not None:
                if prev % self.step != 0:
                    self.keys = self.keys[..., :prev, :]
                    self.values = self.values[..., :prev, :]
                self.keys = torch.cat([self.keys, new_k], dim=2)
                self.values = torch.cat([self.values, new_v], dim=2)
            else:
                self.keys, self.values = new_k, new_v
        self.offset += keys.shape[2]
        self.keys[..., prev: self.offset, :] = keys
        self.values[..., prev: self.offset, :] = values
        return self.keys[..., : self.offset, :], self.values[..., : self.offset, :]

    @property
    def state(self):
        if self.offset == self.keys.shape[2]:
            return self.keys, self.values
        return self.keys[..., : self.offset, :], self.values[..., : self.offset, :]

    @state.setter
    def state(self, v):
        self.keys, self.values = v
        self.offset = self.keys.shape[2]

    def is_trimmable(self) -> bool:
        return True

    def trim(self, n: int) -> int:
        n = min(self.offset, n)
        self.offset -= n
        return n

    def to_quantized(self, group_size: int = 64, bits: int = 4) -> QuantizedKVCache:
        quant_cache = QuantizedKVCache(group_size=group_size, bits=bits)
        quant_cache.offset = self.offset
        if self.keys is not None:
            quant_cache.keys = self._quantize(self.keys, group_size, bits)
            quant_cache.values = self._quantize(self.values, group_size, bits)
        return quant_cache


def maybe_quantize_kv_cache(prompt_cache, quantized_kv_start: int, kv_group_size: int, kv_bits: Optional[int]) -> None:
    """gbx_lm/utils.py:204-214: once the prompt is past `quantized_kv_start`, swap every dense cache for its
    quantized form (in place in the list)."""
    if kv_bits is not None and not isinstance(prompt_cache[0], QuantizedKVCache) and prompt_cache[0].offset > quantized_kv_start:
        for i in range(len(prompt_cache)):
            if isinstance(prompt_cache[i], KVCache):
                prompt_cache[i] = prompt_cache[i].to_quantized(group_size=kv_group_size, bits=kv_bits)
