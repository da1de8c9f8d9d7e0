"""World-size-2 `gloo` test (CPU) of the tensor-parallel host logic: shard-at-load slicing +
TPContext.all_reduce reproduce the unsharded result.  The per-rank matmul is the oracle's (the
product has no CPU path); what is under test is the sharding + collective plumbing."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gbx_lm_b200 import packing, tp as tpmod
        from oracle import mlx_affine as A

        ctx = tpmod.TPContext(rank, world)
        bits, gs, hidden, inter = 4, 64, 256, 512
        to_np = lambda t: t.view(torch.int16).numpy().view(np.uint16)
        u32 = lambda t: t.view(torch.int32).numpy().view(np.uint32)
        up = packing.synth_layer(inter, hidden, bits, gs, seed=1)
        down = packing.synth_layer(hidden, inter, bits, gs, seed=2)
        x = A.synth_x(3, hidden, seed=3)
        f = lambda xx, L: A.quantized_matmul(xx, u32(L["qweight"]), to_np(L["scales"]), to_np(L["zeros"]), gs, bits, "bf16", "f64")
        # unsharded: down(up(x))
        full = f(A.f32_to_bf16_bits(f(x, up)), down)
        # sharded: column-parallel up, row-parallel down, sum all-reduce
        sh = lambda name, L: {k: tpmod.shard_tensor(f"model.layers.0.mlp.{name}.{k}", L[k], bits, gs, rank, world) for k in ("qweight", "scales", "zeros")}
        h_local = f(x, sh("up_proj", up))                                  # [3, inter/world]
        part = f(A.f32_to_bf16_bits(h_local), sh("down_proj", down))       # bf16 partial [3, hidden], as the GPU path emits
        y = ctx.all_reduce(torch.from_numpy(part.astype(np.float32)))
        err = float(np.abs(y.numpy() - full).max() / np.abs(full).max())
        if rank == 0:
            out.put(err)
    finally:
        dist.destroy_process_group()


def test_tp2_column_row_allreduce_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    err = out.get(timeout=5)
    assert err < 2.0 ** -6  # only the bf16 rounding of the partial sums separates the two
