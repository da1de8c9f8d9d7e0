"""CPU tests of the host logic around the path: checkpoint layout (gba2mlx format), layer-mix loader
(strategy -> per-layer bits/group_size), shard slicing for tensor parallelism.  No GPU compute."""
import json
import os

import numpy as np
import pytest
import torch

from gbx_lm_b200 import QuantizedLinear, packing, tp as tpmod, utils, workloads as W
from oracle import mlx_affine as A


@pytest.fixture(scope="module")
def tiny_ckpt(tmp_path_factory):
    d = tmp_path_factory.mktemp("tiny_llama")
    dims = W.MODELS["tiny-llama"]
    strat = W.strategy_bpw22(dims.layers)
    utils.write_synthetic_checkpoint(d, dims, strat, seed=3, default_bits=2, default_gs=128)
    return d, dims, strat


def test_checkpoint_layout_matches_gba2mlx(tiny_ckpt):
    d, dims, strat = tiny_ckpt
    from safetensors import safe_open

    assert (d / "config.json").exists() and (d / "quant_strategy.json").exists()
    assert (d / "model.safetensors").exists() and (d / "model.safetensors.index.json").exists()
    cfg = json.load(open(d / "config.json"))
    assert cfg["quantization"] == {"group_size": 128, "bits": 2} and list(cfg) == sorted(cfg)
    with safe_open(str(d / "model.safetensors"), "pt") as f:
        assert f.metadata() == {"format": "mlx"}
        keys = set(f.keys())
        assert "model.embed_tokens.weight" in keys and "model.norm.weight" in keys and "lm_head.weight" not in keys
        for p in W.PROJS:
            sub = "self_attn" if p in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"
            for leaf in ("qweight", "scales", "zeros"):
                assert f"model.layers.0.{sub}.{p}.{leaf}" in keys
        qw = f.get_tensor("model.layers.0.self_attn.v_proj.qweight")  # 6-bit gs64 in layer 0 of bpw-2.2
        assert qw.dtype == torch.uint32 and qw.shape == (dims.kv_heads * dims.head_dim, dims.hidden * 6 // 32)
        assert f.get_tensor("model.layers.0.self_attn.v_proj.scales").dtype == torch.bfloat16
    idx = json.load(open(d / "model.safetensors.index.json"))
    assert set(idx["weight_map"].values()) == {"model.safetensors"} and idx["metadata"]["total_size"] > 0


def test_load_model_applies_layer_mix_strategy(tiny_ckpt):
    d, dims, strat = tiny_ckpt
    model, cfg = utils.load_model(d, device="cpu")
    plan = {(i, p): (b, g) for (i, p, n, k, b, g) in W.layer_plan(dims, strat)}
    seen = 0
    for name, m in model.named_modules():
        if isinstance(m, QuantizedLinear):
            i = int(name.split(".")[2])
            p = name.split(".")[-1]
            assert (m.bits, m.group_size) == plan[(i, p)], name
            assert m.qweight.shape == (m.output_dims, m.input_dims // 32 * m.bits)
            assert m.scales.shape == (m.output_dims, m.input_dims // m.group_size)
            assert m.scales.dtype == torch.bfloat16 and m.qweight.dtype == torch.uint32
            assert int(m.qweight.view(torch.int32).abs().sum()) != 0  # weights really loaded
            seen += 1
    assert seen == 7 * dims.layers
    bits_seen = {m.bits for m in model.modules() if isinstance(m, QuantizedLinear)}
    assert bits_seen == {2, 3, 6}
    assert model.model.embed_tokens.weight.dtype == torch.bfloat16


def test_load_model_without_strategy_uses_config_quantization(tmp_path):
    dims = W.MODELS["tiny-qwen2"]
    utils.write_synthetic_checkpoint(tmp_path, dims, None, seed=1, default_bits=4, default_gs=128)
    assert not (tmp_path / "quant_strategy.json").exists()
    model, cfg = utils.load_model(tmp_path, device="cpu")
    for m in model.modules():
        if isinstance(m, QuantizedLinear):
            assert (m.bits, m.group_size) == (4, 128)
    attn = model.model.layers[0].self_attn
    assert attn.q_proj.bias is not None and attn.k_proj.bias is not None and attn.o_proj.bias is None  # qqwen2.py:44-47
    assert hasattr(model, "lm_head")  # untied


def test_load_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        utils.load_model(tmp_path, device="cpu")  # no config.json
    json.dump({"model_type": "llama"}, open(tmp_path / "config.json", "w"))
    with pytest.raises(FileNotFoundError):
        utils.load_model(tmp_path, device="cpu")  # no safetensors
    dims = W.MODELS["tiny-llama"]
    utils.write_synthetic_checkpoint(tmp_path, dims, None, seed=1)
    cfg = json.load(open(tmp_path / "config.json"))
    cfg["model_type"] = "mamba"
    json.dump(cfg, open(tmp_path / "config.json", "w"))
    with pytest.raises(ValueError):
        utils.load_model(tmp_path, device="cpu")
    cfg["model_type"] = "llama"
    cfg["quantization"] = {"group_size": 48, "bits": 4}
    json.dump(cfg, open(tmp_path / "config.json", "w"))
    with pytest.raises(AssertionError):
        utils.load_model(tmp_path, device="cpu")


def test_strategy_missing_entry_raises(tiny_ckpt):
    d, dims, strat = tiny_ckpt
    import copy

    bad = copy.deepcopy(strat["measurement"])
    del bad["model.layers.1"]["down_proj"]
    model, _ = utils.load_model(d, device="cpu")
    with pytest.raises(KeyError):
        QuantizedLinear.reinit_module(model, 128, 2, strategy=bad)


def test_make_shards_counts():
    w = {f"w{i}": torch.zeros(1 << 18, dtype=torch.float32) for i in range(8)}  # 1 MiB each
    assert len(utils.make_shards(w, max_file_size_gb=5)) == 1
    shards = utils.make_shards(w, max_file_size_gb=0)  # 0 GiB cap -> one tensor per shard, no empty one
    assert [len(s) for s in shards] == [1] * 8
    assert utils.make_shards({}) == [{}]


def test_multi_shard_layout_round_trip(tmp_path):
    """n > 1 files: names model-0000i-of-0000n.safetensors, {"format": "mlx"} metadata, a name-sorted index whose
    total_size is the byte count, no empty file, every tensor back bit for bit; config written sorted without
    `_name_or_path` and without touching the caller's dict (gbx_lm/utils.py:967-988,1055-1127)."""
    from safetensors import safe_open

    gen = torch.Generator().manual_seed(0)
    w = {f"model.layers.{i}.w": torch.randn((64, 64), generator=gen).to(torch.bfloat16) for i in (2, 0, 1)}
    w["model.layers.0.qweight"] = torch.arange(256, dtype=torch.int32).view(torch.uint32).reshape(16, 16)
    keep = dict(w)
    utils.save_weights(tmp_path, w, max_file_size_gb=0, donate_weights=True)
    assert w == {}                                               # donated
    files = sorted(p.name for p in tmp_path.glob("*.safetensors"))
    assert files == [f"model-{i:05d}-of-00004.safetensors" for i in range(1, 5)]
    idx = json.load(open(tmp_path / "model.safetensors.index.json"))
    assert list(idx["weight_map"]) == sorted(keep) and set(idx["weight_map"].values()) == set(files)
    assert idx["metadata"]["total_size"] == sum(t.numel() * t.element_size() for t in keep.values())
    for name, fname in idx["weight_map"].items():
        with safe_open(str(tmp_path / fname), framework="pt") as f:
            assert f.metadata() == {"format": "mlx"} and list(f.keys()) == [name]
            got = f.get_tensor(name)
        a, b = (got.view(torch.int32), keep[name].view(torch.int32)) if got.dtype == torch.uint32 else (got, keep[name])
        assert torch.equal(a, b)
    cfg = {"z": 1, "_name_or_path": "x", "a": {"k": 2}}
    utils.save_config(cfg, tmp_path / "config.json")
    assert "_name_or_path" in cfg
    assert list(json.load(open(tmp_path / "config.json"))) == ["a", "z"]


def test_workload_bytes_match_baseline_md():
    """8B gate_proj 4-bit gs64 M=1 = 33,067,008 B; k_proj = 2,369,536 B;
    70B down_proj = 132,194,304 B; per-token body bytes 8B uniform 4-bit = 3.931 GB."""
    assert W.qmm_bytes(1, 14336, 4096, 4, 64) == 33067008
    assert W.qmm_bytes(1, 1024, 4096, 4, 64) == 2369536
    assert W.qmm_bytes(1, 8192, 28672, 4, 64) == 132194304
    plan = W.layer_plan(W.MODELS["llama-3-8b"], None, 4, 64)
    assert abs(sum(W.qmm_bytes(1, n, k, b, g) for (_, _, n, k, b, g) in plan) / 1e9 - 3.931) < 0.002
    bp = W.stored_bpw(W.layer_plan(W.MODELS["llama-3-8b"], W.strategy_bpw40(32)))
    assert 3.9 < bp < 4.0


@pytest.mark.parametrize("bits,gs", [(4, 64), (2, 128), (3, 64), (6, 64), (8, 32)])
def test_tp_shards_reassemble_and_sum(bits, gs):
    """Column shards concatenate to the full output; row shards' partial outputs SUM to it (oracle arithmetic)."""
    N, K, world = 32, 512, 4
    L = packing.synth_layer(N, K, bits, gs, seed=5, with_bias=True)
    x = A.synth_x(2, K, seed=6)
    to_np = lambda t: t.view(torch.int16).numpy().view(np.uint16)
    qw_np = L["qweight"].view(torch.int32).numpy().view(np.uint32)
    full = A.quantized_matmul(x, qw_np, to_np(L["scales"]), to_np(L["zeros"]), gs, bits, "f32" if False else "bf16", "f64")
    # column parallel
    outs = []
    for r in range(world):
        sh = {k: tpmod.shard_tensor(f"model.layers.0.mlp.up_proj.{k}", L[k], bits, gs, r, world) for k in ("qweight", "scales", "zeros", "bias")}
        assert sh["qweight"].shape == (N // world, K * bits // 32) and sh["bias"].shape == (N // world,)
        outs.append(A.quantized_matmul(x, sh["qweight"].view(torch.int32).numpy().view(np.uint32), to_np(sh["scales"]), to_np(sh["zeros"]), gs, bits, "bf16", "f64"))
    assert (np.concatenate(outs, axis=1) == full).all()
    # row parallel: fp64 partials sum exactly to the fp64 truth
    xs = A.bf16_bits_to_f32(x)
    tot = np.zeros((2, N))
    for r in range(world):
        sh = {k: tpmod.shard_tensor(f"model.layers.0.mlp.down_proj.{k}", L[k], bits, gs, r, world) for k in ("qweight", "scales", "zeros", "bias")}
        kk = K // world
        assert sh["qweight"].shape == (N, kk * bits // 32) and sh["scales"].shape == (N, kk // gs)
        assert (sh["bias"] == (L["bias"] if r == 0 else torch.zeros_like(L["bias"]))).all()
        q = A.unpack_codes(sh["qweight"].view(torch.int32).numpy().view(np.uint32), bits).astype(np.float64)
        Wd = np.repeat(A.bf16_bits_to_f32(to_np(sh["scales"])), gs, 1).astype(np.float64) * q + np.repeat(A.bf16_bits_to_f32(to_np(sh["zeros"])), gs, 1)
        tot += xs[:, r * kk:(r + 1) * kk].astype(np.float64) @ Wd.T
    assert np.abs(A._round_to(tot, "bf16") - full).max() == 0


def test_tp_illegal_row_split():
    with pytest.raises(ValueError):
        tpmod.check_row_split(4096, 3, 64, 5)
    with pytest.raises(ValueError):
        tpmod.check_row_split(2048, 4, 128, 32)  # 64 codes per rank < one group of 128
    tpmod.check_row_split(27648, 4, 128, 8)  # Qwen2.5-32B down_proj at tp8: 3456 = 27 groups -> legal
    tpmod.check_row_split(28672, 4, 64, 8)   # Llama-3-70B down_proj at tp8: 3584 = 56 groups -> legal


def test_load_model_converts_a_gba_checkpoint(tiny_ckpt, tmp_path):
    """`load_model(..., is_conversion=True)` on an original GBA layout (K-major qweight / scales / zeros, subtractive fp16
    zeros, a q_perm vector; gbx_lm/utils.py:828-843,864-873) gives the same modules as loading the converted checkpoint."""
    import shutil

    from safetensors.torch import load_file

    d, dims, strat = tiny_ckpt
    mlx_w = load_file(str(d / "model.safetensors"))
    gba = {}
    for k, v in mlx_w.items():
        if k.endswith(".qweight"):
            gba[k] = v.view(torch.int32).t().contiguous()
        elif k.endswith(".scales"):
            gba[k] = v.t().contiguous().to(torch.float16)
        elif k.endswith(".zeros"):
            gba[k] = (-v.float()).t().contiguous().to(torch.float16)
        else:
            gba[k] = v.to(torch.float16) if v.is_floating_point() else v
    perm_key = "model.layers.0.mlp.down_proj.q_perm"
    gba[perm_key] = torch.arange(dims.inter, dtype=torch.int16)
    utils.save_weights(tmp_path, gba)
    for f in ("config.json", "quant_strategy.json"):
        shutil.copy(d / f, tmp_path / f)
    ref, _ = utils.load_model(d, device="cpu")
    got, _ = utils.load_model(tmp_path, device="cpu", is_conversion=True)
    n = 0
    for (name, a), (_, b) in zip(ref.named_modules(), got.named_modules()):
        if isinstance(a, QuantizedLinear):
            assert (a.bits, a.group_size) == (b.bits, b.group_size)
            assert torch.equal(a.qweight.view(torch.int32), b.qweight.view(torch.int32)), name
            # synthetic scales / zeros are bf16 values; fp16 holds them exactly unless they are fp16-subnormal
            assert torch.allclose(a.scales.float(), b.scales.float(), rtol=0, atol=1e-7) and b.scales.dtype == torch.bfloat16
            assert torch.allclose(a.zeros.float(), b.zeros.float(), rtol=0, atol=1e-7)
            n += 1
    assert n == 7 * dims.layers
    assert dict(got.named_buffers())[perm_key].shape == (1, 1, dims.inter)
    assert torch.equal(dict(ref.named_parameters())["model.norm.weight"], dict(got.named_parameters())["model.norm.weight"])


def test_expand_statistics_of_an_mlx_oriented_double_quant_checkpoint():
    """`use_double_quantization and not is_conversion` (gbx_lm/utils.py:864-868): the statistics are expanded in place
    of the codes, nothing is transposed or negated; strategy_params finds the group size the expansion needs."""
    from gbx_lm_b200 import gba_convert as G
    from gbx_lm_b200.quantized_linear import strategy_params
    from tests.test_gba_convert import _gba_layer

    n, k, bits, gs = 64, 256, 4, 64
    layer, q, s_bf, z_bf = _gba_layer(n, k, bits, gs, seed=1, double_quant=True)
    p = "model.layers.3.mlp.up_proj."
    w = {p + leaf: t for leaf, t in layer.items()}
    w[p + "qweight"] = w[p + "qweight"].t().contiguous()          # already [N, K*bits/32]
    strategy = {"model.layers.3": {"up_proj": {"bits": [bits], "group_size": {str(bits): gs}}}}
    assert strategy_params(p[:-1], strategy, 2, 128) == (bits, gs) and strategy_params(p[:-1], None, 2, 128) == (2, 128)
    with pytest.raises(KeyError):
        strategy_params("model.layers.3.mlp.down_proj", strategy, 2, 128)
    out = G.expand_statistics(w, lambda m: strategy_params(m, strategy, 2, 128)[1])
    assert set(out) == {p + "qweight", p + "scales", p + "zeros"}
    assert torch.equal(out[p + "qweight"], w[p + "qweight"])
    assert np.array_equal(out[p + "scales"].float().numpy(), s_bf.T) and np.array_equal(out[p + "zeros"].float().numpy(), z_bf.T)


# ---------------------------------------------------------------------------------------------------------------
# the reference's own two tests of this area (/root/reference/tests/test_utils.py:27-54), restated for this loader
# ---------------------------------------------------------------------------------------------------------------
def test_load_eager_and_lazy_agree(tiny_ckpt):
    """test_utils.py:27-36 (`test_load`): the same qweight whether the model is loaded eagerly or lazily."""
    d, _, _ = tiny_ckpt
    model, _ = utils.load(d, device="cpu")
    model_lazy, _ = utils.load(d, lazy=True, device="cpu")
    p1 = model.model.layers[0].mlp.up_proj.qweight
    p2 = model_lazy.model.layers[0].mlp.up_proj.qweight
    assert torch.equal(p1.view(torch.int32), p2.view(torch.int32))


def test_make_shards_count_follows_the_byte_total():
    """test_utils.py:38-54 (`test_make_shards`): with a 1 GiB cap the number of files is the model's size in GiB or one
    more.  Same dimensions as the reference's test (hidden 2048, 32 layers, intermediate 4096, vocabulary 30 000),
    dense fp32 tensors on the meta device (only sizes matter)."""
    h, layers, inter, vocab = 2048, 32, 4096, 30_000
    meta = lambda *shape: torch.empty(shape, dtype=torch.float32, device="meta")  # noqa: E731
    w = {"model.embed_tokens.weight": meta(vocab, h), "lm_head.weight": meta(vocab, h), "model.norm.weight": meta(h)}
    for i in range(layers):
        p = f"model.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            w[p + f"self_attn.{n}.weight"] = meta(h, h)
        w[p + "mlp.gate_proj.weight"] = meta(inter, h)
        w[p + "mlp.up_proj.weight"] = meta(inter, h)
        w[p + "mlp.down_proj.weight"] = meta(h, inter)
        w[p + "input_layernorm.weight"] = meta(h)
        w[p + "post_attention_layernorm.weight"] = meta(h)
    gb = sum(t.numel() * t.element_size() for t in w.values()) // 2 ** 30
    shards = utils.make_shards(w, 1)
    assert gb >= 5 and gb <= len(shards) <= gb + 1
    assert sum(len(s) for s in shards) == len(w) and all(s for s in shards)
    cap = 1 << 30
    assert all(sum(t.numel() * t.element_size() for t in s.values()) <= cap for s in shards)
