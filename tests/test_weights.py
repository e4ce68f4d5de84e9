"""Weight pipeline (llm-inference-engine_b200/weights.py): Hugging Face state dict / the reference's per-tensor
.bin directory (src/weights/layer_weights.cpp:50-80, llama_weights.cpp:49-75) -> fused [N,K] tensors -> packed engine format."""
import importlib
import json
import os

import numpy as np
import pytest

from oracle import oracle
from util import assert_close, rounded, to_np

W = importlib.import_module("llm-inference-engine_b200.weights")
SHAPE = W.shape_of(hidden=256, head_num=2, kv_head_num=1, head_size=128, inter=384, layers=2, vocab=300)


def fake_hf_state_dict(shape, seed=0, bias=False):
    """A random state dict with Hugging Face's LlamaForCausalLM names and [out, in] orientation."""
    rng = np.random.default_rng(seed)
    h, H, Hkv, d, I, V = (shape[k] for k in ("hidden", "head_num", "kv_head_num", "head_size", "inter", "vocab"))
    sd = {"model.embed_tokens.weight": rng.standard_normal((V, h)), "model.norm.weight": 1 + 0.1 * rng.standard_normal(h),
          "lm_head.weight": rng.standard_normal((V, h)) / 16}
    for l in range(shape["layers"]):
        p = f"model.layers.{l}."
        sd[p + "input_layernorm.weight"] = 1 + 0.1 * rng.standard_normal(h)
        sd[p + "post_attention_layernorm.weight"] = 1 + 0.1 * rng.standard_normal(h)
        sd[p + "self_attn.q_proj.weight"] = rng.standard_normal((H * d, h)) / np.sqrt(h)
        sd[p + "self_attn.k_proj.weight"] = rng.standard_normal((Hkv * d, h)) / np.sqrt(h)
        sd[p + "self_attn.v_proj.weight"] = rng.standard_normal((Hkv * d, h)) / np.sqrt(h)
        sd[p + "self_attn.o_proj.weight"] = rng.standard_normal((h, H * d)) / np.sqrt(H * d)
        sd[p + "mlp.gate_proj.weight"] = rng.standard_normal((I, h)) / np.sqrt(h)
        sd[p + "mlp.up_proj.weight"] = rng.standard_normal((I, h)) / np.sqrt(h)
        sd[p + "mlp.down_proj.weight"] = rng.standard_normal((h, I)) / np.sqrt(I)
        if bias:
            for n, rows in (("q", H * d), ("k", Hkv * d), ("v", Hkv * d), ("o", h)):
                sd[p + f"self_attn.{n}_proj.bias"] = 0.05 * rng.standard_normal(rows)
    return {k: v.astype(np.float32) for k, v in sd.items()}


def test_fusing_follows_the_reference_tensor_layout():
    sd = fake_hf_state_dict(SHAPE, bias=True)
    f = W.fuse_hf_state_dict(sd, SHAPE)
    H, Hkv, d, I = SHAPE["head_num"], SHAPE["kv_head_num"], SHAPE["head_size"], SHAPE["inter"]
    w = f["layers"][1]
    assert w["wqkv"].shape == ((H + 2 * Hkv) * d, SHAPE["hidden"]) and w["wgu"].shape == (2 * I, SHAPE["hidden"])
    # q rows, then k rows, then v rows; gate rows then up rows (the layout qkv_bias_and_rope.cu / silu_and_mul.cu index)
    assert np.array_equal(w["wqkv"][:H * d], sd["model.layers.1.self_attn.q_proj.weight"])
    assert np.array_equal(w["wqkv"][H * d:(H + Hkv) * d], sd["model.layers.1.self_attn.k_proj.weight"])
    assert np.array_equal(w["wqkv"][(H + Hkv) * d:], sd["model.layers.1.self_attn.v_proj.weight"])
    assert np.array_equal(w["wgu"][:I], sd["model.layers.1.mlp.gate_proj.weight"]) and np.array_equal(w["wgu"][I:], sd["model.layers.1.mlp.up_proj.weight"])
    assert w["bqkv"].shape == ((H + 2 * Hkv) * d,) and w["bo"].shape == (SHAPE["hidden"],)


def test_reference_bin_directory_round_trip(tmp_path):
    f = W.fuse_hf_state_dict(fake_hf_state_dict(SHAPE, seed=3, bias=True), SHAPE)
    prefix = str(tmp_path / "llama") + "/"
    W.export_reference_bins(f, prefix)
    # file names and sizes the reference's loader expects (layer_weights.cpp:50-80, llama_weights.cpp:49-75)
    names = sorted(os.listdir(prefix))
    assert "model.embed_tokens.weight.bin" in names and "model.norm.weight.bin" in names and "lm_head.weight.bin" in names
    assert "model.layers.1.self_attn.qkv.weight.bin" in names and "model.layers.0.mlp.gate_up_proj.weight.bin" in names
    n, k = W.linear_shapes(SHAPE)["wqkv"]
    assert os.path.getsize(prefix + "model.layers.0.self_attn.qkv.weight.bin") == 4 * n * k
    # the linears are stored as [K, N] memory: launchLinearGemm computes X . Wmem[K,N]
    mem = np.fromfile(prefix + "model.layers.0.self_attn.o_proj.weight.bin", dtype="<f4").reshape(SHAPE["head_num"] * SHAPE["head_size"], SHAPE["hidden"])
    x = np.random.default_rng(0).standard_normal((3, mem.shape[0])).astype(np.float32)
    assert np.allclose(x @ mem, oracle.linear(x, f["layers"][0]["wo"], "nk"), atol=1e-4)
    # lm_head keeps the [V, h] layout (weights.py export_reference_bins: the reference holds no live code that reads it): row v of the file is
    # the output channel of token v, un-transposed
    lm = np.fromfile(prefix + "lm_head.weight.bin", dtype="<f4").reshape(SHAPE["vocab"], SHAPE["hidden"])
    assert np.array_equal(lm, f["lm_head"].astype(np.float32))
    g = W.load_reference_bins(prefix, SHAPE)
    for l in range(SHAPE["layers"]):
        for key in ("g1", "g2", "wqkv", "wo", "wgu", "wd", "bqkv", "bo"):
            assert np.array_equal(g["layers"][l][key], f["layers"][l][key]), key
    for key in ("embed", "final_gamma", "lm_head"):
        assert np.array_equal(g[key], f[key])
    # a truncated file is an error, not a silent partial load
    with open(prefix + "model.norm.weight.bin", "ab") as fh:
        fh.truncate(8)
    with pytest.raises(ValueError):
        W.load_reference_bins(prefix, SHAPE)


def test_packed_directory_round_trip_on_cpu(tmp_path):
    import torch

    f = W.fuse_hf_state_dict(fake_hf_state_dict(SHAPE, seed=4), SHAPE)
    for tp in (1, 2):
        shape = dict(SHAPE, head_num=2, kv_head_num=2) if tp == 2 else SHAPE
        ff = W.fuse_hf_state_dict(fake_hf_state_dict(shape, seed=4), shape)
        for r in range(tp):
            packed = W.pack_model(ff, shape, "bf16", "dense", torch.device("cpu"), tp, r)
            m = W.save_packed(packed, shape, str(tmp_path / f"tp{tp}"), "bf16", "bf16", tp, r)
            assert m["layout"] == "NK" and m["tp"] == tp
            back, m2 = W.load_packed(str(tmp_path / f"tp{tp}"), torch.device("cpu"), r)
            assert json.dumps(m2["shape"], sort_keys=True) == json.dumps(shape, sort_keys=True)
            for l in range(shape["layers"]):
                for key in ("g1", "qkv", "o", "g2", "gate_up", "down"):
                    assert torch.equal(back["layers"][l][key], packed["layers"][l][key]), (tp, r, l, key)
            assert torch.equal(back["lm_head"], packed["lm_head"])
        if tp == 2:  # the two ranks' QKV rows / O columns tile the un-sharded tensor
            p0, _ = W.load_packed(str(tmp_path / "tp2"), torch.device("cpu"), 0)
            p1, _ = W.load_packed(str(tmp_path / "tp2"), torch.device("cpu"), 1)
            full = torch.from_numpy(ff["layers"][0]["wo"]).to(torch.bfloat16)
            assert torch.equal(torch.cat([p0["layers"][0]["o"], p1["layers"][0]["o"]], dim=1), full)
            assert p0["layers"][0]["qkv"].shape[0] * 2 == ff["layers"][0]["wqkv"].shape[0]


@pytest.mark.gpu
@pytest.mark.parametrize("wformat", ["bf16", "fp8", "int4"])
def test_converted_checkpoint_runs_on_the_engine(tmp_path, wformat):
    """reference .bin directory -> load -> pack (bf16 / FP8 / INT4) -> save -> load -> engine step, against the oracle on the same
    fused tensors (for FP8 / INT4: on the exactly dequantised weights)."""
    import torch

    from test_decoder_engine import make_inputs, rel_fro
    from util import to_dev

    mod = importlib.import_module("llm-inference-engine_b200")
    dev = torch.device("cuda")
    f = W.fuse_hf_state_dict(fake_hf_state_dict(SHAPE, seed=8), SHAPE)
    prefix = str(tmp_path / "ref") + "/"
    W.export_reference_bins(f, prefix)
    fused = W.load_reference_bins(prefix, SHAPE)
    packed = W.pack_model(fused, SHAPE, "bf16", "dense" if wformat == "bf16" else wformat, dev)
    W.save_packed(packed, SHAPE, str(tmp_path / "packed"), "bf16", wformat)
    packed2, manifest = W.load_packed(str(tmp_path / "packed"), dev)
    cfg = dict(hidden=SHAPE["hidden"], head_num=SHAPE["head_num"], kv_head_num=SHAPE["kv_head_num"], head_size=SHAPE["head_size"],
               inter=SHAPE["inter"], layers=SHAPE["layers"], max_seq=64, eps=1e-5, base=10000.0)
    dec, keep = W.build_decoder(packed2, SHAPE, dev, "bf16", wformat, cfg["max_seq"], 2)
    B, step = 2, 11
    x, kc, vc = make_inputs(cfg, B, step, 5)
    x, kc, vc = rounded(x, "bf16"), rounded(kc, "bf16"), rounded(vc, "bf16")
    xd, kcd, vcd = to_dev(x, "bf16"), to_dev(kc, "bf16"), to_dev(vc, "bf16")
    dec.step(xd, kcd, vcd, step)
    torch.cuda.synchronize()
    # oracle on the weights the device actually holds
    ocfg = dict(head_num=cfg["head_num"], kv_head_num=cfg["kv_head_num"], head_size=cfg["head_size"], inter=cfg["inter"], eps=cfg["eps"],
                rot_dim=cfg["head_size"], base=cfg["base"])
    ref = x.copy()
    for l, w in enumerate(fused["layers"]):
        ww = {}
        for key, pk in (("wqkv", "qkv"), ("wo", "o"), ("wgu", "gate_up"), ("wd", "down")):
            t = packed2["layers"][l][pk]
            if wformat == "bf16":
                ww[key] = to_np(t)
            elif wformat == "fp8":
                ww[key] = oracle.dequantize_fp8(to_np(t[0]), to_np(t[1]))
            else:
                ww[key] = oracle.dequantize_int4(to_np(t[0]), to_np(t[1]), to_np(t[2]), 128)
        ww.update(g1=rounded(w["g1"], "bf16"), g2=rounded(w["g2"], "bf16"), bqkv=None, bo=None)
        oracle.decoder_layer(ref, ww, kc, vc, ocfg, step, l)
    got = to_np(xd)
    assert rel_fro(got, ref) <= 1e-2, f"{wformat}: {rel_fro(got, ref):.3e}"


def test_converter_cli_on_cpu(tmp_path):
    """scripts/convert_weights.py end to end without a GPU: a Hugging Face style checkpoint file -> the reference's .bin directory ->
    the packed bf16 engine format for two tensor-parallel ranks."""
    import subprocess
    import sys

    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shape = dict(SHAPE, head_num=2, kv_head_num=2)
    sd = {k: torch.from_numpy(v) for k, v in fake_hf_state_dict(shape, seed=21).items()}
    ckpt = tmp_path / "hf"
    ckpt.mkdir()
    torch.save(sd, str(ckpt / "pytorch_model.bin"))
    spec = ",".join(str(shape[k]) for k in ("hidden", "head_num", "kv_head_num", "head_size", "inter", "layers", "vocab"))
    ref_prefix = str(tmp_path / "refbins") + "/"
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(root, "scripts", "convert_weights.py"), "--hf", str(ckpt), "--shape", spec,
                        "--export-ref-bins", ref_prefix], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, timeout=300)
    assert p.returncode == 0, p.stdout.decode()[-2000:]
    assert os.path.exists(ref_prefix + "model.layers.1.mlp.down_proj.weight.bin")
    out = str(tmp_path / "packed")
    p = subprocess.run([sys.executable, os.path.join(root, "scripts", "convert_weights.py"), "--ref-bins", ref_prefix, "--shape", spec, "--out", out,
                        "--tp", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, timeout=300)
    assert p.returncode == 0, p.stdout.decode()[-2000:]
    fused = W.fuse_hf_state_dict(fake_hf_state_dict(shape, seed=21), shape)
    for r in range(2):
        packed, manifest = W.load_packed(out, torch.device("cpu"), r)
        assert manifest["tp"] == 2 and manifest["rank"] == r and manifest["wformat"] == "bf16"
        want = torch.from_numpy(fused["layers"][0]["wd"]).to(torch.bfloat16)
        I2 = shape["inter"] // 2
        assert torch.equal(packed["layers"][0]["down"], want[:, r * I2:(r + 1) * I2].contiguous())
    # FP8 / INT4 packing uses the library's device quantisers: refused (not silently done on the CPU) without a GPU
    p = subprocess.run([sys.executable, os.path.join(root, "scripts", "convert_weights.py"), "--ref-bins", ref_prefix, "--shape", spec, "--out", out + "8",
                        "--wformat", "fp8"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, timeout=300)
    assert p.returncode != 0 and b"needs a GPU" in p.stdout
