"""Weight pipeline: from a checkpoint to what the decode engine streams.

Sources
  * a Hugging Face Llama state dict (`model.layers.N.self_attn.{q,k,v,o}_proj.weight`, `mlp.{gate,up,down}_proj.weight`, ... all
    `[out, in]`), fused here into the reference's tensors: QKV = rows of q, k, v; gate_up = gate rows then up rows;
  * the reference's own on-disk format (src/weights/layer_weights.cpp:50-80, llama_weights.cpp:49-75, src/utils/weight_utils.cu:189-224):
    one raw little-endian fp32 file per tensor under a common prefix,
        model.embed_tokens.weight.bin [V,h]   model.norm.weight.bin [h]   lm_head.weight.bin [V,h]
        model.layers.N.input_layernorm.weight.bin [h]        model.layers.N.post_attention_layernorm.weight.bin [h]
        model.layers.N.self_attn.qkv.weight.bin              model.layers.N.self_attn.o_proj.weight.bin
        model.layers.N.mlp.gate_up_proj.weight.bin           model.layers.N.mlp.down_proj.weight.bin
    whose linears the reference's kernels read as row-major [K, N] memory (`launchLinearGemm` = X . Wmem[K,N]): the
    transpose of the Hugging Face tensor.  RoPE pairs (i, i + d/2) in both (src/kernels/rope.cu:29-42), so no row permutation.

Products
  * `fused` : dict(layers=[dict(g1, wqkv, bqkv, wo, bo, g2, wgu, wd)], embed, final_gamma, lm_head), fp32 numpy, linears in the engine's
    [N, K] orientation -- the same dict the parity tests feed to their CPU checker, so a converted checkpoint needs no second loader;
  * `export_reference_bins`: the reference's directory from `fused` (what its `loadWeights(path)` reads);
  * `save_packed` / `load_packed`: the engine's own format -- per rank, per tensor raw files of the PACKED bytes (bf16 / fp16 / fp32 dense,
    FP8-e4m3 + per-row scale, INT4-g128 + scales + zero points, all [N, K]) plus a JSON manifest: quantisation and tensor-parallel
    sharding happen once, offline; loading is a read + H2D copy;
  * `build_decoder`: a ready `Decoder` (+ embedding, final gamma, LM head on the device) from `fused` or from a packed directory.
Quantisers are the library's (`b200_quantize_fp8` / `b200_quantize_int4`), so this module needs a GPU only
for FP8 / INT4 packing and for `build_decoder`; everything else is numpy.
"""
import importlib
import json
import os

import numpy as np

REF_LAYER_FILES = {  # suffix -> (key in a fused layer dict, is a linear stored [K,N])
    ".input_layernorm.weight.bin": ("g1", False),
    ".post_attention_layernorm.weight.bin": ("g2", False),
    ".self_attn.qkv.weight.bin": ("wqkv", True),
    ".self_attn.o_proj.weight.bin": ("wo", True),
    ".mlp.gate_up_proj.weight.bin": ("wgu", True),
    ".mlp.down_proj.weight.bin": ("wd", True),
}
REF_BIAS_FILES = {".attention.wqkv.bias.bin": "bqkv", ".attention.wo.bias.bin": "bo"}  # layer_weights.cpp:70-77 (attention_bias models)


def shape_of(hidden, head_num, kv_head_num, head_size, inter, layers, vocab):
    return dict(hidden=hidden, head_num=head_num, kv_head_num=kv_head_num, head_size=head_size, inter=inter, layers=layers, vocab=vocab)


LLAMA2_7B = shape_of(4096, 32, 32, 128, 11008, 32, 32000)  # src/models/llama/llama_config.json:2-8


def _np(t):
    if hasattr(t, "detach"):
        t = t.detach().to("cpu").float().numpy()
    return np.ascontiguousarray(np.asarray(t, dtype=np.float32))


def linear_shapes(shape):
    """[N, K] of the four linears of a layer."""
    h, H, Hkv, d, I = shape["hidden"], shape["head_num"], shape["kv_head_num"], shape["head_size"], shape["inter"]
    return dict(wqkv=((H + 2 * Hkv) * d, h), wo=(h, H * d), wgu=(2 * I, h), wd=(h, I))


def fuse_hf_state_dict(sd, shape):
    """Hugging Face LlamaForCausalLM state dict -> `fused`."""
    L = shape["layers"]
    want = linear_shapes(shape)
    layers = []
    for l in range(L):
        p = f"model.layers.{l}."
        w = dict(
            g1=_np(sd[p + "input_layernorm.weight"]),
            wqkv=np.concatenate([_np(sd[p + f"self_attn.{n}_proj.weight"]) for n in ("q", "k", "v")], axis=0),
            bqkv=(np.concatenate([_np(sd[p + f"self_attn.{n}_proj.bias"]) for n in ("q", "k", "v")]) if p + "self_attn.q_proj.bias" in sd else None),
            wo=_np(sd[p + "self_attn.o_proj.weight"]),
            bo=_np(sd[p + "self_attn.o_proj.bias"]) if p + "self_attn.o_proj.bias" in sd else None,
            g2=_np(sd[p + "post_attention_layernorm.weight"]),
            wgu=np.concatenate([_np(sd[p + "mlp.gate_proj.weight"]), _np(sd[p + "mlp.up_proj.weight"])], axis=0),
            wd=_np(sd[p + "mlp.down_proj.weight"]))
        for k, s in want.items():
            if w[k].shape != s:
                raise ValueError(f"layer {l} {k}: shape {w[k].shape}, expected {s}")
        layers.append(w)
    lm = sd["lm_head.weight"] if "lm_head.weight" in sd else sd["model.embed_tokens.weight"]  # tied embeddings
    return dict(layers=layers, embed=_np(sd["model.embed_tokens.weight"]), final_gamma=_np(sd["model.norm.weight"]), lm_head=_np(lm))


def export_reference_bins(fused, prefix):
    """Write the reference's per-tensor fp32 files: `prefix` + name (the reference concatenates, so end it with '/' for a directory)."""
    d = os.path.dirname(prefix)
    if d:
        os.makedirs(d, exist_ok=True)

    def put(name, a):
        np.ascontiguousarray(a, dtype="<f4").tofile(prefix + name)

    put("model.embed_tokens.weight.bin", fused["embed"])
    put("model.norm.weight.bin", fused["final_gamma"])
    # lm_head is written [V, h] -- the layout Hugging Face stores, the one the reference's (dead) model class declares for it
    # (src/models/llama/llama.cpp:271-277 shape {V, h}, trans_b = true) and the engine's NK packing -- NOT transposed like the layer linears,
    # whose files are the [K, N] memory the reference's live launchLinearGemm path actually reads.  The reference never loads or
    # multiplies an LM head in code that runs (src/models is not compiled), so there is no GPU behaviour to follow here; the choice is pinned by
    # tests/test_weights.py::test_reference_bin_directory_round_trip.
    put("lm_head.weight.bin", fused["lm_head"])
    for l, w in enumerate(fused["layers"]):
        base = f"model.layers.{l}"
        for suffix, (key, is_linear) in REF_LAYER_FILES.items():
            put(base + suffix, w[key].T if is_linear else w[key])  # linears as [K, N] memory
        for suffix, key in REF_BIAS_FILES.items():
            if w.get(key) is not None:
                put(base + suffix, w[key])


def load_reference_bins(prefix, shape):
    """The reference's directory -> `fused` (sizes are checked: a file of the wrong length is an error, as in loadWeightFromBin)."""
    h, V = shape["hidden"], shape["vocab"]
    want = linear_shapes(shape)

    def get(name, shp):
        path = prefix + name
        a = np.fromfile(path, dtype="<f4")
        if a.size != int(np.prod(shp)):
            raise ValueError(f"{path}: {a.size} floats, expected {int(np.prod(shp))} for shape {shp}")
        return a.reshape(shp)

    layers = []
    for l in range(shape["layers"]):
        base = f"model.layers.{l}"
        w = dict(bqkv=None, bo=None)
        for suffix, (key, is_linear) in REF_LAYER_FILES.items():
            if is_linear:
                n, k = want[key]
                w[key] = np.ascontiguousarray(get(base + suffix, (k, n)).T)
            else:
                w[key] = get(base + suffix, (h,))
        for suffix, key in REF_BIAS_FILES.items():
            if os.path.exists(prefix + base + suffix):
                w[key] = get(base + suffix, (want["wqkv"][0],) if key == "bqkv" else (h,))
        layers.append(w)
    return dict(layers=layers, embed=get("model.embed_tokens.weight.bin", (V, h)), final_gamma=get("model.norm.weight.bin", (h,)),
                lm_head=get("lm_head.weight.bin", (V, h)))


# ------------------------------------------------------------------ engine format
_TORCH_DT = {"bf16": "bfloat16", "f16": "float16", "f32": "float32"}


def _mod():
    return importlib.import_module(__package__ or "llm-inference-engine_b200")


def _tp():
    return importlib.import_module((__package__ or "llm-inference-engine_b200") + ".tp")


def pack_linear(w_nk, dtype, wformat, device, group=128):
    """fp32 [N, K] -> what a b200_linear_weight_t points at: a tensor (dense) or (q, scales[, zeros]) (FP8 / INT4), on `device`."""
    import torch

    mod = _mod()
    t = torch.from_numpy(np.ascontiguousarray(w_nk)).to(device=device, dtype=getattr(torch, _TORCH_DT[dtype]))
    if wformat == "fp8":
        return mod.quantize_fp8(t)
    if wformat == "int4":
        return mod.quantize_int4(t, group)
    return t


def pack_model(fused, shape, dtype, wformat, device, tp=1, rank=0, group=128):
    """`fused` -> this rank's packed tensors: dict(layers=[dict(g1, qkv, qkv_bias, o, o_bias, g2, gate_up, down)], embed, final_gamma, lm_head).
    Tensor parallelism: QKV / gate_up row-sharded by head / FFN column, O / down K-sharded, everything else replicated (tp.py)."""
    import torch

    tpm = _tp()
    tdt = getattr(torch, _TORCH_DT[dtype])

    def dev(a):
        return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=tdt)

    cfg = dict(head_num=shape["head_num"], kv_head_num=shape["kv_head_num"], head_size=shape["head_size"], inter=shape["inter"])
    layers = []
    for w in fused["layers"]:
        ws = tpm.shard_layer(w, cfg, rank, tp) if tp > 1 else w
        layers.append(dict(g1=dev(ws["g1"]), qkv=pack_linear(ws["wqkv"], dtype, wformat, device, group), qkv_bias=dev(ws.get("bqkv")),
                           o=pack_linear(ws["wo"], dtype, wformat, device, group), o_bias=dev(ws.get("bo")), g2=dev(ws["g2"]),
                           gate_up=pack_linear(ws["wgu"], dtype, wformat, device, group), down=pack_linear(ws["wd"], dtype, wformat, device, group)))
    return dict(layers=layers, embed=dev(fused["embed"]), final_gamma=dev(fused["final_gamma"]), lm_head=dev(fused["lm_head"]))


def _raw(t):
    """bytes of a device / host tensor as stored (bf16 has no numpy dtype: go through a uint8 view)."""
    return t.detach().contiguous().view(-1).view(dtype=__import__("torch").uint8).cpu().numpy()


def save_packed(packed, shape, out_dir, dtype, wformat, tp=1, rank=0, group=128):
    """Write this rank's packed tensors under out_dir/rank{rank}/ + manifest.json (shapes, dtypes, formats, the model shape)."""
    d = os.path.join(out_dir, f"rank{rank}")
    os.makedirs(d, exist_ok=True)
    entries = {}

    def put(name, t):
        if t is None:
            return
        parts = t if isinstance(t, (tuple, list)) else (t,)
        for i, p in enumerate(parts):
            if p is None:
                continue
            fname = f"{name}.{('w', 'scales', 'zeros')[i]}.bin"
            _raw(p).tofile(os.path.join(d, fname))
            entries[fname] = dict(shape=list(p.shape), dtype=str(p.dtype).replace("torch.", ""))

    for l, w in enumerate(packed["layers"]):
        for key in ("g1", "qkv", "qkv_bias", "o", "o_bias", "g2", "gate_up", "down"):
            put(f"layers.{l}.{key}", w.get(key))
    for key in ("embed", "final_gamma", "lm_head"):
        put(key, packed[key])
    manifest = dict(format="b200llm-packed-v1", shape=shape, dtype=dtype, wformat=wformat, group=group, tp=tp, rank=rank, layout="NK", tensors=entries)
    with open(os.path.join(d, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


def load_packed(out_dir, device, rank=0):
    """out_dir/rank{rank}/ -> (packed dict on `device`, manifest)."""
    import torch

    d = os.path.join(out_dir, f"rank{rank}")
    manifest = json.load(open(os.path.join(d, "manifest.json")))
    if manifest.get("format") != "b200llm-packed-v1":
        raise ValueError(f"{d}: not a b200llm packed directory")

    def get(name):
        parts = []
        for part in ("w", "scales", "zeros"):
            fname = f"{name}.{part}.bin"
            e = manifest["tensors"].get(fname)
            if e is None:
                parts.append(None)
                continue
            tdt = getattr(torch, e["dtype"])
            raw = np.fromfile(os.path.join(d, fname), dtype=np.uint8)
            t = torch.from_numpy(raw).view(tdt).reshape(e["shape"]).to(device)
            parts.append(t)
        if parts[0] is None:
            return None
        if parts[1] is None:
            return parts[0]
        return tuple(p for p in parts if p is not None)

    L = manifest["shape"]["layers"]
    layers = [{key: get(f"layers.{l}.{key}") for key in ("g1", "qkv", "qkv_bias", "o", "o_bias", "g2", "gate_up", "down")} for l in range(L)]
    return dict(layers=layers, embed=get("embed"), final_gamma=get("final_gamma"), lm_head=get("lm_head")), manifest


def build_decoder(packed, shape, device, dtype, wformat, max_seq, max_batch, tp=1, rank=0, group=128, eps=1e-5, rope_base=10000.0):
    """A `Decoder` over this rank's packed tensors.  Returns (decoder, packed): keep `packed` alive -- the engine borrows its pointers."""
    mod = _mod()
    H, Hkv, I = shape["head_num"] // tp, shape["kv_head_num"] // tp, shape["inter"] // tp
    dcode = {"f32": mod.F32, "f16": mod.F16, "bf16": mod.BF16}[dtype]
    wcode = {"dense": mod.W_DENSE, "bf16": mod.W_DENSE, "fp8": mod.W_FP8, "int4": mod.W_INT4}[wformat]
    dc = mod.DecoderConfig(shape["hidden"], H, Hkv, shape["head_size"], I, shape["layers"], max_seq, max_batch, dcode, wcode, group, eps,
                           shape["head_size"], rope_base, tp, rank)
    dec = mod.Decoder(dc, device)
    for l, w in enumerate(packed["layers"]):
        dec.set_layer(l, {k: v for k, v in w.items() if v is not None})
    return dec, packed
