#!/usr/bin/env python
"""Offline weight converter.

  convert_weights.py --hf DIR_OR_FILE | --ref-bins PREFIX   --out DIR  [--dtype bf16] [--wformat bf16|fp8|int4] [--tp P]
                     [--export-ref-bins PREFIX]  [--shape hidden,heads,kv_heads,head_size,inter,layers,vocab]

--hf: a Hugging Face Llama checkpoint (directory with *.safetensors / pytorch_model*.bin, or one such file).
--ref-bins: the reference's per-tensor fp32 files (`PREFIX` + model.layers.N....bin).
Writes OUT/rank{r}/ (packed [N,K] tensors + manifest.json) for every tensor-parallel rank; FP8 / INT4 packing needs a GPU.
--export-ref-bins additionally writes the reference's own directory format (what its loadWeights(path) reads)."""
import argparse
import glob
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
W = importlib.import_module("llm-inference-engine_b200.weights")


def load_hf(path):
    import torch

    files = [path] if os.path.isfile(path) else sorted(glob.glob(os.path.join(path, "*.safetensors")) or glob.glob(os.path.join(path, "pytorch_model*.bin")))
    if not files:
        raise SystemExit(f"no checkpoint files under {path}")
    sd = {}
    for f in files:
        if f.endswith(".safetensors"):
            from safetensors.torch import load_file

            sd.update(load_file(f))
        else:
            sd.update(torch.load(f, map_location="cpu", weights_only=True))
    return sd


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--hf")
    ap.add_argument("--ref-bins")
    ap.add_argument("--out")
    ap.add_argument("--export-ref-bins")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f16", "f32"])
    ap.add_argument("--wformat", default="bf16", choices=["bf16", "fp8", "int4"])
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--shape", default=None, help="hidden,heads,kv_heads,head_size,inter,layers,vocab (default: Llama-2-7B)")
    args = ap.parse_args()
    shape = W.shape_of(*[int(v) for v in args.shape.split(",")]) if args.shape else dict(W.LLAMA2_7B)
    if bool(args.hf) == bool(args.ref_bins):
        raise SystemExit("give exactly one of --hf / --ref-bins")
    fused = W.fuse_hf_state_dict(load_hf(args.hf), shape) if args.hf else W.load_reference_bins(args.ref_bins, shape)
    if args.export_ref_bins:
        W.export_reference_bins(fused, args.export_ref_bins)
        print(f"reference format written with prefix {args.export_ref_bins}")
    if args.out:
        import torch

        dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        if args.wformat != "bf16" and dev.type != "cuda":
            raise SystemExit("FP8 / INT4 packing runs the library's device quantisers: needs a GPU")
        for r in range(args.tp):
            packed = W.pack_model(fused, shape, args.dtype, "dense" if args.wformat == "bf16" else args.wformat, dev, args.tp, r)
            m = W.save_packed(packed, shape, args.out, args.dtype, args.wformat, args.tp, r)
            print(f"rank {r}: {len(m['tensors'])} tensors -> {os.path.join(args.out, 'rank%d' % r)}")


if __name__ == "__main__":
    main()
