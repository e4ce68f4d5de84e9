"""Shared helpers for the parity tests: numpy <-> device tensors and the tolerance rules of BASELINE.md section 5."""
import importlib

import numpy as np


def b200():
    return importlib.import_module("llm-inference-engine_b200")


def torch_dtype(name):
    import torch

    return {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[name]


def to_dev(a, dtype="f32"):
    """numpy -> cuda tensor.  float arrays are cast to `dtype`; integer / uint8 arrays keep their type."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype in (torch.float32, torch.float64):
        t = t.to(torch_dtype(dtype))
    return t.cuda().contiguous()


def to_np(t):
    import torch

    if t.dtype in (torch.float16, torch.bfloat16):
        t = t.float()
    return t.detach().cpu().numpy()


def rounded(a, dtype):
    """The fp32 values a device tensor of `dtype` would hold for fp32 input `a` (so that the oracle sees identical inputs)."""
    import torch

    if dtype == "f32":
        return np.ascontiguousarray(a, np.float32)
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch_dtype(dtype)).float().numpy()


def assert_close(got, ref, dtype, what=""):
    """BASELINE.md section 5.
    fp32: |got - ref| <= 1e-5 |ref| + 2e-6 max|ref| (reduction-order differences only).
    16-bit (bf16 / fp16 storage, fp32 accumulate) against the fp32 oracle on identical inputs: relative error <= 1e-2, measured
    (a) over the tensor, ||got - ref||_2 / ||ref||_2 <= 1e-2, and (b) per element with the tensor's scale as the floor,
    |got - ref| <= 1e-2 |ref| + 1e-2 rms(ref) -- an element-wise ratio alone is meaningless where the sum cancels to ~0."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert np.isfinite(got).all(), f"{what}: non-finite output"
    if ref.size == 0:
        return
    scale = np.abs(ref).max()
    err = np.abs(got - ref)
    if dtype == "f32":
        tol = 1e-5 * np.abs(ref) + 2e-6 * max(scale, 1e-30)
        bad = err > tol
        assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} off, max err {err.max():.3e} at scale {scale:.3e}"
    else:
        rms = np.sqrt((ref ** 2).mean())
        fro = np.sqrt((err ** 2).sum()) / max(np.sqrt((ref ** 2).sum()), 1e-30)
        worst = (err / (1e-2 * np.abs(ref) + 1e-2 * rms + 1e-30)).max()
        assert fro <= 1e-2 and worst <= 1.0, (f"{what}: ||err||/||ref|| = {fro:.3e} (limit 1e-2), worst element at {worst:.2f}x of "
                                              f"1e-2|ref| + 1e-2 rms; max abs err {err.max():.3e}, rms(ref) {rms:.3e}, max|ref| {scale:.3e}")
