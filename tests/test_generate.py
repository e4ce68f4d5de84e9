"""The generation loop b200_generate (what the reference's LlamaModel<T>::response intends, src/models/llama/llama.cpp:165-398) against
(1) the CPU oracle composed the same way -- greedy ids are integer work: bit-exact given identical logits -- and (2) the engine's own
primitives called one by one with the reference's step / seed bookkeeping (bit-exact, top-k sampling included)."""
import numpy as np
import pytest

from oracle import oracle
from test_decoder_engine import build_decoder, make_model, oracle_prefill
from util import b200, rounded, to_dev, to_np

CFG = dict(hidden=256, head_num=2, kv_head_num=2, head_size=128, inter=384, layers=2, max_seq=48, eps=1e-6, base=10000.0)
V, END = 500, 2
GOLDEN = dict(model_seed=3, tail_seed=4, prompt_seed=11, batch=2, prompt_len=7, new_tokens=6)  # tests/golden/generate_golden.npz


def tail_weights(seed, dtype):
    rng = np.random.default_rng(seed)
    emb = rounded(rng.standard_normal((V, CFG["hidden"])), dtype)
    gamma = rounded(1 + 0.1 * rng.standard_normal(CFG["hidden"]), dtype)
    lm = rounded(rng.standard_normal((V, CFG["hidden"])) / 16, dtype)
    return emb, gamma, lm


def oracle_generate(model, emb, gamma, lm, prompt, n_new):
    """fp32 restatement of the loop: embedding gather -> context decoder -> last token -> RMSNorm -> LM head -> argmax; then decode steps."""
    cfg = CFG
    B, T = prompt.shape
    L, Hkv, d, S = cfg["layers"], cfg["kv_head_num"], cfg["head_size"], cfg["max_seq"]
    kc = np.zeros((L, B, Hkv, S, d), np.float32)
    vc = np.zeros_like(kc)
    x = emb[prompt.reshape(-1)].astype(np.float32)
    il, hist = np.full(B, T, np.int32), np.zeros(B, np.int32)
    x, kc, vc = oracle_prefill(model, cfg, "f32", x, kc, vc, il, hist)
    h = x.reshape(B, T, -1)[:, -1].copy()
    ocfg = dict(head_num=cfg["head_num"], kv_head_num=cfg["kv_head_num"], head_size=d, inter=cfg["inter"], eps=cfg["eps"], rot_dim=d, base=cfg["base"])
    out, logits_all = [], []
    step = T
    for i in range(n_new):
        xn = h.copy()
        oracle.rmsnorm(xn, None, gamma, cfg["eps"])
        logits = oracle.linear(xn, lm, "nk")
        logits_all.append(logits)
        ids, _ = oracle.topk(logits, 1)
        tok = ids[:, 0].astype(np.int32)
        out.append(tok)
        if i + 1 == n_new:
            break
        step += 1
        h = emb[tok].astype(np.float32).copy()
        for l, w in enumerate(model["layers"]):
            oracle.decoder_layer(h, w, kc, vc, ocfg, step, l)
    return np.stack(out, axis=1), logits_all


def golden():
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generate_golden.npz"))


def test_checker_loop_reproduces_the_golden_fixture():
    """CPU only: the checker's generation loop against the committed fixture (tests/golden/make_generate_golden.py)."""
    g = golden()
    model = make_model(CFG, seed=GOLDEN["model_seed"], bias=False)
    emb, gamma, lm = tail_weights(GOLDEN["tail_seed"], "f32")
    ids, logits = oracle_generate(model, emb, gamma, lm, g["prompt"], GOLDEN["new_tokens"])
    assert np.array_equal(ids, g["ids"])
    assert np.allclose(logits[-1], g["last_logits"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_generate_greedy_matches_oracle_fp32():
    import torch

    dtype = "f32"
    model = make_model(CFG, seed=3, bias=False)
    emb, gamma, lm = tail_weights(4, dtype)
    B, T, N = 2, 7, 6
    rng = np.random.default_rng(11)
    prompt = rng.integers(3, V, size=(B, T)).astype(np.int32)
    dec = build_decoder(model, CFG, dtype, B)
    dev = torch.device("cuda")
    kc = torch.zeros((CFG["layers"], B, CFG["kv_head_num"], CFG["max_seq"], CFG["head_size"]), dtype=torch.float32, device=dev)
    vc = torch.zeros_like(kc)
    ids, ngen = dec.generate(prompt, to_dev(emb, dtype), to_dev(gamma, dtype), to_dev(lm, dtype), kc, vc, N, top_k=1, end_id=END)
    ref, logits_all = oracle_generate(model, emb, gamma, lm, prompt, N)
    # greedy ids are bit-exact unless the oracle's own top two logits are closer than fp32 reduction-order noise
    for i, lg in enumerate(logits_all):
        srt = np.sort(lg, axis=1)
        assert ((srt[:, -1] - srt[:, -2]) > 1e-4 * np.abs(srt[:, -1])).all(), "test model produced a near-tie: pick another seed"
    expect = ref.copy()
    for b in range(B):  # everything from the first end_id on reads end_id
        hit = np.where(expect[b] == END)[0]
        if len(hit):
            expect[b, hit[0]:] = END
    assert np.array_equal(ids, expect), f"{ids} vs {expect}"
    g = golden()  # same seeds as the fixture: the engine also reproduces the committed ids
    assert np.array_equal(prompt, g["prompt"]) and np.array_equal(ref, g["ids"])
    assert np.array_equal(ngen, [(np.where(expect[b] == END)[0][0] if (expect[b] == END).any() else N) for b in range(B)])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_generate_equals_the_primitives_called_one_by_one(dtype):
    """Loop bookkeeping (positions, sampling seeds, cache ownership, id feedback): b200_generate with top-k sampling must reproduce,
    bit for bit, the ids obtained by driving prefill / step / lm_head_topk_sample by hand with the reference's step sequence."""
    import torch

    mod = b200()
    model = make_model(CFG, seed=5, bias=True)
    emb, gamma, lm = tail_weights(6, dtype)
    B, T, N, K = 3, 5, 8, 4
    rng = np.random.default_rng(12)
    prompt = rng.integers(3, V, size=(B, T)).astype(np.int32)
    dev = torch.device("cuda")
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16}[dtype]
    embd, gd, lmd = to_dev(emb, dtype), to_dev(gamma, dtype), to_dev(lm, dtype)

    def caches():
        kc = torch.zeros((CFG["layers"], B, CFG["kv_head_num"], CFG["max_seq"], CFG["head_size"]), dtype=tdt, device=dev)
        return kc, torch.zeros_like(kc)

    dec = build_decoder(model, CFG, dtype, B)
    kc, vc = caches()
    ids, ngen = dec.generate(prompt, embd, gd, lmd, kc, vc, N, top_k=K, end_id=END, check_every=3)

    # by hand
    dec2 = build_decoder(model, CFG, dtype, B)
    kc2, vc2 = caches()
    pid = torch.from_numpy(prompt.reshape(-1)).to(dev)
    x = torch.empty((B * T, CFG["hidden"]), dtype=tdt, device=dev)
    mod.check(mod.lib().b200_input_embedding(mod.ptr(pid), mod.ptr(embd), mod.ptr(x), B * T, CFG["hidden"], mod.dtype_code(x), mod.stream()))
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    hl = torch.zeros(B, dtype=torch.int32, device=dev)
    dec2.prefill(x, kc2, vc2, il, hl, il, T)
    h = x.view(B, T, -1)[:, -1].contiguous()
    bufs = dict(logits=torch.empty((B, V), dtype=torch.float32, device=dev), tmp_ids=torch.empty((B, 8, K), dtype=torch.int32, device=dev),
                tmp_vals=torch.empty((B, 8, K), dtype=torch.float32, device=dev), topk_ids=torch.empty((B, K), dtype=torch.int32, device=dev),
                topk_vals=torch.empty((B, K), dtype=torch.float32, device=dev), seq_len=torch.full((B,), T, dtype=torch.int32, device=dev),
                finished=torch.zeros(B, dtype=torch.uint8, device=dev), output_id=torch.zeros(B, dtype=torch.int32, device=dev))
    step = T
    hand = []
    for i in range(N):
        dec2.lm_head_topk_sample(h, gd, lmd, bufs, K, step, END)
        tok = bufs["output_id"].clone()
        hand.append(to_np(tok))
        if i + 1 == N:
            break
        step += 1
        mod.check(mod.lib().b200_input_embedding(mod.ptr(tok), mod.ptr(embd), mod.ptr(h), B, CFG["hidden"], mod.dtype_code(h), mod.stream()))
        dec2.step(h, kc2, vc2, step)
    hand = np.stack(hand, axis=1)
    expect = hand.copy()
    for b in range(B):
        hit = np.where(expect[b] == END)[0]
        if len(hit):
            expect[b, hit[0]:] = END
    torch.cuda.synchronize()
    # with check_every the loop may stop early once EVERY sequence finished: columns it never produced read end_id, like `expect`
    assert np.array_equal(ids, expect), f"{ids}\nvs\n{expect}"
    assert np.array_equal(to_np(kc), to_np(kc2)) and np.array_equal(to_np(vc), to_np(vc2)) or (expect == END).any()


def test_generate_argument_errors_without_gpu_compute():
    """Argument validation happens before any launch: reachable on a CPU-only box through the C ABI."""
    mod = b200()
    import ctypes as C

    gp = mod.GenerateParams(None, None, None, 10, 1, 2, 4, 0)
    assert mod.lib().b200_generate_workspace_bytes(None, C.byref(gp), 1, 4) == 0
    assert b"null" in mod.lib().b200_last_error_string()
