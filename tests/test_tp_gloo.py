"""World-size-2 check (gloo, CPU) of the tensor-parallel decode path: the sharding of llm-inference-engine_b200/tp.py + one
all-reduce per attention block and per MLP block reproduces the un-sharded decoder layer.  The per-rank blocks are evaluated with
the CPU oracle's ops (the GPU engine evaluates the same blocks with b200_decoder_attn_block / b200_decoder_ffn_block); what is
under test here is the host-side algebra and the collective plumbing that bench.py --gpus N uses."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFG = dict(hidden=128, head_num=4, kv_head_num=2, head_size=32, inter=96, eps=1e-6, base=10000.0)
LAYERS, BATCH, STEP, SEQ = 3, 2, 7, 16


def make_layers(seed):
    rng = np.random.default_rng(seed)
    h, H, Hkv, d, I = CFG["hidden"], CFG["head_num"], CFG["kv_head_num"], CFG["head_size"], CFG["inter"]
    f = lambda *s, sc=1.0: (sc * rng.standard_normal(s)).astype(np.float32)
    return [dict(g1=1 + f(h, sc=0.1), wqkv=f((H + 2 * Hkv) * d, h, sc=h ** -0.5), bqkv=f((H + 2 * Hkv) * d, sc=0.05), wo=f(h, H * d, sc=(H * d) ** -0.5),
                 bo=f(h, sc=0.05), g2=1 + f(h, sc=0.1), wgu=f(2 * I, h, sc=h ** -0.5), wd=f(h, I, sc=I ** -0.5)) for _ in range(LAYERS)]


def make_inputs(seed):
    rng = np.random.default_rng(seed + 1)
    x = rng.standard_normal((BATCH, CFG["hidden"])).astype(np.float32)
    kc = (0.5 * rng.standard_normal((LAYERS, BATCH, CFG["kv_head_num"], SEQ, CFG["head_size"]))).astype(np.float32)
    vc = (0.5 * rng.standard_normal((LAYERS, BATCH, CFG["kv_head_num"], SEQ, CFG["head_size"]))).astype(np.float32)
    return x, kc, vc


def worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist

    from oracle import oracle

    tp = importlib.import_module("llm-inference-engine_b200.tp")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layers = make_layers(3)
        x, kc, vc = make_inputs(3)
        lcfg = tp.local_cfg(CFG, world)
        shards = [tp.shard_layer(w, CFG, rank, world) for w in layers]
        kcl, vcl = tp.shard_kv_cache(kc, CFG["kv_head_num"], rank, world), tp.shard_kv_cache(vc, CFG["kv_head_num"], rank, world)
        H, Hkv, d, I = lcfg["head_num"], lcfg["kv_head_num"], lcfg["head_size"], lcfg["inter"]
        state = dict(res=None)

        def attn_block(l, hidden, pending):
            w = shards[l]
            cur = hidden.copy() if pending is None else pending + state["res"]  # fold the previous FFN output into the stream
            state["res"] = cur.copy()
            xn = cur.copy()
            oracle.rmsnorm(xn, None, w["g1"], CFG["eps"])
            qkv = oracle.linear(xn, w["wqkv"], "nk").reshape(BATCH, H + 2 * Hkv, d)
            oracle.rope_decode(qkv, H, Hkv, STEP, d, CFG["base"])
            mha = oracle.decode_mha(qkv, w["bqkv"], kcl, vcl, H, Hkv, STEP, l)
            return oracle.linear(mha, w["wo"], "nk")  # this rank's partial sum of the row-sharded O projection

        def ffn_block(l, pending):
            w = shards[l]
            out = pending.copy()
            oracle.fused_add_bias_residual_rmsnorm(state["res"], out, w["bo"], w["g2"], CFG["eps"])  # res += attn; + o bias; norm
            gu = oracle.linear(out, w["wgu"], "nk").reshape(BATCH, 2, I)
            act = oracle.silu_and_mul(gu)
            return oracle.linear(act, w["wd"], "nk")

        def fold(hidden, pending):
            return pending + state["res"]

        def all_reduce(a):
            t = torch.from_numpy(a)
            dist.all_reduce(t)

        out = tp.decode_step_tp(LAYERS, x, attn_block, ffn_block, fold, all_reduce)
        if ret is not None:
            ret[rank] = (out, kcl, vcl)
    finally:
        dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_shapes_and_pairing():
    tp = importlib.import_module("llm-inference-engine_b200.tp")
    w = make_layers(0)[0]
    H, Hkv, d, I = CFG["head_num"], CFG["kv_head_num"], CFG["head_size"], CFG["inter"]
    s0, s1 = tp.shard_layer(w, CFG, 0, 2), tp.shard_layer(w, CFG, 1, 2)
    assert s0["wqkv"].shape == ((H + 2 * Hkv) * d // 2, CFG["hidden"]) and s0["wo"].shape == (CFG["hidden"], H * d // 2)
    assert s0["wgu"].shape == (I, CFG["hidden"]) and s0["wd"].shape == (CFG["hidden"], I // 2)
    # gate row i and up row i of a rank are rows (i, I/2 + i) of its shard: the pairing the SwiGLU epilogue relies on
    assert np.array_equal(s1["wgu"][3], w["wgu"][I // 2 + 3]) and np.array_equal(s1["wgu"][I // 2 + 3], w["wgu"][I + I // 2 + 3])
    # the row-sharded linears' K-columns tile the full matrix
    assert np.array_equal(np.concatenate([s0["wo"], s1["wo"]], axis=1), w["wo"])
    assert np.array_equal(np.concatenate([s0["wd"], s1["wd"]], axis=1), w["wd"])
    with pytest.raises(ValueError):
        tp.check_divisible(H, Hkv, I, 3)


def test_two_rank_decode_equals_unsharded_oracle():
    import torch.multiprocessing as mp

    from oracle import oracle

    world, port = 2, free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    # un-sharded reference
    layers = make_layers(3)
    x, kc, vc = make_inputs(3)
    ocfg = dict(head_num=CFG["head_num"], kv_head_num=CFG["kv_head_num"], head_size=CFG["head_size"], inter=CFG["inter"], eps=CFG["eps"],
                rot_dim=CFG["head_size"], base=CFG["base"])
    ref = x.copy()
    for l, w in enumerate(layers):
        oracle.decoder_layer(ref, w, kc, vc, ocfg, STEP, l)
    out0, kc0, vc0 = ret[0]
    out1, kc1, vc1 = ret[1]
    assert np.array_equal(out0, out1)  # every rank holds the same reduced residual stream
    np.testing.assert_allclose(out0, ref, rtol=2e-5, atol=2e-5)
    # head-sharded KV cache: each rank appended exactly its heads at exactly position STEP-1 (indices bit-exact: every other
    # position is untouched); the appended values agree to fp32 reduction-order noise (the all-reduce changes the summation order)
    gk, gv = np.concatenate([kc0, kc1], axis=2), np.concatenate([vc0, vc1], axis=2)
    x0, kc_in, vc_in = make_inputs(3)
    mask = np.ones(kc.shape, bool)
    mask[:, :, :, STEP - 1] = False
    assert np.array_equal(gk[mask], kc_in[mask]) and np.array_equal(gv[mask], vc_in[mask])
    np.testing.assert_allclose(gk, kc, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(gv, vc, rtol=2e-5, atol=2e-5)


VOCAB, TOPK = 1000, 5


def lm_head_inputs():
    rng = np.random.default_rng(21)
    hidden = rng.standard_normal((BATCH, CFG["hidden"])).astype(np.float32)
    lm = (rng.integers(-3, 4, size=(VOCAB, CFG["hidden"])) / 8.0).astype(np.float32)  # few distinct values: exact sums, many TIED logits
    hidden = np.round(hidden * 4) / 4
    # identical rows on different ranks, aligned with a token's hidden state so that they ARE the largest logits of that token: an exact tie
    # for first place between ids (10, 700) for token 0, and across the shard boundary (499, 501) for token 1 -- the lower id must come first
    lm[10] = lm[700] = 0.5 * np.sign(hidden[0])
    lm[499] = lm[501] = 0.375 * np.sign(hidden[1])
    return hidden, lm


def lm_head_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist

    from oracle import oracle

    tp = importlib.import_module("llm-inference-engine_b200.tp")
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hidden, lm = lm_head_inputs()
        lo, _ = tp.vocab_range(VOCAB, rank, world)
        logits = oracle.linear(hidden, tp.shard_lm_head_rows(lm, rank, world), "nk")  # [B, V/P]
        ids, vals = oracle.topk(logits, TOPK)
        pack = torch.from_numpy(np.stack([vals.astype(np.float32), (ids + lo).astype(np.float32)]))  # ids < 2^24: exact in fp32
        gathered = [torch.empty_like(pack) for _ in range(world)]
        dist.all_gather(gathered, pack)
        g = np.stack([t.numpy() for t in gathered])  # [P, 2, B, k]
        ret[rank] = tp.merge_topk(g[:, 0], g[:, 1].astype(np.int64), TOPK)
    finally:
        dist.destroy_process_group()


def test_vocab_sharded_lm_head_topk_is_bit_identical():
    """Vocab-sharded LM head (SURVEY.md 8e refinement): per-rank logits slice -> local top-k -> all-gather of k (value, id) pairs ->
    merge reproduces the un-sharded top-k ids AND values bit for bit, ties included (world size 2, gloo)."""
    import torch.multiprocessing as mp

    from oracle import oracle

    tp = importlib.import_module("llm-inference-engine_b200.tp")
    world, port = 2, free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=lm_head_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    hidden, lm = lm_head_inputs()
    logits = oracle.linear(hidden, lm, "nk")
    want_ids, want_vals = oracle.topk(logits, TOPK)
    assert want_ids[0, :2].tolist() == [10, 700] and want_vals[0, 0] == want_vals[0, 1]  # the fixture does contain the ties it promises
    assert want_ids[1, :2].tolist() == [499, 501] and want_vals[1, 0] == want_vals[1, 1]
    for r in range(world):
        ids, vals = ret[r]
        assert np.array_equal(ids, want_ids) and np.array_equal(vals, want_vals), f"rank {r}: {ids} {vals} vs {want_ids} {want_vals}"
    # the merge's tie rule on its own: equal values, the lower id first, whichever rank it came from
    i, v = tp.merge_topk(np.array([[[2.0, 1.0]], [[2.0, 2.0]]]), np.array([[[900, 5]], [[3, 40]]]), 3)
    assert i.tolist() == [[3, 40, 900]] and v.tolist() == [[2.0, 2.0, 2.0]]
    with pytest.raises(ValueError):
        tp.vocab_range(1001, 0, 2)


@pytest.mark.gpu
def test_two_gpu_engine_matches_unsharded_oracle():
    """Needs >= 2 GPUs (gpurun --gpus 2): the sharded fused engine + NCCL all-reduce against the un-sharded oracle."""
    import subprocess

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "tp_engine_check.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = p.stdout.decode(errors="replace")
    print(out[-3000:])
    assert p.returncode == 0, out[-3000:]
