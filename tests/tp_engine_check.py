"""Run under torchrun (N ranks, one GPU each): the tensor-parallel fused engine (b200_decoder_attn_block / _ffn_block + one NCCL
all-reduce per block) against the UN-SHARDED CPU oracle on the same seeded model.  Exits non-zero on mismatch.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tp_engine_check.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    from oracle import oracle
    from test_decoder_engine import make_inputs, make_model, run_oracle
    from util import assert_close, rounded, to_dev, to_np

    mod = importlib.import_module("llm-inference-engine_b200")
    tp = importlib.import_module("llm-inference-engine_b200.tp")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = dict(hidden=1024, head_num=8, kv_head_num=8 if world <= 8 else world, head_size=128, inter=2048, layers=3, max_seq=160, eps=1e-6, base=10000.0)
    ok = True
    for mode, dtype, batch, step in (("nccl", "f32", 2, 37), ("nccl", "bf16", 1, 130), ("fused", "f32", 2, 37), ("fused", "bf16", 1, 130),
                                     ("fused", "bf16", 4, 64), ("fused", "bf16", 6, 21)):
        model = make_model(cfg, seed=17)
        lcfg = tp.local_cfg(dict(head_num=cfg["head_num"], kv_head_num=cfg["kv_head_num"], head_size=cfg["head_size"], inter=cfg["inter"]), world)
        dc = mod.DecoderConfig(cfg["hidden"], lcfg["head_num"], lcfg["kv_head_num"], cfg["head_size"], lcfg["inter"], cfg["layers"], cfg["max_seq"],
                               batch, {"f32": 0, "f16": 1, "bf16": 2}[dtype], 0, 128, cfg["eps"], cfg["head_size"], cfg["base"], world, rank)
        dec = mod.Decoder(dc, dev)
        for l, w in enumerate(model["layers"]):
            s = tp.shard_layer(w, cfg, rank, world)
            dec.set_layer(l, dict(g1=to_dev(s["g1"], dtype), qkv=to_dev(s["wqkv"], dtype), qkv_bias=to_dev(s["bqkv"], dtype), o=to_dev(s["wo"], dtype),
                                  o_bias=to_dev(s["bo"], dtype), g2=to_dev(s["g2"], dtype), gate_up=to_dev(s["wgu"], dtype), down=to_dev(s["wd"], dtype)))
        x, kc, vc = make_inputs(cfg, batch, step, model["seed"])
        hidden = to_dev(x, dtype)
        kcd = to_dev(tp.shard_kv_cache(kc, cfg["kv_head_num"], rank, world), dtype)
        vcd = to_dev(tp.shard_kv_cache(vc, cfg["kv_head_num"], rank, world), dtype)
        y_attn, y_ffn = torch.empty_like(hidden), torch.empty_like(hidden)

        def attn_block(l, h, pending):
            dec.attn_block(l, h, pending, kcd, vcd, y_attn, step)
            return y_attn

        def ffn_block(l, pending):
            dec.ffn_block(l, pending, y_ffn)
            return y_ffn

        def fold(h, pending):
            dec.fold(h, pending)
            return h

        if mode == "nccl":
            tp.decode_step_tp(cfg["layers"], hidden, attn_block, ffn_block, fold, dist.all_reduce)
        else:  # one-shot all-reduce over NVLink peer memory fused into the consuming kernels: no NCCL call on the path
            dec.tp_attach(dist)
            dec.step_tp(hidden, kcd, vcd, step)
        torch.cuda.synchronize()
        assert mode == "nccl" or dec.tp_error() == 0, "a peer never signalled (exchange timed out)"
        got = to_np(hidden)
        ref, rkc, rvc = run_oracle(model, cfg, dtype, batch, step, storage="f32")
        try:
            if dtype == "f32":
                assert_close(got, ref, "f32", f"TP-{world} {mode} engine")
            else:
                fro = np.linalg.norm(got.astype(np.float64) - ref) / np.linalg.norm(ref)
                assert fro <= 1e-2, f"TP-{world} {dtype} engine vs fp32 oracle: {fro:.3e}"
            mine = to_np(kcd)[:, :, :, step - 1].astype(np.float64)
            want = tp.shard_kv_cache(rkc, cfg["kv_head_num"], rank, world)[:, :, :, step - 1]
            kerr = np.linalg.norm(mine - want) / np.linalg.norm(want)
            assert kerr <= (1e-5 if dtype == "f32" else 1e-2), f"appended K rows of this rank's heads: {kerr:.3e}"
            print(f"[rank {rank}] TP-{world} {mode} {dtype} batch {batch} step {step}: OK", flush=True)
        except AssertionError as e:
            ok = False
            print(f"[rank {rank}] {mode} {dtype} batch {batch} FAILED: {e}", flush=True)
    # ---- tensor-parallel PREFILL (b200_decoder_prefill_tp: this rank's shard + one all-reduce per attention and per MLP block, the
    #      collective supplied by the host) against the UN-SHARDED oracle composition; ragged batch with history; bf16 (tcgen05 path) and fp32
    from test_decoder_engine import oracle_prefill

    for dtype in ("bf16", "f32"):
        try:
            model = make_model(cfg, seed=19, bias=True)
            lcfg = tp.local_cfg(dict(head_num=cfg["head_num"], kv_head_num=cfg["kv_head_num"], head_size=cfg["head_size"], inter=cfg["inter"]), world)
            input_len, hist = np.array([150, 37], np.int32), np.array([0, 9], np.int32)
            B, T = len(input_len), int(input_len.sum())
            dc = mod.DecoderConfig(cfg["hidden"], lcfg["head_num"], lcfg["kv_head_num"], cfg["head_size"], lcfg["inter"], cfg["layers"], cfg["max_seq"],
                                   B, {"f32": 0, "f16": 1, "bf16": 2}[dtype], 0, 128, cfg["eps"], cfg["head_size"], cfg["base"], world, rank)
            dec = mod.Decoder(dc, dev)
            for l, w in enumerate(model["layers"]):
                s = tp.shard_layer(w, cfg, rank, world)
                dec.set_layer(l, dict(g1=to_dev(s["g1"], dtype), qkv=to_dev(s["wqkv"], dtype), qkv_bias=to_dev(s["bqkv"], dtype), o=to_dev(s["wo"], dtype),
                                      o_bias=to_dev(s["bo"], dtype), g2=to_dev(s["g2"], dtype), gate_up=to_dev(s["wgu"], dtype), down=to_dev(s["wd"], dtype)))
            r = np.random.default_rng(23)
            x = rounded(r.standard_normal((T, cfg["hidden"])), dtype)
            kc = rounded(0.5 * r.standard_normal((cfg["layers"], B, cfg["kv_head_num"], cfg["max_seq"], cfg["head_size"])), dtype)
            vc = rounded(0.5 * r.standard_normal(kc.shape), dtype)
            xd = to_dev(x, dtype)
            kcd = to_dev(tp.shard_kv_cache(kc, cfg["kv_head_num"], rank, world), dtype)
            vcd = to_dev(tp.shard_kv_cache(vc, cfg["kv_head_num"], rank, world), dtype)
            ctx = input_len + hist
            dec.prefill_tp(xd, kcd, vcd, to_dev(input_len), to_dev(hist), to_dev(ctx), int(input_len.max()), dist)
            torch.cuda.synchronize()
            oracle.set_threads(oracle.max_threads())
            ref, rkc, rvc = oracle_prefill(model, cfg, dtype, x, kc.copy(), vc.copy(), input_len, hist)
            got = to_np(xd).astype(np.float64)
            fro = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            assert fro <= (1e-5 if dtype == "f32" else 1e-2), f"prefill output: {fro:.3e}"
            mine, want = to_np(kcd).astype(np.float64), tp.shard_kv_cache(rkc, cfg["kv_head_num"], rank, world)
            for b in range(B):
                lo, hi = int(hist[b]), int(ctx[b])
                kerr = np.linalg.norm(mine[:, b, :, lo:hi] - want[:, b, :, lo:hi]) / np.linalg.norm(want[:, b, :, lo:hi])
                assert kerr <= (1e-5 if dtype == "f32" else 1e-2), f"K rows of this rank's heads, sequence {b}: {kerr:.3e}"
            print(f"[rank {rank}] TP-{world} prefill {dtype} (ragged, history): OK  rel err {fro:.2e}", flush=True)
        except (AssertionError, mod.B200Error) as e:
            ok = False
            print(f"[rank {rank}] TP prefill {dtype} FAILED: {e}", flush=True)
    # ---- the whole loop under tensor parallelism (tp.generate_tp: prefill_tp -> vocab-sharded head -> step_tp ...) against b200_generate
    #      of the UN-SHARDED model on this GPU: fp32, greedy -- the same ids
    try:
        dtype, B, T, N, V = "f32", 2, 9, 6, 1000
        model = make_model(cfg, seed=29, bias=True)
        r = np.random.default_rng(31)
        emb = rounded(r.standard_normal((V, cfg["hidden"])), dtype)
        gam = rounded(1 + 0.1 * r.standard_normal(cfg["hidden"]), dtype)
        lm = rounded(r.standard_normal((V, cfg["hidden"])) / 32, dtype)
        prompt = r.integers(3, V, size=(B, T)).astype(np.int32)
        embd, gd, lmd = to_dev(emb, dtype), to_dev(gam, dtype), to_dev(lm, dtype)
        lcfg = tp.local_cfg(dict(head_num=cfg["head_num"], kv_head_num=cfg["kv_head_num"], head_size=cfg["head_size"], inter=cfg["inter"]), world)

        def build(lc, w_of, tpw, tpr):
            dc = mod.DecoderConfig(cfg["hidden"], lc["head_num"], lc["kv_head_num"], cfg["head_size"], lc["inter"], cfg["layers"], cfg["max_seq"], B, 0, 0,
                                   128, cfg["eps"], cfg["head_size"], cfg["base"], tpw, tpr)
            d_ = mod.Decoder(dc, dev)
            for l, w in enumerate(model["layers"]):
                s = w_of(w)
                d_.set_layer(l, dict(g1=to_dev(s["g1"], dtype), qkv=to_dev(s["wqkv"], dtype), qkv_bias=to_dev(s["bqkv"], dtype), o=to_dev(s["wo"], dtype),
                                     o_bias=to_dev(s["bo"], dtype), g2=to_dev(s["g2"], dtype), gate_up=to_dev(s["wgu"], dtype), down=to_dev(s["wd"], dtype)))
            return d_

        full = build(dict(head_num=cfg["head_num"], kv_head_num=cfg["kv_head_num"], inter=cfg["inter"]), lambda w: w, 1, 0)
        kc = torch.zeros((cfg["layers"], B, cfg["kv_head_num"], cfg["max_seq"], cfg["head_size"]), dtype=torch.float32, device=dev)
        want, _ = full.generate(prompt, embd, gd, lmd, kc, torch.zeros_like(kc), N, top_k=1, end_id=-1)
        dec = build(lcfg, lambda w: tp.shard_layer(w, cfg, rank, world), world, rank)
        dec.tp_attach(dist)
        head = tp.VocabShardedHead(mod, dec, tp.shard_lm_head_rows(lmd, rank, world), V, rank, world, B, 1, dev)
        kcs = torch.zeros((cfg["layers"], B, lcfg["kv_head_num"], cfg["max_seq"], cfg["head_size"]), dtype=torch.float32, device=dev)
        got = tp.generate_tp(mod, dec, dist, prompt, embd, gd, head, kcs, torch.zeros_like(kcs), N, end_id=-1)
        torch.cuda.synchronize()
        assert dec.tp_error() == 0, "a peer never signalled (exchange timed out)"
        assert np.array_equal(got, want), f"tensor-parallel ids {got.tolist()} vs un-sharded {want.tolist()}"
        print(f"[rank {rank}] TP-{world} generation loop (prefill_tp + vocab-sharded head + step_tp), {B} x {N} greedy tokens = b200_generate un-sharded: OK", flush=True)
    except (AssertionError, mod.B200Error) as e:
        ok = False
        print(f"[rank {rank}] TP generation loop FAILED: {e}", flush=True)
    # ---- vocab-sharded LM head + top-k + sampling: bit-identical to the un-sharded tail on the same hidden state
    try:
        V, hsz, B, K = 32000, cfg["hidden"], 3, 5
        g = torch.Generator(device=dev)
        g.manual_seed(5)  # same on every rank
        lm = torch.empty((V, hsz), dtype=torch.bfloat16, device=dev).normal_(0.0, 0.02, generator=g)
        hid = torch.randn((B, hsz), device=dev, generator=g).to(torch.bfloat16)
        gam = (1 + 0.1 * torch.randn(hsz, device=dev, generator=g)).to(torch.bfloat16)
        dcfg = mod.DecoderConfig(cfg["hidden"], 8 // world, 8 // world, 128, cfg["inter"] // world, 1, 64, B, 2, 0, 128, 1e-6, 128, 10000.0, world, rank)
        d2 = mod.Decoder(dcfg, dev)
        i32, f32 = torch.int32, torch.float32
        full = dict(logits=torch.empty((B, V), dtype=f32, device=dev), tmp_ids=torch.empty((B, 8, K), dtype=i32, device=dev),
                    tmp_vals=torch.empty((B, 8, K), dtype=f32, device=dev), topk_ids=torch.empty((B, K), dtype=i32, device=dev),
                    topk_vals=torch.empty((B, K), dtype=f32, device=dev), seq_len=torch.full((B,), 9, dtype=i32, device=dev),
                    finished=torch.zeros(B, dtype=torch.uint8, device=dev), output_id=torch.zeros(B, dtype=i32, device=dev))
        d2.lm_head_topk_sample(hid, gam, lm, full, K, 7, 2)
        head = tp.VocabShardedHead(mod, d2, tp.shard_lm_head_rows(lm, rank, world), V, rank, world, B, K, dev)
        seq2, fin2, out2 = torch.full((B,), 9, dtype=i32, device=dev), torch.zeros(B, dtype=torch.uint8, device=dev), torch.zeros(B, dtype=i32, device=dev)
        head.run(dist, hid, gam, seq2, fin2, out2, 7, 2)
        torch.cuda.synchronize()
        same = (torch.equal(head.topk_ids, full["topk_ids"]) and torch.equal(head.topk_vals, full["topk_vals"]) and torch.equal(out2, full["output_id"])
                and torch.equal(seq2, full["seq_len"]) and torch.equal(fin2, full["finished"]))
        assert same, f"sharded ids {head.topk_ids.tolist()} vs {full['topk_ids'].tolist()}"
        print(f"[rank {rank}] TP-{world} vocab-sharded LM head: top-k ids / values / sampled ids bit-identical to the un-sharded tail: OK", flush=True)
    except AssertionError as e:
        ok = False
        print(f"[rank {rank}] vocab-sharded LM head FAILED: {e}", flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    code = 1 if int(flag.item()) else 0
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(code)  # skip NCCL teardown: it is not what is under test and can outlast the job


if __name__ == "__main__":
    main()
