"""Tensor-parallel sharding of one decoder layer (new -- the reference is single-GPU).

Weights are in the engine's packed [N,K] layout.  Rank r of P owns
  * q heads  [r*H/P, (r+1)*H/P)  and kv heads [r*Hkv/P, (r+1)*Hkv/P): the matching ROWS of Wqkv (and of its bias),
  * the matching K-COLUMNS of Wo (row-sharded linear: its output is a partial sum, all-reduced),
  * FFN columns [r*I/P, (r+1)*I/P): the matching gate rows AND up rows of Wgate_up (paired so SwiGLU stays local),
    and the matching K-columns of Wdown (partial sum, all-reduced),
  * KV-cache heads [r*Hkv/P, ...): cache [L, B, Hkv/P, S, d].
Norm gammas, the residual stream and the o-bias are replicated (the bias is added once, after the reduce).
Works on numpy arrays and torch tensors alike (pure slicing + concatenation)."""


def _cat(parts):
    try:
        import torch

        if isinstance(parts[0], torch.Tensor):
            return torch.cat(list(parts), dim=0).contiguous()
    except ImportError:
        pass
    import numpy as np

    return np.ascontiguousarray(np.concatenate(list(parts), axis=0))


def _contig(a):
    if hasattr(a, "contiguous"):
        return a.contiguous()
    import numpy as np

    return np.ascontiguousarray(a)


def check_divisible(head_num, kv_head_num, inter, world):
    if head_num % world or kv_head_num % world or inter % world:
        raise ValueError(f"tensor-parallel degree {world} must divide head_num {head_num}, kv_head_num {kv_head_num} and inter {inter}")


def shard_qkv_rows(wqkv, head_num, kv_head_num, head_size, rank, world):
    """Rows of Wqkv [(H+2Hkv)*d, K] (or entries of the qkv bias [(H+2Hkv)*d]) owned by `rank`: its q heads, k heads, v heads."""
    hq, hk, d = head_num // world, kv_head_num // world, head_size
    q0, k0, v0 = 0, head_num * d, (head_num + kv_head_num) * d
    return _cat([wqkv[q0 + rank * hq * d: q0 + (rank + 1) * hq * d], wqkv[k0 + rank * hk * d: k0 + (rank + 1) * hk * d],
                 wqkv[v0 + rank * hk * d: v0 + (rank + 1) * hk * d]])


def shard_o_cols(wo, head_num, head_size, rank, world):
    """K-columns of Wo [h, H*d] that multiply this rank's attention heads."""
    n = head_num // world * head_size
    return _contig(wo[:, rank * n:(rank + 1) * n])


def shard_gate_up_rows(wgu, inter, rank, world):
    """Rows of Wgate_up [2I, K]: this rank's gate rows followed by its up rows."""
    n = inter // world
    return _cat([wgu[rank * n:(rank + 1) * n], wgu[inter + rank * n: inter + (rank + 1) * n]])


def shard_down_cols(wd, inter, rank, world):
    n = inter // world
    return _contig(wd[:, rank * n:(rank + 1) * n])


def shard_kv_cache(cache, kv_head_num, rank, world):
    """cache [L, B, Hkv, S, d] -> [L, B, Hkv/P, S, d]"""
    n = kv_head_num // world
    return _contig(cache[:, :, rank * n:(rank + 1) * n])


def shard_layer(w, cfg, rank, world):
    """w: dict(g1, wqkv, bqkv|None, wo, bo|None, g2, wgu, wd) in [N,K] layout; cfg: dict(head_num, kv_head_num, head_size, inter).
    Returns this rank's dict with the same keys.  The o bias stays whole: it is applied once, after the all-reduce."""
    H, Hkv, d, I = cfg["head_num"], cfg["kv_head_num"], cfg["head_size"], cfg["inter"]
    check_divisible(H, Hkv, I, world)
    out = dict(w)
    out["wqkv"] = shard_qkv_rows(w["wqkv"], H, Hkv, d, rank, world)
    out["bqkv"] = None if w.get("bqkv") is None else shard_qkv_rows(w["bqkv"], H, Hkv, d, rank, world)
    out["wo"] = shard_o_cols(w["wo"], H, d, rank, world)
    out["wgu"] = shard_gate_up_rows(w["wgu"], I, rank, world)
    out["wd"] = shard_down_cols(w["wd"], I, rank, world)
    return out


def local_cfg(cfg, world):
    """Per-rank shape: what goes into b200_decoder_config_t (head_num, kv_head_num, inter_size are per-rank counts)."""
    check_divisible(cfg["head_num"], cfg["kv_head_num"], cfg["inter"], world)
    out = dict(cfg)
    out["head_num"], out["kv_head_num"], out["inter"] = cfg["head_num"] // world, cfg["kv_head_num"] // world, cfg["inter"] // world
    return out


def decode_step_tp(layers, hidden, attn_block, ffn_block, fold, all_reduce):
    """The tensor-parallel decode step: the call sequence bench.py and the C ABI (b200_decoder_attn_block / _ffn_block / _fold)
    share.  attn_block(l, hidden, pending) -> partial; ffn_block(l, pending) -> partial; all_reduce(t) sums t over ranks in place."""
    pending = None
    for l in range(layers):
        y = attn_block(l, hidden, pending)
        all_reduce(y)
        z = ffn_block(l, y)
        all_reduce(z)
        pending = z
    return fold(hidden, pending)


# ---------------------------------------------------------------- vocab-sharded LM head ("LM head vocab-sharded + allgather of
# per-rank top-k").  Rank r owns rows [r*V/P, (r+1)*V/P) of lm_head [V, h]: it computes its slice of the logits, takes a LOCAL top-k, adds its
# row offset to the ids and all-gathers k (value, id) pairs -- 2*k*B numbers per rank instead of V/P logits.  The global top-k is the top-k of
# the P*k candidates under the same total order the single-GPU kernel uses (value descending, ties to the LOWER id), so ids and values are
# bit-identical to the un-sharded result: every global winner is a local winner of the rank that owns it.
def vocab_range(vocab, rank, world):
    if vocab % world:
        raise ValueError(f"tensor-parallel degree {world} must divide the vocabulary {vocab}")
    n = vocab // world
    return rank * n, (rank + 1) * n


def shard_lm_head_rows(lm_head, rank, world):
    """Rows of lm_head [V, h] owned by `rank`."""
    lo, hi = vocab_range(lm_head.shape[0], rank, world)
    return _contig(lm_head[lo:hi])


def merge_topk(vals, ids, k):
    """vals, ids: numpy [P, B, k'] -- every rank's local top-k' (ids already GLOBAL).  Returns (ids [B, k] int32, vals [B, k] float32): the k
    best candidates per row, value descending, ties to the lower id."""
    import numpy as np

    vals, ids = np.asarray(vals, dtype=np.float32), np.asarray(ids, dtype=np.int64)
    P, B, kk = vals.shape
    if k > P * kk:
        raise ValueError(f"cannot take {k} winners out of {P} x {kk} candidates")
    v = np.transpose(vals, (1, 0, 2)).reshape(B, P * kk)
    i = np.transpose(ids, (1, 0, 2)).reshape(B, P * kk)
    out_i, out_v = np.empty((B, k), np.int32), np.empty((B, k), np.float32)
    for b in range(B):
        order = np.lexsort((i[b], -v[b].astype(np.float64)))[:k]  # primary key: value descending; secondary: id ascending
        out_i[b], out_v[b] = i[b][order], v[b][order]
    return out_i, out_v


class VocabShardedHead:
    """The sampling tail of a tensor-parallel decode step with the LM head vocab-sharded (reference src/models/llama/llama.cpp:247-311 on
    one GPU): final RMSNorm + this rank's rows of lm_head (b200_lm_head_topk_sample on the shard, local top-k) -> ids made global ->
    ONE NCCL all-gather of k (value, id) pairs per row and rank -> b200_topk over the P*k candidates (value descending, ties to the
    lower candidate index = the lower global id: shards are ascending id ranges and a local top-k lists equal values by ascending id)
    -> b200_sampling.  Every buffer is allocated once, so the whole tail can be captured in a CUDA graph.  Bit-identical to the
    un-sharded tail given the same hidden state (tests/tp_engine_check.py)."""

    def __init__(self, mod, dec, lm_head_shard, vocab, rank, world, batch, k, device):
        import torch

        self.mod, self.dec, self.lm, self.vocab, self.rank, self.world, self.k = mod, dec, lm_head_shard, vocab, rank, world, k
        self.lo, _ = vocab_range(vocab, rank, world)
        vl, nb = lm_head_shard.shape[0], mod.TOPK_BLOCKS
        i32, f32 = torch.int32, torch.float32
        self.local = dict(logits=torch.empty((batch, vl), dtype=f32, device=device), tmp_ids=torch.empty((batch, nb, k), dtype=i32, device=device),
                          tmp_vals=torch.empty((batch, nb, k), dtype=f32, device=device), topk_ids=torch.empty((batch, k), dtype=i32, device=device),
                          topk_vals=torch.empty((batch, k), dtype=f32, device=device))
        self.mine = torch.empty((batch, 2 * k), dtype=i32, device=device)            # [vals as int32 bits | global ids]
        self.all = torch.empty((world, batch, 2 * k), dtype=i32, device=device)
        self.cand_vals = torch.empty((batch, world * k), dtype=f32, device=device)
        self.cand_ids = torch.empty((batch, world * k), dtype=i32, device=device)
        self.tmp_i = torch.empty((batch, nb, k), dtype=i32, device=device)
        self.tmp_v = torch.empty((batch, nb, k), dtype=f32, device=device)
        self.idx = torch.empty((batch, k), dtype=i32, device=device)
        self.topk_vals = torch.empty((batch, k), dtype=f32, device=device)
        self.topk_ids = torch.empty((batch, k), dtype=i32, device=device)

    def run(self, dist, hidden, final_gamma, seq_len, finished, output_id, step, end_id):
        import torch

        mod, k, B, P = self.mod, self.k, hidden.shape[0], self.world
        self.dec.lm_head_topk_sample(hidden, final_gamma, self.lm, self.local, k, step, end_id)  # no output_id: logits + local top-k only
        self.mine[:, :k].copy_(self.local["topk_vals"].view(torch.int32))
        torch.add(self.local["topk_ids"], self.lo, out=self.mine[:, k:])
        dist.all_gather_into_tensor(self.all, self.mine)
        self.cand_vals.view(torch.int32).view(B, P, k).copy_(self.all[:, :, :k].permute(1, 0, 2))  # bit copy: values stay float32
        self.cand_ids.view(B, P, k).copy_(self.all[:, :, k:].permute(1, 0, 2))
        p = mod.ptr
        mod.check(mod.lib().b200_topk(p(self.cand_vals), p(self.tmp_i), p(self.tmp_v), p(self.idx), p(self.topk_vals), B, P * k, k, mod.F32, mod.stream()))
        torch.gather(self.cand_ids, 1, self.idx.long(), out=self.topk_ids)
        mod.check(mod.lib().b200_sampling(p(self.topk_ids), p(self.topk_vals), p(seq_len), p(finished), p(output_id), B, k, step, end_id, self.vocab,
                                          mod.F32, mod.stream()))


def generate_tp(mod, dec, dist, prompt_ids, embedding, final_gamma, head, k_cache, v_cache, max_new_tokens, end_id=2):
    """The generation loop on a tensor-parallel engine (what b200_generate does on one GPU, src/models/llama/llama.cpp:165-398 intended):
    embedding -> Decoder.prefill_tp (one NCCL all-reduce per attention and per MLP block) -> last prompt row -> vocab-sharded LM head
    (VocabShardedHead) -> then per token: embedding of the sampled id -> Decoder.step_tp (fused NVLink exchange; call dec.tp_attach
    first) -> head.  prompt_ids: host int [B, T] (equal lengths); k_cache / v_cache: this rank's head shard [L, B, Hkv / P, S, d];
    `head`: a VocabShardedHead built for batch B.  Every rank returns the same ids [B, max_new_tokens] (numpy); everything from a
    sequence's first end_id on reads end_id, as b200_generate reports it."""
    import numpy as np
    import torch

    dev = embedding.device
    prompt = np.ascontiguousarray(np.asarray(prompt_ids, dtype=np.int32))
    B, T = prompt.shape
    x = mod.input_embedding(torch.from_numpy(prompt.reshape(-1)).to(dev), embedding)
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    dec.prefill_tp(x, k_cache, v_cache, il, torch.zeros_like(il), il, T, dist)
    hidden = x.view(B, T, -1)[:, -1].contiguous()
    seq_len, finished = il.clone(), torch.zeros(B, dtype=torch.uint8, device=dev)
    out_id = torch.zeros(B, dtype=torch.int32, device=dev)
    step, out = T, []
    for i in range(max_new_tokens):
        head.run(dist, hidden, final_gamma, seq_len, finished, out_id, step, end_id)
        out.append(out_id.clone())
        if i + 1 == max_new_tokens:
            break
        step += 1
        hidden = mod.input_embedding(out_id, embedding)
        dec.step_tp(hidden, k_cache, v_cache, step)
    ids = torch.stack(out, dim=1).cpu().numpy()
    for b in range(B):
        hit = np.where(ids[b] == end_id)[0]
        if len(hit):
            ids[b, hit[0]:] = end_id
    return ids
