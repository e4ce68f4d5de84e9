// attention_decode.cuh -- internal interface of the fused decode attention (used by the C ABI entry and the engine).
#pragma once
#include "common.cuh"

namespace b200 {

struct DecodeAttnArgs {
    const void *qkv;   // [B, H+2Hkv, d]
    const void *bias;  // [(H+2Hkv)*d] or NULL
    void *k_cache;     // layer base: [B, Hkv, S, d]
    void *v_cache;
    void *out;         // [B, H*d]
    float *partials;   // [B*Hkv][nsplit][G][kAttnPartStride]: per split and q head the un-normalised output o[d], then (max, sum)
    unsigned int *tickets;  // [B*Hkv], zero-initialised, self-resetting
    const float2 *rope_cs;  // optional (cos, sin) table [max_seq_len][rot_dim/2] made by launch_rope_table(); NULL: compute
    int batch, head_num, kv_head_num, head_size, max_seq_len, step;
    const int *steps;  // optional device int[batch]: per-row step (ragged batches), clamped to [1, step]; `step` is then their maximum (the split plan)
    int apply_rope, rot_dim;
    float rot_base;
    int nsplit, chunk;
    // paged cache (optional): k_cache / v_cache are then the LAYER base of a page pool [num_pages, Hkv, kAttnPageSize, d] and position p of
    // batch row b lives in page block_table[b * max_pages + p / kAttnPageSize], row p % kAttnPageSize; max_seq_len = max_pages * kAttnPageSize
    const int *block_table;
    int max_pages;
    int ll_merge;  // 1: `partials` is a ZERO-INITIALISED region of 8-byte {value, flag} words owned by the caller (the engine's scratch): the
                   // splits publish their partials as flagged words and the LAST split of every head polls them, merges and clears
                   // them again -- no fence, no ticket, and every other CTA leaves right after its stores.  0: floats + ticket.
    int prefetch;  // 1: cached K/V rows may be requested before griddepcontrol.wait (fused engine only: the kernel in front
                   // of this one does not write the cache)
};

// positions per page of the paged cache: one page of one kv head = one 16 KiB ring stage of the decode kernel (16-bit), two for fp32
constexpr int kAttnPageSize = 64;
// most KV splits of a (batch row, kv head): the merger keeps the other splits' words in registers
constexpr int kAttnMaxSplits = 8;
// floats per (split, q head) record of the partials: o[128], max, sum, 2 pad (records stay 16-byte aligned)
constexpr int kAttnPartStride = 132;
inline int attn_part_stride(int head_size) { return head_size + 4; }
// number of KV splits (and positions per split) for a decode step
int decode_attn_plan(int batch, int kv_head_num, int step, int *chunk);
size_t decode_attn_partials_floats(int batch, int head_num, int kv_head_num, int head_size, int max_splits);  // x 2 for ll_merge words
int launch_decode_attn(const DecodeAttnArgs &a, int dtype, cudaStream_t st);

// (cos, sin) of pos / base^(2j / rot_dim) for pos < positions, j < rot_dim / 2 -- the values the kernels compute on the fly
int launch_rope_table(float2 *table, int positions, int rot_dim, float rot_base, cudaStream_t st);

}  // namespace b200
