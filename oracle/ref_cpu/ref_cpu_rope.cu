// TEST INFRASTRUCTURE: exposes the CPU reference loop of the reference's own unit test
// tests/unit_tests/test_qkv_bias_and_rope.cu (included from where it lies under $(REF), main() renamed) through a C symbol.
#define main ref_test_main_rope
#define checkResult ref_checkResult_rope
#define checkResults ref_checkResults_rope
#define runTest ref_runTest_rope
#include "tests/unit_tests/test_qkv_bias_and_rope.cu"
#undef main
extern "C" {
void refcpu_qkv_rope(float *q, float *k, float *v, float *qkv, const int *padding_offset, const int *history, const int *input_len, int batch, int seq_len, int token_num, int head_num, int kv_head_num, int head_size, int rot_dim, float base) { CPUfunc(q, k, v, qkv, nullptr, padding_offset, history, input_len, batch, seq_len, token_num, head_num, kv_head_num, head_size, rot_dim, base); }
}
