// TEST INFRASTRUCTURE: exposes the CPU reference loop of the reference's own unit test
// tests/unit_tests/test_build_causal_mask.cu (included from where it lies under $(REF), main() renamed) through a C symbol.
#define main ref_test_main_causal_mask
#define checkResult ref_checkResult_causal_mask
#define checkResults ref_checkResults_causal_mask
#define runTest ref_runTest_causal_mask
#include "tests/unit_tests/test_build_causal_mask.cu"
#undef main
extern "C" {
void refcpu_causal_mask(float *mask, const int *q_lens, const int *k_lens, int max_q, int max_k, int batch) { CPUbuildCausalMask(mask, q_lens, k_lens, max_q, max_k, batch); }
}
