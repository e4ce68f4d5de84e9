// TEST INFRASTRUCTURE: exposes the CPU reference loop of the reference's own unit test
// tests/unit_tests/test_silu_and_mul.cu (included from where it lies under $(REF), main() renamed) through a C symbol.
#define main ref_test_main_swiglu
#define checkResult ref_checkResult_swiglu
#define checkResults ref_checkResults_swiglu
#define runTest ref_runTest_swiglu
#include "tests/unit_tests/test_silu_and_mul.cu"
#undef main
extern "C" {
void refcpu_swiglu(float *in, float *out, int batch, int inter) { CPUSwiGLU<float>(in, out, batch, inter); }
}
