// TEST INFRASTRUCTURE: exposes the CPU reference loop of the reference's own unit test
// tests/unit_tests/test_add_residual.cu (included from where it lies under $(REF), main() renamed) through a C symbol.
#define main ref_test_main_add_residual
#define checkResult ref_checkResult_add_residual
#define checkResults ref_checkResults_add_residual
#define runTest ref_runTest_add_residual
#include "tests/unit_tests/test_add_residual.cu"
#undef main
extern "C" {
void refcpu_add_residual(float *residual, float *out, int hidden, int tokens) { CPUresidual(residual, out, hidden, tokens); }
}
