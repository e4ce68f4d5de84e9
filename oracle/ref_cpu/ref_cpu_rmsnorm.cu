// TEST INFRASTRUCTURE: exposes the CPU reference loop of the reference's own unit test
// tests/unit_tests/test_rmsnorm.cu (included from where it lies under $(REF), main() renamed) through a C symbol.
#define main ref_test_main_rmsnorm
#define checkResult ref_checkResult_rmsnorm
#define checkResults ref_checkResults_rmsnorm
#define runTest ref_runTest_rmsnorm
#include "tests/unit_tests/test_rmsnorm.cu"
#undef main
extern "C" {
void refcpu_rmsnorm(float *x, float *gamma, float eps, int hidden, int tokens) { CPUfusedresidandRMSNorm(x, gamma, eps, hidden, tokens); }
}
