// TEST INFRASTRUCTURE: exposes the CPU reference loop of the reference's own unit test
// tests/unit_tests/test_linear.cu (included from where it lies under $(REF), main() renamed) through a C symbol.
#define main ref_test_main_linear
#define checkResult ref_checkResult_linear
#define checkResults ref_checkResults_linear
#define runTest ref_runTest_linear
#include "tests/unit_tests/test_linear.cu"
#undef main
extern "C" {
void refcpu_linear(float *x, float *w, float *y, int M, int K, int N) { CPUlinear(x, w, y, M, K, N); }
}
