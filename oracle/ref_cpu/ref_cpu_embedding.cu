// TEST INFRASTRUCTURE: exposes the CPU reference loop of the reference's own unit test
// tests/unit_tests/test_input_embedding.cu (included from where it lies under $(REF), main() renamed) through a C symbol.
#define main ref_test_main_embedding
#define checkResult ref_checkResult_embedding
#define checkResults ref_checkResults_embedding
#define runTest ref_runTest_embedding
#include "tests/unit_tests/test_input_embedding.cu"
#undef main
extern "C" {
void refcpu_embedding(const int *ids, float *out, float *table, int tokens, int hidden, int vocab) { cpuEmbedding(ids, out, table, tokens, hidden, vocab); }
}
