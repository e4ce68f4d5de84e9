// Compile check of the shim's device-side printer (src/utils/cuda_debug_utils.cuh; the reference's is cuda_debug_utils.cuh:7-25) for every
// element type the launchers are instantiated for.
#include "src/utils/cuda_debug_utils.cuh"
template __global__ void print_data<float>(float *, bool);
template __global__ void print_data<__half>(__half *, bool);
template __global__ void print_data<__nv_bfloat16>(__nv_bfloat16 *, bool);
void print_first_elements(float *x) { print_data<<<1, 1>>>(x, true); }
