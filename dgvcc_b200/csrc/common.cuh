// Shared device helpers for the dgvcc_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DGVCC_OK 0
#define DGVCC_ERR_ARG (-1)        // bad argument (null pointer, non-positive size ...)
#define DGVCC_ERR_WORKSPACE (-2)  // caller-owned workspace too small
#define DGVCC_ERR_UNSUPPORTED (-3)

#define DGVCC_RETURN_IF_CUDA(expr)            \
    do {                                      \
        cudaError_t _e = (expr);              \
        if (_e != cudaSuccess) return (int)_e; \
    } while (0)

namespace dgvcc {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr float LOG2E = 1.4426950408889634f;

// One MUFU.EX2; results below 2^-126 flush to zero (posteriors < 1.2e-38, far under the 1e-30 atol).
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace dgvcc
