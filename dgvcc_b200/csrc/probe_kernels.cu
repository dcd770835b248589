// Chip-wide issue-rate probes for the roofline denominators MEASURED_PEAKS.json lacks
// (SURVEY.md section 8d): MUFU.EX2 results/s and FFMA/s.  Eight independent dependent-chains
// per thread keep each pipe saturated; 8 CTAs of 256 threads per SM fill every scheduler.
#include "common.cuh"
#include "../../include/dgvcc_b200.h"

namespace dgvcc {

constexpr int PROBE_THREADS = 256;
constexpr int PROBE_CHAINS = 8;

__global__ void __launch_bounds__(PROBE_THREADS) probe_ex2_kernel(float* sink, int iters) {
    float v[PROBE_CHAINS];
#pragma unroll
    for (int c = 0; c < PROBE_CHAINS; ++c) v[c] = -1e-3f * (float)(threadIdx.x + c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < PROBE_CHAINS; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[c]));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < PROBE_CHAINS; ++c) s += v[c];
    if (s == 123.456f) sink[0] = s;  // keeps the chains alive, never true in practice
}

__global__ void __launch_bounds__(PROBE_THREADS) probe_ffma_kernel(float* sink, int iters) {
    float v[PROBE_CHAINS];
    const float a = 0.999f, b = 1e-3f;
#pragma unroll
    for (int c = 0; c < PROBE_CHAINS; ++c) v[c] = (float)(threadIdx.x + c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < PROBE_CHAINS; ++c) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[c]) : "f"(a), "f"(b));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < PROBE_CHAINS; ++c) s += v[c];
    if (s == 123.456f) sink[0] = s;
}

static int probe_grid() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * 8;
}

}  // namespace dgvcc

extern "C" int dgvcc_probe_ex2(float* sink, int iters, int64_t* ops_out, void* stream) {
    if (!sink || iters <= 0) return DGVCC_ERR_ARG;
    const int grid = dgvcc::probe_grid();
    dgvcc::probe_ex2_kernel<<<grid, dgvcc::PROBE_THREADS, 0, (cudaStream_t)stream>>>(sink, iters);
    if (ops_out) *ops_out = (int64_t)grid * dgvcc::PROBE_THREADS * dgvcc::PROBE_CHAINS * iters;
    return (int)cudaGetLastError();
}

extern "C" int dgvcc_probe_ffma(float* sink, int iters, int64_t* ops_out, void* stream) {
    if (!sink || iters <= 0) return DGVCC_ERR_ARG;
    const int grid = dgvcc::probe_grid();
    dgvcc::probe_ffma_kernel<<<grid, dgvcc::PROBE_THREADS, 0, (cudaStream_t)stream>>>(sink, iters);
    if (ops_out) *ops_out = (int64_t)grid * dgvcc::PROBE_THREADS * dgvcc::PROBE_CHAINS * iters;
    return (int)cudaGetLastError();
}
