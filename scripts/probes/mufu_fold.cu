// Microbenchmark 4: instruction count per pair.  Variant 0 = the BL inner loop as shipped (FADD, FFMA, FMUL,
// EX2, FADD per pair); variant 1 folds log2(e) into the FFMA constants (FADD, FFMA, EX2, FADD per pair).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NP = 16;
__device__ __forceinline__ float ex2_mufu(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int V>
__global__ void __launch_bounds__(128) k(float* out, int iters, float inv_s) {
    __shared__ float4 sm[2][128];
    __shared__ float2 sx[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        sm[0][i] = make_float4(i * 1.f, i * 2.f, i * 3.f, i * 4.f);
        sm[1][i] = make_float4(i * 5.f, i * 6.f, i * 7.f, i * 8.f);
        sx[i] = make_float2(i * 0.5f, i * 0.25f);
    }
    __syncthreads();
    float z[NP], na[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) { z[p] = 0.f; na[p] = -1e-3f * (threadIdx.x + p); }
    const float c0 = -2.f * threadIdx.x, c1 = -2.f * (threadIdx.x + 32), cc0 = 1.f * threadIdx.x, cc1 = 2.f * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const int i = it & 127;
        const float2 xs = sx[i];
        const float4 a = sm[0][i], b = sm[1][i];
        const float yd[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const float xd0 = (xs.x * c0 + xs.y) + cc0, xd1 = (xs.x * c1 + xs.y) + cc1;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float dis = __fadd_rn(yd[p & 7], p < 8 ? xd0 : xd1);
            float t;
            if (V == 0) t = __fmul_rn(__fmaf_rn(dis, inv_s, na[p]), 1.4426950408889634f);
            else t = __fmaf_rn(dis, inv_s, na[p]);
            z[p] += ex2_mufu(t);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < NP; ++p) s += z[p];
    if (s == 12345.678f) out[0] = s;
}
template <int V> void run(int ctas) {
    float* out; cudaMalloc(&out, 16);
    const int iters = 3000, grid = 148 * ctas;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<V><<<grid, 128>>>(out, iters, -0.0078125f);
    cudaEventRecord(a);
    k<V><<<grid, 128>>>(out, iters, -0.0078125f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double exps = (double)grid * 128 * iters * NP;
    printf("variant %d warps/SM %2d: %.2f Texp/s (%.1f%% of the 4.65 MUFU peak)\n", V, ctas * 4, exps / ms / 1e9, 100 * exps / ms / 1e9 / 4.65);
    cudaFree(out);
}
int main() { for (int c : {3, 4, 6, 8}) { run<0>(c); run<1>(c); } return 0; }
