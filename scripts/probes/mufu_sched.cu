// Microbenchmark 2: does the ORDER of the per-pair instructions matter for MUFU throughput?
// Same work as the BL inner loop (per point: 3 broadcast LDS, then 16 pairs of FADD, FFMA, FMUL, EX2, FADD);
// variant 0 lets ptxas schedule, variant 1 pins a software-pipelined order with volatile asm
// (arguments of pair i+2 computed between the EX2 of pair i+1 and the accumulate of pair i).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NP = 16;

__device__ __forceinline__ float ex2v(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float addv(float a, float b) { float y; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y; }
__device__ __forceinline__ float fmav(float a, float b, float c) { float y; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
__device__ __forceinline__ float mulv(float a, float b) { float y; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y; }

template <int VARIANT>
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
        if (VARIANT == 0) {
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float dis = __fadd_rn(yd[p & 7], p < 8 ? xd0 : xd1);
                const float d = __fmaf_rn(dis, inv_s, na[p]);
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmul_rn(d, 1.4426950408889634f)));
                z[p] += e;
            }
        } else {
            float t[NP], e[NP];
            auto arg = [&](int p) { t[p] = mulv(fmav(addv(yd[p & 7], p < 8 ? xd0 : xd1), inv_s, na[p]), 1.4426950408889634f); };
            arg(0); arg(1);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                e[p] = ex2v(t[p]);
                if (p + 2 < NP) arg(p + 2);
                if (p >= 1) z[p - 1] = addv(z[p - 1], e[p - 1]);
            }
            z[NP - 1] = addv(z[NP - 1], e[NP - 1]);
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
    printf("variant %d warps/SM %2d: %.2f Texp/s (%.1f%% of 4.65) %s\n", V, ctas * 4, exps / ms / 1e9, 100 * exps / ms / 1e9 / 4.65,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int c : {2, 3, 4, 6, 8}) { run<0>(c); run<1>(c); }
    return 0;
}
