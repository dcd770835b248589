// Microbenchmark: MUFU.EX2 throughput when interleaved with F fp32 ops per exp and L broadcast
// LDS.128 per 8 exps, at several occupancies.  nvcc -arch=sm_100a -O3 -o mufu_mix mufu_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int F, int L, int P>  // F fp32 per exp, L LDS.128 per P exps (P independent chains)
__global__ void __launch_bounds__(128) mix(float* out, int iters, float k0, float k1) {
    __shared__ float4 sm[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = make_float4(i * 1e-3f, i * 2e-3f, i * 3e-3f, i * 4e-3f);
    __syncthreads();
    float acc[P], base[P];
#pragma unroll
    for (int p = 0; p < P; ++p) { acc[p] = 0.f; base[p] = -1e-3f * (threadIdx.x + p); }
    for (int it = 0; it < iters; ++it) {
        float4 v[L > 0 ? L : 1];
#pragma unroll
        for (int l = 0; l < L; ++l) v[l] = sm[(it * L + l) & 255];
        float x = k1;
#pragma unroll
        for (int l = 0; l < L; ++l) x += v[l].x;  // consume the loads (L adds per P exps, negligible)
#pragma unroll
        for (int p = 0; p < P; ++p) {
            float t = base[p];
            if (F >= 1) t = __fadd_rn(t, x);
            if (F >= 2) t = __fmaf_rn(t, k0, k1);
            if (F >= 3) t = __fmul_rn(t, 1.4426950408889634f);
            if (F >= 5) t = __fadd_rn(t, k1);
            float e = ex2(t);
            if (F >= 4) acc[p] = __fmaf_rn(e, k0, acc[p]); else acc[p] += e * 0.f + 0.f, base[p] = e * 1e-9f + base[p];
        }
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p) s += acc[p] + base[p];
    if (s == 12345.678f) out[0] = s;
}

template <int F, int L, int P>
void run(int ctas_per_sm, float peak) {
    int sms = 148;
    float* out; cudaMalloc(&out, 16);
    int iters = 4000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int grid = sms * ctas_per_sm;
    mix<F, L, P><<<grid, 128>>>(out, iters, 0.999f, -0.01f);
    cudaEventRecord(a);
    mix<F, L, P><<<grid, 128>>>(out, iters, 0.999f, -0.01f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double exps = (double)grid * 128 * iters * P;
    printf("F=%d L=%d P=%2d warps/SM=%2d : %.2f Texp/s (%.1f%% of %.2f) %s\n", F, L, P, ctas_per_sm * 4, exps / ms / 1e9,
           100.0 * exps / ms / 1e9 / peak, peak, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    const float peak = 4.65f;
    for (int c : {2, 4, 8}) {
        run<0, 0, 8>(c, peak);
        run<4, 0, 8>(c, peak);
        run<5, 0, 8>(c, peak);
        run<4, 1, 8>(c, peak);
        run<4, 3, 8>(c, peak);
        run<4, 3, 16>(c, peak);
        run<4, 6, 8>(c, peak);
        run<2, 3, 8>(c, peak);
        run<3, 0, 8>(c, peak);
    }
    return 0;
}
