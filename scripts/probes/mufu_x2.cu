// Microbenchmark 5: packed fp32x2 arithmetic (Blackwell add/fma.f32x2) around MUFU.EX2.
// Variant 0 = scalar (FADD, FFMA, EX2, FADD per pair); variant 1 = f32x2 (FADD2, FFMA2, 2 EX2, FADD2 per 2 pairs).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NP = 16;
__device__ __forceinline__ float ex2_mufu(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned long long pack(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

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
    const float c0 = -2.f * threadIdx.x, c1 = -2.f * (threadIdx.x + 32), cc0 = 1.f * threadIdx.x, cc1 = 2.f * threadIdx.x;
    if (V == 0) {
        float z[NP], na[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) { z[p] = 0.f; na[p] = -1e-3f * (threadIdx.x + p); }
        for (int it = 0; it < iters; ++it) {
            const int i = it & 127;
            const float2 xs = sx[i];
            const float4 a = sm[0][i], b = sm[1][i];
            const float yd[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            const float xd0 = (xs.x * c0 + xs.y) + cc0, xd1 = (xs.x * c1 + xs.y) + cc1;
#pragma unroll
            for (int p = 0; p < NP; ++p)
                z[p] += ex2_mufu(__fmaf_rn(__fadd_rn(yd[p & 7], p < 8 ? xd0 : xd1), inv_s, na[p]));
        }
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < NP; ++p) s += z[p];
        if (s == 12345.678f) out[0] = s;
    } else {
        unsigned long long z[NP / 2], na[NP / 2];
#pragma unroll
        for (int p = 0; p < NP / 2; ++p) { z[p] = pack(0.f, 0.f); na[p] = pack(-1e-3f * (threadIdx.x + 2 * p), -1e-3f * (threadIdx.x + 2 * p + 1)); }
        const unsigned long long k1 = pack(inv_s, inv_s);
        for (int it = 0; it < iters; ++it) {
            const int i = it & 127;
            const float2 xs = sx[i];
            const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(&sm[0][i]);  // (yd0,yd1), (yd2,yd3)
            const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(&sm[1][i]);
            const unsigned long long yd[4] = {a.x, a.y, b.x, b.y};
            const float xd0 = (xs.x * c0 + xs.y) + cc0, xd1 = (xs.x * c1 + xs.y) + cc1;
            const unsigned long long x0 = pack(xd0, xd0), x1 = pack(xd1, xd1);
#pragma unroll
            for (int p = 0; p < NP / 2; ++p) {
                const unsigned long long t = fma2(add2(yd[p & 3], p < 4 ? x0 : x1), k1, na[p]);
                float t0, t1;
                unpack(t, t0, t1);
                z[p] = add2(z[p], pack(ex2_mufu(t0), ex2_mufu(t1)));
            }
        }
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < NP / 2; ++p) { float u, v; unpack(z[p], u, v); s += u + v; }
        if (s == 12345.678f) out[0] = s;
    }
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
    printf("variant %d warps/SM %2d: %.2f Texp/s (%.1f%% of the 4.65 MUFU peak) %s\n", V, ctas * 4, exps / ms / 1e9, 100 * exps / ms / 1e9 / 4.65, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() { for (int c : {3, 4, 6, 8}) { run<0>(c); run<1>(c); } return 0; }
