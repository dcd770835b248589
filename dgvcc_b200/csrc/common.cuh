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

// Checked build (`DGVCC_BOUNDS_CHECK=1 python -m dgvcc_b200.build`, nvcc -DDGVCC_BOUNDS_CHECK; loaded instead of the
// product library when the same variable is set at import): every index that comes out of a host-built table or a
// data-dependent computation is tested where it is used; a violation is a device-side assert (site, block, thread
// and condition are reported, the launch fails with cudaErrorAssert) instead of silent memory corruption.
// compute-sanitizer is closed on the GPU pool this was built on -- this is its substitute.  The product build
// compiles every check away (the condition is not evaluated; `cuobjdump -sass` of the objects is identical with and
// without the DGVCC_DEV_CHECK lines).
#ifdef DGVCC_BOUNDS_CHECK
#undef NDEBUG
#include <assert.h>
// device-side assert: the runtime reports file, line, block, thread and the condition when the context is torn down
#define DGVCC_DEV_CHECK(cond) assert(cond)
#define DGVCC_BOUNDS_CHECKED 1
#else
#define DGVCC_DEV_CHECK(cond) ((void)0)
#define DGVCC_BOUNDS_CHECKED 0
#endif

namespace dgvcc {

// Every launcher runs on the device that owns the caller's stream, whatever the calling thread's current
// device is (the reference's configs pass device='cuda:1' etc. without ever calling set_device); the
// previous current device is restored on return.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(void* stream) {
        int cur = 0, want = 0;
        // A stream that is being captured into a CUDA graph belongs to the current device (the capture was begun on it),
        // and cudaStreamGetDevice is one of the calls a global-mode capture forbids: asking would invalidate the graph.
        cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing((cudaStream_t)stream, &capturing) != cudaSuccess) { (void)cudaGetLastError(); return; }
        if (capturing != cudaStreamCaptureStatusNone) return;
        if (cudaGetDevice(&cur) != cudaSuccess) return;
        if (cudaStreamGetDevice((cudaStream_t)stream, &want) != cudaSuccess) { (void)cudaGetLastError(); return; }
        if (want != cur && cudaSetDevice(want) == cudaSuccess) prev = cur;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define DGVCC_DEVICE_GUARD(stream) ::dgvcc::DeviceGuard dgvcc_device_guard_(stream)

// Function attributes (opt-in dynamic shared memory) are per device: `static PerDeviceOnce once;`
// `if (once.first()) cudaFuncSetAttribute(...)` configures a kernel the first time each device launches it.
struct PerDeviceOnce {
    bool done[64] = {};
    bool first() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (done[d]) return false;
        done[d] = true;  // a race only repeats the same idempotent call
        return true;
    }
};

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
