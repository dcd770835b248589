// ISW channel-covariance Gram  G_b = X_b X_b^T  on the Blackwell tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the torch.bmm of models/ISW/instance_whitening.py:37.  X_b is [C, HW] fp32 with HW
// contiguous, i.e. both MMA operands are K-major.  fp32 accuracy (rtol 1e-5 against the reference's
// sgemm) comes from the 3xTF32 split  x = hi + lo,  hi = RN_tf32(x), lo = RN_tf32(x - hi):
//     G ~= hi hi^T + (hi lo^T + lo hi^T)        (fp32 accumulation in TMEM; lo lo^T ~ 2^-22 is dropped)
// The tensor core truncates (does not round) when it adds into the TMEM accumulator, a bias that grows
// with the number of accumulation steps (measured: -4e-8 relative per K=8 step, scripts/probe_tc_accuracy.py).  So the 2^-11-sized cross
// terms get their own accumulator -- the main one then sees a third of the steps -- and the host keeps
// the per-CTA K range short (<= 512) and adds the splits in fp32 round-to-nearest.
// One CTA computes one upper-triangular 128x128 tile of one sample over one split of the K = HW range
// and stores the partial tile; isw_cov_finish_kernel (isw_kernels.cu) adds the splits in order.
//
// Warp roles (320 threads):
//   warp 0   : TMA producer -- cp.async.bulk.tensor of the 128 x 32 fp32 operand tiles (128B swizzle)
//   warp 1   : TMEM allocation; one lane issues tcgen05.mma.kind::tf32 (12 per 32-wide k block)
//   warps 2-9: converters -- rewrite each landed tile as hi in place, lo beside it, then signal the
//              MMA warp; after the k loop they are the epilogue (tcgen05.ld -> global partial tile; warp w
//              reads TMEM lane quarter w % 4, warps 2-5 the left 64 columns, warps 6-9 the right 64)
// Pipeline: full[s] (TMA -> converters), ready[s] (converters -> MMA), empty[s] (MMA -> TMA) over
// STAGES shared-memory stages; accum (MMA -> epilogue).
#include "tc_common.cuh"
#include "../../include/dgvcc_b200.h"

namespace dgvcc {
namespace isw_tc {

constexpr int TILE_M = 128;                 // rows of X per operand tile (= UMMA M = UMMA N)
constexpr int BLOCK_K = 32;                 // fp32 per k block = one 128-byte swizzle row
constexpr int UMMA_K = 8;                   // tf32 MMA k extent (32 bytes)
constexpr int STAGES = 3;
constexpr int TILE_BYTES = TILE_M * BLOCK_K * 4;  // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;       // A_hi, B_hi, A_lo, B_lo
constexpr int CONVERTER_WARPS = 8;                  // two per SM sub-partition
constexpr int THREADS = 64 + 32 * CONVERTER_WARPS;
constexpr int TMEM_COLS = 256;  // two fp32 accumulators: hi*hi^T, and the small cross terms
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

using namespace dgvcc::tc;

// K-major operand tiles: rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO); LBO is unused (=16 B).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) { return umma_desc_sw128(smem_addr, 16, 1024); }

// kind::tf32, fp32 accumulate, A and B K-major, M = N = 128.
constexpr uint32_t IDESC = umma_idesc_tf32(TILE_M, TILE_M, false);

struct Args {
    int c, hw, splits, k_per_split, tiles_1d, n_tiles;
    float* part;
};

__global__ void __launch_bounds__(THREADS, 1)
isw_gram_tc_kernel(const __grid_constant__ CUtensorMap tmap, const Args a) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzled tiles
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + STAGES * STAGE_BYTES;  // full[S], ready[S], empty[S], accum, tmem slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 8 * (3 * STAGES + 1));
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto ready_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (3 * STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int ti = 0, rem = blockIdx.x;
    while (rem >= a.tiles_1d - ti) { rem -= a.tiles_1d - ti; ++ti; }
    const int tj = ti + rem;
    const bool diag = ti == tj;
    const int split = blockIdx.y, b = blockIdx.z;
    const int k0 = split * a.k_per_split;
    const int k1 = min(a.hw, k0 + a.k_per_split);
    const int n_kb = (k1 - k0 + BLOCK_K - 1) / BLOCK_K;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(ready_bar(s), CONVERTER_WARPS);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 128 fp32 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            for (int kb = 0; kb < n_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(empty_bar(s), ph ^ 1u);
                const uint32_t stage = base + s * STAGE_BYTES;
                mbar_arrive_expect_tx(full_bar(s), diag ? TILE_BYTES : 2 * TILE_BYTES);
                tma_load_3d(stage, &tmap, full_bar(s), k0 + kb * BLOCK_K, ti * TILE_M, b);
                if (!diag) tma_load_3d(stage + TILE_BYTES, &tmap, full_bar(s), k0 + kb * BLOCK_K, tj * TILE_M, b);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (int kb = 0; kb < n_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(ready_bar(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t stage = base + s * STAGE_BYTES;
                const uint64_t a_hi = umma_desc(stage);
                const uint64_t b_hi = umma_desc(stage + (diag ? 0 : TILE_BYTES));
                const uint64_t a_lo = umma_desc(stage + 2 * TILE_BYTES);
                const uint64_t b_lo = umma_desc(stage + (diag ? 2 : 3) * TILE_BYTES);
#pragma unroll
                for (int ks = 0; ks < BLOCK_K / UMMA_K; ++ks) {
                    const uint64_t adv = (uint64_t)((ks * UMMA_K * 4) >> 4);  // +32 B inside the swizzle row
                    umma_tf32(tmem_d, a_hi + adv, b_hi + adv, IDESC, (kb | ks) != 0);
                    umma_tf32(tmem_d + TILE_M, a_hi + adv, b_lo + adv, IDESC, (kb | ks) != 0);
                    umma_tf32(tmem_d + TILE_M, a_lo + adv, b_hi + adv, IDESC, 1u);
                }
                umma_commit(empty_bar(s));  // frees the stage once these MMAs have read it
            }
            umma_commit(accum_bar);
        }
    } else {
        // ===== converters: hi in place, lo beside =====
        const int ctid = threadIdx.x - 64;  // 0..255
        for (int kb = 0; kb < n_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(full_bar(s), ph);
            uint8_t* stage = base_ptr + s * STAGE_BYTES;
            const int n_vec = (diag ? 1 : 2) * (TILE_BYTES / 16);  // float4s of A_hi (and B_hi, contiguous after it)
            float4* hi = reinterpret_cast<float4*>(stage);
            float4* lo = reinterpret_cast<float4*>(stage + 2 * TILE_BYTES);
#pragma unroll 4
            for (int i = ctid; i < n_vec; i += 32 * CONVERTER_WARPS) {
                const float4 v = hi[i];
                float4 h, l;
                h.x = tf32_round(v.x); l.x = tf32_round(v.x - h.x);
                h.y = tf32_round(v.y); l.y = tf32_round(v.y - h.y);
                h.z = tf32_round(v.z); l.z = tf32_round(v.z - h.z);
                h.w = tf32_round(v.w); l.w = tf32_round(v.w - h.w);
                hi[i] = h;
                lo[i] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(ready_bar(s));
        }
        // ===== epilogue: TMEM -> global partial tile =====
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lane_base = 32 * (warp & 3);  // a warp may only touch its own quarter of the TMEM lanes
        const int row = lane_base + lane;
        float* out = a.part + ((((size_t)b * a.splits + split) * a.n_tiles + blockIdx.x) * TILE_M + row) * TILE_M;
        const bool row_ok = ti * TILE_M + row < a.c;
        const int col_begin = (warp - 2) / 4 * (TILE_M / 2);
#pragma unroll 1
        for (int c0 = col_begin; c0 < col_begin + TILE_M / 2; c0 += 32) {
            uint32_t r[32], x[32];
            const uint32_t taddr = tmem_d + ((uint32_t)lane_base << 16) + (uint32_t)c0;
            tmem_ld32(taddr, r);           // hi hi^T
            tmem_ld32(taddr + TILE_M, x);  // hi lo^T + lo hi^T
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 32; ++q) r[q] = __float_as_uint(__uint_as_float(r[q]) + __uint_as_float(x[q]));
            if (row_ok && tj * TILE_M + c0 < a.c) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(out + c0 + 4 * q) =
                        make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                    __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
    }
}

}  // namespace isw_tc
}  // namespace dgvcc

using namespace dgvcc;
using namespace dgvcc::isw_tc;

extern "C" int dgvcc_isw_gram_tc_partials(const float* x, int batch, int c, int hw, int splits, int k_per_split,
                                          float* part, void* stream) {
    if (!x || !part || batch <= 0 || c <= 0 || hw <= 0 || splits <= 0 || k_per_split <= 0) return DGVCC_ERR_ARG;
    // TMA needs 16-byte global strides and a 16-byte aligned base; rows past C are zero-filled by the
    // tensor map, so any C works, but tiny channel counts waste the 128-wide tile
    if (hw % 4 != 0 || c < 32 || k_per_split % BLOCK_K != 0 || ((uintptr_t)x & 15u)) return DGVCC_ERR_UNSUPPORTED;

    CUtensorMap tmap;
    if (!make_tmap_f32_3d(&tmap, x, (uint64_t)hw, (uint64_t)c, (uint64_t)batch, TILE_M)) return DGVCC_ERR_UNSUPPORTED;

    static bool attr_set = false;  // idempotent; a race only repeats the same call
    if (!attr_set) {
        DGVCC_RETURN_IF_CUDA(cudaFuncSetAttribute(isw_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    Args a;
    a.c = c; a.hw = hw; a.splits = splits; a.k_per_split = k_per_split;
    a.tiles_1d = ceil_div(c, TILE_M);
    a.n_tiles = a.tiles_1d * (a.tiles_1d + 1) / 2;
    a.part = part;
    isw_gram_tc_kernel<<<dim3(a.n_tiles, splits, batch), THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tmap, a);
    return (int)cudaGetLastError();
}
