// Bayesian-dataset target preparation for sm_100a (the step right before BL).
//
// Replaces the per-point numpy work of datasets/bay_dataset.py:
//   bay_knn_mean_kernel     : BayesianDataset._cal_dists (bay_dataset.py:38-48) -- mean distance to the 3
//                             nearest heads.  The reference materialises the N x N matrix
//                             sqrt(max(sq_i - 2 p_i.p_j + sq_j, 0)) (1.15 GB at N = 12 000) and partitions every
//                             row; here each thread scans the heads through a shared-memory tile and keeps the
//                             4 smallest values of the same expansion in registers.
//   bay_crop_targets_kernel : crop block of _train_transform (bay_dataset.py:85-107) -- clipped box / crop
//                             overlap ratio (utils/misc.py:39-45), the >= 0.3 filter as an ordered block
//                             compaction, shift into crop coordinates, unconditional mirror in x.
// Arithmetic follows numpy's dtype rules: T = double for float64 annotations (JHU), float for float32 (QNRF).
#include <math.h>

#include "common.cuh"
#include "../../include/dgvcc_b200.h"

namespace dgvcc {
namespace bay {

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }

constexpr int KNN_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(KNN_THREADS)
bay_knn_mean_kernel(const typename Vec2<T>::type* __restrict__ pts, int n, T* __restrict__ dists) {
    using V = typename Vec2<T>::type;
    __shared__ V cand[KNN_THREADS];
    __shared__ T cand_sq[KNN_THREADS];
    const int i = blockIdx.x * KNN_THREADS + threadIdx.x;
    const V q = pts[min(i, n - 1)];
    const T sq_i = add_rn(mul_rn(q.x, q.x), mul_rn(q.y, q.y));  // np.sum(pts*pts, axis=1)
    T best[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) best[k] = (T)INFINITY;
    T small_sum = 0;  // N < 4: mean over columns 1.. of the unsorted row (bay_dataset.py:45-46)

    for (int j0 = 0; j0 < n; j0 += KNN_THREADS) {
        __syncthreads();
        if (j0 + threadIdx.x < n) {
            const V c = pts[j0 + threadIdx.x];
            cand[threadIdx.x] = c;
            cand_sq[threadIdx.x] = add_rn(mul_rn(c.x, c.x), mul_rn(c.y, c.y));
        }
        __syncthreads();
        const int lim = min(KNN_THREADS, n - j0);
        for (int t = 0; t < lim; ++t) {
            const T dot = add_rn(mul_rn(q.x, cand[t].x), mul_rn(q.y, cand[t].y));
            // (sq_i - 2*dot) + sq_j, the reference's evaluation order
            const T v = add_rn(add_rn(sq_i, -mul_rn((T)2, dot)), cand_sq[t]);
            if (n < 4) {
                if (j0 + t >= 1) small_sum = add_rn(small_sum, sqrt_rn(v > 0 ? v : (T)0));
            } else if (v < best[3]) {
                best[3] = v;
#pragma unroll
                for (int k = 3; k > 0; --k)
                    if (best[k] < best[k - 1]) { const T tmp = best[k]; best[k] = best[k - 1]; best[k - 1] = tmp; }
            }
        }
    }
    if (i >= n) return;
    if (n < 4) {
        dists[i] = small_sum / (T)(n - 1);
    } else {
        // np.partition(dists, 3)[:, 1:4]: the 2nd..4th smallest (monotone max/sqrt applied after the selection)
        T s = 0;
#pragma unroll
        for (int k = 1; k < 4; ++k) s = add_rn(s, sqrt_rn(best[k] > 0 ? best[k] : (T)0));
        dists[i] = s / (T)3;
    }
}

constexpr int CROP_THREADS = 1024;

// One CTA (ordered compaction): kept points keep their index order like boolean-mask indexing.
// gt_out is float64 in every case: `gt[mask] - [j, i]` promotes float32 annotations to float64.
template <typename T>
__global__ void __launch_bounds__(CROP_THREADS)
bay_crop_targets_kernel(const typename Vec2<T>::type* __restrict__ gt, const T* __restrict__ dists, int n, double c_left,
                        double c_up, double c_right, double c_down, double* __restrict__ gt_out,
                        T* __restrict__ targ_out, int* __restrict__ kept_out) {
    __shared__ int warp_cnt[CROP_THREADS / 32];
    __shared__ int running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    const T cl = (T)c_left, cu = (T)c_up, cr = (T)c_right, cd = (T)c_down;
    const double w = c_right - c_left;
    for (int base = 0; base < n; base += CROP_THREADS) {
        const int i = base + tid;
        bool keep = false;
        T ratio = 0;
        typename Vec2<T>::type p;
        p.x = 0; p.y = 0;
        if (i < n) {
            p = gt[i];
            T nd = dists[i];
            nd = nd < (T)4 ? (T)4 : (nd > (T)128 ? (T)128 : nd);   // np.clip(dists, 4.0, 128.0)
            const T half = nd / (T)2;
            const T il = fmax(cl, p.x - half), iu = fmax(cu, p.y - half);
            const T ir = fmin(cr, p.x + half), idn = fmin(cd, p.y + half);
            const T area = mul_rn(fmax(ir - il, (T)0), fmax(idn - iu, (T)0));
            const T r = area / mul_rn(nd, nd);                        // 1.0 * inner_area / origin_area
            ratio = r < (T)0 ? (T)0 : (r > (T)1 ? (T)1 : r);
            keep = ratio >= (T)0.3;
        }
        const unsigned int ballot = __ballot_sync(FULL_MASK, keep);
        if (lane == 0) warp_cnt[warp] = __popc(ballot);
        __syncthreads();
        int before = running, total = 0;
        for (int k = 0; k < CROP_THREADS / 32; ++k) {
            if (k < warp) before += warp_cnt[k];
            total += warp_cnt[k];
        }
        if (keep) {
            const int pos = before + __popc(ballot & ((1u << lane) - 1u));
            DGVCC_DEV_CHECK(pos >= 0 && pos < n);
            const double x = (double)p.x - c_left;   // gt - [j, i]
            gt_out[2 * pos] = w - x;                 // gt[:, 0] = w - gt[:, 0]   (bay_dataset.py:104-105)
            gt_out[2 * pos + 1] = (double)p.y - c_up;
            targ_out[pos] = ratio;
        }
        __syncthreads();
        if (tid == 0) running += total;
        __syncthreads();
    }
    if (tid == 0) *kept_out = running;
}

template <typename T>
int launch_knn(const void* pts, int n, void* dists, cudaStream_t st) {
    bay_knn_mean_kernel<T><<<ceil_div(n, KNN_THREADS), KNN_THREADS, 0, st>>>(
        (const typename Vec2<T>::type*)pts, n, (T*)dists);
    return (int)cudaGetLastError();
}

template <typename T>
int launch_crop(const void* gt, const void* dists, int n, double l, double u, double r, double d, double* gt_out,
                void* targ_out, int* kept, cudaStream_t st) {
    bay_crop_targets_kernel<T><<<1, CROP_THREADS, 0, st>>>((const typename Vec2<T>::type*)gt, (const T*)dists, n, l, u, r, d,
                                                           gt_out, (T*)targ_out, kept);
    return (int)cudaGetLastError();
}

}  // namespace bay
}  // namespace dgvcc

using namespace dgvcc;

extern "C" int dgvcc_bay_knn_mean(const void* pts_xy, int n, int is_double, void* dists, void* stream) {
    DGVCC_DEVICE_GUARD(stream);
    if (n < 2) return n < 0 ? DGVCC_ERR_ARG : DGVCC_OK;  // the N = 0 / N = 1 constants are the host wrapper's
    if (!pts_xy || !dists) return DGVCC_ERR_ARG;
    return is_double ? bay::launch_knn<double>(pts_xy, n, dists, (cudaStream_t)stream)
                     : bay::launch_knn<float>(pts_xy, n, dists, (cudaStream_t)stream);
}

extern "C" int dgvcc_bay_crop_targets(const void* gt_xy, const void* dists, int n, int is_double, double crop_left,
                                      double crop_up, double crop_right, double crop_down, double* gt_out,
                                      void* targ_out, int* kept_out, void* stream) {
    DGVCC_DEVICE_GUARD(stream);
    if (n <= 0 || !gt_xy || !dists || !gt_out || !targ_out || !kept_out) return DGVCC_ERR_ARG;
    return is_double ? bay::launch_crop<double>(gt_xy, dists, n, crop_left, crop_up, crop_right, crop_down, gt_out,
                                                targ_out, kept_out, (cudaStream_t)stream)
                     : bay::launch_crop<float>(gt_xy, dists, n, crop_left, crop_up, crop_right, crop_down, gt_out,
                                               targ_out, kept_out, (cudaStream_t)stream);
}
