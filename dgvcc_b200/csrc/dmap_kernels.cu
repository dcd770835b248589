// Geometry-adaptive density-map generation for sm_100a.
//
// Replaces utils/dmap_gen.py:14-81 of the reference (scipy KDTree k=4 query + one full-image
// scipy.ndimage.gaussian_filter per head) by
//   dmap_knn_kernel     : brute-force tiled 4-nearest search in fp64 (self included, like
//   + dmap_knn_merge      KDTree.query(points, k=4)), candidates split into slices across CTAs and merged;
//                         sigma = 0.1*(d1+d2+d3)                                     dmap_gen.py:34-48
//   dmap_prepare_kernel : per head: truncated pixel index, in-bounds test, kernel radius
//                         int(truncate*sigma+0.5) and the normaliser of scipy's 1-D Gaussian kernel,
//                         summed in numpy's pairwise order                             dmap_gen.py:41-49
//   dmap_splat_kernel   : one CTA per 32x32 output tile gathers the heads whose stamp overlaps it, in
//                         index order, and accumulates fl32(f64(fl32(w[dy])) * w[dx]) per pixel in
//                         fp32 -- the closed form of gaussian_filter on a one-hot image.
//                         Every output pixel is written exactly once (zero fill fused), coalesced.
// All fp64 arithmetic uses explicit _rn intrinsics (no FMA contraction): scipy's wheels are baseline
// x86-64 without FMA, and the neighbour indices must match bit for bit.
#include <math.h>

#include "common.cuh"
#include "../../include/dgvcc_b200.h"

namespace dgvcc {
namespace dmap {

constexpr int KNN_THREADS = 256;

// ---------------------------------------------------------------------------------- kNN + sigma
// Each CTA finds, for its 256 query heads, the 4 nearest candidates inside one slice of the head list
// (grid.y slices, so that small and large N both fill the chip); dmap_knn_merge_kernel then merges the
// per-slice lists.  Order is (squared distance, index) everywhere, so ties keep the lower index.
__device__ __forceinline__ void knn_insert(double (&best)[4], int (&bidx)[4], double d2, int j) {
    best[3] = d2; bidx[3] = j;
#pragma unroll
    for (int k = 3; k > 0; --k) {
        if (best[k] < best[k - 1] || (best[k] == best[k - 1] && bidx[k] < bidx[k - 1])) {
            const double tb = best[k]; best[k] = best[k - 1]; best[k - 1] = tb;
            const int ti = bidx[k]; bidx[k] = bidx[k - 1]; bidx[k - 1] = ti;
        }
    }
}

__global__ void __launch_bounds__(KNN_THREADS)
dmap_knn_kernel(const double2* __restrict__ pts, int n, int slice_len, double* __restrict__ part_d2,
                int32_t* __restrict__ part_idx) {
    __shared__ double2 cand[KNN_THREADS];
    const int i = blockIdx.x * KNN_THREADS + threadIdx.x;
    const double2 q = pts[min(i, n - 1)];
    double best[4];
    int bidx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { best[k] = INFINITY; bidx[k] = n; }  // scipy: missing neighbour = (inf, N)
    const int j_begin = blockIdx.y * slice_len, j_end = min(n, j_begin + slice_len);
    for (int j0 = j_begin; j0 < j_end; j0 += KNN_THREADS) {
        __syncthreads();
        if (j0 + threadIdx.x < j_end) cand[threadIdx.x] = pts[j0 + threadIdx.x];
        __syncthreads();
        const int lim = min(KNN_THREADS, j_end - j0);
#pragma unroll 4
        for (int t = 0; t < lim; ++t) {
            const double dx = __dsub_rn(q.x, cand[t].x);
            const double dy = __dsub_rn(q.y, cand[t].y);
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            // candidates arrive in index order, so on ties the earlier (lower) index stays in front
            if (d2 < best[3]) knn_insert(best, bidx, d2, j0 + t);
        }
    }
    if (i >= n) return;
    const size_t o = ((size_t)blockIdx.y * n + i) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) { part_d2[o + k] = best[k]; part_idx[o + k] = bidx[k]; }
}

__global__ void __launch_bounds__(KNN_THREADS)
dmap_knn_merge_kernel(const double* __restrict__ part_d2, const int32_t* __restrict__ part_idx, int n, int slices,
                      int32_t* __restrict__ nn_idx, double* __restrict__ nn_dist, double* __restrict__ sigma) {
    const int i = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (i >= n) return;
    double best[4];
    int bidx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { best[k] = INFINITY; bidx[k] = n; }
    for (int s = 0; s < slices; ++s) {
        const size_t o = ((size_t)s * n + i) * 4;
        for (int k = 0; k < 4; ++k) {
            const double d2 = part_d2[o + k];
            const int j = part_idx[o + k];
            if (j < n && (d2 < best[3] || (d2 == best[3] && j < bidx[3]))) knn_insert(best, bidx, d2, j);
        }
    }
    double d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        d[k] = __dsqrt_rn(best[k]);
        nn_idx[4 * i + k] = bidx[k];
        nn_dist[4 * i + k] = d[k];
    }
    // dmap_gen.py:45-48: (d1 + d2 + d3) * 0.1 when there are more than 3 heads, else 15
    sigma[i] = (n > 3) ? __dmul_rn(__dadd_rn(__dadd_rn(d[1], d[2]), d[3]), 0.1) : 15.0;
}

// ---------------------------------------------------------------------------- per-head stamp data
struct Stamp {
    int ix, iy;      // pixel of the one-hot write (dmap_gen.py:42), after numpy's negative-index wrap
    int radius;      // int(truncate * sigma + 0.5); -1 marks a skipped head (dmap_gen.py:41-44)
    int pad_;
    double coef;     // -0.5 / sigma^2
    double norm;     // sum of exp(coef * i^2), i = -radius..radius, in numpy's pairwise order
};

__device__ __forceinline__ double phi(double coef, int i) {
    return exp(__dmul_rn(coef, (double)((long long)i * i)));
}

// Leaf of numpy's pairwise summation (DOUBLE_pairwise_sum, n <= 128) over phi(first), phi(first+1), ...
__device__ __forceinline__ double pairwise_leaf(double coef, int first, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, phi(coef, first + i));
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = phi(coef, first + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], phi(coef, first + i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, phi(coef, first + i));
    return res;
}

// numpy's pairwise summation: blocks of <= 128 are leaves, larger ranges split at n/2 rounded down to a
// multiple of 8 and the halves are added.  Explicit stack instead of recursion (depth <= 24 for n < 2^31),
// so the kernel needs no enlarged device stack however large a sparse image's sigma gets.
__device__ double pairwise_phi_sum(double coef, int first, int n) {
    constexpr int DEPTH = 26;
    int s_first[DEPTH], s_n[DEPTH], s_stage[DEPTH];
    double s_left[DEPTH];
    int sp = 0;
    s_first[0] = first; s_n[0] = n; s_stage[0] = 0;
    double ret = 0.0;
    while (sp >= 0) {
        const int cn = s_n[sp];
        int n2 = cn / 2;
        n2 -= n2 % 8;
        if (s_stage[sp] == 0) {
            if (cn <= 128) { ret = pairwise_leaf(coef, s_first[sp], cn); --sp; continue; }
            s_stage[sp] = 1;
            s_first[sp + 1] = s_first[sp]; s_n[sp + 1] = n2; s_stage[sp + 1] = 0;
            ++sp;
        } else if (s_stage[sp] == 1) {
            s_left[sp] = ret;
            s_stage[sp] = 2;
            s_first[sp + 1] = s_first[sp] + n2; s_n[sp + 1] = cn - n2; s_stage[sp + 1] = 0;
            ++sp;
        } else {
            ret = __dadd_rn(s_left[sp], ret);
            --sp;
        }
    }
    return ret;
}

// ---- the same sum, one warp per head -------------------------------------------------------------
// Groups of 8 lanes evaluate one leaf each: lane j of a group owns numpy's accumulator r[j] (elements
// j, j+8, ... of the leaf), the group leader combines them in numpy's order and adds the tail.  The tree above
// the leaves is then combined by lane 0 in numpy's recursion order.  Bit-identical to pairwise_phi_sum.
constexpr int MAX_LEAVES = 256;

struct LeafList {
    int first[MAX_LEAVES];
    int len[MAX_LEAVES];
    double sum[MAX_LEAVES];
};

// DFS enumeration of the leaves of numpy's recursion over [first, first + n); returns their number
// (or -1 when there are more than MAX_LEAVES).  Pure integer work, done redundantly by every lane.
__device__ int enumerate_leaves(int first, int n, LeafList& ll, bool write) {
    constexpr int DEPTH = 26;
    int s_first[DEPTH], s_n[DEPTH], s_stage[DEPTH];
    int sp = 0, count = 0;
    s_first[0] = first; s_n[0] = n; s_stage[0] = 0;
    while (sp >= 0) {
        const int cn = s_n[sp];
        int n2 = cn / 2;
        n2 -= n2 % 8;
        if (s_stage[sp] == 0) {
            if (cn <= 128) {
                if (count >= MAX_LEAVES) return -1;
                if (write) { ll.first[count] = s_first[sp]; ll.len[count] = cn; }
                ++count; --sp;
                continue;
            }
            s_stage[sp] = 1;
            s_first[sp + 1] = s_first[sp]; s_n[sp + 1] = n2; s_stage[sp + 1] = 0;
            ++sp;
        } else if (s_stage[sp] == 1) {
            s_stage[sp] = 2;
            s_first[sp + 1] = s_first[sp] + n2; s_n[sp + 1] = cn - n2; s_stage[sp + 1] = 0;
            ++sp;
        } else {
            --sp;
        }
    }
    return count;
}

// One leaf (n <= 128) by a group of 8 lanes; the result is valid on the group's lane 0.
__device__ __forceinline__ double leaf_by_group(double coef, int first, int n, int sub /*0..7*/, unsigned group_mask) {
    if (n < 8) {  // numpy's plain loop: nothing to parallelise
        double res = 0.0;
        if (sub == 0)
            for (int i = 0; i < n; ++i) res = __dadd_rn(res, phi(coef, first + i));
        return res;
    }
    const int body = n - (n % 8);
    double r = phi(coef, first + sub);
    for (int i = 8; i < body; i += 8) r = __dadd_rn(r, phi(coef, first + i + sub));
    // ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)): a butterfly over the 8 lanes is exactly this tree
    r = __dadd_rn(r, __shfl_xor_sync(group_mask, r, 1));
    r = __dadd_rn(r, __shfl_xor_sync(group_mask, r, 2));
    r = __dadd_rn(r, __shfl_xor_sync(group_mask, r, 4));
    if (sub == 0)
        for (int i = body; i < n; ++i) r = __dadd_rn(r, phi(coef, first + i));
    return r;
}

// Combine the leaf sums in numpy's recursion order (lane 0 only).
__device__ double combine_leaves(int first, int n, const LeafList& ll) {
    constexpr int DEPTH = 26;
    int s_n[DEPTH], s_stage[DEPTH];
    double s_left[DEPTH];
    int sp = 0, next = 0;
    s_n[0] = n; s_stage[0] = 0;
    double ret = 0.0;
    (void)first;
    while (sp >= 0) {
        const int cn = s_n[sp];
        int n2 = cn / 2;
        n2 -= n2 % 8;
        if (s_stage[sp] == 0) {
            if (cn <= 128) { ret = ll.sum[next++]; --sp; continue; }
            s_stage[sp] = 1; s_n[sp + 1] = n2; s_stage[sp + 1] = 0; ++sp;
        } else if (s_stage[sp] == 1) {
            s_left[sp] = ret; s_stage[sp] = 2; s_n[sp + 1] = cn - n2; s_stage[sp + 1] = 0; ++sp;
        } else {
            ret = __dadd_rn(s_left[sp], ret); --sp;
        }
    }
    return ret;
}

constexpr int PREP_WARPS = 4;

__global__ void __launch_bounds__(PREP_WARPS * 32)
dmap_prepare_kernel(const double2* __restrict__ pts, const double* __restrict__ sigma, double fixed_sigma,
                    double truncate, int n, int height, int width, Stamp* __restrict__ stamps) {
    __shared__ LeafList leaves[PREP_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * PREP_WARPS + warp;
    if (i >= n) return;
    LeafList& ll = leaves[warp];
    const double2 p = pts[i];
    Stamp s;
    s.pad_ = 0;
    s.ix = (int)p.x;  // Python int(): truncation toward zero
    s.iy = (int)p.y;
    const bool keep = s.iy < height && s.ix < width;
    if (s.iy < 0) s.iy += height;  // numpy wraps negative indices; the host wrapper rejects < -size like numpy
    if (s.ix < 0) s.ix += width;
    const double sd = sigma ? sigma[i] : fixed_sigma;
    if (!keep || s.iy < 0 || s.ix < 0) {
        s.radius = -1; s.coef = 0.0; s.norm = 1.0;
    } else if (!(sd > 1e-15)) {  // scipy: sigma <= 1e-15 copies the input (identity filter)
        s.radius = 0; s.coef = 0.0; s.norm = 1.0;
    } else {
        s.radius = (int)__dadd_rn(__dmul_rn(truncate, sd), 0.5);
        s.coef = -0.5 / __dmul_rn(sd, sd);
        const int n_el = 2 * s.radius + 1;
        const int n_leaves = enumerate_leaves(-s.radius, n_el, ll, lane == 0);
        if (n_leaves < 0) {
            s.norm = (lane == 0) ? pairwise_phi_sum(s.coef, -s.radius, n_el) : 0.0;  // absurdly wide kernel
        } else {
            __syncwarp();
            const int group = lane >> 3, sub = lane & 7;
            const unsigned group_mask = 0xffu << (8 * group);
            for (int l0 = 0; l0 < n_leaves; l0 += 4) {
                const int l = l0 + group;
                if (l < n_leaves) {  // uniform inside a group of 8 lanes
                    const double v = leaf_by_group(s.coef, ll.first[l], ll.len[l], sub, group_mask);
                    if (sub == 0) ll.sum[l] = v;
                }
            }
            __syncwarp();
            s.norm = (lane == 0) ? combine_leaves(-s.radius, n_el, ll) : 0.0;
        }
    }
    if (lane == 0) stamps[i] = s;
}

// ------------------------------------------------------------------------------------------ splat
constexpr int TILE = 32;
constexpr int SPLAT_THREADS = 256;
constexpr int GROUP = 8;  // heads whose tile weights are staged together

__global__ void __launch_bounds__(SPLAT_THREADS)
dmap_splat_kernel(const Stamp* __restrict__ stamps, int n, int height, int width, float* __restrict__ density) {
    __shared__ int list[SPLAT_THREADS];
    __shared__ int warp_cnt[SPLAT_THREADS / 32];
    __shared__ double wy[GROUP][TILE];  // fl32-rounded row weights, widened again (exact)
    __shared__ double wx[GROUP][TILE];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * TILE, y0 = blockIdx.y * TILE;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};  // pixels (y0 + warp + 8k, x0 + lane)

    for (int base = 0; base < n; base += SPLAT_THREADS) {
        // heads of this batch whose stamp overlaps the tile, compacted in index order
        const int i = base + tid;
        bool hit = false;
        if (i < n) {
            const Stamp s = stamps[i];
            hit = s.radius >= 0 && s.ix + s.radius >= x0 && s.ix - s.radius < x0 + TILE &&
                  s.iy + s.radius >= y0 && s.iy - s.radius < y0 + TILE;
        }
        const unsigned int ballot = __ballot_sync(FULL_MASK, hit);
        if (lane == 0) warp_cnt[warp] = __popc(ballot);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < SPLAT_THREADS / 32; ++w) {
            if (w < warp) before += warp_cnt[w];
            total += warp_cnt[w];
        }
        if (hit) list[before + __popc(ballot & ((1u << lane) - 1u))] = i;
        __syncthreads();

        for (int g0 = 0; g0 < total; g0 += GROUP) {
            const int gcnt = min(GROUP, total - g0);
            // stage the tile's 32 row and 32 column weights of each head of the group
            for (int k = tid; k < gcnt * 2 * TILE; k += SPLAT_THREADS) {
                const int h = k / (2 * TILE), which = (k / TILE) & 1, off = k % TILE;
                const Stamp s = stamps[list[g0 + h]];
                const int d = which ? (x0 + off - s.ix) : (y0 + off - s.iy);
                double w = 0.0;
                if (d >= -s.radius && d <= s.radius) w = __ddiv_rn(phi(s.coef, d), s.norm);
                if (which) wx[h][off] = w;
                else wy[h][off] = (double)(float)w;  // first filter pass stores float32 (output dtype)
            }
            __syncthreads();
            for (int h = 0; h < gcnt; ++h) {
                const double cx = wx[h][lane];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    acc[k] = __fadd_rn(acc[k], (float)__dmul_rn(wy[h][warp + 8 * k], cx));
            }
            __syncthreads();
        }
    }
    const int x = x0 + lane;
    if (x < width) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int y = y0 + warp + 8 * k;
            if (y < height) density[(size_t)y * width + x] = acc[k];
        }
    }
}

}  // namespace dmap
}  // namespace dgvcc

using namespace dgvcc;
using namespace dgvcc::dmap;

static int knn_slices(int n) {
    const int query_ctas = ceil_div(n, KNN_THREADS);
    int s = ceil_div(4 * 148, query_ctas);       // ~4 CTAs per SM in flight
    const int max_s = ceil_div(n, KNN_THREADS);  // a slice is at least one candidate tile
    if (s > max_s) s = max_s;
    return s < 1 ? 1 : s;
}

extern "C" size_t dgvcc_dmap_knn_workspace_bytes(int n) {
    if (n <= 0) return 16;
    return (size_t)knn_slices(n) * n * 4 * (sizeof(double) + sizeof(int32_t));
}

extern "C" int dgvcc_dmap_knn_sigma(const double* pts_xy, int n, int32_t* nn_idx, double* nn_dist, double* sigma,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    if (n < 0) return DGVCC_ERR_ARG;
    if (n == 0) return DGVCC_OK;
    if (!pts_xy || !nn_idx || !nn_dist || !sigma || !workspace) return DGVCC_ERR_ARG;
    if (workspace_bytes < dgvcc_dmap_knn_workspace_bytes(n)) return DGVCC_ERR_WORKSPACE;
    const int slices = knn_slices(n);
    const int slice_len = ceil_div(ceil_div(n, slices), KNN_THREADS) * KNN_THREADS;
    const int used = ceil_div(n, slice_len);
    double* part_d2 = (double*)workspace;
    int32_t* part_idx = (int32_t*)(part_d2 + (size_t)slices * n * 4);
    cudaStream_t st = (cudaStream_t)stream;
    dmap_knn_kernel<<<dim3(ceil_div(n, KNN_THREADS), used), KNN_THREADS, 0, st>>>((const double2*)pts_xy, n, slice_len,
                                                                                 part_d2, part_idx);
    DGVCC_RETURN_IF_CUDA(cudaGetLastError());
    dmap_knn_merge_kernel<<<ceil_div(n, KNN_THREADS), KNN_THREADS, 0, st>>>(part_d2, part_idx, n, used, nn_idx, nn_dist, sigma);
    return (int)cudaGetLastError();
}

extern "C" size_t dgvcc_dmap_workspace_bytes(int n) { return (size_t)(n > 0 ? n : 1) * sizeof(Stamp); }

extern "C" int dgvcc_dmap_splat(const double* pts_xy, const double* sigma, double fixed_sigma, double truncate, int n,
                                int height, int width, void* workspace, size_t workspace_bytes, float* density,
                                void* stream) {
    if (n < 0 || height <= 0 || width <= 0 || !density) return DGVCC_ERR_ARG;
    if (n > 0 && (!pts_xy || !workspace)) return DGVCC_ERR_ARG;
    if (workspace_bytes < dgvcc_dmap_workspace_bytes(n)) return DGVCC_ERR_WORKSPACE;
    if (!sigma && !(fixed_sigma >= 0.0)) return DGVCC_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    Stamp* stamps = (Stamp*)workspace;
    if (n > 0) {
        dmap_prepare_kernel<<<ceil_div(n, PREP_WARPS), PREP_WARPS * 32, 0, st>>>((const double2*)pts_xy, sigma, fixed_sigma, truncate, n,
                                                              height, width, stamps);
        DGVCC_RETURN_IF_CUDA(cudaGetLastError());
    }
    dmap_splat_kernel<<<dim3(ceil_div(width, TILE), ceil_div(height, TILE)), SPLAT_THREADS, 0, st>>>(
        stamps, n, height, width, density);
    return (int)cudaGetLastError();
}
