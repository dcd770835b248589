// Fused Bayesian-loss forward/backward for sm_100a.
//
// Replaces losses/bl.py (Post_Prob + Bay_Loss + autograd) of the reference without ever
// materialising the [points x pixels] posterior.  Work decomposition ("pixel owner"):
//   warp task  = (image, 32-column block, band of R grid rows); lane <-> grid column,
//                each thread owns R pixels of one column in registers;
//   the warp streams ALL points of its image through a private shared-memory tile
//   (x, x*x and the R per-row y-distances of each point), so the inner loop per
//   (point, pixel) pair is  FADD, FFMA, FMUL, MUFU.EX2, FADD/FFMA  -- MUFU-bound.
// Sweeps over the points (dense, no culling):
//   K1 bl_minz_kernel   : pass A min_n dis (bl.py:39), pass B softmax denominator (bl.py:44)
//   K2 bl_counts_kernel : expected counts c_n = sum_m D[m] p[n,m]   (bl.py:73), per-tile partials
//   K3 bl_select_kernel : deterministic reduction of the partials, |t-c|, trimmed top-k
//                         (radix select), loss (bl.py:75-79)
//   K4 bl_grad_kernel   : dL/dD[m] = g * sum_n w_n p[n,m]            (autograd of bl.py:73-79)
//
// Rounding contract: the softmax arguments reproduce the
// reference's fp32 sequence bit for bit -- no FMA contraction in the distance expansion
// (explicit _rn intrinsics), IEEE sqrt, IEEE division by 2*sigma^2 (exact scaling when it is
// a power of two, otherwise a Markstein-corrected reciprocal multiply that is correctly
// rounded).  Only exp (MUFU.EX2 of a rounded product) and summation order differ, ~1e-6.
#include "common.cuh"
#include "../../include/dgvcc_b200.h"

namespace dgvcc {
namespace bl {

constexpr int WARPS_PER_CTA = 4;
constexpr int CTA_THREADS = WARPS_PER_CTA * 32;
constexpr int TILE_PTS = 128;  // points staged per warp per tile

struct Scale {
    float s;          // fl32(2*sigma^2)                     bl.py:42
    float r;          // RN(1/s)
    float neg_inv_s;  // -1/s (exact when s is a power of two)
};

struct Geom {
    int hp, wp;        // grid rows, columns
    int col_blocks;    // ceil(wp / 32)
    int tiles;         // warp tasks per image = col_blocks * ceil(hp / R)
    float stride;      // image pixels per grid cell
    float half;        // stride / 2
};

// -dis / s, rounded exactly like the reference's true division.
template <bool POW2>
__device__ __forceinline__ float neg_div(float dis, const Scale& k) {
    if (POW2) return __fmul_rn(dis, k.neg_inv_s);
    const float x = -dis;
    const float q = __fmul_rn(x, k.r);
    const float rem = __fmaf_rn(-q, k.s, x);
    return __fmaf_rn(rem, k.r, q);
}

// exp(a - amax) with a = -dis/s:  fl(a - amax) is formed exactly as torch.softmax does.
template <bool POW2>
__device__ __forceinline__ float pair_exp(float dis, float neg_amax, const Scale& k) {
    float d;
    if (POW2)
        d = __fmaf_rn(dis, k.neg_inv_s, neg_amax);  // a is exact, so one rounding == fl(a - amax)
    else
        d = __fadd_rn(neg_div<false>(dis, k), neg_amax);
    return ex2_ftz(__fmul_rn(d, LOG2E));
}

// ((-2 * fl(p*c)) + fl(p*p)) + fl(c*c)   bl.py:27-28;  cm2 = -2c (scaling by 2 commutes with rounding)
__device__ __forceinline__ float axis_sqdist(float p, float pp, float cm2, float cc) {
    return __fadd_rn(__fadd_rn(__fmul_rn(p, cm2), pp), cc);
}

// Grid-cell centre, bl.py:14-15: arange(0, c_size, stride) + stride/2 (exact for integer strides).
__device__ __forceinline__ float cell_centre(int idx, const Geom& g) {
    return __fadd_rn(__fmul_rn((float)idx, g.stride), g.half);
}

template <int R>
struct __align__(16) WarpTile {
    float2 xs[TILE_PTS];      // (x, x*x)
    float yd[TILE_PTS][R];    // y-axis squared distance to each of the warp's R rows
    float aux[TILE_PTS];      // K2: staged counts; K4: signed weights
};

struct TaskInfo {
    int img, n_pts, pt0, row0, n_rows;
    int col, row_base;   // this lane's column, first row of the band
    bool col_ok;
    int task;            // tile index inside the image
};

template <int R>
__device__ __forceinline__ bool decode_task(const int32_t* __restrict__ meta, int batch, const Geom& g,
                                            TaskInfo& t) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    t.task = blockIdx.x * WARPS_PER_CTA + warp;
    if (t.task >= g.tiles) return false;
    const int32_t* pt_off = meta;
    const int32_t* row_off = meta + (batch + 1);
    const int32_t* order = meta + 2 * (batch + 1) + batch;
    t.img = order[blockIdx.y];
    t.pt0 = pt_off[t.img];
    t.n_pts = pt_off[t.img + 1] - t.pt0;
    t.row0 = row_off[t.img];
    t.n_rows = row_off[t.img + 1] - t.row0;
    const int jb = t.task % g.col_blocks, kb = t.task / g.col_blocks;
    t.col = jb * 32 + lane;
    t.col_ok = t.col < g.wp;
    if (!t.col_ok) t.col = g.wp - 1;  // clamp: compute a duplicate, never store it
    t.row_base = kb * R;
    return true;
}

// Stage points [n0, n0+cnt) of the image: lane-strided, coalesced float2 loads; entries past the
// image's last point (padding up to `padded`) repeat the last point so they stay finite.
template <int R>
__device__ __forceinline__ void stage_points(WarpTile<R>& tile, const float2* __restrict__ pts, int n0,
                                             int n_pts, int padded, const float (&cym2)[R],
                                             const float (&cyy)[R]) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < padded; i += 32) {
        const int n = min(n0 + i, n_pts - 1);
        const float2 p = __ldg(&pts[n]);
        tile.xs[i] = make_float2(p.x, __fmul_rn(p.x, p.x));
        const float yy = __fmul_rn(p.y, p.y);
#pragma unroll
        for (int r = 0; r < R; ++r) tile.yd[i][r] = axis_sqdist(p.y, yy, cym2[r], cyy[r]);
    }
}

template <int R>
__device__ __forceinline__ void load_yd(const WarpTile<R>& tile, int i, float (&yd)[R]) {
    if (R == 2) {
        const float2 v = *reinterpret_cast<const float2*>(&tile.yd[i][0]);
        yd[0] = v.x; yd[1] = v.y;
    } else {
#pragma unroll
        for (int q = 0; q < R / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(&tile.yd[i][4 * q]);
            yd[4 * q + 0] = v.x; yd[4 * q + 1] = v.y; yd[4 * q + 2] = v.z; yd[4 * q + 3] = v.w;
        }
    }
}

template <int R>
__device__ __forceinline__ void row_constants(const TaskInfo& t, const Geom& g, float (&cym2)[R],
                                              float (&cyy)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float cy = cell_centre(min(t.row_base + r, g.hp - 1), g);
        cym2[r] = -2.0f * cy;
        cyy[r] = __fmul_rn(cy, cy);
    }
}

// ------------------------------------------------------------------------------------------ K1
template <int R, bool POW2>
__global__ void __launch_bounds__(CTA_THREADS)
bl_minz_kernel(const float2* __restrict__ pts_all, const int32_t* __restrict__ meta,
               const float* __restrict__ st_sizes, int batch, Geom g, Scale k, float bg_ratio, int use_bg,
               float* __restrict__ amax_out, float* __restrict__ rz_out, float* __restrict__ pbg_out) {
    __shared__ WarpTile<R> tiles[WARPS_PER_CTA];
    TaskInfo t;
    if (!decode_task<R>(meta, batch, g, t)) return;
    WarpTile<R>& tile = tiles[threadIdx.x >> 5];
    const size_t img_base = (size_t)t.img * g.hp * g.wp;

    if (t.n_pts == 0) {  // bl.py:63-65: the only row is "sum of density", posterior == 1
        if (t.col_ok)
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (t.row_base + r < g.hp) {
                    const size_t m = img_base + (size_t)(t.row_base + r) * g.wp + t.col;
                    amax_out[m] = 0.f; rz_out[m] = 1.f; pbg_out[m] = 1.f;
                }
        return;
    }

    float cym2[R], cyy[R];
    row_constants<R>(t, g, cym2, cyy);
    const float cx = cell_centre(t.col, g);
    const float cxm2 = -2.0f * cx, cxx = __fmul_rn(cx, cx);
    const float2* pts = pts_all + t.pt0;

    // pass A: min over points of the squared distance (bl.py:39)
    float mind[R];
#pragma unroll
    for (int r = 0; r < R; ++r) mind[r] = __int_as_float(0x7f800000);
    for (int n0 = 0; n0 < t.n_pts; n0 += TILE_PTS) {
        const int cnt = min(TILE_PTS, t.n_pts - n0);
        __syncwarp();
        stage_points<R>(tile, pts, n0, t.n_pts, cnt, cym2, cyy);
        __syncwarp();
#pragma unroll 4
        for (int i = 0; i < cnt; ++i) {
            const float2 xs = tile.xs[i];
            float yd[R];
            load_yd<R>(tile, i, yd);
            const float xd = axis_sqdist(xs.x, xs.y, cxm2, cxx);
#pragma unroll
            for (int r = 0; r < R; ++r) mind[r] = fminf(mind[r], __fadd_rn(yd[r], xd));
        }
    }

    // background row (bl.py:39-43) and the softmax max
    float neg_amax[R], a_bg[R];
    const float dbg = __fmul_rn(st_sizes[t.img], bg_ratio);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float amax = neg_div<POW2>(mind[r], k);
        a_bg[r] = 0.f;
        if (use_bg) {
            const float root = __fsqrt_rn(fmaxf(mind[r], 0.0f));
            const float diff = __fadd_rn(dbg, -root);
            a_bg[r] = neg_div<POW2>(__fmul_rn(diff, diff), k);
            amax = fmaxf(amax, a_bg[r]);
        }
        neg_amax[r] = -amax;
    }

    // pass B: softmax denominator, accumulated in point order like torch's dim-0 softmax
    float z[R];
#pragma unroll
    for (int r = 0; r < R; ++r) z[r] = 0.f;
    for (int n0 = 0; n0 < t.n_pts; n0 += TILE_PTS) {
        const int cnt = min(TILE_PTS, t.n_pts - n0);
        __syncwarp();
        stage_points<R>(tile, pts, n0, t.n_pts, cnt, cym2, cyy);
        __syncwarp();
#pragma unroll 2
        for (int i = 0; i < cnt; ++i) {
            const float2 xs = tile.xs[i];
            float yd[R];
            load_yd<R>(tile, i, yd);
            const float xd = axis_sqdist(xs.x, xs.y, cxm2, cxx);
#pragma unroll
            for (int r = 0; r < R; ++r) z[r] += pair_exp<POW2>(__fadd_rn(yd[r], xd), neg_amax[r], k);
        }
    }

    if (t.col_ok) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (t.row_base + r >= g.hp) continue;
            float e_bg = 0.f;
            if (use_bg) {
                e_bg = ex2_ftz(__fmul_rn(__fadd_rn(a_bg[r], neg_amax[r]), LOG2E));
                z[r] += e_bg;
            }
            const float rz = 1.0f / z[r];
            const size_t m = img_base + (size_t)(t.row_base + r) * g.wp + t.col;
            amax_out[m] = -neg_amax[r];
            rz_out[m] = rz;
            pbg_out[m] = e_bg * rz;
        }
    }
}

// ------------------------------------------------------------------------------------------ K2
template <int R, bool POW2>
__global__ void __launch_bounds__(CTA_THREADS)
bl_counts_kernel(const float2* __restrict__ pts_all, const int32_t* __restrict__ meta,
                 const float* __restrict__ density, int batch, Geom g, Scale k, int use_bg,
                 const float* __restrict__ amax_in, const float* __restrict__ rz_in,
                 const float* __restrict__ pbg_in, int64_t total_rows, float* __restrict__ cpart) {
    __shared__ WarpTile<R> tiles[WARPS_PER_CTA];
    TaskInfo t;
    if (!decode_task<R>(meta, batch, g, t)) return;
    WarpTile<R>& tile = tiles[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const size_t img_base = (size_t)t.img * g.hp * g.wp;
    float* part = cpart + (size_t)t.task * total_rows + t.row0;

    // per-pixel weights D[m]/Z[m]; pixels outside the grid get weight 0
    float neg_amax[R], wd[R], bg_part = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool ok = t.col_ok && (t.row_base + r < g.hp);
        const size_t m = img_base + (size_t)min(t.row_base + r, g.hp - 1) * g.wp + t.col;
        const float d = ok ? density[m] : 0.f;
        neg_amax[r] = -amax_in[m];
        wd[r] = d * rz_in[m];
        bg_part = fmaf(d, pbg_in[m], bg_part);
    }
    if (use_bg || t.n_pts == 0) {  // background row, or the sum-of-density row of an empty image
        bg_part = warp_sum(bg_part);
        if (lane == 0) part[t.n_rows - 1] = bg_part;
    }
    if (t.n_pts == 0) return;

    float cym2[R], cyy[R];
    row_constants<R>(t, g, cym2, cyy);
    const float cx = cell_centre(t.col, g);
    const float cxm2 = -2.0f * cx, cxx = __fmul_rn(cx, cx);
    const float2* pts = pts_all + t.pt0;

    for (int n0 = 0; n0 < t.n_pts; n0 += TILE_PTS) {
        const int cnt = min(TILE_PTS, t.n_pts - n0);
        const int padded = (cnt + 7) & ~7;
        __syncwarp();
        stage_points<R>(tile, pts, n0, t.n_pts, padded, cym2, cyy);
        __syncwarp();
        for (int i0 = 0; i0 < padded; i0 += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float2 xs = tile.xs[i0 + u];
                float yd[R];
                load_yd<R>(tile, i0 + u, yd);
                const float xd = axis_sqdist(xs.x, xs.y, cxm2, cxx);
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    s = fmaf(pair_exp<POW2>(__fadd_rn(yd[r], xd), neg_amax[r], k), wd[r], s);
                v[u] = s;
            }
            // transpose-reduce: 8 per-lane partials -> lane quad q holds the warp total of point i0+q
#pragma unroll
            for (int h = 4, bit = 16; h >= 1; h >>= 1, bit >>= 1) {
                const bool up = (lane & bit) != 0;
#pragma unroll
                for (int u = 0; u < h; ++u) {
                    const float send = up ? v[u] : v[u + h];
                    const float keep = up ? v[u + h] : v[u];
                    v[u] = keep + __shfl_xor_sync(FULL_MASK, send, bit);
                }
            }
            v[0] += __shfl_xor_sync(FULL_MASK, v[0], 2);
            v[0] += __shfl_xor_sync(FULL_MASK, v[0], 1);
            if ((lane & 3) == 0) tile.aux[i0 + (lane >> 2)] = v[0];
        }
        __syncwarp();
        for (int i = lane; i < cnt; i += 32) part[n0 + i] = tile.aux[i];
    }
}

// ------------------------------------------------------------------------------------------ K3
constexpr int SELECT_THREADS = 512;

__device__ __forceinline__ float block_sum_ordered(float v, float* scratch) {
    // deterministic: warp butterfly, then warp partials added in warp order by thread 0
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float total = 0.f;
    if (threadIdx.x == 0)
        for (int w = 0; w < SELECT_THREADS / 32; ++w) total += scratch[w];
    return total;  // valid on thread 0
}

__global__ void __launch_bounds__(SELECT_THREADS)
bl_select_kernel(const float* __restrict__ cpart, int tiles, int64_t total_rows,
                 const int32_t* __restrict__ meta, const float* __restrict__ targets, int batch,
                 float inv_batch, float* __restrict__ counts, float* __restrict__ residual,
                 float* __restrict__ wsel, float* __restrict__ loss_img, float* __restrict__ loss_out,
                 unsigned int* __restrict__ ticket) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int sh_prefix, sh_rank, sh_equal;
    __shared__ unsigned int warp_cnt[SELECT_THREADS / 32];
    __shared__ float scratch[SELECT_THREADS / 32];

    const int img = blockIdx.x, tid = threadIdx.x;
    const int32_t* pt_off = meta;
    const int32_t* row_off = meta + (batch + 1);
    const int32_t* keep = meta + 2 * (batch + 1);
    const int pt0 = pt_off[img], n_pts = pt_off[img + 1] - pt0;
    const int row0 = row_off[img], n_rows = row_off[img + 1] - row0;
    const int n_cand = n_rows - 1;  // res[:-1]; the last row is always kept (bl.py:77-78)
    const int n_keep = keep[img];

    // expected counts: fixed-order sum of the per-tile partials; residual |t - c|  (bl.py:73-75)
    for (int j = tid; j < n_rows; j += SELECT_THREADS) {
        const float* p = cpart + row0 + j;
        float c = 0.f;
        for (int tl = 0; tl < tiles; ++tl) c += p[(size_t)tl * total_rows];
        const float tgt = (j < n_pts) ? targets[pt0 + j] : 0.f;
        counts[row0 + j] = c;
        residual[row0 + j] = fabsf(__fadd_rn(tgt, -c));
    }
    __syncthreads();

    // k-th smallest residual among the candidates: MSB-first radix select on the float bits
    unsigned int thr = 0xffffffffu;  // keep everything
    unsigned int take_equal = 0xffffffffu, n_equal = 0;
    if (n_keep <= 0) {
        thr = 0u; take_equal = 0u;  // keep nothing (residuals are >= +0, none is < 0)
    } else if (n_keep < n_cand) {
        if (tid == 0) { sh_prefix = 0u; sh_rank = (unsigned)n_keep; }
        unsigned int mask = 0u;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int b = tid; b < 256; b += SELECT_THREADS) hist[b] = 0u;
            __syncthreads();
            const unsigned int prefix = sh_prefix;
            for (int j = tid; j < n_cand; j += SELECT_THREADS) {
                const unsigned int bits = __float_as_uint(residual[row0 + j]);
                if ((bits & mask) == prefix) atomicAdd(&hist[(bits >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned int rank = sh_rank, cum = 0u;
                int b = 0;
                for (; b < 255; ++b) {
                    if (cum + hist[b] >= rank) break;
                    cum += hist[b];
                }
                sh_rank = rank - cum;
                sh_prefix = prefix | ((unsigned)b << shift);
                sh_equal = hist[b];
            }
            mask |= 255u << shift;
            __syncthreads();
        }
        thr = sh_prefix; take_equal = sh_rank; n_equal = sh_equal;
    }
    const bool ordered_ties = (n_keep > 0 && n_keep < n_cand && take_equal < n_equal);

    // selection, signed weights for backward, per-image loss
    float lsum = 0.f;
    unsigned int ties_seen = 0u;
    for (int base = 0; base < n_rows; base += SELECT_THREADS) {
        const int j = base + tid;
        const bool live = j < n_rows;
        float res = 0.f, c = 0.f, tgt = 0.f;
        unsigned int bits = 0u;
        if (live) {
            res = residual[row0 + j];
            c = counts[row0 + j];
            tgt = (j < n_pts) ? targets[pt0 + j] : 0.f;
            bits = __float_as_uint(res);
        }
        const bool cand = live && j < n_cand;
        bool tie_ok = true;
        if (ordered_ties) {  // equal residuals straddle the cut: keep the first `take_equal` in index order
            const bool is_tie = cand && bits == thr;
            const unsigned int ballot = __ballot_sync(FULL_MASK, is_tie);
            if ((tid & 31) == 0) warp_cnt[tid >> 5] = __popc(ballot);
            __syncthreads();
            unsigned int before = ties_seen, total = 0u;
            for (int w = 0; w < SELECT_THREADS / 32; ++w) {
                if (w < (tid >> 5)) before += warp_cnt[w];
                total += warp_cnt[w];
            }
            before += __popc(ballot & ((1u << (tid & 31)) - 1u));
            tie_ok = before < take_equal;
            ties_seen += total;
            __syncthreads();
        }
        bool sel = false;
        if (live) {
            if (j == n_rows - 1) sel = true;
            else sel = (bits < thr) || (bits == thr && take_equal > 0u && tie_ok);
            const float x = __fadd_rn(tgt, -c);  // d|x|/dc = -sign(x)
            const float w = sel ? (x > 0.f ? -1.f : (x < 0.f ? 1.f : 0.f)) : 0.f;
            wsel[row0 + j] = w;
            if (sel) lsum += res;
        }
    }
    const float l_img = block_sum_ordered(lsum, scratch);

    // last CTA to finish adds the per-image losses in image order (deterministic), bl.py:79
    if (tid == 0) {
        loss_img[img] = l_img;
        __threadfence();
        const unsigned int done = atomicAdd(ticket, 1u);
        if (done == (unsigned)batch - 1u) {
            __threadfence();
            float total = 0.f;
            const volatile float* li = loss_img;
            for (int i = 0; i < batch; ++i) total += li[i];
            loss_out[0] = total * inv_batch;
            *ticket = 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------ K4
template <int R, bool POW2>
__global__ void __launch_bounds__(CTA_THREADS)
bl_grad_kernel(const float2* __restrict__ pts_all, const int32_t* __restrict__ meta, int batch, Geom g,
               Scale k, int use_bg, float inv_batch, const float* __restrict__ grad_loss,
               const float* __restrict__ amax_in, const float* __restrict__ rz_in,
               const float* __restrict__ pbg_in, const float* __restrict__ wsel,
               float* __restrict__ grad_density) {
    __shared__ WarpTile<R> tiles[WARPS_PER_CTA];
    TaskInfo t;
    if (!decode_task<R>(meta, batch, g, t)) return;
    WarpTile<R>& tile = tiles[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const size_t img_base = (size_t)t.img * g.hp * g.wp;
    const float gscale = grad_loss[0] * inv_batch;
    const float* w_rows = wsel + t.row0;
    const bool has_bg_row = use_bg || t.n_pts == 0;
    const float w_bg = has_bg_row ? w_rows[t.n_rows - 1] : 0.f;

    float acc[R], neg_amax[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const size_t m = img_base + (size_t)min(t.row_base + r, g.hp - 1) * g.wp + t.col;
        neg_amax[r] = -amax_in[m];
        acc[r] = 0.f;
    }

    if (t.n_pts > 0) {
        float cym2[R], cyy[R];
        row_constants<R>(t, g, cym2, cyy);
        const float cx = cell_centre(t.col, g);
        const float cxm2 = -2.0f * cx, cxx = __fmul_rn(cx, cx);
        const float2* pts = pts_all + t.pt0;
        for (int n0 = 0; n0 < t.n_pts; n0 += TILE_PTS) {
            const int cnt = min(TILE_PTS, t.n_pts - n0);
            __syncwarp();
            stage_points<R>(tile, pts, n0, t.n_pts, cnt, cym2, cyy);
            for (int i = lane; i < cnt; i += 32) tile.aux[i] = w_rows[n0 + i];
            __syncwarp();
#pragma unroll 2
            for (int i = 0; i < cnt; ++i) {
                const float w = tile.aux[i];
                if (w == 0.f) continue;  // trimmed by the top-k (bl.py:77): contributes nothing
                const float2 xs = tile.xs[i];
                float yd[R];
                load_yd<R>(tile, i, yd);
                const float xd = axis_sqdist(xs.x, xs.y, cxm2, cxx);
#pragma unroll
                for (int r = 0; r < R; ++r)
                    acc[r] = fmaf(pair_exp<POW2>(__fadd_rn(yd[r], xd), neg_amax[r], k), w, acc[r]);
            }
        }
    }

    if (t.col_ok) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (t.row_base + r >= g.hp) continue;
            const size_t m = img_base + (size_t)(t.row_base + r) * g.wp + t.col;
            grad_density[m] = gscale * fmaf(acc[r], rz_in[m], w_bg * pbg_in[m]);
        }
    }
}

// ------------------------------------------------------------------------- posterior (API parity)
template <int R, bool POW2>
__global__ void __launch_bounds__(CTA_THREADS)
bl_posterior_kernel(const float2* __restrict__ pts_all, const int32_t* __restrict__ meta, int batch, Geom g,
                    Scale k, int use_bg, const float* __restrict__ amax_in, const float* __restrict__ rz_in,
                    const float* __restrict__ pbg_in, float* __restrict__ prob_out) {
    __shared__ WarpTile<R> tiles[WARPS_PER_CTA];
    TaskInfo t;
    if (!decode_task<R>(meta, batch, g, t)) return;
    WarpTile<R>& tile = tiles[threadIdx.x >> 5];
    const size_t M = (size_t)g.hp * g.wp;
    const size_t img_base = (size_t)t.img * M;
    float* prob = prob_out + (size_t)t.row0 * M;

    float neg_amax[R], rz[R];
    size_t pix[R];
    bool ok[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        ok[r] = t.col_ok && (t.row_base + r < g.hp);
        pix[r] = (size_t)min(t.row_base + r, g.hp - 1) * g.wp + t.col;
        neg_amax[r] = -amax_in[img_base + pix[r]];
        rz[r] = rz_in[img_base + pix[r]];
        if (ok[r] && (use_bg || t.n_pts == 0)) prob[(size_t)(t.n_rows - 1) * M + pix[r]] = pbg_in[img_base + pix[r]];
    }
    if (t.n_pts == 0) return;

    float cym2[R], cyy[R];
    row_constants<R>(t, g, cym2, cyy);
    const float cx = cell_centre(t.col, g);
    const float cxm2 = -2.0f * cx, cxx = __fmul_rn(cx, cx);
    const float2* pts = pts_all + t.pt0;
    for (int n0 = 0; n0 < t.n_pts; n0 += TILE_PTS) {
        const int cnt = min(TILE_PTS, t.n_pts - n0);
        __syncwarp();
        stage_points<R>(tile, pts, n0, t.n_pts, cnt, cym2, cyy);
        __syncwarp();
        for (int i = 0; i < cnt; ++i) {
            const float2 xs = tile.xs[i];
            float yd[R];
            load_yd<R>(tile, i, yd);
            const float xd = axis_sqdist(xs.x, xs.y, cxm2, cxx);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float p = pair_exp<POW2>(__fadd_rn(yd[r], xd), neg_amax[r], k) * rz[r];
                if (ok[r]) prob[(size_t)(n0 + i) * M + pix[r]] = p;
            }
        }
    }
}

// ------------------------------------------------------- Bay_Loss on materialised posteriors
// counts[row] = sum_m D[m] * prob[row, m]: one warp per row, coalesced, fixed order.
__global__ void __launch_bounds__(256)
bl_prob_counts_kernel(const float* __restrict__ prob, const float* __restrict__ density,
                      const int32_t* __restrict__ meta, int batch, int M, int64_t total_rows,
                      float* __restrict__ cpart) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= total_rows) return;
    const int lane = threadIdx.x & 31;
    const int32_t* row_off = meta + (batch + 1);
    int img = 0;
    while (img + 1 < batch && row >= row_off[img + 1]) ++img;
    const float* d = density + (size_t)img * M;
    const float* p = prob + (size_t)row * M;
    float s = 0.f;
    for (int m = lane; m < M; m += 32) s = fmaf(d[m], p[m], s);
    s = warp_sum(s);
    if (lane == 0) cpart[row] = s;
}

// grad[b,m] = g * sum_rows w[row] * prob[row,m]
__global__ void __launch_bounds__(256)
bl_prob_grad_kernel(const float* __restrict__ prob, const int32_t* __restrict__ meta, int batch, int M,
                    float inv_batch, const float* __restrict__ grad_loss, const float* __restrict__ wsel,
                    float* __restrict__ grad_density) {
    const int img = blockIdx.y;
    const int m = blockIdx.x * 256 + threadIdx.x;
    if (m >= M) return;
    const int32_t* row_off = meta + (batch + 1);
    const int r0 = row_off[img], r1 = row_off[img + 1];
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) {
        const float w = wsel[r];
        if (w != 0.f) acc = fmaf(w, prob[(size_t)r * M + m], acc);
    }
    grad_density[(size_t)img * M + m] = grad_loss[0] * inv_batch * acc;
}

// ------------------------------------------------------------------------------- host side
static bool is_pow2(float s) {
    int e;
    return s > 0.f && frexpf(s, &e) == 0.5f;
}

static Scale make_scale(float sigma) {
    Scale k;
    k.s = (float)(2.0 * (double)sigma * (double)sigma);  // python: 2.0 * sigma ** 2, then cast to fp32
    k.r = 1.0f / k.s;
    k.neg_inv_s = -k.r;
    return k;
}

static int pick_rows_per_thread(int batch, int hp, int wp) {
    const long col_blocks = ceil_div(wp, 32);
    const long want = 148L * 16;  // at least ~16 warps per SM
    for (int r : {8, 4}) {
        if ((long)batch * col_blocks * ceil_div(hp, r) >= want) return r;
    }
    return 2;
}

static Geom make_geom(int hp, int wp, int R, float stride) {
    Geom g;
    g.hp = hp; g.wp = wp;
    g.col_blocks = ceil_div(wp, 32);
    g.tiles = g.col_blocks * ceil_div(hp, R);
    g.stride = stride;
    g.half = stride / 2.0f;
    return g;
}

static int layout(int64_t total_rows, int batch, int hp, int wp, int tiles_override, dgvcc_bl_layout* L) {
    if (!L || total_rows <= 0 || batch <= 0 || hp <= 0 || wp <= 0) return DGVCC_ERR_ARG;
    const int R = pick_rows_per_thread(batch, hp, wp);
    const int tiles = tiles_override > 0 ? tiles_override : make_geom(hp, wp, R, 1.f).tiles;
    const size_t pix = (size_t)batch * hp * wp * sizeof(float);
    const size_t rows = (size_t)total_rows * sizeof(float);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return (int64_t)o; };
    L->amax = take(pix); L->rz = take(pix); L->pbg = take(pix);
    L->counts = take(rows); L->wsel = take(rows); L->residual = take(rows);
    L->loss_img = take((size_t)batch * sizeof(float));
    L->ticket = take(sizeof(unsigned int));
    L->cpart = take((size_t)tiles * rows);
    L->total = (int64_t)off;
    L->tiles = tiles;
    L->rows_per_thread = R;
    return DGVCC_OK;
}

template <typename T>
static T* at(const void* ws, int64_t off) { return reinterpret_cast<T*>((char*)ws + off); }

}  // namespace bl
}  // namespace dgvcc

using namespace dgvcc;
using namespace dgvcc::bl;

extern "C" int dgvcc_abi_version(void) { return 1; }

extern "C" int dgvcc_bl_workspace_layout(int64_t total_rows, int batch, int hp, int wp, dgvcc_bl_layout* out) {
    return layout(total_rows, batch, hp, wp, 0, out);
}

#define BL_DISPATCH(R_, POW2_, KERNEL, GRID, STREAM, ...)                                      \
    do {                                                                                       \
        if ((R_) == 8) {                                                                       \
            if (POW2_) KERNEL<8, true><<<GRID, CTA_THREADS, 0, STREAM>>>(__VA_ARGS__);         \
            else KERNEL<8, false><<<GRID, CTA_THREADS, 0, STREAM>>>(__VA_ARGS__);              \
        } else if ((R_) == 4) {                                                                \
            if (POW2_) KERNEL<4, true><<<GRID, CTA_THREADS, 0, STREAM>>>(__VA_ARGS__);         \
            else KERNEL<4, false><<<GRID, CTA_THREADS, 0, STREAM>>>(__VA_ARGS__);              \
        } else {                                                                               \
            if (POW2_) KERNEL<2, true><<<GRID, CTA_THREADS, 0, STREAM>>>(__VA_ARGS__);         \
            else KERNEL<2, false><<<GRID, CTA_THREADS, 0, STREAM>>>(__VA_ARGS__);              \
        }                                                                                      \
    } while (0)

static int check_common(const void* a, const void* b, const void* ws, int batch, int hp, int wp,
                        int64_t total_rows, float stride, float sigma) {
    if (!a || !b || !ws) return DGVCC_ERR_ARG;
    if (batch <= 0 || hp <= 0 || wp <= 0 || total_rows < batch) return DGVCC_ERR_ARG;
    if (!(stride > 0.f) || !(sigma > 0.f)) return DGVCC_ERR_ARG;
    return DGVCC_OK;
}

static int launch_minz(const float* pts_xy, const int32_t* meta, const float* st_sizes, int batch,
                       const Geom& g, const Scale& k, int R, float bg_ratio, int use_bg,
                       const dgvcc_bl_layout& L, void* ws, cudaStream_t st) {
    const dim3 grid(ceil_div(g.tiles, WARPS_PER_CTA), batch);
    BL_DISPATCH(R, is_pow2(k.s), bl_minz_kernel, grid, st, (const float2*)pts_xy, meta, st_sizes, batch, g, k,
                bg_ratio, use_bg, at<float>(ws, L.amax), at<float>(ws, L.rz), at<float>(ws, L.pbg));
    return (int)cudaGetLastError();
}

static int launch_select(const float* targets, const int32_t* meta, int batch, int64_t total_rows,
                         float inv_batch, int tiles, const dgvcc_bl_layout& L, void* ws, float* loss_out,
                         cudaStream_t st) {
    bl_select_kernel<<<batch, SELECT_THREADS, 0, st>>>(
        at<float>(ws, L.cpart), tiles, total_rows, meta, targets, batch, inv_batch, at<float>(ws, L.counts),
        at<float>(ws, L.residual), at<float>(ws, L.wsel), at<float>(ws, L.loss_img), loss_out,
        at<unsigned int>(ws, L.ticket));
    return (int)cudaGetLastError();
}

extern "C" int dgvcc_bl_forward(const float* pts_xy, const float* targets, const int32_t* meta,
                                const float* st_sizes, const float* density, int batch, int hp, int wp,
                                int64_t total_rows, float stride, float sigma, float bg_ratio, int use_bg,
                                float inv_batch, void* workspace, size_t workspace_bytes, float* loss_out,
                                void* stream) {
    int rc = check_common(meta, density, workspace, batch, hp, wp, total_rows, stride, sigma);
    if (rc) return rc;
    if (!st_sizes || !loss_out) return DGVCC_ERR_ARG;
    dgvcc_bl_layout L;
    if ((rc = layout(total_rows, batch, hp, wp, 0, &L))) return rc;
    if (workspace_bytes < (size_t)L.total) return DGVCC_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int R = L.rows_per_thread;
    const Geom g = make_geom(hp, wp, R, stride);
    const Scale k = make_scale(sigma);
    if ((rc = launch_minz(pts_xy, meta, st_sizes, batch, g, k, R, bg_ratio, use_bg, L, workspace, st))) return rc;
    const dim3 grid(ceil_div(g.tiles, WARPS_PER_CTA), batch);
    BL_DISPATCH(R, is_pow2(k.s), bl_counts_kernel, grid, st, (const float2*)pts_xy, meta, density, batch, g, k,
                use_bg, at<float>(workspace, L.amax), at<float>(workspace, L.rz), at<float>(workspace, L.pbg),
                total_rows, at<float>(workspace, L.cpart));
    DGVCC_RETURN_IF_CUDA(cudaGetLastError());
    return launch_select(targets, meta, batch, total_rows, inv_batch, L.tiles, L, workspace, loss_out, st);
}

extern "C" int dgvcc_bl_backward(const float* pts_xy, const int32_t* meta, int batch, int hp, int wp,
                                 int64_t total_rows, float stride, float sigma, int use_bg, float inv_batch,
                                 const float* grad_loss, const void* workspace, size_t workspace_bytes,
                                 float* grad_density, void* stream) {
    int rc = check_common(meta, grad_loss, workspace, batch, hp, wp, total_rows, stride, sigma);
    if (rc) return rc;
    if (!grad_density) return DGVCC_ERR_ARG;
    dgvcc_bl_layout L;
    if ((rc = layout(total_rows, batch, hp, wp, 0, &L))) return rc;
    if (workspace_bytes < (size_t)L.total) return DGVCC_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int R = L.rows_per_thread;
    const Geom g = make_geom(hp, wp, R, stride);
    const Scale k = make_scale(sigma);
    const dim3 grid(ceil_div(g.tiles, WARPS_PER_CTA), batch);
    BL_DISPATCH(R, is_pow2(k.s), bl_grad_kernel, grid, st, (const float2*)pts_xy, meta, batch, g, k, use_bg,
                inv_batch, grad_loss, at<float>(workspace, L.amax), at<float>(workspace, L.rz),
                at<float>(workspace, L.pbg), at<float>(workspace, L.wsel), grad_density);
    return (int)cudaGetLastError();
}

extern "C" int dgvcc_bl_posterior(const float* pts_xy, const int32_t* meta, const float* st_sizes, int batch,
                                  int hp, int wp, int64_t total_rows, float stride, float sigma, float bg_ratio,
                                  int use_bg, void* workspace, size_t workspace_bytes, float* prob_out,
                                  void* stream) {
    int rc = check_common(meta, st_sizes, workspace, batch, hp, wp, total_rows, stride, sigma);
    if (rc) return rc;
    if (!prob_out) return DGVCC_ERR_ARG;
    dgvcc_bl_layout L;
    if ((rc = layout(total_rows, batch, hp, wp, 0, &L))) return rc;
    if (workspace_bytes < (size_t)L.total) return DGVCC_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int R = L.rows_per_thread;
    const Geom g = make_geom(hp, wp, R, stride);
    const Scale k = make_scale(sigma);
    if ((rc = launch_minz(pts_xy, meta, st_sizes, batch, g, k, R, bg_ratio, use_bg, L, workspace, st))) return rc;
    const dim3 grid(ceil_div(g.tiles, WARPS_PER_CTA), batch);
    BL_DISPATCH(R, is_pow2(k.s), bl_posterior_kernel, grid, st, (const float2*)pts_xy, meta, batch, g, k, use_bg,
                at<float>(workspace, L.amax), at<float>(workspace, L.rz), at<float>(workspace, L.pbg), prob_out);
    return (int)cudaGetLastError();
}

extern "C" int dgvcc_bl_bayloss_forward(const float* prob, const float* targets, const int32_t* meta,
                                        const float* density, int batch, int hp, int wp, int64_t total_rows,
                                        float inv_batch, void* workspace, size_t workspace_bytes,
                                        float* loss_out, void* stream) {
    if (!prob || !meta || !density || !workspace || !loss_out) return DGVCC_ERR_ARG;
    if (batch <= 0 || hp <= 0 || wp <= 0 || total_rows < batch) return DGVCC_ERR_ARG;
    dgvcc_bl_layout L;
    int rc;
    if ((rc = layout(total_rows, batch, hp, wp, 0, &L))) return rc;
    if (workspace_bytes < (size_t)L.total) return DGVCC_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int M = hp * wp;
    bl_prob_counts_kernel<<<(unsigned)((total_rows + 7) / 8), 256, 0, st>>>(prob, density, meta, batch, M, total_rows,
                                                                            at<float>(workspace, L.cpart));
    DGVCC_RETURN_IF_CUDA(cudaGetLastError());
    return launch_select(targets, meta, batch, total_rows, inv_batch, /*tiles=*/1, L, workspace, loss_out, st);
}

extern "C" int dgvcc_bl_bayloss_backward(const float* prob, const int32_t* meta, int batch, int hp, int wp,
                                         int64_t total_rows, float inv_batch, const float* grad_loss,
                                         const void* workspace, size_t workspace_bytes, float* grad_density,
                                         void* stream) {
    if (!prob || !meta || !grad_loss || !workspace || !grad_density) return DGVCC_ERR_ARG;
    if (batch <= 0 || hp <= 0 || wp <= 0 || total_rows < batch) return DGVCC_ERR_ARG;
    dgvcc_bl_layout L;
    int rc;
    if ((rc = layout(total_rows, batch, hp, wp, 0, &L))) return rc;
    if (workspace_bytes < (size_t)L.total) return DGVCC_ERR_WORKSPACE;
    const int M = hp * wp;
    bl_prob_grad_kernel<<<dim3(ceil_div(M, 256), batch), 256, 0, (cudaStream_t)stream>>>(
        prob, meta, batch, M, inv_batch, grad_loss, at<float>(workspace, L.wsel), grad_density);
    return (int)cudaGetLastError();
}
