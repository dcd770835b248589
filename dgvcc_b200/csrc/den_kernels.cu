// Dataset-side density handling for sm_100a.
//
// Replaces the density part of DenClsDataset._train_transform (datasets/den_cls_dataset.py:109-150; the same
// block is in datasets/den_dataset.py:86-127) and the block-occupancy map of DenClsDataset.__getitem__
// (den_cls_dataset.py:60-61): zero padding, crop, d x d sum-pool, horizontal flip, then 16 x 16 block sums > 0.
// One launch for a whole batch of full-resolution maps (e.g. straight from dmap_splat_kernel, without the
// *_dmap.npy round trip).  HBM-bound: every source pixel of a crop is read once, every target written once.
#include "common.cuh"
#include "../../include/dgvcc_b200.h"

namespace dgvcc {
namespace den {

constexpr int BLOCK = 16;  // den_cls_dataset.py:60: reshape(1, H/16, 16, W/16, 16)

enum MetaCol { M_SRC = 0, M_H, M_W, M_LEFT, M_TOP, M_I, M_J, M_FLIP, META_COLS = DGVCC_DEN_META_COLS };
static_assert(META_COLS == 8, "header and kernel disagree on the meta row width");

// One CTA = one 16 x 16 block of pooled pixels of one image = one occupancy cell.
__global__ void __launch_bounds__(BLOCK * BLOCK)
den_train_targets_kernel(const float* __restrict__ maps, const int64_t* __restrict__ meta, int dh, int dw, int d,
                         float* __restrict__ out_dmap, float* __restrict__ out_bmap) {
    __shared__ float warp_part[BLOCK * BLOCK / 32];
    const int64_t* m = meta + (size_t)blockIdx.z * META_COLS;
    const float* src = maps + m[M_SRC];
    const int height = (int)m[M_H], width = (int)m[M_W];
    const int ox = blockIdx.x * BLOCK + threadIdx.x, oy = blockIdx.y * BLOCK + threadIdx.y;
    float sum = 0.f;
    if (ox < dw && oy < dh) {
        // source pixel of pooled (oy, ox), tap (r, c): padded (i + oy*d + r, j + ox*d + c) - (top, left)
        const int sy0 = (int)m[M_I] + oy * d - (int)m[M_TOP], sx0 = (int)m[M_J] + ox * d - (int)m[M_LEFT];
        for (int r = 0; r < d; ++r) {
            const int sy = sy0 + r;
            if (sy < 0 || sy >= height) continue;  // zero padding (F.pad, den_cls_dataset.py:116)
            const float* row = src + (size_t)sy * width;
            for (int c = 0; c < d; ++c) {
                const int sx = sx0 + c;
                if (sx >= 0 && sx < width) sum = __fadd_rn(sum, __ldg(row + sx));
            }
        }
        const int tx = m[M_FLIP] ? dw - 1 - ox : ox;  // F.hflip, den_cls_dataset.py:144
        DGVCC_DEV_CHECK(m[M_SRC] >= 0 && height > 0 && width > 0 && d >= 1 && tx >= 0 && tx < dw);
        out_dmap[((size_t)blockIdx.z * dh + oy) * dw + tx] = sum;
    }
    if (!out_bmap) return;  // uniform
    // block sum > 0 (den_cls_dataset.py:60-61); fixed-order tree, the sign test is order-independent for the
    // non-negative maps the generator produces
    float s = warp_sum(sum);
    const int tid = threadIdx.y * BLOCK + threadIdx.x;
    if ((tid & 31) == 0) warp_part[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < BLOCK * BLOCK / 32; ++w) t += warp_part[w];
        const int bw = dw / BLOCK, bx = m[M_FLIP] ? bw - 1 - blockIdx.x : blockIdx.x;
        out_bmap[((size_t)blockIdx.z * (dh / BLOCK) + blockIdx.y) * bw + bx] = t > 0.f ? 1.f : 0.f;
    }
}

}  // namespace den
}  // namespace dgvcc

using namespace dgvcc;
using namespace dgvcc::den;

extern "C" int dgvcc_den_train_targets(const float* maps, const int64_t* meta, int batch, int crop_h, int crop_w,
                                       int downsample, float* out_dmap, float* out_bmap, void* stream) {
    DGVCC_DEVICE_GUARD(stream);
    if (!maps || !meta || !out_dmap || batch <= 0 || crop_h <= 0 || crop_w <= 0 || downsample <= 0) return DGVCC_ERR_ARG;
    // the reference's reshape needs exact divisibility (den_cls_dataset.py:138, :60)
    if (crop_h % downsample || crop_w % downsample) return DGVCC_ERR_ARG;
    const int dh = crop_h / downsample, dw = crop_w / downsample;
    if (out_bmap && (dh % BLOCK || dw % BLOCK)) return DGVCC_ERR_ARG;
    if (batch > 65535) return DGVCC_ERR_UNSUPPORTED;
    den_train_targets_kernel<<<dim3(ceil_div(dw, BLOCK), ceil_div(dh, BLOCK), batch), dim3(BLOCK, BLOCK), 0,
                               (cudaStream_t)stream>>>(maps, meta, dh, dw, downsample, out_dmap, out_bmap);
    return (int)cudaGetLastError();
}
