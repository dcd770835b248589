// Host-side CSR packing of a ragged Bayesian-loss batch (no CUDA call in this file).
//
// What bl.py:21-22 does with torch.cat on the device -- concatenate the per-image point lists -- plus the small
// int32 table the kernels read (include/dgvcc_b200.h), done in one pass over the host arrays straight into the
// caller's (pinned) staging buffer, so that the training step uploads ONE buffer.  The Python path
// (dgvcc_b200/losses/bl.py: build_meta + numpy concatenate) produces the same bytes; tests/test_abi.py compares them.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "../../include/dgvcc_b200.h"

static inline size_t align16(size_t n) { return (n + 15) / 16 * 16; }

extern "C" int dgvcc_bl_pack_host(const float* const* points, const float* const* targets, const int32_t* counts, int batch,
                                  int use_background, int chunk, void* dst, size_t dst_bytes, dgvcc_bl_packed* info) {
    if (!counts || !info || batch <= 0 || chunk <= 0) return DGVCC_ERR_ARG;
    int64_t total_points = 0, total_rows = 0, total_chunks = 0, max_chunks = 0;
    for (int i = 0; i < batch; ++i) {
        if (counts[i] < 0) return DGVCC_ERR_ARG;
        const int64_t n = counts[i], nc = std::max<int64_t>(1, (n + chunk - 1) / chunk);
        total_points += n;
        total_rows += n == 0 ? 1 : n + (use_background ? 1 : 0);
        total_chunks += nc;
        max_chunks = std::max(max_chunks, nc);
    }
    if (total_points > 0x7fffffffLL || total_rows > 0x7fffffffLL) return DGVCC_ERR_UNSUPPORTED;
    const int b = batch;
    const size_t meta_ints = (size_t)4 * b + 3 + 4 * (size_t)total_chunks;
    const size_t n_pts = (size_t)std::max<int64_t>(total_points, 1);
    info->total_points = total_points;
    info->total_rows = total_rows;
    info->total_chunks = total_chunks;
    info->multi_chunk = max_chunks > 1;
    info->meta_bytes = (int64_t)(meta_ints * 4);
    info->off_points = (int64_t)align16(meta_ints * 4);
    info->off_targets = info->off_points + (int64_t)align16(8 * n_pts);
    info->total_bytes = info->off_targets + (int64_t)align16(4 * n_pts);
    if (!dst) return DGVCC_OK;  // size query
    if (dst_bytes < (size_t)info->total_bytes) return DGVCC_ERR_WORKSPACE;
    if (total_points > 0 && !points) return DGVCC_ERR_ARG;

    int32_t* meta = (int32_t*)dst;
    int32_t* pt_off = meta, * row_off = meta + (b + 1), * keep = meta + (2 * b + 2), * icb = meta + (3 * b + 2);
    int32_t* table = meta + (4 * b + 3);
    pt_off[0] = row_off[0] = icb[0] = 0;
    float* out_pts = (float*)((char*)dst + info->off_points);
    float* out_tgt = (float*)((char*)dst + info->off_targets);
    int32_t c = 0;
    for (int i = 0; i < b; ++i) {
        const int64_t n = counts[i], rows = n == 0 ? 1 : n + (use_background ? 1 : 0);
        const int64_t nc = std::max<int64_t>(1, (n + chunk - 1) / chunk);
        pt_off[i + 1] = pt_off[i] + (int32_t)n;
        row_off[i + 1] = row_off[i] + (int32_t)rows;
        keep[i] = (int32_t)ceil(0.9 * (double)(rows - 1));  // bl.py:76, the same double arithmetic as Python's
        icb[i + 1] = icb[i] + (int32_t)nc;
        for (int64_t k = 0; k < nc; ++k, ++c) {  // near-equal slices
            const int64_t start = n * k / nc, stop = n * (k + 1) / nc;
            table[4 * c] = i; table[4 * c + 1] = (int32_t)start; table[4 * c + 2] = (int32_t)(stop - start);
        }
        if (n > 0) {
            if (!points[i]) return DGVCC_ERR_ARG;
            memcpy(out_pts + 2 * (size_t)pt_off[i], points[i], (size_t)n * 8);
            if (targets) {
                if (!targets[i]) return DGVCC_ERR_ARG;
                memcpy(out_tgt + pt_off[i], targets[i], (size_t)n * 4);
            }
        }
    }
    // schedule: longest chunks first (the sweeps hand the launch slots out through a work queue: a
    // longest-processing-time schedule, the tail of a sweep is made of the shortest tasks)
    std::vector<int32_t> order(total_chunks);
    for (int32_t k = 0; k < (int32_t)total_chunks; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return table[4 * x + 2] > table[4 * y + 2]; });
    for (int32_t slot = 0; slot < (int32_t)total_chunks; ++slot) table[4 * slot + 3] = order[slot];
    return DGVCC_OK;
}
