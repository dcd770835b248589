// Tensor-core (tcgen05 + TMA) Gram partials for the ISW covariance -- placeholder until the kernel lands.
#include "common.cuh"
#include "../../include/dgvcc_b200.h"

extern "C" int dgvcc_isw_gram_tc_partials(const float* x, int batch, int c, int hw, int splits, int k_per_split,
                                          float* part, void* stream) {
    (void)x; (void)batch; (void)c; (void)hw; (void)splits; (void)k_per_split; (void)part; (void)stream;
    return DGVCC_ERR_UNSUPPORTED;
}
