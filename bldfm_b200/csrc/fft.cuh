// fft.cuh -- pruned in-house back-transform (K9+K10+K11 fused).  Placeholder until the shared-memory
// mixed-radix kernels land: reports "unsupported" so that api.cu takes the library (cuFFT) path.
#pragma once

#include "../../include/bldfm_b200.h"
#include "common.cuh"

namespace bldfm {

inline bool pruned_fft_supported(const bldfm_geometry&, bool) { return false; }

inline size_t pruned_fft_work_bytes(const bldfm_geometry&, bool, int64_t) { return 0; }

inline int pruned_fft_run(cudaStream_t, int, size_t, const bldfm_geometry&, bool, bool, const void*,
                          const void*, int64_t, void*, void*, void*, int*)
{
    return BLDFM_ERR_INVALID;
}

}  // namespace bldfm
