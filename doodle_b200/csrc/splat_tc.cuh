// K2 / K3, tcgen05 path (3xTF32 tensor-core contraction).  Placeholder until the kernels land.
#pragma once
#include "helio_common.cuh"

namespace helio {
inline bool splat_tc_fwd_supported(int, int, int) { return false; }
inline bool splat_tc_fwd_preferred(int, int, int) { return false; }
inline bool splat_tc_bwd_supported(int, int, int) { return false; }
inline bool splat_tc_bwd_preferred(int, int, int) { return false; }
inline cudaError_t splat_tc_fwd(const float*, float*, int, int, int, float, float, int, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t splat_tc_bwd(const float*, const float*, float*, int, int, int, float, float, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace helio
