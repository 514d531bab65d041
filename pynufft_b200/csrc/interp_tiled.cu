// placeholder: tiled kernels land in the next commit
#include "common.cuh"
bool tiled_supported(const Geom& g) { (void)g; return false; }
int interp_tiled_launch(b200nufft_plan_t, const float2*, float2*, int, cudaStream_t) { return B200_ERR_UNSUPPORTED; }
int gridding_tiled_launch(b200nufft_plan_t, const float2*, float2*, int, cudaStream_t) { return B200_ERR_UNSUPPORTED; }
