// Training step: forward + backward of one network's photometric-MSE term (placeholder until the
// fused backward lands).
#include "common.cuh"
using namespace nerfb200;
extern "C" {
size_t nerf_b200_train_workspace_bytes(int n_rays, int n_samples) { (void)n_rays; (void)n_samples; return 0; }
int nerf_b200_train_fwd_bwd(const void *, const nerf_b200_params *, const nerf_b200_params *, const float *,
                            const float *, const float *, int, int, float, float, const float *, int, int,
                            void *, float *, float *, void *)
{
    return NERF_B200_EUNSUPPORTED;
}
}
