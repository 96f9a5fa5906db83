// extern "C" entry points that dispatch on the precision mode (include/nerf_b200.h).
#include "common.cuh"

namespace nerfb200 {
int simt_query(const void *, const float *, const float *, long long, float *, float *, cudaStream_t);
int simt_render_pose(const void *, const float *, int, int, float, float, float, int, int, int, float *, float *, cudaStream_t);
int simt_render_rays(const void *, const float *, const float *, int, int, float, float, const float *, const float *, float *, float *, float *, float *, cudaStream_t);
int tc_render_pose(const void *, const float *, int, int, float, float, float, int, int, int, bool, float *, float *, unsigned int *, cudaStream_t);
int tc_render_rays(const void *, const float *, const float *, int, int, float, float, const float *, const float *, bool, float *, float *, float *, float *, unsigned int *, cudaStream_t);
int tc_query_points(const void *, const float *, const float *, long long, float *, float *, unsigned int *, cudaStream_t);
int tc_render_pose_fp8(const void *, const float *, int, int, float, float, float, int, int, int, float *, float *, unsigned int *, cudaStream_t);
int tc_render_rays_fp8(const void *, const float *, const float *, int, int, float, float, const float *, const float *, float *, float *, float *, float *, unsigned int *, cudaStream_t);
}
using namespace nerfb200;

// Debug hooks are kept PER DEVICE: a word allocated on cuda:0 is never handed to a kernel on cuda:1.  A pointer is
// filed under the device that owns it (cudaPointerGetAttributes); NULL detaches the current device's hook.
namespace nerfb200 {
constexpr int kMaxDevices = 64;
static std::atomic<unsigned int *> g_watchdog[kMaxDevices];
static std::atomic<long long *> g_trace[kMaxDevices];
static int device_of(const void *p)
{
    int dev = 0;
    if (p) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice) return at.device;
        cudaGetLastError();
    }
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return dev;
}
unsigned int *watchdog_word()
{
    const int d = device_of(nullptr);
    return d >= 0 && d < kMaxDevices ? g_watchdog[d].load(std::memory_order_acquire) : nullptr;
}
long long *trace_buffer()
{
    const int d = device_of(nullptr);
    return d >= 0 && d < kMaxDevices ? g_trace[d].load(std::memory_order_acquire) : nullptr;
}
}  // namespace nerfb200

extern "C" {

void nerf_b200_set_watchdog_word(unsigned int *device_word)
{
    const int d = device_of(device_word);
    if (d >= 0 && d < kMaxDevices) g_watchdog[d].store(device_word, std::memory_order_release);
}
void nerf_b200_set_trace_buffer(long long *device_buf)
{
    const int d = device_of(device_buf);
    if (d >= 0 && d < kMaxDevices) g_trace[d].store(device_buf, std::memory_order_release);
}

int nerf_b200_query_network(const void *packed, const float *positions, const float *directions,
                            int64_t n, int mode, float *sigma, float *rgb, void *stream)
{
    if (!packed || !positions || !directions || !sigma || !rgb || n <= 0) return NERF_B200_EINVAL;
    if (mode == NERF_B200_FP32) return simt_query(packed, positions, directions, n, sigma, rgb, (cudaStream_t)stream);
    if ((uintptr_t)packed & 1023) return NERF_B200_EALIGN;
    // BF16: the fused tensor-core kernel with a (point, direction) pair per row; the direction part of colour layer 0
    // is an fp32 per-row bias built by the back warps
    if (mode == NERF_B200_BF16) return tc_query_points(packed, positions, directions, n, sigma, rgb, watchdog_word(), (cudaStream_t)stream);
    return NERF_B200_EUNSUPPORTED;
}

int nerf_b200_render_image(const void *packed, const float *c2w_host, int width, int height, float focal,
                           float near, float far, int n_samples, int row0, int n_rows, int mode,
                           float *rgb_out, float *depth_out, void *stream)
{
    if (!packed || !c2w_host || !rgb_out || !depth_out || width <= 0 || height <= 0 || n_samples <= 0 ||
        n_rows <= 0 || row0 < 0 || row0 + n_rows > height || !(focal > 0.f))
        return NERF_B200_EINVAL;
    if ((uintptr_t)packed & 1023) return NERF_B200_EALIGN;
    if (mode == NERF_B200_FP32)
        return simt_render_pose(packed, c2w_host, width, height, focal, near, far, n_samples, row0, n_rows,
                                rgb_out, depth_out, (cudaStream_t)stream);
    if (mode == NERF_B200_BF16 || mode == NERF_B200_BF16X3)
        return tc_render_pose(packed, c2w_host, width, height, focal, near, far, n_samples, row0, n_rows,
                              mode == NERF_B200_BF16X3, rgb_out, depth_out, watchdog_word(), (cudaStream_t)stream);
    return NERF_B200_EINVAL;
}

int nerf_b200_render_rays_ex(const void *packed, const float *rays_o, const float *rays_d, int n_rays,
                             int n_samples, float near, float far, const float *t_rand, const float *z_vals, int mode,
                             float *rgb_out, float *depth_out, float *acc_out, float *weights_out, void *stream)
{
    if (!packed || !rays_o || !rays_d || !rgb_out || !depth_out || n_rays <= 0 || n_samples <= 0)
        return NERF_B200_EINVAL;
    if ((uintptr_t)packed & 1023) return NERF_B200_EALIGN;
    if (mode == NERF_B200_FP32)
        return simt_render_rays(packed, rays_o, rays_d, n_rays, n_samples, near, far, t_rand, z_vals, rgb_out,
                                depth_out, acc_out, weights_out, (cudaStream_t)stream);
    if (mode == NERF_B200_BF16 || mode == NERF_B200_BF16X3)
        return tc_render_rays(packed, rays_o, rays_d, n_rays, n_samples, near, far, t_rand, z_vals,
                              mode == NERF_B200_BF16X3, rgb_out, depth_out, acc_out, weights_out, watchdog_word(),
                              (cudaStream_t)stream);
    return NERF_B200_EINVAL;
}

int nerf_b200_render_image_fp8(const void *packed_fp8, const float *c2w_host, int width, int height, float focal, float near,
                               float far, int n_samples, int row0, int n_rows, float *rgb_out, float *depth_out, void *stream)
{
    if (!packed_fp8 || !c2w_host || !rgb_out || !depth_out || width <= 0 || height <= 0 || n_samples <= 0 || n_rows <= 0 ||
        row0 < 0 || row0 + n_rows > height || !(focal > 0.f))
        return NERF_B200_EINVAL;
    if ((uintptr_t)packed_fp8 & 1023) return NERF_B200_EALIGN;
    return tc_render_pose_fp8(packed_fp8, c2w_host, width, height, focal, near, far, n_samples, row0, n_rows, rgb_out, depth_out,
                              watchdog_word(), (cudaStream_t)stream);
}

int nerf_b200_render_rays_fp8(const void *packed_fp8, const float *rays_o, const float *rays_d, int n_rays, int n_samples,
                              float near, float far, const float *t_rand, const float *z_vals, float *rgb_out, float *depth_out,
                              float *acc_out, float *weights_out, void *stream)
{
    if (!packed_fp8 || !rays_o || !rays_d || !rgb_out || !depth_out || n_rays <= 0 || n_samples <= 0) return NERF_B200_EINVAL;
    if ((uintptr_t)packed_fp8 & 1023) return NERF_B200_EALIGN;
    return tc_render_rays_fp8(packed_fp8, rays_o, rays_d, n_rays, n_samples, near, far, t_rand, z_vals, rgb_out, depth_out, acc_out,
                              weights_out, watchdog_word(), (cudaStream_t)stream);
}

int nerf_b200_render_rays(const void *packed, const float *rays_o, const float *rays_d, int n_rays,
                          int n_samples, float near, float far, const float *t_rand, int mode,
                          float *rgb_out, float *depth_out, float *acc_out, void *stream)
{
    return nerf_b200_render_rays_ex(packed, rays_o, rays_d, n_rays, n_samples, near, far, t_rand, nullptr, mode,
                                    rgb_out, depth_out, acc_out, nullptr, stream);
}

}  // extern "C"
