// Data path on the device: what SyntheticDataset does per image on the host in float64
// (src/data/loader.py:42-64) and what NeRFTrainer.train_step does per iteration with full-image tensors
// (src/training/trainer.py:100-118, 271-292).
//
//   composite_white_kernel   RGBA8 pixels -> fp32 RGB on a white background, bit for bit the reference's
//                            float64 arithmetic rounded once to fp32: c = r/255, a = A/255, fl32(c*a + (1 - a)).
//                            4 bytes in, 12 bytes out per pixel: the images cross PCIe as uint8, a sixth of the float64
//                            arrays the reference builds.
//   ray_batch_kernel         the ray batch of a training step straight from the pixel indices: origins, directions
//                            (the same bits as _get_rays(pose)[index]) and target colours gathered from the
//                            device-resident image -- the full-image ray tensors (24 B per pixel per step) never exist.
#include "common.cuh"
#include <algorithm>

namespace nerfb200 {

__global__ void __launch_bounds__(256) composite_white_kernel(const uchar4 *__restrict__ rgba, size_t n, float *__restrict__ rgb)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uchar4 p = __ldg(rgba + i);
        // numpy: img / 255.0 (float64), rgb * alpha + (1 - alpha): three roundings, no contraction
        const double a = __ddiv_rn((double)p.w, 255.0), ia = __dsub_rn(1.0, a);
        rgb[3 * i + 0] = __double2float_rn(__dadd_rn(__dmul_rn(__ddiv_rn((double)p.x, 255.0), a), ia));
        rgb[3 * i + 1] = __double2float_rn(__dadd_rn(__dmul_rn(__ddiv_rn((double)p.y, 255.0), a), ia));
        rgb[3 * i + 2] = __double2float_rn(__dadd_rn(__dmul_rn(__ddiv_rn((double)p.z, 255.0), a), ia));
    }
}

// Vector path (16-byte aligned input, whole groups of four pixels): a thread takes four pixels with one 16-byte load;
// c / 255 is formed WITHOUT a division -- q0 = c * (1/255), r = fma(-q0, 255, c), q = fma(r, 1/255, q0) is the correctly
// rounded quotient for every c in 0..255 (checked exhaustively against exact rational arithmetic; q0 alone is off by an ulp
// for 24 of them) -- where the scalar kernel spends most of its time in four double-precision divisions per pixel, and a
// 256-entry table of the quotients is bound by its shared-memory bank conflicts (ncu: L1 96 % busy, 4.1 TB/s).  The 48 bytes
// a thread produces go through a per-warp staging buffer so that every store instruction writes 512 contiguous bytes.
// Same bits as the scalar kernel.
__device__ __forceinline__ double div255(uint32_t c)
{
    const double x = (double)c, inv = 1.0 / 255.0;
    const double q0 = __dmul_rn(x, inv);
    return __fma_rn(__fma_rn(-q0, 255.0, x), inv, q0);
}
__global__ void __launch_bounds__(256) composite_white4_kernel(const uint4 *__restrict__ rgba4, size_t n4, float *__restrict__ rgb)
{
    __shared__ __align__(16) float stage[8][32 * 12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 *mine = reinterpret_cast<float4 *>(stage[warp]);
    const size_t n_warp_items = (n4 + 31) / 32;                  // a warp item = 32 groups = 128 pixels
    for (size_t item = (size_t)blockIdx.x * 8 + warp; item < n_warp_items; item += (size_t)gridDim.x * 8) {
        const size_t g = item * 32 + lane;
        if (g < n4) {
            const uint4 q = __ldg(rgba4 + g);
            const uint32_t px[4] = {q.x, q.y, q.z, q.w};
            float o[12];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double a = div255(px[i] >> 24), ia = __dsub_rn(1.0, a);
                o[3 * i + 0] = __double2float_rn(__dadd_rn(__dmul_rn(div255(px[i] & 255u), a), ia));
                o[3 * i + 1] = __double2float_rn(__dadd_rn(__dmul_rn(div255((px[i] >> 8) & 255u), a), ia));
                o[3 * i + 2] = __double2float_rn(__dadd_rn(__dmul_rn(div255((px[i] >> 16) & 255u), a), ia));
            }
            mine[3 * lane + 0] = make_float4(o[0], o[1], o[2], o[3]);
            mine[3 * lane + 1] = make_float4(o[4], o[5], o[6], o[7]);
            mine[3 * lane + 2] = make_float4(o[8], o[9], o[10], o[11]);
        }
        __syncwarp();
        const size_t left = n4 - item * 32;                        // groups of this item: 3 float4 each
        const int n_out = (int)(left < 32 ? left : 32) * 3;
        float4 *dst = reinterpret_cast<float4 *>(rgb) + item * 96;
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (j * 32 + lane < n_out) dst[j * 32 + lane] = mine[j * 32 + lane];
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) ray_batch_kernel(Pose pose, int width, float half_w, float half_h, float focal,
                                                        const long long *__restrict__ index, int n, const float *__restrict__ image,
                                                        float *__restrict__ rays_o, float *__restrict__ rays_d, float *__restrict__ target)
{
    const int total = 3 * n;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int r = e / 3, c = e - 3 * r;
        const long long pix = __ldg(index + r);
        const int j = (int)(pix / width), i = (int)(pix - (long long)j * width);
        float dx, dy;
        pixel_dir(i, j, half_w, half_h, focal, dx, dy);
        rays_d[e] = rotate_dir(pose, c, dx, dy);
        rays_o[e] = pose.t[c];
        if (target) target[e] = __ldg(image + 3 * pix + c);
    }
}

}  // namespace nerfb200

using namespace nerfb200;

static int blocks_for(size_t items)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)std::max<size_t>(1, std::min<size_t>((items + 255) / 256, (size_t)sms * 8));
}

extern "C" {

int nerf_b200_composite_white(const unsigned char *rgba, int64_t n_pixels, float *rgb_out, void *stream)
{
    if (!rgba || !rgb_out || n_pixels <= 0) return NERF_B200_EINVAL;
    if ((uintptr_t)rgba & 3) return NERF_B200_EALIGN;
    size_t done = 0;
    if (n_pixels >= 4 && (((uintptr_t)rgba | (uintptr_t)rgb_out) & 15) == 0) {
        const size_t n4 = (size_t)n_pixels / 4;
        composite_white4_kernel<<<blocks_for(n4), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4 *>(rgba), n4, rgb_out);
        done = n4 * 4;
        const int rc = launch_status();
        if (rc != 0 || done == (size_t)n_pixels) return rc;
    }
    composite_white_kernel<<<blocks_for((size_t)n_pixels - done), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uchar4 *>(rgba) + done, (size_t)n_pixels - done, rgb_out + 3 * done);
    return launch_status();
}

int nerf_b200_ray_batch(const float *c2w_host, int width, int height, float focal, const int64_t *pixel_index, int n,
                        const float *image, float *rays_o, float *rays_d, float *target, void *stream)
{
    if (!c2w_host || !pixel_index || !rays_o || !rays_d || width <= 0 || height <= 0 || n <= 0 || !(focal > 0.f) ||
        (target && !image))
        return NERF_B200_EINVAL;
    ray_batch_kernel<<<blocks_for((size_t)n * 3), 256, 0, (cudaStream_t)stream>>>(
        pose_from_c2w(c2w_host), width, (float)((double)width * 0.5), (float)((double)height * 0.5), focal,
        reinterpret_cast<const long long *>(pixel_index), n, image, rays_o, rays_d, target);
    return launch_status();
}

}  // extern "C"
