// Shared host/device helpers for the nerf_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>

#include "../../include/nerf_b200.h"
#include "packed_layout.h"

namespace nerfb200 {

// launch accounting (nerf_b200_launch_count)
extern std::atomic<unsigned long long> g_launches;
// per-device debug hooks (api.cu): the watchdog word / timeline buffer registered for the CURRENT device, or nullptr
unsigned int *watchdog_word();
long long *trace_buffer();
inline int launch_status()
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return 0;
}

struct Pose {           // rotation rows + translation, passed by value as a kernel argument
    float r[3][3];
    float t[3];
};
inline Pose pose_from_c2w(const float *c2w)
{
    Pose p;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) p.r[i][j] = c2w[4 * i + j];
        p.t[i] = c2w[4 * i + 3];
    }
    return p;
}

// ---- bit-exact recipes (mirrors of oracle/scalar_oracle.c; every rounding is explicit) ----

// torch.linspace(0,1,n)[i]  (reference src/benchmark/base_renderer.py:274)
__device__ __forceinline__ float linspace01(int i, int n, float step)
{
    if (n == 1) return 0.0f;
    return (i < n / 2) ? fmaf(step, (float)i, 0.0f) : fmaf(-step, (float)(n - 1 - i), 1.0f);
}
__device__ __forceinline__ float linspace_step(int n)
{
    return n > 1 ? __fdiv_rn(1.0f, (float)(n - 1)) : 0.0f;
}
// z_i = fl(fl(near*fl(1-t)) + fl(far*t))   (base_renderer.py:275)
__device__ __forceinline__ float depth_uniform(int i, int n, float step, float near, float far)
{
    float t = linspace01(i, n, step);
    return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
}
// stratified jitter (src/utils/rendering.py:42-47) for sample i given t in [0,1)
__device__ __forceinline__ float depth_jittered(int i, int n, float step, float near, float far, float t)
{
    float zc = depth_uniform(i, n, step, near, far);
    float lo = zc, hi = zc;
    if (i > 0) lo = __fmul_rn(0.5f, __fadd_rn(zc, depth_uniform(i - 1, n, step, near, far)));
    if (i < n - 1) hi = __fmul_rn(0.5f, __fadd_rn(depth_uniform(i + 1, n, step, near, far), zc));
    return __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), t));
}
// camera-space direction of pixel (col i, row j)   (base_renderer.py:247-251)
__device__ __forceinline__ void pixel_dir(int i, int j, float half_w, float half_h, float focal,
                                          float &dx, float &dy)
{
    dx = __fdiv_rn(__fsub_rn((float)i, half_w), focal);
    dy = -__fdiv_rn(__fsub_rn((float)j, half_h), focal);
}
// world direction component c: torch.sum over 3 products from a +0 accumulator (base_renderer.py:255)
__device__ __forceinline__ float rotate_dir(const Pose &p, int c, float dx, float dy)
{
    float s = __fadd_rn(0.0f, __fmul_rn(dx, p.r[c][0]));
    s = __fadd_rn(s, __fmul_rn(dy, p.r[c][1]));
    return __fadd_rn(s, __fmul_rn(-1.0f, p.r[c][2]));
}
// point = fl(o + fl(d*z))   (base_renderer.py:279)
__device__ __forceinline__ float point_on_ray(float o, float d, float z)
{
    return __fadd_rn(o, __fmul_rn(d, z));
}

constexpr float kPiF = 3.14159274101257324f;   // fl32(pi): `freq * torch.pi` with a 0-dim fp32 freq (nerf.py:42)

}  // namespace nerfb200
