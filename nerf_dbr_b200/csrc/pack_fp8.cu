// FP8-mode weight packing (fp8_layout.h): calibration of the activation scales with the fp32 network, per-row weight
// scales, quantisation to e4m3 and the scaled fp32 tables.  Setup-time work (reference pattern: the compression pass
// of CompressedNeRFRenderer.setup, src/benchmark/compressed_renderer.py:37-87).
#include "common.cuh"
#include "fp8_layout.h"
#include <cuda_fp8.h>

namespace nerfb200 {

int simt_calibrate(const void *packed, const float *positions, const float *directions, long long n, float *amax, cudaStream_t stream);

__device__ __forceinline__ float pow2_floor_scale(float target, float maxabs)
{
    if (!(maxabs > 0.f)) return 1.0f;
    return exp2f(floorf(log2f(target / maxabs)));
}
__device__ __forceinline__ unsigned char to_e4m3(float v)
{
    return (unsigned char)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3);
}

// weight row n of quantised layer l (1..7: trunk, hidden columns; 8: colour layer 0 rows 0..127 and the density row 128)
__device__ __forceinline__ const float *q_row(const nerf_b200_params &p, int l, int n, int &ld)
{
    if (l == 8) {
        if (n == 128) { ld = 256; return p.density_w; }
        ld = 283;
        return p.color0_w + (size_t)n * 283;
    }
    ld = l == 4 ? 319 : 256;
    return p.layer_w[l] + (size_t)n * ld;
}

// pass 1: activation scales from the calibration maxima, per-row weight scales (one warp per row)
__global__ void __launch_bounds__(256) fp8_scales_kernel(nerf_b200_params p, float *__restrict__ q)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x < 8) q[Q_SA + threadIdx.x] = pow2_floor_scale(240.0f, q[Q_SA + 8 + threadIdx.x]);
    const int rows = 7 * 256 + 129;
    if (warp >= rows) return;
    const int l = warp < 7 * 256 ? 1 + warp / 256 : 8, n = warp < 7 * 256 ? warp % 256 : warp - 7 * 256;
    int ld;
    const float *w = q_row(p, l, n, ld);
    float mx = 0.f;
    for (int k = lane; k < 256; k += 32) mx = fmaxf(mx, fabsf(__ldg(w + k)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) q[Q_SW + (size_t)l * 256 + n] = pow2_floor_scale(448.0f, mx);
}

// pass 2: tables and the operand stream
__global__ void __launch_bounds__(256) fp8_pack_kernel(nerf_b200_params p, unsigned char *__restrict__ out)
{
    float *q = reinterpret_cast<float *>(out);
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    auto sa = [&](int l) { return l < 0 ? 1.0f : q[Q_SA + l]; };
    auto sw = [&](int l, int n) { return q[Q_SW + (size_t)l * 256 + n]; };
    // ---- fp32 tables (the BF16 kernel's offsets, scaled contents)
    for (size_t i = tid; i < 8 * 256; i += nth) {
        const int l = (int)(i / 256), n = (int)(i % 256);
        q[F_BIAS + i] = p.layer_b[l][n] * sa(l);
        q[Q_MUL + i] = l == 0 ? sa(0) : sa(l) / (sa(l - 1) * sw(l, n));
    }
    for (size_t i = tid; i < 128; i += nth) {
        const float s = sa(7) * sw(8, (int)i);
        q[F_BC0 + i] = p.color0_b[i] * s;
        for (int c = 0; c < 3; ++c) q[F_WC1 + c * 128 + i] = p.color1_w[c * 128 + i] / s;
        for (int j = 0; j < 32; ++j) q[F_WC0D + (size_t)j * 128 + i] = j < kDirFeat ? p.color0_w[i * 283 + 256 + j] * s : 0.f;
    }
    if (tid == 0) {
        for (int c = 0; c < 3; ++c) q[F_BC1 + c] = p.color1_b[c];
        q[F_BC1 + 3] = 0.f;
        q[F_BSIG] = p.density_b[0];
        q[Q_SIGINV] = 1.0f / (sa(7) * sw(8, 128));
    }
    unsigned char *st = out + Q_OFFSET;
    // ---- bf16 stages: layer 0 (stage 0) and layer 4's skip part (stage 7), chunk = half, [128 n x 64 k], 16-byte units
    for (size_t u = tid; u < 2 * 2 * 128 * 8; u += nth) {
        const int which = (int)(u / 2048), half = (int)((u / 1024) & 1), n = (int)((u / 8) & 127), unit = (int)(u & 7);
        const int row = 128 * half + n;
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = unit * 8 + j;
            float w = 0.f;
            if (k < kPosFeat) w = which == 0 ? p.layer_w[0][row * 63 + k] : p.layer_w[4][(size_t)row * 319 + 256 + k] * sa(3) * sw(4, row);
            v[j] = __float2bfloat16_rn(w);
        }
        *reinterpret_cast<uint4 *>(st + q_stage_offset(which == 0 ? 0 : q_layer_stage(4)) + (size_t)half * kChunkBytes + swz128((uint32_t)n, (uint32_t)unit * 8)) =
            *reinterpret_cast<const uint4 *>(v);
    }
    // ---- e4m3 stages of trunk layers 1..7: stage = first + (layer 4: 1) + kp, chunk = half, 16-byte units of 16 k
    for (size_t u = tid; u < (size_t)7 * 2 * 2 * 128 * 8; u += nth) {
        const int l = 1 + (int)(u / 4096), kp = (int)((u / 2048) & 1), half = (int)((u / 1024) & 1), n = (int)((u / 8) & 127), unit = (int)(u & 7);
        const int row = 128 * half + n, ld = l == 4 ? 319 : 256;
        const float s = sw(l, row);
        __align__(16) unsigned char v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = to_e4m3(p.layer_w[l][(size_t)row * ld + 128 * kp + unit * 16 + j] * s);
        *reinterpret_cast<uint4 *>(st + q_stage_offset(q_layer_stage(l) + (l == 4 ? 1 : 0) + kp) + (size_t)half * kChunkBytes +
                                   swz128_u8((uint32_t)n, (uint32_t)unit * 16)) = *reinterpret_cast<const uint4 *>(v);
    }
    // ---- colour layer 0 (+ density row): stage 16, chunk = kp, [144 n x 128 k]
    for (size_t u = tid; u < (size_t)2 * kC0Rows * 8; u += nth) {
        const int kp = (int)(u / (kC0Rows * 8)), n = (int)((u / 8) % kC0Rows), unit = (int)(u & 7);
        __align__(16) unsigned char v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = 128 * kp + unit * 16 + j;
            float w = 0.f;
            if (n < 128) w = p.color0_w[(size_t)n * 283 + k] * sw(8, n);
            else if (n == 128) w = p.density_w[k] * sw(8, 128);
            v[j] = to_e4m3(w);
        }
        *reinterpret_cast<uint4 *>(st + q_stage_offset(16) + (size_t)kp * (kC0Rows * 128) + swz128_u8((uint32_t)n, (uint32_t)unit * 16)) =
            *reinterpret_cast<const uint4 *>(v);
    }
}

}  // namespace nerfb200

using namespace nerfb200;

extern "C" {

size_t nerf_b200_packed_fp8_bytes(void) { return PACKED_FP8_BYTES; }

int nerf_b200_pack_weights_fp8(const nerf_b200_params *params_host, const void *packed, const float *calib_positions,
                               const float *calib_directions, int64_t n_calib, void *packed_fp8, void *stream_)
{
    if (!params_host || !packed || !calib_positions || !calib_directions || n_calib <= 0 || !packed_fp8) return NERF_B200_EINVAL;
    if (((uintptr_t)packed | (uintptr_t)packed_fp8) & 1023) return NERF_B200_EALIGN;
    const nerf_b200_params &p = *params_host;
    for (int l = 0; l < 8; ++l)
        if (!p.layer_w[l] || !p.layer_b[l]) return NERF_B200_EINVAL;
    if (!p.density_w || !p.density_b || !p.color0_w || !p.color0_b || !p.color1_w || !p.color1_b) return NERF_B200_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    float *q = reinterpret_cast<float *>(packed_fp8);
    cudaError_t e = cudaMemsetAsync(q + Q_SA, 0, 16 * sizeof(float), stream);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    int rc = simt_calibrate(packed, calib_positions, calib_directions, n_calib, q + Q_SA + 8, stream);
    if (rc) return rc;
    fp8_scales_kernel<<<((7 * 256 + 129) * 32 + 255) / 256, 256, 0, stream>>>(p, q);
    if ((rc = launch_status())) return rc;
    fp8_pack_kernel<<<296, 256, 0, stream>>>(p, reinterpret_cast<unsigned char *>(packed_fp8));
    return launch_status();
}

}  // extern "C"
