// FP32 mode: the NeRF MLP on CUDA cores (FFMA), fused with ray generation / sampling /
// positional encoding in front and alpha compositing behind.  This is the <=1e-4 max-abs
// parity mode (single-pass TF32 fails that gate); the throughput mode is the
// tcgen05 kernel in mlp_tc.cu.
//
// One CTA (256 threads) processes tiles of 64 samples.  Activations stay in shared memory,
// K-major ([k][m]); each layer streams its K-major fp32 weights (packed_layout.h) through a
// [16][N] shared stage; every thread owns an 8(m) x 8|4(n) register tile.
//
// reference: NeRFModel.forward src/models/nerf.py:92-131; _render_ray_chunk
// src/benchmark/pytorch_renderers.py:156-170; execute_volume_rendering :105-125.
#include "common.cuh"
#include "simt_tile.cuh"

namespace nerfb200 {


enum { SRC_POINTS = 0, SRC_RAYS = 1, SRC_POSE = 2 };

struct SimtArgs {
    const float *wf;                      // fp32 region of the packed weights
    // SRC_POINTS
    const float *positions, *directions;
    long long n_points;
    float *sigma_out, *rgb_out;
    float *amax;                          // SRC_POINTS, optional: per-layer activation maxima (FP8 calibration)
    // SRC_RAYS / SRC_POSE
    Pose pose;
    int width, row0;
    float half_w, half_h, focal;
    const float *rays_o, *rays_d, *t_rand;
    const float *z_vals;                  // optional explicit depths [n_rays, n_samples]
    float *weights;                       // optional compositing weights out [n_rays, n_samples]
    int n_rays, n_samples, rays_per_item;
    float near, far;
    float *rgb_map, *depth, *acc;
};

// alpha compositing of one ray by one warp from shared (sigma,r,g,b); depths recomputed.
// reference pytorch_renderers.py:105-125
__device__ void simt_composite_ray(const float4 *__restrict__ smp, int n_samples, float near, float far,
                                   const float *__restrict__ t_rand_ray, const float *__restrict__ z_ray,
                                   float *__restrict__ w_ray, float dnorm, int lane,
                                   float &o_r, float &o_g, float &o_b, float &o_d, float &o_a)
{
    const float step = linspace_step(n_samples);
    double carry = 1.0;
    float cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f, ca = 0.f;
    for (int s0 = 0; s0 < n_samples; s0 += 32) {
        int s = s0 + lane;
        bool on = s < n_samples;
        float z = 0.f, zn = 0.f;
        if (on) {
            z = z_ray ? __ldg(z_ray + s)
                : t_rand_ray ? depth_jittered(s, n_samples, step, near, far, __ldg(t_rand_ray + s))
                             : depth_uniform(s, n_samples, step, near, far);
            if (s + 1 < n_samples)
                zn = z_ray ? __ldg(z_ray + s + 1)
                     : t_rand_ray ? depth_jittered(s + 1, n_samples, step, near, far, __ldg(t_rand_ray + s + 1))
                                  : depth_uniform(s + 1, n_samples, step, near, far);
        }
        float4 v = on ? smp[s] : make_float4(0.f, 0.f, 0.f, 0.f);
        float dist = __fmul_rn((s + 1 < n_samples) ? __fsub_rn(zn, z) : 1e10f, dnorm);
        float alpha = (on && n_samples > 1) ? __fsub_rn(1.0f, expf(__fmul_rn(-v.x, dist))) : 0.f;   // S == 1: see composite_kernel
        double keep = on ? (double)__fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double nb = __shfl_up_sync(0xffffffffu, keep, o);
            if (lane >= o) keep *= nb;
        }
        double excl = __shfl_up_sync(0xffffffffu, keep, 1);
        float trans = (float)(carry * (lane == 0 ? 1.0 : excl));
        carry *= __shfl_sync(0xffffffffu, keep, 31);
        float w = __fmul_rn(alpha, trans);
        if (w_ray && on) w_ray[s] = w;
        cr = fmaf(w, v.y, cr); cg = fmaf(w, v.z, cg); cb = fmaf(w, v.w, cb);
        cd = fmaf(w, z, cd); ca += w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cr += __shfl_xor_sync(0xffffffffu, cr, o); cg += __shfl_xor_sync(0xffffffffu, cg, o);
        cb += __shfl_xor_sync(0xffffffffu, cb, o); cd += __shfl_xor_sync(0xffffffffu, cd, o);
        ca += __shfl_xor_sync(0xffffffffu, ca, o);
    }
    o_r = cr; o_g = cg; o_b = cb; o_d = cd; o_a = ca;
}

template <int SRC>
__global__ void __launch_bounds__(kSimtThreads, 1) simt_mlp_kernel(SimtArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SimtSmem &sm = *reinterpret_cast<SimtSmem *>(smem_raw);
    const int tid = threadIdx.x;

    if (SRC == SRC_POINTS) {
        const long long n_tiles = (a.n_points + TM - 1) / TM;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            __syncthreads();
            if (tid < TM) {
                long long g = tile * TM + tid;
                bool on = g < a.n_points;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    sm.pos[c * TM + tid] = on ? __ldg(a.positions + 3 * g + c) : 0.f;
                    sm.dir[c * TM + tid] = on ? __ldg(a.directions + 3 * g + c) : 0.f;
                }
            }
            __syncthreads();
            simt_encode(sm, tid);
            float4 r;
            simt_network(sm, a.wf, tid, r, a.amax);
            if (tid < TM && a.sigma_out) {
                long long g = tile * TM + tid;
                if (g < a.n_points) {
                    a.sigma_out[g] = r.x;
                    a.rgb_out[3 * g + 0] = r.y; a.rgb_out[3 * g + 1] = r.z; a.rgb_out[3 * g + 2] = r.w;
                }
            }
        }
        return;
    }

    const int S = a.n_samples;
    const float step = linspace_step(S);
    const int n_items = (a.n_rays + a.rays_per_item - 1) / a.rays_per_item;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int r0 = item * a.rays_per_item;
        const int nr = min(a.rays_per_item, a.n_rays - r0);
        const int n_smp = nr * S;
        for (int t0 = 0; t0 < n_smp; t0 += TM) {
            __syncthreads();
            if (tid < TM) {
                int g = t0 + tid;
                float p[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
                if (g < n_smp) {
                    int ray = r0 + g / S, s = g - (g / S) * S;
                    float o[3];
                    if (SRC == SRC_POSE) {
                        int j = a.row0 + ray / a.width, i = ray % a.width;
                        float dx, dy;
                        pixel_dir(i, j, a.half_w, a.half_h, a.focal, dx, dy);
#pragma unroll
                        for (int c = 0; c < 3; ++c) { d[c] = rotate_dir(a.pose, c, dx, dy); o[c] = a.pose.t[c]; }
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) { d[c] = __ldg(a.rays_d + 3 * (size_t)ray + c); o[c] = __ldg(a.rays_o + 3 * (size_t)ray + c); }
                    }
                    float z = a.z_vals ? __ldg(a.z_vals + (size_t)ray * S + s)
                              : a.t_rand ? depth_jittered(s, S, step, a.near, a.far, __ldg(a.t_rand + (size_t)ray * S + s))
                                         : depth_uniform(s, S, step, a.near, a.far);
#pragma unroll
                    for (int c = 0; c < 3; ++c) p[c] = point_on_ray(o[c], d[c], z);
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) { sm.pos[c * TM + tid] = p[c]; sm.dir[c * TM + tid] = d[c]; }
            }
            __syncthreads();
            simt_encode(sm, tid);
            float4 r;
            simt_network(sm, a.wf, tid, r);
            if (tid < TM && t0 + tid < n_smp) sm.out[t0 + tid] = r;
        }
        __syncthreads();
        const int warp = tid >> 5, lane = tid & 31;
        for (int q = warp; q < nr; q += kSimtThreads / 32) {
            int ray = r0 + q;
            float d[3];
            if (SRC == SRC_POSE) {
                int j = a.row0 + ray / a.width, i = ray % a.width;
                float dx, dy;
                pixel_dir(i, j, a.half_w, a.half_h, a.focal, dx, dy);
#pragma unroll
                for (int c = 0; c < 3; ++c) d[c] = rotate_dir(a.pose, c, dx, dy);
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) d[c] = __ldg(a.rays_d + 3 * (size_t)ray + c);
            }
            float dn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
            float cr, cg, cb, cd, ca;
            simt_composite_ray(sm.out + (size_t)q * S, S, a.near, a.far,
                               a.t_rand ? a.t_rand + (size_t)ray * S : nullptr,
                               a.z_vals ? a.z_vals + (size_t)ray * S : nullptr,
                               a.weights ? a.weights + (size_t)ray * S : nullptr, dn, lane, cr, cg, cb, cd, ca);
            if (lane == 0) {
                a.rgb_map[3 * (size_t)ray + 0] = cr; a.rgb_map[3 * (size_t)ray + 1] = cg; a.rgb_map[3 * (size_t)ray + 2] = cb;
                a.depth[ray] = cd;
                if (a.acc) a.acc[ray] = ca;
            }
        }
    }
}

static int simt_grid(long long work)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)(work < sms ? (work > 0 ? work : 1) : sms);   // persistent: one CTA per SM
}

template <int SRC>
static int simt_launch(const SimtArgs &a, long long work, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(simt_mlp_kernel<SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(SimtSmem));
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    simt_mlp_kernel<SRC><<<simt_grid(work), kSimtThreads, sizeof(SimtSmem), stream>>>(a);
    return launch_status();
}

int simt_query(const void *packed, const float *positions, const float *directions, long long n,
               float *sigma, float *rgb, cudaStream_t stream)
{
    SimtArgs a = {};
    a.wf = reinterpret_cast<const float *>(packed);
    a.positions = positions; a.directions = directions; a.n_points = n;
    a.sigma_out = sigma; a.rgb_out = rgb;
    return simt_launch<SRC_POINTS>(a, (n + TM - 1) / TM, stream);
}

// FP8-mode calibration: run the fp32 network over `n` (position, direction) rows and keep the per-layer activation
// maxima in amax[8] (device, zeroed by the caller).  Rows past n are zero-padded points: they only add small values.
int simt_calibrate(const void *packed, const float *positions, const float *directions, long long n, float *amax, cudaStream_t stream)
{
    SimtArgs a = {};
    a.wf = reinterpret_cast<const float *>(packed);
    a.positions = positions; a.directions = directions; a.n_points = n;
    a.amax = amax;
    return simt_launch<SRC_POINTS>(a, (n + TM - 1) / TM, stream);
}

static int rays_per_item_for(int n_samples)
{
    int rpi = 1024 / n_samples;
    return rpi < 1 ? 1 : rpi;
}

int simt_render_pose(const void *packed, const float *c2w, int width, int height, float focal,
                     float near, float far, int n_samples, int row0, int n_rows, float *rgb_out,
                     float *depth_out, cudaStream_t stream)
{
    if (n_samples > kItemMax) return NERF_B200_EUNSUPPORTED;
    SimtArgs a = {};
    a.wf = reinterpret_cast<const float *>(packed);
    a.pose = pose_from_c2w(c2w);
    a.width = width; a.row0 = row0;
    a.half_w = (float)((double)width * 0.5); a.half_h = (float)((double)height * 0.5); a.focal = focal;
    a.n_rays = n_rows * width; a.n_samples = n_samples; a.rays_per_item = rays_per_item_for(n_samples);
    a.near = near; a.far = far;
    a.rgb_map = rgb_out; a.depth = depth_out; a.acc = nullptr;
    return simt_launch<SRC_POSE>(a, (a.n_rays + a.rays_per_item - 1) / a.rays_per_item, stream);
}

int simt_render_rays(const void *packed, const float *rays_o, const float *rays_d, int n_rays,
                     int n_samples, float near, float far, const float *t_rand, const float *z_vals, float *rgb_out,
                     float *depth_out, float *acc_out, float *weights_out, cudaStream_t stream)
{
    if (n_samples > kItemMax) return NERF_B200_EUNSUPPORTED;
    SimtArgs a = {};
    a.wf = reinterpret_cast<const float *>(packed);
    a.rays_o = rays_o; a.rays_d = rays_d; a.t_rand = t_rand; a.z_vals = z_vals; a.weights = weights_out;
    a.n_rays = n_rays; a.n_samples = n_samples; a.rays_per_item = rays_per_item_for(n_samples);
    a.near = near; a.far = far;
    a.rgb_map = rgb_out; a.depth = depth_out; a.acc = acc_out;
    return simt_launch<SRC_RAYS>(a, (n_rays + a.rays_per_item - 1) / a.rays_per_item, stream);
}

}  // namespace nerfb200
