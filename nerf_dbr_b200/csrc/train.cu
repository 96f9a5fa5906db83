// Training step: forward + backward of ONE network's photometric-MSE term (host entry point for both modes; the
// FP32-mode kernels live here, the BF16-mode tensor-core kernels in mlp_tc.cu / train_dgrad_tc.cu / train_tc.cu)
//   loss_term = mean_{R x 3} (C - target)^2,   C = composite(MLP(points on rays))
// reference: NeRFTrainer.train_step / _render_rays / _query_network (src/training/trainer.py:83-138,
// 294-351), VolumeRenderer.volume_render (src/utils/rendering.py:102-143); backward math: SURVEY App. B.
//
// Rays are processed in chunks (bounded caller-owned workspace).  Per chunk, FP32 mode:
//   1. train_fwd_kernel   64-sample tiles: encode, 8x256 trunk, heads; every layer's activations are
//                         stored K-major ([feature][sample]) for the backward
//   2. train_ray_kernel   one warp per ray: compositing forward, loss, dL/dC, compositing backward
//                         (double precision) -> dL/dsigma_pre, dL/d(colour pre-sigmoid) per sample
//   3. train_bwd_kernel   64-sample tiles: heads and the dgrad chain layer 7..1 (dX = dY . W with the
//                         nn.Linear-layout weight copies), masked by the stored activations (ReLU')
//   4. wgrad_kernel       per parameter tensor: dW[n][k] += sum_s dY[n][s] X[k][s] (+ bias row sums),
//                         64x64 output tiles, split over samples, fp32 atomics into the caller's grads
// Gradients are ACCUMULATED (+=), pre-scaled by 2 / (3 R_global) so data-parallel ranks all-reduce-sum.
//
// FP32 mode is the gradient parity reference on the device (relative error <= 5e-4 against the reference's
// autograd).  BF16 mode runs the forward (the fused tcgen05 kernel's TRAIN variant, which stores the
// activations it multiplies), the dgrad chain (train_dgrad_tc.cu) and the weight-gradient GEMMs (train_tc.cu)
// on the tensor cores; only the per-ray compositing backward and the two tiny head gradients stay on CUDA cores.
#include <algorithm>
#include "common.cuh"
#include "simt_tile.cuh"
#include "train_layout.h"

namespace nerfb200 {

constexpr int kChunkSamples = 524288;        // samples per chunk (multiple of 64): 9.5 GB of workspace at most

int tc_train_forward(const void *packed, const float *rays_o, const float *rays_d, int n_rays, int n_samples, float near,
                     float far, const float *t_rand, float *ws, int ws_ch, unsigned int *dbg, int sm_limit, cudaStream_t stream);
int dgrad_chain_tc(const void *packed, float *ws, int ch, int n_samples, unsigned int *dbg, int sm_limit, cudaStream_t stream);
size_t wgrad_skinny_scratch_bytes();
int wgrad_skinny(const float *A, int rows_a, int ch, const __nv_bfloat16 *ws, int row_b, int rows_b, float *dW, int ld, float *dbias, float *scratch, int sm_limit,
                 cudaStream_t stream);
size_t wgrad_tc_scratch_bytes(int splits);
int wgrad_tc_batch(const __nv_bfloat16 *ws, int ch, const WgradJob *jobs, int n_jobs, float *scratch, int ctas, cudaStream_t stream);

struct TrainArgs {
    const float *wf;                         // fp32 region of the packed weights
    float *ws;                               // workspace [R_TOTAL][ch]
    int ch;                                  // padded samples in this chunk (multiple of 64)
    const float *rays_o, *rays_d, *t_rand, *target;
    int ray0, n_rays, n_samples;             // chunk = rays [ray0, ray0 + n_rays)
    float near, far;
    float grad_scale;                        // 2 / (3 R_global)
    float *loss_sum, *rgb_out;
};

__device__ __forceinline__ void copy_tile_to_global(const float *__restrict__ smem, int rows, float *__restrict__ g,
                                                    int ch, int col0, int tid)
{
    // smem [rows][TM] -> g[row * ch + col0 + m], float4 along m
    for (int i = tid; i < rows * (TM / 4); i += kSimtThreads) {
        int r = i / (TM / 4), c4 = i % (TM / 4);
        reinterpret_cast<float4 *>(g + (size_t)r * ch + col0)[c4] = reinterpret_cast<const float4 *>(smem + r * TM)[c4];
    }
}

// ------------------------------------------------------------------------------------------ 1. forward
__global__ void __launch_bounds__(kSimtThreads, 1) train_fwd_kernel(TrainArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SimtSmem &sm = *reinterpret_cast<SimtSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const int S = a.n_samples, n_smp = a.n_rays * S;
    const float step = linspace_step(S);
    const float *wf = a.wf;
    const int m0 = (tid & 7) * 8, ng = tid >> 3, n0 = ng * 8;
    for (int tile = blockIdx.x; tile * TM < a.ch; tile += gridDim.x) {
        const int col0 = tile * TM;
        __syncthreads();
        if (tid < TM) {
            int g = col0 + tid;
            float p[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
            if (g < n_smp) {
                int ray = a.ray0 + g / S, s = g % S;
                float o[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) { d[c] = __ldg(a.rays_d + 3 * (size_t)ray + c); o[c] = __ldg(a.rays_o + 3 * (size_t)ray + c); }
                float z = a.t_rand ? depth_jittered(s, S, step, a.near, a.far, __ldg(a.t_rand + (size_t)ray * S + s))
                                   : depth_uniform(s, S, step, a.near, a.far);
#pragma unroll
                for (int c = 0; c < 3; ++c) p[c] = point_on_ray(o[c], d[c], z);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) { sm.pos[c * TM + tid] = p[c]; sm.dir[c * TM + tid] = d[c]; }
        }
        __syncthreads();
        simt_encode(sm, tid);
        __syncthreads();
        copy_tile_to_global(sm.pe, 64, a.ws + (size_t)R_PE * a.ch, a.ch, col0, tid);
        copy_tile_to_global(sm.de, 32, a.ws + (size_t)R_DE * a.ch, a.ch, col0, tid);
        float acc[8][8];
        zero_acc(acc);
        simt_accumulate<8>(acc, sm.pe, 64, wf + F_W0T, sm.ws, tid, m0, n0);
        simt_store_relu<8>(acc, wf + F_BIAS, sm.actA, m0, n0);
        __syncthreads();
        copy_tile_to_global(sm.actA, 256, a.ws + (size_t)R_H * a.ch, a.ch, col0, tid);
        float *in = sm.actA, *out = sm.actB;
        for (int l = 1; l < 8; ++l) {
            zero_acc(acc);
            simt_accumulate<8>(acc, in, 256, wf + f_wt(l), sm.ws, tid, m0, n0);
            if (l == 4) simt_accumulate<8>(acc, sm.pe, 64, wf + F_W4P, sm.ws, tid, m0, n0);
            simt_store_relu<8>(acc, wf + F_BIAS + l * 256, out, m0, n0);
            __syncthreads();
            copy_tile_to_global(out, 256, a.ws + (size_t)(R_H + 256 * l) * a.ch, a.ch, col0, tid);
            float *t = in; in = out; out = t;
        }
        {
            float c[8][4];
            const int n4 = ng * 4;
            zero_acc(c);
            simt_accumulate<4>(c, in, 256, wf + F_WC0H, sm.ws, tid, m0, n4);
            simt_accumulate<4>(c, sm.de, 32, wf + F_WC0D, sm.ws, tid, m0, n4);
            simt_store_relu<4>(c, wf + F_BC0, out, m0, n4);
        }
        __syncthreads();
        copy_tile_to_global(out, 128, a.ws + (size_t)R_C0H * a.ch, a.ch, col0, tid);
        if (tid < TM) {
            float s = 0.f;
            for (int k = 0; k < 256; ++k) s = fmaf(in[k * TM + tid], __ldg(wf + F_WSIG + k), s);
            a.ws[(size_t)R_SIGPRE * a.ch + col0 + tid] = s + __ldg(wf + F_BSIG);
#pragma unroll
            for (int chn = 0; chn < 3; ++chn) {
                float t = 0.f;
                for (int k = 0; k < 128; ++k) t = fmaf(out[k * TM + tid], __ldg(wf + F_WC1 + chn * 128 + k), t);
                t += __ldg(wf + F_BC1 + chn);
                a.ws[(size_t)(R_RGB + chn) * a.ch + col0 + tid] = 1.0f / (1.0f + expf(-t));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ 2. rays
// one warp per ray; the transmittance product, the loss gradient and the reverse scan run in double
__global__ void train_ray_kernel(TrainArgs a)
{
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = a.n_samples;
    const float step = linspace_step(S);
    const float *sigpre = a.ws + (size_t)R_SIGPRE * a.ch;
    const float *rgb = a.ws + (size_t)R_RGB * a.ch;
    float *dsig = a.ws + (size_t)R_DSIG * a.ch, *dy = a.ws + (size_t)R_DY * a.ch;
    for (int q = blockIdx.x * warps + warp; q < a.n_rays; q += gridDim.x * warps) {
        const int ray = a.ray0 + q;
        const size_t base = (size_t)q * S;
        const float dx = __ldg(a.rays_d + 3 * (size_t)ray), dyy = __ldg(a.rays_d + 3 * (size_t)ray + 1), dz = __ldg(a.rays_d + 3 * (size_t)ray + 2);
        const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dyy, dyy)), __fmul_rn(dz, dz)));
        const float *tr = a.t_rand ? a.t_rand + (size_t)ray * S : nullptr;
        auto zof = [&](int s) { return tr ? depth_jittered(s, S, step, a.near, a.far, __ldg(tr + s)) : depth_uniform(s, S, step, a.near, a.far); };
        // pass 1: C = sum w c
        double carry = 1.0, C[3] = {0.0, 0.0, 0.0};
        for (int s0 = 0; s0 < S; s0 += 32) {
            int s = s0 + lane;
            bool on = s < S && S > 1;
            float z = on ? zof(s) : 0.f;
            float dist = __fmul_rn((s + 1 < S) ? __fsub_rn(zof(min(s + 1, S - 1)), z) : 1e10f, nrm);
            float sg = on ? fmaxf(sigpre[base + s], 0.f) : 0.f;
            float alpha = on ? __fsub_rn(1.0f, expf(__fmul_rn(-sg, dist))) : 0.f;
            double keep = on ? (double)__fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0;
            double incl = keep;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { double nb = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl *= nb; }
            double excl = __shfl_up_sync(0xffffffffu, incl, 1);
            double T = carry * (lane == 0 ? 1.0 : excl);
            carry *= __shfl_sync(0xffffffffu, incl, 31);
            double w = (double)alpha * T;
            if (on) {
                C[0] += w * rgb[base + s]; C[1] += w * rgb[(size_t)a.ch + base + s]; C[2] += w * rgb[2 * (size_t)a.ch + base + s];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) C[c] += __shfl_xor_sync(0xffffffffu, C[c], o);
        double g[3], err2 = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double e = (double)(float)C[c] - (double)__ldg(a.target + 3 * (size_t)ray + c);
            err2 += e * e;
            g[c] = e * (double)a.grad_scale;
        }
        if (lane == 0) {
            atomicAdd(a.loss_sum, (float)err2);
            if (a.rgb_out) { a.rgb_out[3 * (size_t)ray] = (float)C[0]; a.rgb_out[3 * (size_t)ray + 1] = (float)C[1]; a.rgb_out[3 * (size_t)ray + 2] = (float)C[2]; }
        }
        // total of G_k w_k, then pass 2 with a running prefix of it
        // dL/d alpha_i = G_i T_i - (1/p_i) sum_{k>i} G_k w_k ;  dL/d sigma_pre = dL/d alpha * (1-alpha) * dist * [sigma_pre > 0]
        double total_gw = g[0] * C[0] + g[1] * C[1] + g[2] * C[2];
        carry = 1.0;
        double prefix = 0.0;
        for (int s0 = 0; s0 < S; s0 += 32) {
            int s = s0 + lane;
            bool on = s < S && S > 1;
            float z = on ? zof(s) : 0.f;
            float dist = __fmul_rn((s + 1 < S) ? __fsub_rn(zof(min(s + 1, S - 1)), z) : 1e10f, nrm);
            float sp = on ? sigpre[base + s] : 0.f;
            float sg = fmaxf(sp, 0.f);
            float ex = on ? expf(__fmul_rn(-sg, dist)) : 1.f;
            float alpha = on ? __fsub_rn(1.0f, ex) : 0.f;
            double keep = on ? (double)__fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0;
            double incl = keep;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { double nb = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl *= nb; }
            double excl = __shfl_up_sync(0xffffffffu, incl, 1);
            double T = carry * (lane == 0 ? 1.0 : excl);
            carry *= __shfl_sync(0xffffffffu, incl, 31);
            double w = (double)alpha * T;
            float c0 = on ? rgb[base + s] : 0.f, c1 = on ? rgb[(size_t)a.ch + base + s] : 0.f, c2 = on ? rgb[2 * (size_t)a.ch + base + s] : 0.f;
            double G = g[0] * c0 + g[1] * c1 + g[2] * c2;
            double gw = G * w, run = gw;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { double nb = __shfl_up_sync(0xffffffffu, run, o); if (lane >= o) run += nb; }
            double suffix = total_gw - (prefix + run);
            prefix += __shfl_sync(0xffffffffu, run, 31);
            double dalpha = G * T - suffix / keep;
            if (on) {
                dsig[base + s] = sp > 0.f ? (float)(dalpha * (double)ex * (double)dist) : 0.f;
                dy[base + s] = (float)(w * g[0] * c0 * (1.0 - c0));
                dy[(size_t)a.ch + base + s] = (float)(w * g[1] * c1 * (1.0 - c1));
                dy[2 * (size_t)a.ch + base + s] = (float)(w * g[2] * c2 * (1.0 - c2));
            } else if (s < S) {
                dsig[base + s] = 0.f; dy[base + s] = 0.f; dy[(size_t)a.ch + base + s] = 0.f; dy[2 * (size_t)a.ch + base + s] = 0.f;
            }
        }
    }
    // padding samples of the chunk carry no gradient
    const int n_smp = a.n_rays * S;
    for (int i = n_smp + blockIdx.x * blockDim.x + threadIdx.x; i < a.ch; i += gridDim.x * blockDim.x) {
        dsig[i] = 0.f; dy[i] = 0.f; dy[(size_t)a.ch + i] = 0.f; dy[2 * (size_t)a.ch + i] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------ 3. dgrad chain
// acc (dL/d h of the layer below) masked by that layer's stored activation (ReLU') -> dpre, to shared
// (next GEMM's A operand, K-major) and to the workspace (wgrad operand)
__device__ __forceinline__ void store_masked(const float (&acc)[8][8], const float *__restrict__ hmask, float *out_s,
                                             float *__restrict__ out_g, int ch, int col0, int m0, int n0)
{
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 *hp = reinterpret_cast<const float4 *>(hmask + (size_t)(n0 + j) * ch + col0 + m0);
        float4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
        float4 lo = make_float4(h0.x > 0.f ? acc[0][j] : 0.f, h0.y > 0.f ? acc[1][j] : 0.f, h0.z > 0.f ? acc[2][j] : 0.f, h0.w > 0.f ? acc[3][j] : 0.f);
        float4 hi = make_float4(h1.x > 0.f ? acc[4][j] : 0.f, h1.y > 0.f ? acc[5][j] : 0.f, h1.z > 0.f ? acc[6][j] : 0.f, h1.w > 0.f ? acc[7][j] : 0.f);
        float4 *op = reinterpret_cast<float4 *>(out_s + (size_t)(n0 + j) * TM + m0);
        op[0] = lo; op[1] = hi;
        float4 *gp = reinterpret_cast<float4 *>(out_g + (size_t)(n0 + j) * ch + col0 + m0);
        gp[0] = lo; gp[1] = hi;
    }
}

__global__ void __launch_bounds__(kSimtThreads, 1) train_bwd_kernel(TrainArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SimtSmem &sm = *reinterpret_cast<SimtSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const float *wf = a.wf;
    const int m0 = (tid & 7) * 8, ng = tid >> 3, n0 = ng * 8;
    float *ws = a.ws;
    const int ch = a.ch;
    for (int tile = blockIdx.x; tile * TM < ch; tile += gridDim.x) {
        const int col0 = tile * TM;
        __syncthreads();
        // dpre_c0[k][m] = (sum_c dy[c][m] W_c1[c][k]) * [c0h > 0]   -> sm.actA rows 0..127, and workspace
        for (int i = tid; i < 128 * TM; i += kSimtThreads) {
            int k = i / TM, m = i % TM;
            float v = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) v = fmaf(ws[(size_t)(R_DY + c) * ch + col0 + m], __ldg(wf + F_WC1 + c * 128 + k), v);
            if (!(ws[(size_t)(R_C0H + k) * ch + col0 + m] > 0.f)) v = 0.f;
            sm.actA[k * TM + m] = v;
            ws[(size_t)(R_DPREC0 + k) * ch + col0 + m] = v;
        }
        // dh7 = dpre_c0 . W_c0[:, :256] + dsig (x) w_sigma
        float acc[8][8];
        zero_acc(acc);
        simt_accumulate<8>(acc, sm.actA, 128, wf + F_WC0O, sm.ws, tid, m0, n0);
        {
            const float4 *dp = reinterpret_cast<const float4 *>(ws + (size_t)R_DSIG * ch + col0 + m0);
            float4 d0 = dp[0], d1 = dp[1];
            float ds[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float w = __ldg(wf + F_WSIG + n0 + j);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i][j] = fmaf(ds[i], w, acc[i][j]);
            }
        }
        float *cur = sm.actB, *nxt = sm.actA;
        __syncthreads();   // all reads of sm.actA (dpre_c0) done before it is reused below
        store_masked(acc, ws + (size_t)(R_H + 7 * 256) * ch, cur, ws + (size_t)(R_DPRE + 7 * 256) * ch, ch, col0, m0, n0);
        for (int l = 7; l >= 1; --l) {
            // dh_{l-1} = dpre_l . W_l[:, :256]   (K = 256 outputs of layer l, N = 256 inputs)
            zero_acc(acc);
            simt_accumulate<8>(acc, cur, 256, wf + f_wo(l), sm.ws, tid, m0, n0);
            store_masked(acc, ws + (size_t)(R_H + (l - 1) * 256) * ch, nxt, ws + (size_t)(R_DPRE + (l - 1) * 256) * ch, ch, col0, m0, n0);
            float *t = cur; cur = nxt; nxt = t;
        }
    }
}

// ------------------------------------------------------------------------------------------ 4. wgrad
// dW[n][col_off + k] += sum_s A[n][s] B[k][s]  (n < rows_a, k < rows_b), bias[n] += sum_s A[n][s].
// 64 x 64 output tile per block, 4 x 4 per thread, samples split over blockIdx.z.
__global__ void __launch_bounds__(256) wgrad_kernel(const float *__restrict__ A, int rows_a, const float *__restrict__ B,
                                                    int rows_b, int ch, float *__restrict__ dW, int ld, int col_off,
                                                    float *__restrict__ dbias)
{
    __shared__ __align__(16) float As[32][68], Bs[32][68];
    const int tid = threadIdx.x, tn = tid >> 4, tk = tid & 15;
    const int nb = blockIdx.x * 64, kb = blockIdx.y * 64;
    const int per = ((ch + gridDim.z - 1) / gridDim.z + 31) & ~31;
    const int s_begin = blockIdx.z * per, s_end = min(ch, s_begin + per);
    float acc[4][4] = {};
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s0 = s_begin; s0 < s_end; s0 += 32) {
        __syncthreads();
        for (int i = tid; i < 64 * 8; i += 256) {          // 64 rows x 32 samples = 8 float4 per row
            int r = i >> 3, c4 = i & 7;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (nb + r < rows_a) va = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(nb + r) * ch + s0) + c4);
            if (kb + r < rows_b) vb = __ldg(reinterpret_cast<const float4 *>(B + (size_t)(kb + r) * ch + s0) + c4);
            As[c4 * 4 + 0][r] = va.x; As[c4 * 4 + 1][r] = va.y; As[c4 * 4 + 2][r] = va.z; As[c4 * 4 + 3][r] = va.w;
            Bs[c4 * 4 + 0][r] = vb.x; Bs[c4 * 4 + 1][r] = vb.y; Bs[c4 * 4 + 2][r] = vb.z; Bs[c4 * 4 + 3][r] = vb.w;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 32; ++s) {
            float4 av = *reinterpret_cast<const float4 *>(&As[s][tn * 4]);
            float4 bv = *reinterpret_cast<const float4 *>(&Bs[s][tk * 4]);
            float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                rs[i] += aa[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int n = nb + tn * 4 + i;
        if (n >= rows_a) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = kb + tk * 4 + j;
            if (k < rows_b) atomicAdd(dW + (size_t)n * ld + col_off + k, acc[i][j]);
        }
        if (dbias && blockIdx.y == 0 && tk == 0) atomicAdd(dbias + n, rs[i]);
    }
}

static int sm_count()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

// BF16 mode runs the per-tensor wgrad launches (independent of each other) alternately on two auxiliary streams:
// every launch has ~19 us of fixed cost (ramp, pipeline fill, partial write-out, tail, then its reduce) during
// which HBM idles; with two streams the next tensor's CTAs take over SMs as the previous one's drain.  Forked from
// and joined back into the caller's stream with events, one set per device.
struct WgradStreams {
    cudaStream_t caller = nullptr;               // the stream this set forks from (a set is never shared between caller streams)
    cudaStream_t s[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
    bool ok = false;
};
// One set per (device, caller stream), up to kSetsPerDevice caller streams per device: two training calls that run
// concurrently on different streams (the coarse and the fine network of a small batch) must not share fork / join events.
constexpr int kSetsPerDevice = 8;
static WgradStreams *wgrad_streams(cudaStream_t caller)
{
    static WgradStreams per_dev[64][kSetsPerDevice];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    WgradStreams *w = nullptr;
    for (int i = 0; i < kSetsPerDevice && !w; ++i)
        if (per_dev[dev][i].ok && per_dev[dev][i].caller == caller) w = &per_dev[dev][i];
    for (int i = 0; i < kSetsPerDevice && !w; ++i)
        if (!per_dev[dev][i].ok) w = &per_dev[dev][i];
    if (!w) return nullptr;                      // more caller streams than sets
    if (!w->ok) {
        bool good = cudaEventCreateWithFlags(&w->fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < 2 && good; ++i)
            good = cudaStreamCreateWithFlags(&w->s[i], cudaStreamNonBlocking) == cudaSuccess &&
                   cudaEventCreateWithFlags(&w->join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!good) { cudaGetLastError(); return nullptr; }
        w->caller = caller;
        w->ok = true;
    }
    return w;
}

static int chunk_rays(int n_samples)
{
    int r = kChunkSamples / n_samples;
    return r < 1 ? 1 : r;
}

}  // namespace nerfb200

using namespace nerfb200;

extern "C" {

size_t nerf_b200_train_workspace_bytes(int n_rays, int n_samples)
{
    if (n_rays <= 0 || n_samples <= 0) return 0;
    long long per = (long long)std::min(n_rays, chunk_rays(n_samples)) * n_samples;
    long long ch = (per + 63) / 64 * 64;
    return (size_t)ch * R_TOTAL * sizeof(float) + wgrad_tc_scratch_bytes(256) + 2 * wgrad_skinny_scratch_bytes();   // partials of one batched wgrad launch + the two skinny jobs
}

int nerf_b200_train_fwd_bwd(const void *packed, const nerf_b200_params *params, const nerf_b200_params *grads,
                            const float *rays_o, const float *rays_d, const float *target, int n_rays,
                            int n_samples, float near, float far, const float *t_rand, int n_rays_global,
                            int mode, void *workspace, float *loss_sum, float *rgb_out, void *stream_)
{
    return nerf_b200_train_fwd_bwd_ex(packed, params, grads, rays_o, rays_d, target, n_rays, n_samples, near, far, t_rand,
                                      n_rays_global, mode, workspace, loss_sum, rgb_out, NERF_B200_TRAIN_ALL, 0, stream_);
}

int nerf_b200_train_fwd_bwd_ex(const void *packed, const nerf_b200_params *params, const nerf_b200_params *grads,
                               const float *rays_o, const float *rays_d, const float *target, int n_rays,
                               int n_samples, float near, float far, const float *t_rand, int n_rays_global,
                               int mode, void *workspace, float *loss_sum, float *rgb_out, int phases, int sm_limit,
                               void *stream_)
{
    (void)params;
    if (phases != NERF_B200_TRAIN_ALL && phases != NERF_B200_TRAIN_ACTIVATIONS && phases != NERF_B200_TRAIN_WEIGHT_GRADS)
        return NERF_B200_EINVAL;
    if (sm_limit < 0) return NERF_B200_EINVAL;
    // split phases keep the activations in the workspace between two calls: BF16 mode, one chunk
    if (phases != NERF_B200_TRAIN_ALL && (mode != NERF_B200_BF16 || n_rays > chunk_rays(n_samples > 0 ? n_samples : 1)))
        return NERF_B200_EUNSUPPORTED;
    if (!packed || !grads || !rays_o || !rays_d || !target || !workspace || !loss_sum || n_rays <= 0 ||
        n_samples <= 0 || n_rays_global <= 0)
        return NERF_B200_EINVAL;
    if (mode != NERF_B200_FP32 && mode != NERF_B200_BF16) return NERF_B200_EUNSUPPORTED;
    const bool tc = mode == NERF_B200_BF16;      // BF16: weight gradients on the tensor cores (bf16 operands, fp32 accumulate)
    if (n_samples > kChunkSamples) return NERF_B200_EUNSUPPORTED;
    if (((uintptr_t)packed & 1023) || ((uintptr_t)workspace & 15)) return NERF_B200_EALIGN;
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaError_t e = cudaFuncSetAttribute(train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SimtSmem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SimtSmem));
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    const int sms = sm_count();
    const int per_chunk = chunk_rays(n_samples);
    const nerf_b200_params &g = *grads;
    for (int r0 = 0; r0 < n_rays; r0 += per_chunk) {
        TrainArgs a = {};
        a.wf = reinterpret_cast<const float *>(packed);
        a.ws = reinterpret_cast<float *>(workspace);
        a.ray0 = r0;
        a.n_rays = std::min(per_chunk, n_rays - r0);
        a.n_samples = n_samples;
        a.ch = (a.n_rays * n_samples + 63) / 64 * 64;
        a.rays_o = rays_o; a.rays_d = rays_d; a.t_rand = t_rand; a.target = target;
        a.near = near; a.far = far;
        a.grad_scale = 2.0f / (3.0f * (float)n_rays_global);
        a.loss_sum = loss_sum; a.rgb_out = rgb_out;
        const int tiles = a.ch / TM;
        int rc;
        if (phases & NERF_B200_TRAIN_ACTIVATIONS) {
        if (tc) {
            // forward on the tensor cores (TRAIN variant of the fused kernel); pad columns of the stored
            // activations must be finite zeros: wgrad multiplies them by zero gradients
            const int n_smp = a.n_rays * n_samples;
            if (a.ch != n_smp) {                        // the last slab of the bf16 operand rows (activations and dpre)
                cudaError_t me = cudaMemsetAsync(reinterpret_cast<__nv_bfloat16 *>(a.ws) + big_tile(0, a.ch / 64 - 1), 0, (size_t)G_TOTAL * 128, stream);
                if (me != cudaSuccess) { cudaGetLastError(); return (int)me; }
            }
            const float *tr = t_rand ? t_rand + (size_t)r0 * n_samples : nullptr;
            if ((rc = tc_train_forward(packed, rays_o + 3 * (size_t)r0, rays_d + 3 * (size_t)r0, a.n_rays, n_samples, near, far,
                                       tr, a.ws, a.ch, watchdog_word(), sm_limit, stream)))
                return rc;
        } else {
            train_fwd_kernel<<<std::min(tiles, sms), kSimtThreads, sizeof(SimtSmem), stream>>>(a);
            if ((rc = launch_status())) return rc;
        }
        train_ray_kernel<<<std::min((a.n_rays + 7) / 8, sms * 8), 256, 0, stream>>>(a);
        if ((rc = launch_status())) return rc;
        if (tc) {
            // dgrad chain on the tensor cores; the pad columns of the dpre rows are zero since the memset above
            const int n_smp = a.n_rays * n_samples;
            if (a.ch != n_smp) {
                cudaError_t me = cudaMemset2DAsync(a.ws + (size_t)R_DSIG * a.ch + n_smp, (size_t)a.ch * sizeof(float), 0,
                                                   (size_t)(a.ch - n_smp) * sizeof(float), 4, stream);       // dsig, dy (fp32 rows)
                if (me != cudaSuccess) { cudaGetLastError(); return (int)me; }
            }
            if ((rc = dgrad_chain_tc(packed, a.ws, a.ch, n_smp, watchdog_word(), sm_limit, stream))) return rc;
        } else {
            train_bwd_kernel<<<std::min(tiles, sms), kSimtThreads, sizeof(SimtSmem), stream>>>(a);
            if ((rc = launch_status())) return rc;
        }
        }                                               // phase: activations
        if (!(phases & NERF_B200_TRAIN_WEIGHT_GRADS)) continue;
        float *ws = a.ws;
        const size_t ch = a.ch;
        auto row = [&](int r) { return ws + (size_t)r * ch; };
        const int split = std::max(1, std::min(64, (int)(ch / 2048)));
        float *scratch = ws + (size_t)R_TOTAL * ch;
        WgradStreams *wst = tc ? wgrad_streams(stream) : nullptr;
        if (tc && !wst) return (int)cudaErrorUnknown;
        cudaError_t ce;
        if (tc) {                                       // fork: both wgrad streams wait for the dgrad chain
            if ((ce = cudaEventRecord(wst->fork, stream)) != cudaSuccess) return (int)ce;
            for (int i = 0; i < 2; ++i)
                if ((ce = cudaStreamWaitEvent(wst->s[i], wst->fork, 0)) != cudaSuccess) return (int)ce;
        }
        // BF16 mode: the tensor-core jobs of the chunk are collected and go out as ONE launch (+ one reduce) on
        // auxiliary stream 0; the two skinny head gradients run beside it on auxiliary stream 1
        WgradJob jobs[16];
        int n_jobs = 0;
        auto wgrad = [&](const float *A, int rows_a, const float *B, int rows_b, const float *dW, int ld, int col_off,
                         const float *db) -> int {
            if (tc) {
                // the big operand rows live in bf16 blocks (train_layout.h: G_* feature numbering)
                const __nv_bfloat16 *wsb = reinterpret_cast<const __nv_bfloat16 *>(ws);
                const int row_a = big_feature((int)((A - ws) / ch)), row_b = big_feature((int)((B - ws) / ch));
                if (rows_a <= 4) {                   // density head (1 row), colour layer 1 (3 rows): own partial areas
                    float *sk = scratch + wgrad_tc_scratch_bytes(256) / sizeof(float) + (size_t)(rows_a == 1 ? 0 : 1) * (wgrad_skinny_scratch_bytes() / sizeof(float));
                    return wgrad_skinny(A, rows_a, (int)ch, wsb, row_b, rows_b, const_cast<float *>(dW), ld, const_cast<float *>(db),
                                        sk, sm_limit, wst->s[1]);
                }
                jobs[n_jobs++] = WgradJob{row_a, rows_a, row_b, rows_b, const_cast<float *>(dW), ld, col_off, const_cast<float *>(db)};
                return 0;
            }
            dim3 grid((rows_a + 63) / 64, (rows_b + 63) / 64, split);
            wgrad_kernel<<<grid, 256, 0, stream>>>(A, rows_a, B, rows_b, (int)ch, const_cast<float *>(dW), ld, col_off,
                                                   const_cast<float *>(db));
            return launch_status();
        };
        // the auxiliary streams are joined back into the caller's stream on EVERY path out of here, error or not:
        // an unjoined fork would leave later work on `stream` unordered against launches already queued on them
        bool forked = tc;
        auto join = [&]() -> int {
            if (!forked) return 0;
            forked = false;
            int bad = 0;
            for (int i = 0; i < 2; ++i) {
                cudaError_t je = cudaEventRecord(wst->join[i], wst->s[i]);
                if (je == cudaSuccess) je = cudaStreamWaitEvent(stream, wst->join[i], 0);
                if (je != cudaSuccess && !bad) { cudaGetLastError(); bad = (int)je; }
            }
            return bad;
        };
        auto all_wgrads = [&]() -> int {
            int rc2;
            if ((rc2 = wgrad(row(R_DPRE), 256, row(R_PE), 63, g.layer_w[0], 63, 0, g.layer_b[0]))) return rc2;
            for (int l = 1; l < 8; ++l) {
                const int ld = l == 4 ? 319 : 256;
                if ((rc2 = wgrad(row(R_DPRE + 256 * l), 256, row(R_H + 256 * (l - 1)), 256, g.layer_w[l], ld, 0, g.layer_b[l]))) return rc2;
                if (l == 4 && (rc2 = wgrad(row(R_DPRE + 256 * 4), 256, row(R_PE), 63, g.layer_w[4], 319, 256, nullptr))) return rc2;
            }
            if ((rc2 = wgrad(row(R_DSIG), 1, row(R_H + 256 * 7), 256, g.density_w, 256, 0, g.density_b))) return rc2;
            if ((rc2 = wgrad(row(R_DPREC0), 128, row(R_H + 256 * 7), 256, g.color0_w, 283, 0, g.color0_b))) return rc2;
            if ((rc2 = wgrad(row(R_DPREC0), 128, row(R_DE), 27, g.color0_w, 283, 256, nullptr))) return rc2;
            if ((rc2 = wgrad(row(R_DY), 3, row(R_C0H), 128, g.color1_w, 128, 0, g.color1_b))) return rc2;
            if (tc)
                return wgrad_tc_batch(reinterpret_cast<const __nv_bfloat16 *>(ws), (int)ch, jobs, n_jobs, scratch,
                                      sm_limit > 0 ? std::min(sms, sm_limit) : sms, wst->s[0]);
            return 0;
        };
        rc = all_wgrads();
        const int jrc = join();
        if (rc) return rc;
        if (jrc) return jrc;
    }
    return 0;
}

}  // extern "C"
