// Shared CUDA-core tile primitives of the FP32 mode: 64-sample tiles, K-major activations in
// shared memory, [16][N] weight stages, 8(m) x 8|4(n) register tiles.  Used by mlp_simt.cu
// (inference) and train.cu (forward with stored activations, dgrad chain).
#pragma once
#include "common.cuh"

namespace nerfb200 {

constexpr int TM = 64;
constexpr int kSimtThreads = 256;

constexpr int kItemMax = 2048;           // samples per render work item kept in shared memory

struct SimtSmem {
    float actA[256 * TM];
    float actB[256 * TM];
    float pe[64 * TM];
    float de[32 * TM];
    float ws[16 * 256];
    float pos[3 * TM];
    float dir[3 * TM];
    float4 out[kItemMax];                 // (sigma, r, g, b) per sample of the current item
};

// acc[8][CN] += act[K][TM] (shared, K-major) x Wg[K][N] (global, K-major)
template <int CN>
__device__ __forceinline__ void simt_accumulate(float (&acc)[8][CN], const float *__restrict__ act,
                                                int K, const float *__restrict__ Wg, float *ws,
                                                int tid, int m0, int n0)
{
    constexpr int N = CN * 32;
    for (int k0 = 0; k0 < K; k0 += 16) {
        __syncthreads();                                   // previous stage fully consumed
        const float4 *src = reinterpret_cast<const float4 *>(Wg + (size_t)k0 * N);
        for (int i = tid; i < 4 * N; i += kSimtThreads) reinterpret_cast<float4 *>(ws)[i] = __ldg(src + i);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[8], w[CN];
            const float4 *ap = reinterpret_cast<const float4 *>(act + (size_t)(k0 + kk) * TM + m0);
            float4 a0 = ap[0], a1 = ap[1];
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            const float4 *wp = reinterpret_cast<const float4 *>(ws + kk * N + n0);
#pragma unroll
            for (int q = 0; q < CN / 4; ++q) {
                float4 t = wp[q];
                w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < CN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
    }
}

template <int CN>
__device__ __forceinline__ void simt_store_relu(const float (&acc)[8][CN], const float *__restrict__ bias,
                                                float *out, int m0, int n0)
{
#pragma unroll
    for (int j = 0; j < CN; ++j) {
        float b = __ldg(bias + n0 + j);
        float4 lo = make_float4(fmaxf(acc[0][j] + b, 0.f), fmaxf(acc[1][j] + b, 0.f),
                                fmaxf(acc[2][j] + b, 0.f), fmaxf(acc[3][j] + b, 0.f));
        float4 hi = make_float4(fmaxf(acc[4][j] + b, 0.f), fmaxf(acc[5][j] + b, 0.f),
                                fmaxf(acc[6][j] + b, 0.f), fmaxf(acc[7][j] + b, 0.f));
        float4 *op = reinterpret_cast<float4 *>(out + (size_t)(n0 + j) * TM + m0);
        op[0] = lo; op[1] = hi;
    }
}

template <int CN>
__device__ __forceinline__ void zero_acc(float (&acc)[8][CN])
{
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < CN; ++j) acc[i][j] = 0.f;
}

// Encoded inputs for the 64 samples whose positions / directions sit in sm.pos / sm.dir.
static __device__ void simt_encode(SimtSmem &sm, int tid)
{
    const int m = tid & (TM - 1), g = tid >> 6;            // 4 feature groups
    for (int j = g; j < 64; j += 4) {
        float v = 0.f;
        if (j < 3) v = sm.pos[j * TM + m];
        else if (j < 63) {
            int k = (j - 3) / 6, w = (j - 3) - 6 * k, c = w % 3;
            float arg = __fmul_rn(kPiF * (float)(1 << k), sm.pos[c * TM + m]);
            v = w < 3 ? sinf(arg) : cosf(arg);
        }
        sm.pe[j * TM + m] = v;
    }
    for (int j = g; j < 32; j += 4) {
        float v = 0.f;
        if (j < 3) v = sm.dir[j * TM + m];
        else if (j < 27) {
            int k = (j - 3) / 6, w = (j - 3) - 6 * k, c = w % 3;
            float arg = __fmul_rn(kPiF * (float)(1 << k), sm.dir[c * TM + m]);
            v = w < 3 ? sinf(arg) : cosf(arg);
        }
        sm.de[j * TM + m] = v;
    }
}

// The network on one 64-sample tile: sm.pe / sm.de -> (sigma, rgb) for row m < 64 returned to
// thread tid == m (others return garbage).
// amax (optional, device float[8], calibration of the FP8 mode): amax[l] = max over everything seen of trunk layer l's
// activations (non-negative floats order like their bit patterns, so an integer atomicMax does it)
template <int CN>
__device__ __forceinline__ void simt_track_max(const float (&acc)[8][CN], const float *__restrict__ bias, int n0, float *amax_l)
{
    float mx = 0.f;
#pragma unroll
    for (int j = 0; j < CN; ++j) {
        const float b = __ldg(bias + n0 + j);
#pragma unroll
        for (int i = 0; i < 8; ++i) mx = fmaxf(mx, acc[i][j] + b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(amax_l), __float_as_int(mx));
}

static __device__ void simt_network(SimtSmem &sm, const float *__restrict__ wf, int tid, float4 &result, float *amax = nullptr)
{
    const int m0 = (tid & 7) * 8, ng = tid >> 3;
    float acc[8][8];
    const int n0 = ng * 8;
    // layer 0
    zero_acc(acc);
    simt_accumulate<8>(acc, sm.pe, 64, wf + F_W0T, sm.ws, tid, m0, n0);
    simt_store_relu<8>(acc, wf + F_BIAS, sm.actA, m0, n0);
    if (amax) simt_track_max<8>(acc, wf + F_BIAS, n0, amax);
    float *in = sm.actA, *out = sm.actB;
    for (int l = 1; l < 8; ++l) {
        zero_acc(acc);
        simt_accumulate<8>(acc, in, 256, wf + f_wt(l), sm.ws, tid, m0, n0);
        if (l == 4) simt_accumulate<8>(acc, sm.pe, 64, wf + F_W4P, sm.ws, tid, m0, n0);
        // `out` was last read two layers ago; the __syncthreads inside accumulate order it
        simt_store_relu<8>(acc, wf + F_BIAS + l * 256, out, m0, n0);
        if (amax) simt_track_max<8>(acc, wf + F_BIAS + l * 256, n0, amax + l);
        float *t = in; in = out; out = t;
    }
    // `in` = layer-7 activations h; colour layer 0 (N = 128) -> `out`
    {
        float c[8][4];
        const int n4 = ng * 4;
        zero_acc(c);
        simt_accumulate<4>(c, in, 256, wf + F_WC0H, sm.ws, tid, m0, n4);
        simt_accumulate<4>(c, sm.de, 32, wf + F_WC0D, sm.ws, tid, m0, n4);
        simt_store_relu<4>(c, wf + F_BC0, out, m0, n4);
    }
    __syncthreads();
    if (tid < TM) {
        float s = 0.f;
        for (int k = 0; k < 256; ++k) s = fmaf(in[k * TM + tid], __ldg(wf + F_WSIG + k), s);
        s = fmaxf(s + __ldg(wf + F_BSIG), 0.f);
        float y[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float t = 0.f;
            for (int k = 0; k < 128; ++k) t = fmaf(out[k * TM + tid], __ldg(wf + F_WC1 + ch * 128 + k), t);
            t += __ldg(wf + F_BC1 + ch);
            y[ch] = 1.0f / (1.0f + expf(-t));
        }
        result = make_float4(s, y[0], y[1], y[2]);
    }
}


}  // namespace nerfb200
