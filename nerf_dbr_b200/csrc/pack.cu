// Weight packing: 22 fp32 state-dict tensors -> the packed buffer of packed_layout.h.
// Runs once per checkpoint in setup() (reference pattern: cpu_optimized_renderer.py:31-52).
#include "common.cuh"

namespace nerfb200 {

struct ChunkSrc { const float *w; int ld; int n0; int k0; int k_valid; int rows; const float *extra; };
// rows n0.. (n < rows), columns k0..k0+63; `extra` = one more row (the density head inside colour layer 0's chunks)

__constant__ ChunkTable kPackTable = make_chunk_table();
__constant__ ChunkTable kPackDgTable = make_dgrad_table();

// source of bf16 chunk `ci` of the consumption-ordered stream (packed_layout.h)
__device__ __forceinline__ ChunkSrc chunk_source(const nerf_b200_params &p, int ci)
{
    const ChunkInfo c = kPackTable.c[ci];
    const int n0 = 128 * c.half;
    if (c.layer == 0) return {p.layer_w[0], 63, n0, 0, 63, 128, nullptr};
    if (c.layer == 8) return {p.color0_w, 283, n0, 64 * c.asrc, 64, 128, p.density_w + 64 * c.asrc};
    if (c.layer == 4) return c.asrc == 4 ? ChunkSrc{p.layer_w[4], 319, n0, 256, 63, 128, nullptr}
                                         : ChunkSrc{p.layer_w[4], 319, n0, 64 * c.asrc, 64, 128, nullptr};
    return {p.layer_w[c.layer], 256, n0, 64 * c.asrc, 64, 128, nullptr};
}

__global__ void pack_kernel(nerf_b200_params p, unsigned char *__restrict__ packed, int what)
{
    float *f = reinterpret_cast<float *>(packed);
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;

    // ---- fp32 region ----  (without PACK_FP32_MATRICES: only what the tensor-core kernels read -- biases, head weights
    // [0, F_W0T) and colour layer 0's direction block [F_WC0D, F_WO))
    const bool mats = (what & NERF_B200_PACK_FP32_MATRICES) != 0;
    const size_t lean = F_W0T + (F_WO - F_WC0D);
    for (size_t ii = tid; ii < (mats ? F_END : lean); ii += nth) {
        const size_t i = mats ? ii : (ii < F_W0T ? ii : F_WC0D + (ii - F_W0T));
        float v = 0.f;
        if (i < F_WSIG) { int l = (int)(i / 256), n = (int)(i % 256); v = p.layer_b[l][n]; }
        else if (i < F_BSIG) v = p.density_w[i - F_WSIG];
        else if (i < F_BC0) v = (i == F_BSIG) ? p.density_b[0] : 0.f;
        else if (i < F_WC1) v = p.color0_b[i - F_BC0];
        else if (i < F_BC1) v = p.color1_w[i - F_WC1];
        else if (i < F_W0T) v = (i - F_BC1 < 3) ? p.color1_b[i - F_BC1] : 0.f;
        else if (i < F_WT) { size_t j = i - F_W0T; int k = (int)(j / 256), n = (int)(j % 256); v = k < 63 ? p.layer_w[0][n * 63 + k] : 0.f; }
        else if (i < F_W4P) {
            size_t j = i - F_WT; int l = 1 + (int)(j / 65536); j %= 65536;
            int k = (int)(j / 256), n = (int)(j % 256);
            v = p.layer_w[l][(size_t)n * (l == 4 ? 319 : 256) + k];
        }
        else if (i < F_WC0H) { size_t j = i - F_W4P; int k = (int)(j / 256), n = (int)(j % 256); v = k < 63 ? p.layer_w[4][n * 319 + 256 + k] : 0.f; }
        else if (i < F_WC0D) { size_t j = i - F_WC0H; int k = (int)(j / 128), n = (int)(j % 128); v = p.color0_w[n * 283 + k]; }
        else if (i < F_WO) { size_t j = i - F_WC0D; int k = (int)(j / 128), n = (int)(j % 128); v = k < 27 ? p.color0_w[n * 283 + 256 + k] : 0.f; }
        else if (i < F_WC0O) {
            size_t j = i - F_WO; int l = 1 + (int)(j / 65536); j %= 65536;
            int n = (int)(j / 256), k = (int)(j % 256);
            v = p.layer_w[l][(size_t)n * (l == 4 ? 319 : 256) + k];
        }
        else { size_t j = i - F_WC0O; int n = (int)(j / 256), k = (int)(j % 256); v = p.color0_w[n * 283 + k]; }
        f[i] = v;
    }

    // ---- bf16 region: one 16-byte unit (8 consecutive k of one row) per thread-iteration ----
    constexpr size_t units_trunk = (size_t)(kChunksPerTile - kChunksC0) * 128 * 8;
    constexpr size_t units = units_trunk + (size_t)kChunksC0 * kC0Rows * 8;
    for (size_t uidx = tid; uidx < units; uidx += nth) {
        int ci, n, unit;
        if (uidx < units_trunk) { ci = (int)(uidx / 1024); n = (int)((uidx % 1024) / 8); unit = (int)(uidx % 8); }
        else {
            size_t r = uidx - units_trunk;
            ci = (kChunksPerTile - kChunksC0) + (int)(r / (kC0Rows * 8)); n = (int)((r % (kC0Rows * 8)) / 8); unit = (int)(r % 8);
        }
        const ChunkSrc src = chunk_source(p, ci);
        __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int k = unit * 8 + j;
            float w = 0.f;
            if (k < src.k_valid) {
                if (n < src.rows) w = src.w[(size_t)(src.n0 + n) * src.ld + src.k0 + k];
                else if (n == src.rows && src.extra) w = src.extra[k];
            }
            hi[j] = __float2bfloat16_rn(w);
            lo[j] = __float2bfloat16_rn(w - __bfloat162float(hi[j]));
        }
        size_t off = chunk_offset(ci) + swz128((uint32_t)n, (uint32_t)unit * 8);
        *reinterpret_cast<uint4 *>(packed + B_OFFSET + off) = *reinterpret_cast<const uint4 *>(hi);
        if (what & NERF_B200_PACK_BF16_LO) *reinterpret_cast<uint4 *>(packed + B_LO_OFFSET + off) = *reinterpret_cast<const uint4 *>(lo);
    }

    // ---- dgrad stream: chunk (g, half, kb), element (k, n) = W[64 kb + n][128 half + k] ----
    const size_t dg_units = (what & NERF_B200_PACK_DGRAD) ? (size_t)kDgChunks * 128 * 8 : 0;
    for (size_t uidx = tid; uidx < dg_units; uidx += nth) {
        const int ci = (int)(uidx / 1024), k = (int)((uidx % 1024) / 8), unit = (int)(uidx % 8);
        const ChunkInfo c = kPackDgTable.c[ci];
        const int kin = 128 * c.half + k;                     // input feature of the layer = output of the dgrad GEMM
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = 64 * c.asrc + unit * 8 + j;         // output feature of the layer = K of the dgrad GEMM
            float w = 0.f;
            if (c.layer == 0) {
                if (n < 128) w = p.color0_w[(size_t)n * 283 + kin];
                else if (n == 128) w = p.density_w[kin];
            } else {
                const int l = 8 - c.layer;
                w = p.layer_w[l][(size_t)n * (l == 4 ? 319 : 256) + kin];
            }
            v[j] = __float2bfloat16_rn(w);
        }
        *reinterpret_cast<uint4 *>(packed + B_DG_OFFSET + (size_t)ci * kChunkBytes + swz128((uint32_t)k, (uint32_t)unit * 8)) =
            *reinterpret_cast<const uint4 *>(v);
    }
}

}  // namespace nerfb200

using namespace nerfb200;

extern "C" {

size_t nerf_b200_packed_bytes(void) { return PACKED_BYTES; }

int nerf_b200_pack_weights(const nerf_b200_params *params_host, void *packed, void *stream)
{
    return nerf_b200_pack_weights_ex(params_host, packed, NERF_B200_PACK_ALL, stream);
}

int nerf_b200_pack_weights_ex(const nerf_b200_params *params_host, void *packed, int what, void *stream)
{
    if (!params_host || !packed || (what & ~NERF_B200_PACK_ALL)) return NERF_B200_EINVAL;
    const nerf_b200_params &p = *params_host;
    for (int l = 0; l < 8; ++l)
        if (!p.layer_w[l] || !p.layer_b[l]) return NERF_B200_EINVAL;
    if (!p.density_w || !p.density_b || !p.color0_w || !p.color0_b || !p.color1_w || !p.color1_b)
        return NERF_B200_EINVAL;
    if ((uintptr_t)packed & 1023) return NERF_B200_EALIGN;
    pack_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(p, reinterpret_cast<unsigned char *>(packed), what);
    return launch_status();
}

}  // extern "C"
