// Layout of the FP8-mode weight buffer (nerf_b200_pack_weights_fp8) and the schedule constants of its kernel.
//
// FP8 mode = the fused render kernel with e4m3 operands for the 256-wide contractions (tcgen05.mma.kind::f8f6f4, twice
// the bf16 rate), bf16 kept where precision matters most and costs least (layer 0 and the skip part of layer 4, whose A
// operand is the encoded position), fp32 accumulation, fp32 heads and compositing as in BF16 mode.  It is the B200
// counterpart of the reference's CompressedNeRFRenderer (src/benchmark/compressed_renderer.py:89-211: per-tensor int8
// weights dequantised to fp16), judged against that renderer rather than against the 1e-4 / 0.05 dB gates.
//
// Scaling (all scales are powers of two, so applying them is exact):
//   activations of trunk layer l:  hq_l = e4m3(relu(pre_l) * sa_l),   sa_l = 2^floor(log2(240 / max|h_l|)) from a
//                                  calibration pass of the fp32 network over sample points of the scene (pack time)
//   weights of layer l >= 1:       Wq_l[n][k] = e4m3(W_l[n][k] * sw_l[n]),   sw_l[n] = 2^floor(log2(448 / max_k |W_l[n][k]|))
//   so the accumulator of layer l is sa_{l-1} sw_l[n] pre_l[n]; the epilogue computes
//                                  hq_l = e4m3(relu(acc * m_l[n] + b'_l[n])),   m_l[n] = sa_l / (sa_{l-1} sw_l[n]),  b'_l = b_l sa_l
//   layer 4's bf16 skip weights are pre-multiplied by sa_3 sw_4[n]; colour layer 0's column scale s[n] = sa_7 sw_c0[n]
//   is folded into its fp32 per-ray bias (x s) and into colour layer 1's weights (/ s); the density column's inverse
//   scale is applied by the back warps.
//
// Buffer:
//   [ fp32 region, Q_F32 floats ]  the BF16 kernel's small tables AT THE SAME OFFSETS (packed_layout.h: F_BIAS .. F_BC1,
//                                  F_WC0D) holding the scaled values above, so the shared front / back code reads them
//                                  unchanged; then Q_MUL [8][256] multipliers, Q_SA [16] activation scales (+ amax),
//                                  Q_SW [9][256] weight scales, Q_SIGINV.
//   [ operand stream at Q_OFFSET ] 17 stages per tile in consumption order:
//       stage 0            layer 0: two bf16 chunks [128 n x 64 k] (h0, h1)
//       2 per trunk layer  e4m3 chunks [128 n x 128 k] (16 KB, rows of 128 B, 128-byte swizzle): (h0 kp0, h1 kp0),
//                          (h0 kp1, h1 kp1) with kp = pair of 64-wide K-blocks
//       layer 4            first a stage of two bf16 skip chunks (h0, h1), then its two e4m3 stages
//       stage 16           colour layer 0: two e4m3 chunks [144 n x 128 k] (rows 0..127 colour, 128 density, rest 0)
#pragma once
#include "packed_layout.h"

namespace nerfb200 {

constexpr size_t Q_MUL = F_WO;                                // [8][256]  (F_WO = end of the F_WC0D block)
constexpr size_t Q_SA = Q_MUL + 8 * 256;                      // [8] activation scales sa_l, [8..16) calibration maxima
constexpr size_t Q_SW = Q_SA + 16;                            // [9][256] weight scales (row 8: colour layer 0, [8][128] = density)
constexpr size_t Q_SIGINV = Q_SW + 9 * 256;                   // [1] (+3 pad) 1 / (sa_7 sw_sigma)
constexpr size_t Q_F32 = Q_SIGINV + 4;
constexpr size_t Q_OFFSET = ((Q_F32 * 4 + 1023) / 1024) * 1024;

constexpr int kQStages = 17;
constexpr int kQStageBytes = 32768, kQStageBytesC0 = 2 * kC0Rows * 128;   // 36864
constexpr size_t Q_STREAM_BYTES = 16 * (size_t)kQStageBytes + kQStageBytesC0;
constexpr size_t PACKED_FP8_BYTES = Q_OFFSET + Q_STREAM_BYTES;
__host__ __device__ constexpr uint32_t q_stage_bytes(int st) { return st < 16 ? kQStageBytes : kQStageBytesC0; }
__host__ __device__ constexpr size_t q_stage_offset(int st) { return (size_t)st * kQStageBytes; }
// first stage of a layer (0..7 trunk, 8 colour layer 0); layer 4's first stage is its bf16 skip stage
__host__ __device__ constexpr int q_layer_stage(int layer)
{
    return layer == 0 ? 0 : layer <= 4 ? 1 + 2 * (layer - 1) : layer < 8 ? 10 + 2 * (layer - 5) : 16;
}
// byte offset of e4m3 element (n, k) inside one [rows x 128 k] chunk (rows of 128 B, 16-byte units XOR-swizzled by n & 7)
__host__ __device__ constexpr uint32_t swz128_u8(uint32_t n, uint32_t k)
{
    return n * 128u + ((((k >> 4) ^ (n & 7u)) << 4) | (k & 15u));
}

}  // namespace nerfb200
