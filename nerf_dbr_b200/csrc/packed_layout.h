// Layout of one packed network (the caller-owned buffer nerf_b200_pack_weights fills).
//
//   [ fp32 region ]  biases, head weights, and K-major ([K][N]) fp32 copies of every
//                    weight matrix for the FP32 (CUDA-core) mode
//   [ bf16 region ]  the tcgen05 operand stream: a fixed sequence of K-chunks, each chunk
//                    [N rows x 64 K] bf16 in the 128-byte-swizzled K-major canonical layout
//                    (row n at n*128 B, 16-byte unit c of the row stored at unit c ^ (n & 7)),
//                    i.e. byte-for-byte what a B operand tile must look like in shared
//                    memory, so one cp.async.bulk per chunk stages it.
//
// Source tensors: reference src/models/nerf.py:72-90 (nn.Linear weights are [out,in]).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace nerfb200 {

constexpr int kHidden = 256;
constexpr int kPosFreq = 10, kDirFreq = 4;
constexpr int kPosFeat = 63, kDirFeat = 27;    // 3 + 6L
constexpr int kPosPad = 64, kDirPad = 32;      // padded K of the encoded inputs
constexpr int kColorHidden = 128;

// ---- fp32 region (offsets in floats) ----------------------------------------------------
constexpr size_t F_BIAS = 0;                                  // [8][256] trunk biases
constexpr size_t F_WSIG = F_BIAS + 8 * 256;                   // [256] density head weight
constexpr size_t F_BSIG = F_WSIG + 256;                       // [1] (+3 pad)
constexpr size_t F_BC0 = F_BSIG + 4;                          // [128] colour-0 bias
constexpr size_t F_WC1 = F_BC0 + 128;                         // [3][128] colour-1 weight
constexpr size_t F_BC1 = F_WC1 + 3 * 128;                     // [3] (+1 pad)
constexpr size_t F_W0T = F_BC1 + 4;                           // [64][256]  layer 0, K-major, row 63 = 0
constexpr size_t F_WT = F_W0T + 64 * 256;                     // 7 x [256][256]: layers 1..7 (layer 4: hidden part)
constexpr size_t F_W4P = F_WT + 7 * 256 * 256;                // [64][256]  layer 4, encoded-position part
constexpr size_t F_WC0H = F_W4P + 64 * 256;                   // [256][128] colour-0, hidden part
constexpr size_t F_WC0D = F_WC0H + 256 * 128;                 // [32][128]  colour-0, encoded-direction part (rows 27..31 = 0)
constexpr size_t F_END = F_WC0D + 32 * 128;
__host__ __device__ constexpr size_t f_wt(int layer) { return F_WT + (size_t)(layer - 1) * 256 * 256; }

// ---- bf16 region --------------------------------------------------------------------------
constexpr size_t kChunkBytes256 = 256 * 128;                  // [256 x 64] bf16
constexpr size_t kChunkBytes128 = 128 * 128;                  // [128 x 64] bf16
// chunk sequence per sample tile: L0 (1: pe) | L1..L3 (4 each) | L4 (4 hidden + 1 pe) |
// L5..L7 (4 each) | C0 (4 chunks of N=128)
constexpr int kChunks256 = 1 + 12 + 5 + 12;                   // 30
constexpr int kChunks128 = 4;
constexpr size_t B_OFFSET = ((F_END * 4 + 1023) / 1024) * 1024;   // byte offset of the bf16 region
constexpr size_t B_C0 = (size_t)kChunks256 * kChunkBytes256;      // byte offset of colour-0 chunks inside it
constexpr size_t B_BYTES = B_C0 + (size_t)kChunks128 * kChunkBytes128;
// optional low-order bf16 stream for the split-precision mode follows (same layout)
constexpr size_t B_LO_OFFSET = B_OFFSET + B_BYTES;
constexpr size_t PACKED_BYTES = B_LO_OFFSET + B_BYTES;

// byte offset of element (n, k) inside one swizzled chunk
__host__ __device__ constexpr uint32_t swz128(uint32_t n, uint32_t k)
{
    return n * 128u + ((((k >> 3) ^ (n & 7u)) << 4) | ((k & 7u) << 1));
}

}  // namespace nerfb200
