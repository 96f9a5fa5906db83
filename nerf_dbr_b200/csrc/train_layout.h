// Training workspace layout shared by train.cu (FP32 kernels), train_tc.cu and the TRAIN variant of the fused
// tensor-core kernel (mlp_tc.cu): [row][ch] fp32, K-major for the weight-gradient GEMMs (sample contiguous).
#pragma once
// BF16 mode keeps the big operand rows (R_PE, R_H, R_C0H, R_DE, R_DPRE, R_DPREC0) as bf16 with the SAME row numbers and
// pitch `ch` elements, i.e. row r at byte r * ch * 2 of the same buffer (they end at byte 8912 ch, below the first fp32
// row it still uses, R_SIGPRE at byte 9088 ch); the small per-sample rows and the masks stay fp32 / uint64.
namespace nerfb200 {
constexpr int R_PE = 0;                      // 64   encoded position (row 63 = 0)
constexpr int R_H = R_PE + 64;               // 8 x 256   post-ReLU trunk activations
constexpr int R_C0H = R_H + 8 * 256;         // 128  post-ReLU colour layer 0
constexpr int R_DE = R_C0H + 128;            // 32   encoded direction (27 used)
constexpr int R_SIGPRE = R_DE + 32;          // 1    density head pre-activation
constexpr int R_RGB = R_SIGPRE + 1;          // 3    post-sigmoid colour
constexpr int R_DSIG = R_RGB + 3;            // 1    dL/d sigma_pre
constexpr int R_DY = R_DSIG + 1;             // 3    dL/d colour pre-sigmoid
constexpr int R_DPRE = R_DY + 3;             // 8 x 256   dL/d pre-activation of trunk layers
constexpr int R_DPREC0 = R_DPRE + 8 * 256;   // 128
// ReLU masks written by the tensor-core forward (bit j of word (layer, part) = activation 64*part + j > 0), so the
// dgrad chain reads 8 bytes per thread and half instead of 64 strided floats: uint64 [(layer*4 + part)][ch]
constexpr int R_MASK = R_DPREC0 + 128;       // 8 layers x 4 parts x 2 floats (one uint64 per sample)
constexpr int R_MASKC0 = R_MASK + 64;        // colour layer 0: 2 parts x 2 floats
constexpr int R_TOTAL = R_MASKC0 + 4;
}  // namespace nerfb200
