// Training workspace layout shared by train.cu (FP32 kernels), train_tc.cu and the TRAIN variant of the fused
// tensor-core kernel (mlp_tc.cu): [row][ch] fp32, K-major for the weight-gradient GEMMs (sample contiguous).
#pragma once
// BF16 mode keeps the big operand rows (R_PE, R_H, R_C0H, R_DE, R_DPRE, R_DPREC0) as bf16 in the first 8912 ch bytes of
// the same buffer (slab-major, see big_off below; the first fp32 row it still uses, R_SIGPRE, starts at byte 9088 ch);
// the small per-sample rows and the masks stay fp32 / uint64 [row][ch].
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

// BF16 mode: slab-major, pre-swizzled operand rows.  Samples are cut into slabs of 64; a slab holds every big row as
// 128 bytes (64 bf16) in row order, and the eight 16-byte units of a row are XOR-swizzled by (row & 7) -- byte for byte
// the 128B-swizzled K-major shared-memory image of a wgrad operand tile.  Rows [r0, r0 + n) of one slab are therefore
// one contiguous n x 128 B block: wgrad stages an operand tile with a single cp.async.bulk, and the forward / dgrad
// epilogues' stores land in a compact region instead of 64 pages.  (Every operand group starts at a multiple of 8
// rows, so the swizzle by the global row number equals the swizzle by the row inside the tile.)
constexpr int R_BIG = R_DPREC0 + 128;        // rows that exist as bf16 operand rows (R_PE .. R_DPREC0)
static_assert(R_H % 8 == 0 && R_C0H % 8 == 0 && R_DE % 8 == 0 && R_DPRE % 8 == 0 && R_DPREC0 % 8 == 0, "tile-local swizzle");
// element (bf16) offset of (row r, sample col)
__host__ __device__ constexpr size_t big_off(int r, int col)
{
    return ((size_t)(col >> 6) * R_BIG + r) * 64 + (size_t)(((((col & 63) >> 3) ^ (r & 7)) << 3) | (col & 7));
}
// element offset of the tile rows [r0, ..) of slab `slab`
__host__ __device__ constexpr size_t big_tile(int r0, int slab) { return ((size_t)slab * R_BIG + r0) * 64; }
}  // namespace nerfb200
