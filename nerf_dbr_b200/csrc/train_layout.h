// Training workspace layout shared by train.cu (FP32 kernels), train_tc.cu and the TRAIN variant of the fused
// tensor-core kernel (mlp_tc.cu): [row][ch] fp32, K-major for the weight-gradient GEMMs (sample contiguous).
#pragma once
// BF16 mode keeps the big operand rows (R_PE, R_H, R_C0H, R_DE, R_DPRE, R_DPREC0) as bf16 in the first 8960 ch bytes of
// the same buffer (slab-major blocks, see G_* below; the first fp32 row it still uses, R_SIGPRE, starts at byte 9088 ch);
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
// ReLU masks written by the tensor-core forward, so the dgrad chain reads 8 bytes per thread and half instead of 64
// activations: uint64 [(layer*4 + part)][ch], one 32-bit word per 32 features (relu_mask_word below)
constexpr int R_MASK = R_DPREC0 + 128;       // 8 layers x 4 parts x 2 floats (one uint64 per sample)
constexpr int R_MASKC0 = R_MASK + 64;        // colour layer 0: 2 parts x 2 floats
constexpr int R_TOTAL = R_MASKC0 + 4;

// BF16 mode: slab-major, sample-major, pre-swizzled operand blocks.  The big operand rows are grouped in blocks of 64
// features (G_* below, in features; every group starts on a block boundary) and the samples in slabs of 64.  One
// (slab, block) is an 8 KB tile [64 samples][64 features] of bf16: a sample's 64 features are 128 contiguous bytes
// whose eight 16-byte units are XOR-swizzled by (sample & 7).  That is byte for byte the 128B-swizzled *MN-major*
// shared-memory image tcgen05 takes for a GEMM whose K dimension is the sample axis, so
//   * a forward / dgrad epilogue thread (one sample, 64 features in registers) stores its 128 bytes with eight
//     16-byte stores and a warp covers 4 KB contiguously (the [feature][sample] alternative costs 64 two-byte
//     stores per thread), and
//   * the operand tile of any row group for one slab is contiguous (blocks x 8 KB): wgrad stages it with a single
//     cp.async.bulk and multiplies it in place (descriptor: LBO = 8 KB between feature blocks, SBO = 1 KB between
//     groups of 8 samples).
constexpr int G_PE = 0;                      // 64   encoded position (feature 63 = 0)
constexpr int G_H = 64;                      // 8 x 256
constexpr int G_C0H = G_H + 8 * 256;         // 128
constexpr int G_DE = G_C0H + 128;            // 64   encoded direction (27 used, 32 written)
constexpr int G_DPRE = G_DE + 64;            // 8 x 256
constexpr int G_DPREC0 = G_DPRE + 8 * 256;   // 128
constexpr int G_TOTAL = G_DPREC0 + 128;      // 4480 features = 70 blocks
constexpr int kBlocksPerSlab = G_TOTAL / 64;
static_assert((size_t)G_TOTAL * 2 <= (size_t)R_SIGPRE * 4, "bf16 operand blocks must end below the first fp32 row BF16 mode uses");
// element (bf16) offset of the 128-byte row of sample `col` in the block that starts at feature g0 (multiple of 64)
__host__ __device__ constexpr size_t big_row(int g0, int col)
{
    return (((size_t)(col >> 6) * kBlocksPerSlab + (g0 >> 6)) * 64 + (col & 63)) * 64;
}
// element offset of the contiguous tile holding the blocks from feature g0 on, for slab `slab`
__host__ __device__ constexpr size_t big_tile(int g0, int slab) { return ((size_t)slab * kBlocksPerSlab + (g0 >> 6)) * 4096; }
// map a fp32-layout row number (R_*) of a big operand group to its feature number (G_*)
__host__ __device__ constexpr int big_feature(int r)
{
    return r < R_SIGPRE ? (r < R_DE ? r : G_DE + (r - R_DE)) : G_DPRE + (r - R_DPRE);
}

// one tensor-core weight-gradient job: dW[n][col_off + k] += sum_s A[n][s] B[k][s], dbias[n] += sum_s A[n][s]
struct WgradJob {
    int row_a, rows_a, row_b, rows_b_valid;      // feature groups in G_* numbering
    float *dW;
    int ld, col_off;
    float *dbias;                                // or nullptr
};

#ifdef __CUDACC__
// Warp-cooperative version for the hot epilogues: lane l holds the 64 features of the sample in row l of the warp's
// 32 rows.  Storing them directly makes every store instruction touch 32 different lines with 16 bytes each (half a
// sector: the L2 sees one partial-sector write per lane).  Going through a 2 KB warp-private staging buffer, four
// units at a time, each store instruction instead writes 8 rows x 64 contiguous bytes -- whole sectors only.
// col_r[j] = workspace sample of row (lane >> 2) + 8 j (or -1), i.e. the rows this lane writes back.
__device__ __forceinline__ void store_block_rows_staged(void *ws, int g0, const int (&col_r)[4], const uint32_t (&pk)[32],
                                                        uint32_t stage, int lane)
{
    const uint32_t wr = stage + lane * 64, sx = (lane >> 1) & 3;
    const int q = lane & 3;
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 4; ++p)                    // slot p of row `lane` (rotated: conflict-free both ways)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wr + ((p ^ sx) << 4)), "r"(pk[16 * ph + 4 * p]),
                         "r"(pk[16 * ph + 4 * p + 1]), "r"(pk[16 * ph + 4 * p + 2]), "r"(pk[16 * ph + 4 * p + 3]) : "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = (lane >> 2) + 8 * j;
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(stage + r * 64 + (((q ^ (r >> 1)) & 3) << 4)) : "memory");
            if (col_r[j] >= 0) {
                uint4 *row = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned short *>(ws) + big_row(g0, col_r[j]));
                row[(4 * ph + q) ^ (col_r[j] & 7)] = v;
            }
        }
    }
}

// The same store through the TMA engine, for a warp whose 32 rows are the 32 consecutive samples col0 .. col0 + 31 with
// col0 a multiple of 32 (every full warp of the training kernels): its rows are 4 KB contiguous in the workspace.  Each
// lane writes its 128-byte row into a 4 KB staging buffer (unit u of row r at u ^ (r & 7): the workspace image, and
// conflict-free), one lane hands the buffer to cp.async.bulk shared -> global.  Against the staged version this drops
// 8 ld.shared + 8 st.global + their address arithmetic per thread, and the L2 sees whole 128-byte lines in 4 KB bursts
// instead of eight half lines per store instruction.  `stage` must alternate between two buffers of the warp from call
// to call: before a buffer is rewritten, the copy issued from it two calls ago has finished reading it
// (wait_group.read 1; bulk groups of a thread complete in order); with a single buffer per warp (kPending = 0) the
// previous copy must have read it.  The issuing lane is always lane 0: it must run bulk_store_drain() before the kernel ends.
template <int kPending = 1>
__device__ __forceinline__ void store_block_rows_bulk(void *ws, int g0, int col0, const uint32_t (&pk)[32], uint32_t stage, int lane)
{
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
    __syncwarp();
    const uint32_t wr = stage + lane * 128;
    const int sw = lane & 7;                                   // (col0 + lane) & 7
#pragma unroll
    for (int u = 0; u < 8; ++u)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wr + ((u ^ sw) << 4)), "r"(pk[4 * u]), "r"(pk[4 * u + 1]),
                     "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3]) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        unsigned short *dst = reinterpret_cast<unsigned short *>(ws) + big_row(g0, col0);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(dst), "r"(stage) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}
__device__ __forceinline__ void bulk_store_drain() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// before a staging buffer is reused by plain stores: no bulk copy of this thread is still reading shared memory
__device__ __forceinline__ void bulk_store_reads_done() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ReLU mask of 32 post-ReLU activations held as 16 bf16 pairs (all >= +0): pair i's flags land at bit i (even
// feature) and bit 16 + i (odd feature).  h + 0x7fff carries into bit 15 exactly when h != 0: 3 instructions a pair.
__device__ __forceinline__ uint32_t relu_mask_word(const uint32_t *pk)
{
    uint32_t acc[4] = {0u, 0u, 0u, 0u};          // four independent chains: the epilogue is latency-bound, not issue-bound
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i & 3] |= ((pk[i] + 0x7fff7fffu) >> (15 - i)) & (0x00010001u << i);
    return (acc[0] | acc[1]) | (acc[2] | acc[3]);
}
// ... and the AND-mask (0xffff per active half) of pair i: shift the flags to the byte sign bits, replicate them
__device__ __forceinline__ uint32_t relu_pair_mask(uint32_t word, int i)
{
    uint32_t r;                                  // prmt selector nibble 8 | b: replicate the sign bit of byte b
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(word << (15 - i)), "r"(0u), "r"(0xBB99u));
    return r;
}
// one sample's 64 features of one block: eight 16-byte stores (pk[4u .. 4u+3] = features 8u .. 8u+7 as bf16 pairs)
__device__ __forceinline__ void store_block_row(void *ws, int g0, int col, const uint32_t (&pk)[32], int units = 8)
{
    uint4 *row = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned short *>(ws) + big_row(g0, col));
    const int sw = col & 7;
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (u < units) row[u ^ sw] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
}
#endif
}  // namespace nerfb200
