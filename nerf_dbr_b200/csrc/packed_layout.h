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
constexpr size_t F_WO = F_WC0D + 32 * 128;                    // 7 x [256 out][256 in]: layers 1..7 in nn.Linear order (layer 4:
                                                              // hidden columns only) -- the dgrad operand of the training step
constexpr size_t F_WC0O = F_WO + 7 * 256 * 256;               // [128 out][256 in] colour-0, hidden columns
constexpr size_t F_END = F_WC0O + 128 * 256;
__host__ __device__ constexpr size_t f_wo(int layer) { return F_WO + (size_t)(layer - 1) * 256 * 256; }
__host__ __device__ constexpr size_t f_wt(int layer) { return F_WT + (size_t)(layer - 1) * 256 * 256; }

// ---- bf16 region --------------------------------------------------------------------------
// The tcgen05 kernel computes every 256-wide layer as two N = 128 output-column halves (one MMA
// instruction = M128 x N128 x K16 = 64 tensor-pipe cycles; narrower instructions fall below the
// pipe's ~46-cycle issue floor, tools/probe/mma_probe.cu) and streams the weights as
// [128 n x 64 k] bf16 chunks (16 KB, 128B-swizzled K-major) in EXACTLY the order its MMA issuer
// consumes them, so the producer warp just copies consecutive 32 KB stages (2 chunks).
// Per 128-sample tile: 64 chunks = 32 stages = 1 MiB.
//
// Order inside a layer:  h0k0 h0k1 h1k0 h0k2 h0k3 h1k1 h1k2 h1k3   (h = output half, k = K-block)
//   * K-blocks 0,1 of the input are the previous layer's half 0, K-blocks 2,3 its half 1 (whose
//     epilogue finishes last): nothing before the 4th chunk needs them -> ~770 cycles of slack.
//   * half 0 completes after the 5th chunk, so its epilogue (bias, ReLU, bf16, write-back) is done
//     before the layer ends; half 1's epilogue overlaps the next layer's first three chunks.
//   layer 4 adds an encoded-position chunk per half (no dependency); layer 0 is those alone;
//   colour layer 0 has one half of N = 144: rows 0..127 are colour layer 0's hidden-input weights,
//   row 128 is the DENSITY HEAD (its output lands in accumulator column 128 -- 256 FMAs per sample
//   that would otherwise run on CUDA cores), rows 129..143 are zero padding (N must be a multiple of 16).
constexpr int kChunkBytes = 128 * 128;                        // [128 x 64] bf16
constexpr int kC0Rows = 144;
constexpr int kChunkBytesC0 = kC0Rows * 128;                  // [144 x 64] bf16
constexpr int kChunksPerTile = 64;
constexpr int kChunksC0 = 4;                                  // the last four chunks
constexpr int kStageChunks = 2;
constexpr int kStageBytes = kStageChunks * kChunkBytes;       // 32 KB (trunk stages)
constexpr int kStageBytesC0 = kStageChunks * kChunkBytesC0;   // 36 KB (the two colour-layer-0 stages)
constexpr int kStagesPerTile = kChunksPerTile / kStageChunks; // 32
constexpr int kStagesC0 = kChunksC0 / kStageChunks;           // 2
constexpr int kStageSlotBytes = kStageBytesC0;                // ring slot size
__host__ __device__ constexpr size_t chunk_offset(int ci)
{
    return ci < kChunksPerTile - kChunksC0 ? (size_t)ci * kChunkBytes
                                           : (size_t)(kChunksPerTile - kChunksC0) * kChunkBytes + (size_t)(ci - (kChunksPerTile - kChunksC0)) * kChunkBytesC0;
}
__host__ __device__ constexpr size_t stage_offset(int st) { return chunk_offset(st * kStageChunks); }
__host__ __device__ constexpr uint32_t stage_bytes(int st) { return st < kStagesPerTile - kStagesC0 ? kStageBytes : kStageBytesC0; }

struct ChunkInfo {
    uint8_t layer;   // 0..7 trunk, 8 = colour layer 0
    uint8_t half;    // output half: columns [128 half, 128 half + 128)
    uint8_t asrc;    // A operand: 0..3 = hidden K-block of the previous layer, 4 = encoded position
    uint8_t flags;   // 1 = first chunk of this half (overwrite), 2 = last chunk of this half,
                     // 4 = first use of this A K-block in the layer (wait for its epilogue)
};
struct ChunkTable { ChunkInfo c[kChunksPerTile]; };

constexpr ChunkTable make_chunk_table()
{
    ChunkTable t{};
    int n = 0;
    for (int layer = 0; layer < 9; ++layer) {
        const int begin = n;
        const int halves = layer == 8 ? 1 : 2;
        if (layer == 0) {
            t.c[n++] = ChunkInfo{0, 0, 4, 0};
            t.c[n++] = ChunkInfo{0, 1, 4, 0};
        } else if (layer == 8) {
            for (int kb = 0; kb < 4; ++kb) t.c[n++] = ChunkInfo{8, 0, (uint8_t)kb, 0};
        } else {
            const uint8_t L = (uint8_t)layer;
            if (layer == 4) t.c[n++] = ChunkInfo{L, 0, 4, 0};
            t.c[n++] = ChunkInfo{L, 0, 0, 0};
            t.c[n++] = ChunkInfo{L, 0, 1, 0};
            if (layer == 4) t.c[n++] = ChunkInfo{L, 1, 4, 0};
            t.c[n++] = ChunkInfo{L, 1, 0, 0};
            t.c[n++] = ChunkInfo{L, 0, 2, 0};
            t.c[n++] = ChunkInfo{L, 0, 3, 0};
            t.c[n++] = ChunkInfo{L, 1, 1, 0};
            t.c[n++] = ChunkInfo{L, 1, 2, 0};
            t.c[n++] = ChunkInfo{L, 1, 3, 0};
        }
        for (int h = 0; h < halves; ++h) {
            int first = -1, last = -1;
            for (int i = begin; i < n; ++i)
                if (t.c[i].half == h) { if (first < 0) first = i; last = i; }
            t.c[first].flags |= 1;
            t.c[last].flags |= 2;
        }
        for (int kb = 0; kb < 4; ++kb)
            for (int i = begin; i < n; ++i)
                if (t.c[i].asrc == kb) { t.c[i].flags |= 4; break; }
    }
    return t;
}
static_assert(make_chunk_table().c[kChunksPerTile - 1].layer == 8, "chunk table must fill exactly 64 entries");

// ---- dgrad stream (training backward on the tensor cores) ---------------------------------
// dX = dY . W per layer, run as the same N = 128-half / K-block schedule as the forward: 8 GEMMs per tile
//   G0: dh7 = [dpre_c0 (128) | dsigma_pre (1)] . [W_c0[:, :256] ; w_sigma]      (K-blocks 0,1 = colour rows, 2 = the density row, 3 = 0)
//   Gg: dh_{7-g} = dpre_{8-g} . W_{8-g}[:, :256]   for g = 1..7
// chunk (g, half, kb) = [128 k x 64 n] bf16, element (k, n) = W[64 kb + n][128 half + k]: the transpose of the forward
// operand, again pre-swizzled and stored in consumption order (8 chunks per GEMM, same pattern as a trunk layer).
constexpr int kDgGemms = 8;
constexpr int kDgChunks = kDgGemms * 8;                        // 64 x 16 KB = 1 MiB
constexpr ChunkTable make_dgrad_table()
{
    ChunkTable t{};
    int n = 0;
    for (int g = 0; g < kDgGemms; ++g) {
        const uint8_t G = (uint8_t)g;
        const int begin = n;
        t.c[n++] = ChunkInfo{G, 0, 0, 0};
        t.c[n++] = ChunkInfo{G, 0, 1, 0};
        t.c[n++] = ChunkInfo{G, 1, 0, 0};
        t.c[n++] = ChunkInfo{G, 0, 2, 0};
        t.c[n++] = ChunkInfo{G, 0, 3, 0};
        t.c[n++] = ChunkInfo{G, 1, 1, 0};
        t.c[n++] = ChunkInfo{G, 1, 2, 0};
        t.c[n++] = ChunkInfo{G, 1, 3, 0};
        for (int h = 0; h < 2; ++h) {
            int first = -1, last = -1;
            for (int i = begin; i < n; ++i)
                if (t.c[i].half == h) { if (first < 0) first = i; last = i; }
            t.c[first].flags |= 1;
            t.c[last].flags |= 2;
        }
        for (int kb = 0; kb < 4; ++kb)
            for (int i = begin; i < n; ++i)
                if (t.c[i].asrc == kb) { t.c[i].flags |= 4; break; }
    }
    return t;
}

constexpr size_t B_OFFSET = ((F_END * 4 + 1023) / 1024) * 1024;   // byte offset of the bf16 region
constexpr size_t B_BYTES = chunk_offset(kChunksPerTile);            // ~1 MiB
// low-order bf16 stream for a split-precision mode follows (same layout)
constexpr size_t B_LO_OFFSET = B_OFFSET + B_BYTES;
constexpr size_t B_DG_OFFSET = ((B_LO_OFFSET + B_BYTES + 1023) / 1024) * 1024;   // dgrad stream
constexpr size_t B_DG_BYTES = (size_t)kDgChunks * kChunkBytes;
constexpr size_t PACKED_BYTES = B_DG_OFFSET + B_DG_BYTES;

// byte offset of element (n, k) inside one swizzled chunk
__host__ __device__ constexpr uint32_t swz128(uint32_t n, uint32_t k)
{
    return n * 128u + ((((k >> 3) ^ (n & 7u)) << 4) | ((k & 7u) << 1));
}

}  // namespace nerfb200
