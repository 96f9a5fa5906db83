// Training step, BF16 mode: the dgrad chain on the tensor cores.
//
// The backward of the 8x256 trunk is the forward run in reverse -- eight GEMMs per 128-sample tile,
//   G0: dh7 = [dpre_c0 | dsigma_pre] . [W_c0[:, :256] ; w_sigma]        Gg: dh_{7-g} = dpre_{8-g} . W_{8-g}[:, :256]
// -- so it reuses the fused kernel's machinery unchanged: N = 128 halves, activations (here: gradients) resident in
// TMEM and rewritten in place as the next GEMM's bf16 A operand, two alternating 256-column regions, the chunk order
// that hides the epilogue latency, a straight-line compile-time MMA schedule, and the consumption-ordered weight
// stream (packed_layout.h "dgrad stream") fed by cp.async.bulk through a 4-slot mbarrier ring.
//   * front warps (12-15): per row, dpre_c0 = (dy . W_c1) * [c0h > 0] and dsigma_pre -> stored for wgrad and written
//     to TMEM as G0's A operand (K-blocks 0,1: colour, 2: density, 3: zeros);
//   * epilogue warps (4-11): dpre = dh * [h > 0] with the mask read from the stored forward activations; dpre is
//     stored K-major ([feature][sample], the wgrad operand) and written back to TMEM in place;
//   * no shared-memory operand at all besides the weights.
// reference: autograd through NeRFModel.forward (src/models/nerf.py:105-129) in NeRFTrainer.train_step
// (src/training/trainer.py:125-126); math: SURVEY Appendix B.
#include <utility>
#include "common.cuh"
#include "ptx.cuh"
#include "train_layout.h"

namespace nerfb200 {
namespace dg {

using namespace ptx;

constexpr int kThreads = 512;
constexpr int kSlots = 4;
constexpr uint32_t kSlotBytes = 2 * kChunkBytes;              // 32 KB stages (2 chunks)
constexpr int kStages = kDgChunks / 2;                        // 32 per tile
static_assert(kStages % kSlots == 0, "compile-time ring parities");
constexpr uint32_t SM_W = 0;
constexpr uint32_t SM_WC1 = kSlots * kSlotBytes;              // [3][128] f32
constexpr uint32_t SM_BAR = SM_WC1 + 1536;
constexpr uint32_t SM_TMEM = SM_BAR + 256;
constexpr int kStageBufs = 2;                                     // 4 KB store staging buffers per epilogue warp
constexpr uint32_t SM_STAGE = SM_TMEM + 256;
constexpr uint32_t kSmem = SM_STAGE + 8 * kStageBufs * 4096 + 1024;
static_assert(SM_STAGE % 128 == 0 && kSmem <= 232448, "staging buffers: 128-byte aligned, within the shared memory of an SM");

enum { B_WFULL = 0, B_WEMPTY = 4, B_ACCFULL = 8, B_AREADY = 10, B_R1FREE = 14, B_COUNT = 15 };
// a_ready[kb]: phase g of a tile = K-block kb of GEMM g's A operand is in TMEM (g = 0: written by the front warps,
// g >= 1: by the epilogue of GEMM g-1); 8 phases per tile.  acc_full[h]: 8 phases per tile.  r1_free: the epilogue
// of G7 has read region 1, the front warps may write the next tile's G0 operand there.

constexpr ChunkTable kDg = make_dgrad_table();

struct Args {
    const unsigned char *packed;
    float *ws;
    int ch, n_samples_total;              // workspace pitch, valid samples (columns) in this chunk
    int n_tiles, tiles_per_cta;
    unsigned int *dbg;
};

__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity, unsigned int *dbg, uint32_t code)
{
    mbar_wait_bounded(bar, parity, dbg, 0x90000000u, code);
}

struct IssueCtx { uint32_t bars; uint32_t region[2]; uint64_t wdesc; unsigned int *dbg; };

template <int CI>
__device__ __forceinline__ void issue_chunk(const IssueCtx &x)
{
    constexpr ChunkInfo c = kDg.c[CI];
    constexpr int stage = CI / 2, slot = stage % kSlots;
    constexpr uint32_t idesc = idesc_bf16(128, 128);
    if constexpr (CI % 2 == 0) wait_bar(x.bars + 8u * (B_WFULL + slot), (stage / kSlots) & 1, x.dbg, 4);
    if constexpr ((c.flags & 4) != 0) wait_bar(x.bars + 8u * (B_AREADY + c.asrc), c.layer & 1, x.dbg, 3);
    tc_fence_after_sync();
    if (elect_one()) {
        const uint32_t d_tmem = x.region[c.layer & 1] + c.half * 128;
        const uint32_t a_tmem = x.region[(c.layer & 1) ^ 1] + c.asrc * 64;
        const uint64_t bdesc = x.wdesc + (uint64_t)((slot * kSlotBytes + (CI % 2) * kChunkBytes) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            mma_bf16_ts(d_tmem, a_tmem + 8 * k, bdesc + 2 * k, idesc, !((c.flags & 1) && k == 0));
        if constexpr ((c.flags & 2) != 0) mma_commit(x.bars + 8u * (B_ACCFULL + c.half));
        if constexpr (CI % 2 == 1) mma_commit(x.bars + 8u * (B_WEMPTY + slot));
    }
    __syncwarp();
}
template <int... CI>
__device__ __forceinline__ void issue_tile(const IssueCtx &x, std::integer_sequence<int, CI...>) { (issue_chunk<CI>(x), ...); }

__global__ void __launch_bounds__(kThreads, 1) dgrad_chain_kernel(const Args a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_base = smem_u32(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bars = sm_base + SM_BAR;
    auto bar = [&](int i) { return bars + 8u * i; };
    const float *wf = reinterpret_cast<const float *>(a.packed);
    const unsigned char *wb = a.packed + B_DG_OFFSET;
    const int tile_begin = blockIdx.x * a.tiles_per_cta;
    const int my_tiles = max(0, min(a.n_tiles, tile_begin + a.tiles_per_cta) - tile_begin);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) mbar_init(bar(B_ACCFULL + i), 1);
        for (int i = 0; i < 4; ++i) mbar_init(bar(B_AREADY + i), 4);
        mbar_init(bar(B_R1FREE), 8);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(sm_base + SM_TMEM);
    {
        float *wc1 = reinterpret_cast<float *>(sm + SM_WC1);
        for (int i = threadIdx.x; i < 384; i += kThreads) wc1[i] = __ldg(wf + F_WC1 + i);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + SM_TMEM);
    const size_t ch = (size_t)a.ch;

    if (warp == 0) {
        // ---------------- weight producer
        if (lane == 0) {
            uint32_t sg = 0;
            for (int t = 0; t < my_tiles; ++t)
                for (int st = 0; st < kStages; ++st, ++sg) {
                    const uint32_t slot = sg % kSlots, round = sg / kSlots;
                    if (round > 0) wait_bar(bar(B_WEMPTY + slot), (round - 1) & 1, a.dbg, 1);
                    mbar_arrive_expect_tx(bar(B_WFULL + slot), kSlotBytes);
                    bulk_g2s(sm_base + SM_W + slot * kSlotBytes, wb + (size_t)st * kSlotBytes, kSlotBytes, bar(B_WFULL + slot));
                }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (eight GEMMs per tile: the region parity does not flip across tiles)
        IssueCtx x;
        x.bars = bars;
        x.region[0] = tmem_base;
        x.region[1] = tmem_base + 256;
        x.wdesc = smem_desc_sw128(sm_base + SM_W);
        x.dbg = a.dbg;
        for (int t = 0; t < my_tiles; ++t) issue_tile(x, std::make_integer_sequence<int, kDgChunks>{});
    } else if (warp >= 4 && warp < 12) {
        // ---------------- epilogue: dpre = dh * [h > 0]; store for wgrad; write back bf16 in place
        const int ew = warp - 4, q = ew & 3, w2 = ew >> 2;
        const int row = q * 32 + lane;
        for (int t = 0; t < my_tiles; ++t) {
            const int col = (tile_begin + t) * 128 + row;
            const bool on = col < a.n_samples_total;
            int col_r[4];                                                  // the rows this lane writes back (staged stores)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = (tile_begin + t) * 128 + q * 32 + (lane >> 2) + 8 * j;
                col_r[j] = c < a.n_samples_total ? c : -1;
            }
            // a warp whose 32 samples all exist (every warp but those of the last tile) stores through the TMA engine
            const int col0 = (tile_begin + t) * 128 + q * 32;
            const bool whole = col0 + 32 <= a.n_samples_total;
            uint32_t sbuf = 0;
            for (int g = 0; g < kDgGemms; ++g) {
                const int layer = 7 - g;                                   // dh of trunk layer `layer`
                const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (g & 1) * 256 + 64 * w2;
                // the masks do not depend on the chain: fetch both halves' words before waiting for the accumulator
                const unsigned long long *mrow = reinterpret_cast<const unsigned long long *>(a.ws + (size_t)R_MASK * ch) +
                                                 (size_t)(layer * 4 + w2) * ch + col;
                unsigned long long mbits[2] = {0ull, 0ull};
                if (on) { mbits[0] = mrow[0]; mbits[1] = mrow[2 * ch]; }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    wait_bar(bar(B_ACCFULL + hh), g & 1, a.dbg, 5);
                    tc_fence_after_sync();
                    const uint32_t t_cols = t_lane + hh * 128;
                    const int n0 = hh * 128 + 64 * w2;
                    uint32_t xa[32], xb[32], pk[32];
                    tmem_ld32(t_cols, xa);
                    tmem_ld32(t_cols + 32, xb);
                    tmem_ld_wait();
                    const uint32_t m_lo = on ? (uint32_t)mbits[hh] : 0u, m_hi = on ? (uint32_t)(mbits[hh] >> 32) : 0u;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        // bf16 pairs: the values wgrad multiplies are exactly the ones the next GEMM of the chain sees;
                        // dpre = dh where the forward's activation was > 0 (train_layout.h: relu_mask_word)
                        pk[i] = pack_bf16(__uint_as_float(xa[2 * i]), __uint_as_float(xa[2 * i + 1])) & relu_pair_mask(m_lo, i);
                        pk[16 + i] = pack_bf16(__uint_as_float(xb[2 * i]), __uint_as_float(xb[2 * i + 1])) & relu_pair_mask(m_hi, i);
                    }
                    if (g < kDgGemms - 1) {                                // hand the operand to the next GEMM first
                        tmem_st32(t_cols, pk);
                        tmem_st_wait();
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar(B_AREADY + 2 * hh + w2));
                    }
                    if (whole) {
                        store_block_rows_bulk<kStageBufs - 1>(a.ws, G_DPRE + layer * 256 + n0, col0, pk,
                                                              sm_base + SM_STAGE + (ew * kStageBufs + sbuf) * 4096, lane);
                        sbuf = (sbuf + 1u) % kStageBufs;
                    } else {
                        if (lane == 0) bulk_store_reads_done();            // an earlier tile's copies may still read the buffer
                        __syncwarp();
                        store_block_rows_staged(a.ws, G_DPRE + layer * 256 + n0, col_r, pk, sm_base + SM_STAGE + ew * kStageBufs * 4096, lane);
                    }
                }
            }
            // region 1 (G7's accumulator) has been read: the next tile's G0 operand may be written there
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_R1FREE));
        }
        if (lane == 0) bulk_store_drain();
    } else if (warp >= 12) {
        // ---------------- front: G0's A operand
        const int row = (warp - 12) * 32 + lane;
        const uint32_t wc1 = sm_base + SM_WC1;
        const uint32_t t_row = tmem_base + ((uint32_t)((warp - 12) * 32) << 16) + 256;       // region 1
        for (int t = 0; t < my_tiles; ++t) {
            const int col = (tile_begin + t) * 128 + row;
            const bool on = col < a.n_samples_total;
            float dy[3] = {0.f, 0.f, 0.f}, dsig = 0.f;
            unsigned long long mc0[2] = {0ull, 0ull};
            if (on) {
#pragma unroll
                for (int c = 0; c < 3; ++c) dy[c] = a.ws[(size_t)(R_DY + c) * ch + col];
                dsig = a.ws[(size_t)R_DSIG * ch + col];
                const unsigned long long *mp = reinterpret_cast<const unsigned long long *>(a.ws + (size_t)R_MASKC0 * ch);
                mc0[0] = mp[col]; mc0[1] = mp[ch + col];
            }
            if (t > 0) wait_bar(bar(B_R1FREE), (t - 1) & 1, a.dbg, 8);
            tc_fence_after_sync();
#pragma unroll 1
            for (int kb = 0; kb < 2; ++kb) {
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 w0 = ld_shared_f4(wc1 + (64 * kb + 4 * i) * 4);
                    const float4 w1 = ld_shared_f4(wc1 + 512 + (64 * kb + 4 * i) * 4);
                    const float4 w2 = ld_shared_f4(wc1 + 1024 + (64 * kb + 4 * i) * 4);
                    float v[4] = {dy[0] * w0.x + dy[1] * w1.x + dy[2] * w2.x, dy[0] * w0.y + dy[1] * w1.y + dy[2] * w2.y,
                                  dy[0] * w0.z + dy[1] * w1.z + dy[2] * w2.z, dy[0] * w0.w + dy[1] * w1.w + dy[2] * w2.w};
                    const uint32_t mw = (2 * i) >> 4 ? (uint32_t)(mc0[kb] >> 32) : (uint32_t)mc0[kb];
                    pk[2 * i] = pack_bf16(v[0], v[1]) & relu_pair_mask(mw, (2 * i) & 15);
                    pk[2 * i + 1] = pack_bf16(v[2], v[3]) & relu_pair_mask(mw, (2 * i + 1) & 15);
                }
                if (on) store_block_row(a.ws, G_DPREC0 + 64 * kb, col, pk);
                tmem_st32(t_row + 64 * kb, pk);
            }
            {
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) pk[i] = 0u;
                tmem_st32(t_row + 192, pk);                               // K-block 3: zeros
                pk[0] = pack_bf16(dsig, 0.f);
                tmem_st32(t_row + 128, pk);                               // K-block 2: the density-head gradient
            }
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int kb = 0; kb < 4; ++kb) mbar_arrive(bar(B_AREADY + kb));
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) { tc_fence_after_sync(); tmem_dealloc<512>(tmem_base); }
}

}  // namespace dg

// dgrad chain over the chunk's samples (columns [0, n_samples) of the workspace; `ch` = pitch, multiple of 64)
int dgrad_chain_tc(const void *packed, float *ws, int ch, int n_samples, unsigned int *dbg, int sm_limit, cudaStream_t stream)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sm_limit > 0 && sm_limit < sms) sms = sm_limit;
    dg::Args a = {};
    a.packed = reinterpret_cast<const unsigned char *>(packed);
    a.ws = ws; a.ch = ch; a.n_samples_total = n_samples; a.dbg = dbg;
    a.n_tiles = (n_samples + 127) / 128;
    a.tiles_per_cta = (a.n_tiles + sms - 1) / sms;
    const int grid = (a.n_tiles + a.tiles_per_cta - 1) / a.tiles_per_cta;
    cudaError_t e = cudaFuncSetAttribute(dg::dgrad_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dg::kSmem);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    dg::dgrad_chain_kernel<<<grid, dg::kThreads, dg::kSmem, stream>>>(a);
    return launch_status();
}

}  // namespace nerfb200
