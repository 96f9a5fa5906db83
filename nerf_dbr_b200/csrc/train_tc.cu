// Training step, BF16 mode: tensor-core kernels.
//
// wgrad_tc_kernel -- dW[n][k] = sum_s dY[n][s] X[k][s] for one parameter tensor, as a tcgen05 GEMM with
// M = n (128 or 256 output features), N = k (up to 256 input features), K = samples.  Both operands sit in
// the training workspace as bf16 -- exactly the values the forward and the dgrad chain multiplied -- in slab-major,
// sample-major, pre-swizzled blocks (train_layout.h), so the operand tile of a 64-sample slab is one contiguous
// run of 8 KB blocks that is already the 128B-swizzled MN-major shared-memory image: one thread stages it with
// two cp.async.bulk copies (A: rows_a/64 blocks, B: rows_b/64 blocks, three stages in flight -- the kernel is
// HBM-bound) and the MMAs take both operands MN-major.  The fp32 accumulators of
// the whole 256 x 256 tensor fill the 512 TMEM columns.  ONE launch covers every tensor of a chunk (a job table in the
// kernel parameters; CTAs apportioned to jobs by operand bytes); a job's slab range is split over its CTAs, each
// writes its partial to scratch and ONE wgrad_reduce_kernel launch folds the partials of all jobs into the caller's
// gradient tensors.  Bias gradients (row sums of A) are taken from the staged tiles in shared memory by the
// otherwise idle epilogue warps.
//
// reference: the autograd backward of the ten nn.Linear layers in NeRFModel (src/models/nerf.py:72-90)
// inside NeRFTrainer.train_step (src/training/trainer.py:125-126).
#include "common.cuh"
#include "ptx.cuh"
#include "train_layout.h"
#include <algorithm>

namespace nerfb200 {
namespace wg {

using namespace ptx;

constexpr int kThreads = 384;                    // warp 0 producer, 1 MMA issuer, 2 TMEM owner, 4..11 bias sums + epilogue
constexpr int kStages = 3;
constexpr uint32_t kTileA = 256 * 128;           // [256 x 64] bf16
constexpr uint32_t kTileB = 256 * 128;
constexpr uint32_t kStage = kTileA + kTileB;     // 64 KB
constexpr uint32_t SM_BAR = kStages * kStage;
constexpr uint32_t SM_TMEM = SM_BAR + 64;
constexpr uint32_t kSmem = SM_TMEM + 16 + 1024;

enum { B_FULL = 0, B_EMPTY = kStages, B_DONE = 2 * kStages };

constexpr int kMaxJobs = 16;

// One launch covers every tensor of a chunk: CTA b works for job j with cta_begin[j] <= b < cta_begin[j + 1] on the
// (b - cta_begin[j])-th part of the slab range.  CTAs are apportioned to jobs by operand bytes (host side).
struct Job {
    int row_a, rows_a;                           // A = features [row_a, row_a + rows_a) (G_* numbering), rows_a = 128 or 256
    int row_b, rows_b;                           // B features, staged in whole blocks of 64, <= 256
    int n_b;                                     // MMA N: the valid B features rounded up to 16 (pad features hold zeros)
    float *partial;                              // [splits][rows_a][n_b]
    float *bias_partial;                         // optional: [splits][rows_a] row sums of A over this CTA's slabs
};
struct Args {
    const __nv_bfloat16 *ws;                     // bf16 operand blocks
    int n_slabs;                                 // samples / 64
    int n_jobs;
    int cta_begin[kMaxJobs + 1];
    unsigned int *dbg;                           // optional watchdog word (ptx.cuh: mbar_wait_bounded)
    Job job[kMaxJobs];
};
struct ReduceArgs {                              // block b reduces 64 elements of job j, blk_begin[j] <= b < blk_begin[j + 1]
    int n_jobs;
    int blk_begin[kMaxJobs + 1];
    const float *partial[kMaxJobs];
    int splits[kMaxJobs], rows_a[kMaxJobs], n_b[kMaxJobs], rows_b_valid[kMaxJobs];
    float *dW[kMaxJobs];
    const float *bias_partial[kMaxJobs];         // [splits][rows_a] or nullptr
    float *dbias[kMaxJobs];
    int bias_blk[kMaxJobs];                      // first of the job's bias blocks (one per 256 rows), after its element blocks
    int ld[kMaxJobs], col_off[kMaxJobs];
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ Args args)
{
    int j = 0;
    while (j + 1 < args.n_jobs && (int)blockIdx.x >= args.cta_begin[j + 1]) ++j;
    const Job &a = args.job[j];
    const int split = blockIdx.x - args.cta_begin[j], splits = args.cta_begin[j + 1] - args.cta_begin[j];
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_base = smem_u32(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bar = [&](int i) { return sm_base + SM_BAR + 8u * i; };
    auto wait = [&](uint32_t b, uint32_t parity) { mbar_wait_bounded(b, parity, args.dbg, 0xA0000000u, 0u); };
    const int per = (args.n_slabs + splits - 1) / splits;
    const int c_begin = split * per, c_end = min(args.n_slabs, c_begin + per);
    const int my_slabs = max(0, c_end - c_begin);
    const bool want_bias = a.bias_partial != nullptr;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(bar(B_FULL + i), 1); mbar_init(bar(B_EMPTY + i), want_bias ? 9 : 1); }
        mbar_init(bar(B_DONE), 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(sm_base + SM_TMEM);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + SM_TMEM);
    const int m_blocks = a.rows_a / 128;

    if (warp == 0) {
        // ---------------- producer: two bulk copies per slab
        if (lane == 0) {
            const uint32_t bytes_a = (uint32_t)a.rows_a * 128u, bytes_b = (uint32_t)a.rows_b * 128u;
            for (int c = 0; c < my_slabs; ++c) {
                const int s = c % kStages, use = c / kStages;
                if (use > 0) wait(bar(B_EMPTY + s), (use - 1) & 1);
                mbar_arrive_expect_tx(bar(B_FULL + s), bytes_a + bytes_b);
                bulk_g2s(sm_base + s * kStage, args.ws + big_tile(a.row_a, c_begin + c), bytes_a, bar(B_FULL + s));
                bulk_g2s(sm_base + s * kStage + kTileA, args.ws + big_tile(a.row_b, c_begin + c), bytes_b, bar(B_FULL + s));
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer
        const uint32_t idesc = idesc_bf16_mn(128, (uint32_t)a.n_b);
        for (int c = 0; c < my_slabs; ++c) {
            const int s = c % kStages, use = c / kStages;
            wait(bar(B_FULL + s), use & 1);
            tc_fence_after_sync();
            if (elect_one()) {
                const uint64_t adesc = smem_desc_sw128_mn(sm_base + s * kStage, 8192);
                const uint64_t bdesc = smem_desc_sw128_mn(sm_base + s * kStage + kTileA, 8192);
                for (int mb = 0; mb < m_blocks; ++mb)         // 128 output features = two blocks; 16 samples = 2048 B
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_bf16_ss(tmem_base + mb * 256, adesc + (uint64_t)((mb * 16384 + k * 2048) >> 4),
                                    bdesc + (uint64_t)((k * 2048) >> 4), idesc, (c | k) != 0);
                mma_commit(bar(B_EMPTY + s));
                if (c == my_slabs - 1) mma_commit(bar(B_DONE));
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;                       // 8 warps
        // ---------------- bias gradients: thread t sums feature t of every staged A tile (a warp reads 32 consecutive
        // features of one sample per step: conflict-free)
        if (want_bias) {
            const int t = ew * 32 + lane;
            float bs = 0.f;
            for (int c = 0; c < my_slabs; ++c) {
                const int s = c % kStages, use = c / kStages;
                wait(bar(B_FULL + s), use & 1);
                if (t < a.rows_a) {
                    const unsigned short *blk = reinterpret_cast<const unsigned short *>(sm + s * kStage + (t >> 6) * 8192);
                    const int u = (t & 63) >> 3, e = t & 7;
#pragma unroll 16
                    for (int k = 0; k < 64; ++k)
                        bs += __uint_as_float((uint32_t)blk[k * 64 + ((u ^ (k & 7)) << 3) + e] << 16);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_EMPTY + s));
            }
            if (t < a.rows_a) a.bias_partial[(size_t)split * a.rows_a + t] = bs;      // folded in CTA order by the reduce kernel
        }
        // ---------------- epilogue: accumulators -> this CTA's partial
        if (my_slabs > 0) {
            wait(bar(B_DONE), 0);
            tc_fence_after_sync();
        }
        if (ew < 4 * m_blocks) {
            const int mb = ew >> 2, q = warp & 3;      // TMEM lane quadrant = warp % 4
            const int n = mb * 128 + q * 32 + lane;
            float *dst = a.partial + ((size_t)split * a.rows_a + n) * a.n_b;
            for (int col = 0; col < a.n_b; col += 32) {
                uint32_t v[32];
                if (my_slabs > 0) {
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + mb * 256 + col, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                }
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    if (col + i < a.n_b)
                        *reinterpret_cast<uint4 *>(dst + col + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) { tc_fence_after_sync(); tmem_dealloc<512>(tmem_base); }
}

// grads[n][col_off + k] += sum_split partial[split][n][k]   (k < rows_b_valid), dbias[n] += sum_split bias_partial[split][n]:
// every job of the chunk in one launch, every sum in split order (no atomics: the result does not depend on timing).
// A block owns 256 consecutive elements of a job's [rows_a][n_b] partial as 64 float4 (n_b is a multiple of 16, so a
// float4 never straddles rows); its four thread groups each take every fourth split with four 16-byte loads in flight
// (the partials were just written: they sit in L2, latency is the cost).  A job's bias blocks follow its element blocks.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ ReduceArgs r)
{
    __shared__ float4 part[4][64];
    int j = 0;
    while (j + 1 < r.n_jobs && (int)blockIdx.x >= r.blk_begin[j + 1]) ++j;
    const int rows_a = r.rows_a[j], rows_b = r.n_b[j], splits = r.splits[j];
    if (r.dbias[j] && (int)blockIdx.x >= r.bias_blk[j]) {
        const int n = ((int)blockIdx.x - r.bias_blk[j]) * 256 + threadIdx.x;
        if (n < rows_a) {
            const float *p = r.bias_partial[j] + n;
            float s = 0.f;
#pragma unroll 4
            for (int sp = 0; sp < splits; ++sp) s += __ldg(p + (size_t)sp * rows_a);
            r.dbias[j][n] += s;
        }
        return;
    }
    const int o = threadIdx.x & 63, sg = threadIdx.x >> 6;
    const int i = ((blockIdx.x - r.blk_begin[j]) * 64 + o) * 4, total = rows_a * rows_b;
    const size_t stride = (size_t)rows_a * rows_b;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
    auto acc = [](float4 &d, const float4 v) { d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w; };
    if (i < total) {
        const float *p = r.partial[j] + i;
        int sp = sg;
        for (; sp + 12 < splits; sp += 16) {
            const float4 v0 = __ldg(reinterpret_cast<const float4 *>(p + (size_t)sp * stride));
            const float4 v1 = __ldg(reinterpret_cast<const float4 *>(p + (size_t)(sp + 4) * stride));
            const float4 v2 = __ldg(reinterpret_cast<const float4 *>(p + (size_t)(sp + 8) * stride));
            const float4 v3 = __ldg(reinterpret_cast<const float4 *>(p + (size_t)(sp + 12) * stride));
            acc(s0, v0); acc(s1, v1); acc(s2, v2); acc(s3, v3);
        }
        for (; sp < splits; sp += 4) acc(s0, __ldg(reinterpret_cast<const float4 *>(p + (size_t)sp * stride)));
    }
    part[sg][o] = make_float4((s0.x + s1.x) + (s2.x + s3.x), (s0.y + s1.y) + (s2.y + s3.y), (s0.z + s1.z) + (s2.z + s3.z),
                              (s0.w + s1.w) + (s2.w + s3.w));
    __syncthreads();
    if (sg == 0 && i < total) {
        const int n = i / rows_b, k = i % rows_b;
        const float4 a0 = part[0][o], a1 = part[1][o], a2 = part[2][o], a3 = part[3][o];
        const float v[4] = {(a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z),
                            (a0.w + a1.w) + (a2.w + a3.w)};
        float *dst = r.dW[j] + (size_t)n * r.ld[j] + r.col_off[j] + k;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (k + e < r.rows_b_valid[j]) dst[e] += v[e];
    }
}

// Skinny weight gradients (density head: 1 output row; colour layer 1: 3): dW[a][k] += sum_s A[a][s] B[k][s],
// dbias[a] += sum_s A[a][s].  A is fp32 [row][ch]; B is a group of bf16 operand blocks (train_layout.h).  A block
// walks slabs: it stages the slab's contiguous B tile (coalesced 16-byte loads) and the 64 A values per row in
// shared memory, thread k takes the dot products of B feature k; the block's totals go to its row of `partial`
// ([block][rows_a][rows_b + 1], last column = bias sums) and skinny_reduce_kernel folds the rows in block order.
__global__ void __launch_bounds__(256) wgrad_skinny_kernel(const float *__restrict__ A, int rows_a, int ch,
                                                           const __nv_bfloat16 *__restrict__ ws, int row_b, int rows_b,
                                                           float *__restrict__ partial)
{
    __shared__ __align__(16) uint4 tile[256 * 8];          // [block][sample][64 features], swizzled as stored
    __shared__ __align__(16) float as[4][64];
    const int k = threadIdx.x, n_slabs = ch / 64;
    const unsigned short *blk = reinterpret_cast<const unsigned short *>(tile) + (k >> 6) * 4096;
    const int u = (k & 63) >> 3, e = k & 7;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, bs = 0.f;
    for (int slab = blockIdx.x; slab < n_slabs; slab += gridDim.x) {
        const uint4 *src = reinterpret_cast<const uint4 *>(ws + big_tile(row_b, slab));
        for (int i = threadIdx.x; i < rows_b * 8; i += 256) tile[i] = __ldg(src + i);
        if (threadIdx.x < 16 * rows_a) {
            const int r = threadIdx.x >> 4, j = threadIdx.x & 15;
            reinterpret_cast<float4 *>(as[r])[j] = __ldg(reinterpret_cast<const float4 *>(A + (size_t)r * ch + (size_t)slab * 64) + j);
        }
        __syncthreads();
        if (k < rows_b) {
#pragma unroll 8
            for (int s = 0; s < 64; ++s) {
                const float b = __uint_as_float((uint32_t)blk[s * 64 + ((u ^ (s & 7)) << 3) + e] << 16);
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (r < rows_a) acc[r] = fmaf(as[r][s], b, acc[r]);
            }
        }
        if (threadIdx.x >= 224 && threadIdx.x - 224 < rows_a) {   // bias: the last warp's first lanes
            const float *ar = as[threadIdx.x - 224];
#pragma unroll 8
            for (int j = 0; j < 64; ++j) bs += ar[j];
        }
        __syncthreads();
    }
    float *row = partial + (size_t)blockIdx.x * rows_a * (rows_b + 1);
    if (k < rows_b)
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r < rows_a) row[r * (rows_b + 1) + k] = acc[r];
    if (threadIdx.x >= 224 && threadIdx.x - 224 < rows_a) row[(threadIdx.x - 224) * (rows_b + 1) + rows_b] = bs;
}

// dW[a][k] += sum_block partial[block][a][k], dbias[a] += sum_block partial[block][a][rows_b].  One WARP per element:
// lane l adds blocks l, l + 32, ... (independent loads), then a fixed shuffle tree -- the order never depends on timing.
__global__ void __launch_bounds__(256) skinny_reduce_kernel(const float *__restrict__ partial, int blocks, int rows_a, int rows_b,
                                                            float *__restrict__ dW, int ld, float *__restrict__ dbias)
{
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, width = rows_b + 1, total = rows_a * width;
    if (i >= total) return;
    float s = 0.f;
#pragma unroll 4
    for (int b = lane; b < blocks; b += 32) s += __ldg(partial + (size_t)b * total + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane != 0) return;
    const int r = i / width, k = i % width;
    if (k < rows_b) dW[(size_t)r * ld + k] += s;
    else if (dbias) dbias[r] += s;
}

}  // namespace wg

constexpr int kSkinnyMaxBlocks = 1024;
size_t wgrad_skinny_scratch_bytes() { return (size_t)kSkinnyMaxBlocks * 4 * 257 * sizeof(float); }

int wgrad_skinny(const float *A, int rows_a, int ch, const __nv_bfloat16 *ws, int row_b, int rows_b, float *dW, int ld,
                 float *dbias, float *scratch, int sm_limit, cudaStream_t stream)
{
    if (rows_a > 4 || rows_b > 256 || (row_b & 63) || !scratch) return NERF_B200_EINVAL;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sm_limit > 0 && sm_limit < sms) sms = sm_limit;
    const int grid = std::max(1, std::min(std::min(ch / 64, 4 * sms), kSkinnyMaxBlocks));
    wg::wgrad_skinny_kernel<<<grid, 256, 0, stream>>>(A, rows_a, ch, ws, row_b, rows_b, scratch);
    int rc = launch_status();
    if (rc) return rc;
    wg::skinny_reduce_kernel<<<(rows_a * (rows_b + 1) + 7) / 8, 256, 0, stream>>>(scratch, grid, rows_a, rows_b, dW, ld, dbias);
    return launch_status();
}

size_t wgrad_tc_scratch_bytes(int ctas) { return (size_t)(ctas + wg::kMaxJobs) * (256 * 256 + 256) * sizeof(float); }   // element + bias partials

// dW (+)= A-by-B^T over the chunk's samples on the tensor cores for every job in one launch, then one reduce launch.
// A and B are feature groups (G_* numbering, starting on a block boundary) of the bf16 operand blocks at `ws`;
// `ctas` CTAs are apportioned to the jobs by operand bytes (largest remainder, at least one each).
int wgrad_tc_batch(const __nv_bfloat16 *ws, int ch, const WgradJob *jobs, int n_jobs, float *scratch, int ctas, cudaStream_t stream)
{
    if (n_jobs < 1 || n_jobs > wg::kMaxJobs) return NERF_B200_EINVAL;
    wg::Args a = {};
    wg::ReduceArgs r = {};
    a.ws = ws; a.n_slabs = ch / 64; a.n_jobs = r.n_jobs = n_jobs;
    a.dbg = watchdog_word();
    ctas = std::max(ctas, n_jobs);
    int weight[wg::kMaxJobs], share[wg::kMaxJobs], total_w = 0, given = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const WgradJob &q = jobs[j];
        if ((q.row_a | q.row_b | q.rows_a) & 63) return NERF_B200_EINVAL;
        weight[j] = q.rows_a + (q.rows_b_valid + 63) / 64 * 64;
        total_w += weight[j];
    }
    double frac[wg::kMaxJobs];
    for (int j = 0; j < n_jobs; ++j) {
        const double exact = (double)ctas * weight[j] / total_w;
        share[j] = std::max(1, std::min(a.n_slabs, (int)exact));
        frac[j] = exact - share[j];
        given += share[j];
    }
    while (given < ctas) {                                            // largest remainder first
        int best = -1;
        for (int j = 0; j < n_jobs; ++j)
            if (share[j] < a.n_slabs && (best < 0 || frac[j] > frac[best])) best = j;
        if (best < 0) break;
        ++share[best]; frac[best] -= 1.0; ++given;
    }
    size_t off = 0;
    int blk = 0;
    float *bias_base = scratch + (size_t)(ctas + wg::kMaxJobs) * 256 * 256;      // bias partials behind the element partials
    size_t boff = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const WgradJob &q = jobs[j];
        wg::Job &d = a.job[j];
        d.row_a = q.row_a; d.rows_a = q.rows_a; d.row_b = q.row_b;
        d.rows_b = (q.rows_b_valid + 63) / 64 * 64;
        d.n_b = (q.rows_b_valid + 15) / 16 * 16;
        d.partial = scratch + off;
        d.bias_partial = q.dbias ? bias_base + boff : nullptr;
        a.cta_begin[j] = j ? a.cta_begin[j - 1] + share[j - 1] : 0;
        r.blk_begin[j] = blk;
        r.partial[j] = d.partial; r.splits[j] = share[j]; r.rows_a[j] = q.rows_a; r.n_b[j] = d.n_b;
        r.rows_b_valid[j] = q.rows_b_valid; r.dW[j] = q.dW; r.ld[j] = q.ld; r.col_off[j] = q.col_off;
        r.bias_partial[j] = d.bias_partial; r.dbias[j] = q.dbias;
        off += (size_t)share[j] * q.rows_a * d.n_b;
        blk += (q.rows_a * d.n_b + 255) / 256;
        r.bias_blk[j] = blk;
        if (q.dbias) { boff += (size_t)share[j] * q.rows_a; blk += (q.rows_a + 255) / 256; }
    }
    a.cta_begin[n_jobs] = a.cta_begin[n_jobs - 1] + share[n_jobs - 1];
    r.blk_begin[n_jobs] = blk;
    cudaError_t e = cudaFuncSetAttribute(wg::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg::kSmem);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    wg::wgrad_tc_kernel<<<a.cta_begin[n_jobs], wg::kThreads, wg::kSmem, stream>>>(a);
    int rc = launch_status();
    if (rc) return rc;
    wg::wgrad_reduce_kernel<<<blk, 256, 0, stream>>>(r);
    return launch_status();
}

}  // namespace nerfb200
