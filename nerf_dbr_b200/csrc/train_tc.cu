// Training step, BF16 mode: tensor-core kernels.
//
// wgrad_tc_kernel -- dW[n][k] = sum_s dY[n][s] X[k][s] for one parameter tensor, as a tcgen05 GEMM with
// M = n (128 or 256 output features), N = k (up to 256 input features), K = samples.  Both operands sit in
// the training workspace K-major ([feature][sample], bf16 -- exactly the values the forward and the dgrad chain
// multiplied); loader warps copy 64-sample slabs straight into the 128B-swizzled K-major shared-memory tiles the MMA reads (A: [rows_a x 64],
// B: [rows_b x 64], two stages); the fp32 accumulators of the whole 256 x 256 tensor fill the 512 TMEM
// columns.  The sample range is split over CTAs; each writes its partial to scratch and
// wgrad_reduce_kernel folds the partials into the caller's gradient tensor (no atomics).
//
// reference: the autograd backward of the ten nn.Linear layers in NeRFModel (src/models/nerf.py:72-90)
// inside NeRFTrainer.train_step (src/training/trainer.py:125-126).
#include "common.cuh"
#include "ptx.cuh"

namespace nerfb200 {
namespace wg {

using namespace ptx;

constexpr int kThreads = 512;
constexpr int kLoaderWarps = 12;                 // warps 4..15
constexpr uint32_t kTileA = 256 * 128;           // [256 x 64] bf16
constexpr uint32_t kTileB = 256 * 128;
constexpr uint32_t kStage = kTileA + kTileB;     // 64 KB
constexpr uint32_t SM_BAR = 2 * kStage;
constexpr uint32_t SM_TMEM = SM_BAR + 64;
constexpr uint32_t kSmem = SM_TMEM + 16 + 1024;

enum { B_FULL = 0, B_EMPTY = 2, B_DONE = 4 };

struct Args {
    const __nv_bfloat16 *A; int rows_a;          // 128 or 256
    const __nv_bfloat16 *B; int rows_b;                  // padded to a multiple of 16, <= 256; rows >= rows_b_valid read as 0
    int rows_b_valid;
    int ch;                                      // samples (multiple of 64)
    float *partial;                              // [gridDim.x][rows_a][rows_b]
    float *dbias;                                // optional: dbias[n] += sum_s A[n][s] (fp32, from the loader's registers)
};

__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity)
{
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000LL) __trap();
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const Args a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_base = smem_u32(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bar = [&](int i) { return sm_base + SM_BAR + 8u * i; };
    const int n_chunks = a.ch / 64;
    const int per = (n_chunks + gridDim.x - 1) / gridDim.x;
    const int c_begin = blockIdx.x * per, c_end = min(n_chunks, c_begin + per);
    const int my_chunks = max(0, c_end - c_begin);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(bar(B_FULL + i), kLoaderWarps); mbar_init(bar(B_EMPTY + i), 1); }
        mbar_init(bar(B_DONE), 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(sm_base + SM_TMEM);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + SM_TMEM);
    const int m_blocks = a.rows_a / 128;

    if (warp == 1) {
        // ---------------- MMA issuer
        const uint32_t idesc = idesc_bf16(128, (uint32_t)a.rows_b);
        for (int c = 0; c < my_chunks; ++c) {
            const int s = c & 1;
            wait(bar(B_FULL + s), (c >> 1) & 1);
            tc_fence_after_sync();
            if (elect_one()) {
                const uint64_t adesc = smem_desc_sw128(sm_base + s * kStage);
                const uint64_t bdesc = smem_desc_sw128(sm_base + s * kStage + kTileA);
                for (int mb = 0; mb < m_blocks; ++mb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_bf16_ss(tmem_base + mb * 256, adesc + (uint64_t)((mb * 16384) >> 4) + 2 * k, bdesc + 2 * k, idesc,
                                    (c | k) != 0);
                mma_commit(bar(B_EMPTY + s));
                if (c == my_chunks - 1) mma_commit(bar(B_DONE));
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ---------------- loaders: fp32 [row][sample] -> bf16 swizzled K-major tiles
        const int lt = (warp - 4) * 32 + lane, n_lt = kLoaderWarps * 32;
        const int units_a = a.rows_a * 8, units_b = a.rows_b * 8;      // 16-byte units (8 samples) per 64-sample slab
        // a thread meets the same A rows in every slab (unit index -> row is slab-independent): bias partial sums
        // ride in registers, one per batch slot
        float bsum[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (int c = 0; c < my_chunks; ++c) {
            const int s = c & 1;
            if (c >= 2) wait(bar(B_EMPTY + s), ((c >> 1) - 1) & 1);
            const size_t s0 = (size_t)(c_begin + c) * 64;
            uint8_t *ta = sm + s * kStage, *tb = ta + kTileA;
            // batches of 4 units per thread: 8 independent 16-byte loads in flight before the first conversion
            // (the kernel is HBM-bound; a load-convert-store loop leaves the memory system idle)
            int batch = 0;
            for (int u0 = lt; u0 < units_a + units_b; u0 += 4 * n_lt, ++batch) {
                uint4 x[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int u = u0 + b * n_lt;
                    const bool is_a = u < units_a;
                    const int v = is_a ? u : u - units_a;
                    const int row = v >> 3, cu = v & 7;
                    x[b] = make_uint4(0u, 0u, 0u, 0u);
                    if (u < units_a + units_b && (is_a || row < a.rows_b_valid))
                        x[b] = __ldg(reinterpret_cast<const uint4 *>((is_a ? a.A : a.B) + (size_t)row * a.ch + s0 + cu * 8));
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int u = u0 + b * n_lt;
                    if (u >= units_a + units_b) continue;
                    const bool is_a = u < units_a;
                    const int v = is_a ? u : u - units_a;
                    const int row = v >> 3, cu = v & 7;
                    *reinterpret_cast<uint4 *>((is_a ? ta : tb) + row * 128 + ((cu ^ (row & 7)) << 4)) = x[b];
                    if (is_a && batch < 2) {
                        const uint32_t w[4] = {x[b].x, x[b].y, x[b].z, x[b].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) bsum[batch][b] += __uint_as_float(w[j] << 16) + __uint_as_float(w[j] & 0xffff0000u);
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_FULL + s));
        }
        if (a.dbias) {                                  // units_a <= 2048 = 2 batches of 4 x 384 threads (less 1024)
#pragma unroll
            for (int batch = 0; batch < 2; ++batch)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int u = lt + (batch * 4 + b) * n_lt;
                    if (u < units_a && bsum[batch][b] != 0.f) atomicAdd(a.dbias + (u >> 3), bsum[batch][b]);
                }
        }
        // ---------------- epilogue: accumulators -> this CTA's partial
        if (my_chunks > 0) {
            wait(bar(B_DONE), 0);
            tc_fence_after_sync();
        }
        const int ew = warp - 4;                       // 12 warps; warps 0..7 cover (m block, lane quadrant)
        if (ew < 4 * m_blocks) {
            const int mb = ew >> 2, q = warp & 3;      // TMEM lane quadrant = warp % 4
            const int n = mb * 128 + q * 32 + lane;
            float *dst = a.partial + ((size_t)blockIdx.x * a.rows_a + n) * a.rows_b;
            for (int col = 0; col < a.rows_b; col += 32) {
                uint32_t v[32];
                if (my_chunks > 0) {
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + mb * 256 + col, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                }
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    if (col + i < a.rows_b)
                        *reinterpret_cast<uint4 *>(dst + col + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) { tc_fence_after_sync(); tmem_dealloc<512>(tmem_base); }
}

// grads[n][col_off + k] += sum_split partial[split][n][k]   (k < rows_b_valid)
__global__ void wgrad_reduce_kernel(const float *__restrict__ partial, int splits, int rows_a, int rows_b, int rows_b_valid,
                                    float *__restrict__ dW, int ld, int col_off)
{
    const int total = rows_a * rows_b_valid;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / rows_b_valid, k = i % rows_b_valid;
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += partial[((size_t)sp * rows_a + n) * rows_b + k];
        dW[(size_t)n * ld + col_off + k] += s;
    }
}

// dbias[n] += sum_s A[n][s]; one block per row (a row is up to 1 MB: the whole GPU has to pull)
__global__ void __launch_bounds__(256) rowsum_kernel(const float *__restrict__ A, int rows, int ch, float *__restrict__ dbias)
{
    __shared__ float part[8];
    const int r = blockIdx.x;
    if (r >= rows) return;
    const float4 *p = reinterpret_cast<const float4 *>(A + (size_t)r * ch);
    float s = 0.f;
    for (int i = threadIdx.x; i < ch / 4; i += 256) { float4 v = __ldg(p + i); s += (v.x + v.y) + (v.z + v.w); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i];
        dbias[r] += t;
    }
}

// Skinny weight gradients (density head: 1 output row; colour layer 1: 3): dW[a][k] += sum_s A[a][s] B[k][s],
// dbias[a] += sum_s A[a][s].  One block per input feature k streams B's row once; A (<= 4 rows) stays in L2.
__global__ void __launch_bounds__(256) wgrad_skinny_kernel(const float *__restrict__ A, int rows_a, const __nv_bfloat16 *__restrict__ B,
                                                           int rows_b, int ch, float *__restrict__ dW, int ld,
                                                           float *__restrict__ dbias)
{
    __shared__ float part[8][5];
    const int k = blockIdx.x;                      // k == rows_b: the bias block (B row of ones)
    const uint2 *bp = k < rows_b ? reinterpret_cast<const uint2 *>(B + (size_t)k * ch) : nullptr;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < ch / 4; i += 256) {
        float4 b = make_float4(1.f, 1.f, 1.f, 1.f);
        if (bp) {
            const uint2 r = __ldg(bp + i);
            b = make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                            __uint_as_float(r.y & 0xffff0000u));
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r < rows_a) {
                const float4 av = __ldg(reinterpret_cast<const float4 *>(A + (size_t)r * ch) + i);
                s[r] += (av.x * b.x + av.y * b.y) + (av.z * b.z + av.w * b.w);
            }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int r = 0; r < 4; ++r) part[threadIdx.x >> 5][r] = s[r];
    __syncthreads();
    if (threadIdx.x < rows_a) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
        if (k < rows_b) dW[(size_t)threadIdx.x * ld + k] += t;
        else if (dbias) dbias[threadIdx.x] += t;
    }
}

}  // namespace wg

int wgrad_skinny(const float *A, int rows_a, const __nv_bfloat16 *B, int rows_b, int ch, float *dW, int ld, float *dbias,
                 cudaStream_t stream)
{
    wg::wgrad_skinny_kernel<<<rows_b + 1, 256, 0, stream>>>(A, rows_a, B, rows_b, ch, dW, ld, dbias);
    return launch_status();
}

size_t wgrad_tc_scratch_bytes(int splits) { return (size_t)splits * 256 * 256 * sizeof(float); }

// dW (+)= A^T-by-B over the chunk's samples on the tensor cores; bias row sums on CUDA cores.
int wgrad_tc(const __nv_bfloat16 *A, int rows_a, const __nv_bfloat16 *B, int rows_b_valid, int ch, float *dW, int ld, int col_off,
             float *dbias, float *scratch, int splits, cudaStream_t stream)
{
    wg::Args a = {};
    a.A = A; a.rows_a = rows_a; a.B = B; a.rows_b_valid = rows_b_valid;
    a.rows_b = (rows_b_valid + 15) / 16 * 16;
    a.ch = ch; a.partial = scratch; a.dbias = dbias;
    const int n_chunks = ch / 64;
    if (splits > n_chunks) splits = n_chunks;
    cudaError_t e = cudaFuncSetAttribute(wg::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg::kSmem);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    wg::wgrad_tc_kernel<<<splits, wg::kThreads, wg::kSmem, stream>>>(a);
    int rc = launch_status();
    if (rc) return rc;
    const int total = rows_a * rows_b_valid;
    wg::wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(scratch, splits, rows_a, a.rows_b, rows_b_valid, dW, ld, col_off);
    if ((rc = launch_status())) return rc;
    return rc;
}

}  // namespace nerfb200
