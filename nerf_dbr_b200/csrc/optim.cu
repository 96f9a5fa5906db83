// Optimizer step of the training loop, fused with the data-parallel gradient exchange over NVLink.
//
// reference: NeRFTrainer.train_step, src/training/trainer.py:125-136 -- loss.backward(); clip_grad_norm_(both
// networks, max_norm); Adam.step() (lr, weight_decay as L2-in-gradient, betas 0.9 / 0.999, eps 1e-8); ExponentialLR.
// The reference has no data parallelism; here every rank holds a replica and the gradients of the ray shards are
// summed.
//
// All 2 x 22 parameter tensors of a step live in ONE flat fp32 bucket (P), likewise gradients (G), Adam moments
// (M, V).  Two launches do what torch does in ~10 (all-reduce, norm, clip, multi-tensor Adam, zero_grad):
//
//   dp_reduce_kernel   (K_A)  wait until every rank's G is final (flag barrier over NVLink peer memory), sum THIS
//                             rank's 1/world shard of G over all ranks in rank order (peer loads, or one
//                             multimem.ld_reduce through the NVSwitch), store the sums into EVERY rank's Gsum
//                             (peer stores / multimem.st: the all-gather), and publish the shard's sum of squares.
//   dp_adam_kernel     (K_B)  wait until every shard has arrived, total norm = sqrt(sum of the per-rank parts, in rank
//                             order), clip coefficient, Adam on the whole bucket (replicated: every rank computes the
//                             same bits), zero G for the next step, loss and norm out.
//
// Every rank reduces in the same order and applies the same update, so the replicas stay bit-identical.  With
// world == 1 the same two kernels run on local memory (K_A degenerates to a copy + sum of squares).  The symmetric
// allocation [ctl | G | Gsum] comes from the caller (torch symmetric memory: cuMem + fabric handles); this file
// only sees the mapped peer addresses.
//
// No atomics on data: block partials are combined by the last block in block order -> deterministic.
#include <algorithm>
#include <cstddef>
#include "common.cuh"

namespace nerfb200 {
namespace opt {

constexpr int kMaxWorld = NERF_B200_DP_MAX_WORLD;
constexpr int kThreads = 256;
constexpr int kMaxBlocksA = 256;
constexpr long long kSpinLimit = 40000000000LL;           // ~20 s of SM clocks: a rank that never arrives is a dead job

// ctl block at the head of the symmetric allocation (1024 bytes)
struct Ctl {
    unsigned int flag_a[kMaxWorld];                       // [q] = last step for which rank q's G was final
    unsigned int flag_b[kMaxWorld];                       // [q] = last step for which rank q's shard (and norm part) arrived
    float norm_part[kMaxWorld];                           // sum of squares of rank q's shard of the summed gradient
};
static_assert(sizeof(Ctl) <= NERF_B200_DP_CTL_BYTES, "ctl block");
static_assert(kMaxBlocksA <= kThreads, "the last block fetches one partial per thread");

struct Local {                                            // per-rank private state (device memory, zeroed by the caller)
    unsigned int step;                                    // steps completed: the flags carry step + 1
    unsigned int ticket_a, ticket_b;
    unsigned int opt_step;                                // optimizer steps taken (Adam's `step`, the scheduler's last_epoch): host-settable on resume
    float block_part[kMaxBlocksA];                        // (the state block is NERF_B200_DP_STATE_BYTES)
};
static_assert(sizeof(Local) <= NERF_B200_DP_STATE_BYTES && offsetof(Local, opt_step) == NERF_B200_DP_STATE_OPT_STEP, "state block");

struct Args {
    int rank, world;
    long long n;                                          // floats in the bucket (multiple of 4 * world)
    long long n_opt;                                      // leading floats that are parameters (norm + Adam); the tail rides along (loss slot)
    unsigned char *peer[kMaxWorld];                       // symmetric allocation of every rank as mapped here ([rank] = own)
    unsigned char *mc;                                    // multicast mapping of the same allocation, or nullptr
    Local *local;
    int barriers;                                         // 0 = single-GPU emulation of several ranks by sequential launches (tests)
    // K_B only
    float *p, *m, *v;
    const double *hyper;                                  // device: lr0, gamma, beta1, beta2, eps, weight_decay, max_norm, loss_scale
    float *loss_out;                                      // [0] = tail slot 0 of the summed bucket * loss_scale, [1] = gradient norm before clipping, [2] = lr used
};

__device__ __forceinline__ Ctl *ctl_of(unsigned char *base) { return reinterpret_cast<Ctl *>(base); }
__device__ __forceinline__ float *g_of(unsigned char *base) { return reinterpret_cast<float *>(base + NERF_B200_DP_CTL_BYTES); }
__device__ __forceinline__ float *gsum_of(unsigned char *base, long long n) { return g_of(base) + n; }

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_sys_v4(const float *p)
{
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_v4(float *p, float4 v)
{
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float ld_sys_f32(const float *p)
{
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
// NVSwitch in-network reduction / broadcast on a multicast address (NVLS)
__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float *mc)
{
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_v4(float *mc, float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// thread q < world waits until rank q's flag reaches `want`; then the block goes on
__device__ __forceinline__ void wait_flags(const unsigned int *flags, int world, unsigned int want)
{
    if ((int)threadIdx.x < world) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(flags + threadIdx.x) - want) < 0) {
            __nanosleep(64);
            if (clock64() - t0 > kSpinLimit) __trap();    // a peer never arrived: fail the launch instead of hanging the node
        }
    }
    __syncthreads();
}

// deterministic block sum (fixed tree), result valid in thread 0
__device__ __forceinline__ float block_sum(float x)
{
    __shared__ float wsum[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    float s = 0.f;
    if (threadIdx.x == 0)
#pragma unroll
        for (int i = 0; i < kThreads / 32; ++i) s += wsum[i];
    return s;
}

__global__ void __launch_bounds__(kThreads) dp_reduce_kernel(const __grid_constant__ Args a)
{
    unsigned char *self = a.peer[a.rank];
    const unsigned int epoch = a.local->step + 1;
    if (a.barriers) {
        // barrier A: my G is final (it was written by earlier kernels of this stream) -> tell everyone, wait for everyone
        if (blockIdx.x == 0 && (int)threadIdx.x < a.world) {
            __threadfence_system();
            st_release_sys(ctl_of(a.peer[threadIdx.x])->flag_a + a.rank, epoch);
        }
        wait_flags(ctl_of(self)->flag_a, a.world, epoch);
    }
    const long long shard4 = a.n / 4 / a.world, first4 = shard4 * a.rank;          // float4 units
    float sq = 0.f;
    constexpr int kBatch = 4;                             // independent 16-byte loads in flight per thread and peer
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i0 = blockIdx.x * (long long)kThreads + threadIdx.x; i0 < shard4; i0 += stride * kBatch) {
        float4 s[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const long long i = i0 + b * stride;
            s[b] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < shard4) s[b] = a.mc ? multimem_ld_reduce_v4(g_of(a.mc) + (first4 + i) * 4)      // the switch sums the replicas
                                        : ld_sys_v4(g_of(a.peer[0]) + (first4 + i) * 4);
        }
        if (!a.mc)
            for (int q = 1; q < a.world; ++q) {                                     // rank order: same bits on every rank
                float4 t[kBatch];
#pragma unroll
                for (int b = 0; b < kBatch; ++b) {
                    const long long i = i0 + b * stride;
                    t[b] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < shard4) t[b] = ld_sys_v4(g_of(a.peer[q]) + (first4 + i) * 4);
                }
#pragma unroll
                for (int b = 0; b < kBatch; ++b) { s[b].x += t[b].x; s[b].y += t[b].y; s[b].z += t[b].z; s[b].w += t[b].w; }
            }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const long long i = i0 + b * stride;
            if (i >= shard4) continue;
            const long long e = (first4 + i) * 4;
            if (a.mc) {
                multimem_st_v4(gsum_of(a.mc, a.n) + e, s[b]);                       // one store, delivered to every replica
            } else {
                for (int q = 0; q < a.world; ++q) st_sys_v4(gsum_of(a.peer[q], a.n) + e, s[b]);
            }
            if (e + 0 < a.n_opt) sq = fmaf(s[b].x, s[b].x, sq);
            if (e + 1 < a.n_opt) sq = fmaf(s[b].y, s[b].y, sq);
            if (e + 2 < a.n_opt) sq = fmaf(s[b].z, s[b].z, sq);
            if (e + 3 < a.n_opt) sq = fmaf(s[b].w, s[b].w, sq);
        }
    }
    const float bs = block_sum(sq);
    __shared__ bool last;
    __threadfence_system();                               // this block's shard stores are ordered before its ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        a.local->block_part[blockIdx.x] = bs;
        __threadfence();
        last = atomicAdd(&a.local->ticket_a, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    // last block: the shard's sum of squares in block order (partials fetched in parallel, added by one thread), then
    // publish it and signal "shard delivered"
    __shared__ float parts[kMaxBlocksA];
    __threadfence();
    if (threadIdx.x < gridDim.x) parts[threadIdx.x] = *reinterpret_cast<volatile float *>(&a.local->block_part[threadIdx.x]);
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (unsigned int b = 0; b < gridDim.x; ++b) tot += parts[b];
        a.local->ticket_a = 0;
        for (int q = 0; q < a.world; ++q)
            *reinterpret_cast<volatile float *>(&ctl_of(a.peer[q])->norm_part[a.rank]) = tot;
        __threadfence_system();
        if (a.barriers)
            for (int q = 0; q < a.world; ++q) st_release_sys(ctl_of(a.peer[q])->flag_b + a.rank, epoch);
    }
}

__global__ void __launch_bounds__(kThreads) dp_adam_kernel(const __grid_constant__ Args a)
{
    unsigned char *self = a.peer[a.rank];
    const unsigned int epoch = a.local->step + 1;
    if (a.barriers) wait_flags(ctl_of(self)->flag_b, a.world, epoch);      // every shard of Gsum and every norm part is here
    float total = 0.f;
    for (int q = 0; q < a.world; ++q) total += ld_sys_f32(&ctl_of(self)->norm_part[q]);
    // Adam's bias corrections and ExponentialLR's rate for this step, from the device-resident step count (no host
    // traffic per step): lr_t = lr0 gamma^t, 1 - beta^(t+1)   (trainer.py:55-64, 136)
    __shared__ float sched[3];
    if (threadIdx.x == 0) {
        const double t = (double)a.local->opt_step;
        sched[0] = (float)(a.hyper[0] * pow(a.hyper[1], t));
        sched[1] = (float)(1.0 - pow(a.hyper[2], t + 1.0));
        sched[2] = (float)(1.0 - pow(a.hyper[3], t + 1.0));
    }
    __syncthreads();
    const float lr = sched[0], bc1 = sched[1], bc2 = sched[2];
    const float beta1 = (float)a.hyper[2], beta2 = (float)a.hyper[3], eps = (float)a.hyper[4], wd = (float)a.hyper[5];
    const float max_norm = (float)a.hyper[6], loss_scale = (float)a.hyper[7];
    const float omb1 = (float)(1.0 - a.hyper[2]), omb2 = (float)(1.0 - a.hyper[3]);      // formed in double, as torch's python scalars are
    const float norm = sqrtf(total);
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
    const float coef = max_norm > 0.f ? fminf(1.0f, max_norm / (norm + 1e-6f)) : 1.0f;
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const float *gs = gsum_of(self, a.n);
    float *g = g_of(self);
    const long long n4 = a.n / 4;
    for (long long i = blockIdx.x * (long long)kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads) {
        const long long e = i * 4;
        const float4 gv = ld_sys_v4(gs + e);              // written by peers: not through the non-coherent path
        *reinterpret_cast<float4 *>(g + e) = make_float4(0.f, 0.f, 0.f, 0.f);      // zero_grad for the next step
        if (e >= a.n_opt) continue;                       // tail slots (loss): reduced, not optimised
        float4 pv = *reinterpret_cast<const float4 *>(a.p + e), mv = *reinterpret_cast<const float4 *>(a.m + e),
               vv = *reinterpret_cast<const float4 *>(a.v + e);
        float gg[4] = {gv.x, gv.y, gv.z, gv.w}, pp[4] = {pv.x, pv.y, pv.z, pv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w},
              v2[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float grad = gg[j] * coef;
            grad = fmaf(wd, pp[j], grad);                 // Adam(weight_decay): L2 in the gradient, not AdamW
            mm[j] = fmaf(grad - mm[j], omb1, mm[j]);               // exp_avg.lerp_(grad, 1 - beta1)
            v2[j] = fmaf(v2[j], beta2, omb2 * grad * grad);      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            const float denom = sqrtf(v2[j]) * inv_sqrt_bc2 + eps;
            pp[j] -= step_size * (mm[j] / denom);
        }
        *reinterpret_cast<float4 *>(a.p + e) = make_float4(pp[0], pp[1], pp[2], pp[3]);
        *reinterpret_cast<float4 *>(a.m + e) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        *reinterpret_cast<float4 *>(a.v + e) = make_float4(v2[0], v2[1], v2[2], v2[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.loss_out) {
        a.loss_out[0] = ld_sys_f32(gs + a.n_opt) * loss_scale;
        a.loss_out[1] = norm;
        a.loss_out[2] = lr;
    }
    // the last block to finish closes the step (every block has read `step` by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&a.local->ticket_b, 1u) == gridDim.x - 1) {
            a.local->ticket_b = 0;
            a.local->step = epoch;
            a.local->opt_step += 1;
            __threadfence();
        }
    }
}

static int fill(Args &a, const nerf_b200_dp *dp)
{
    if (!dp || dp->world < 1 || dp->world > kMaxWorld || dp->rank < 0 || dp->rank >= dp->world || !dp->state) return NERF_B200_EINVAL;
    if (dp->n <= 0 || dp->n % (4 * dp->world) || dp->n_opt < 0 || dp->n_opt >= dp->n || dp->n_opt % 4) return NERF_B200_EINVAL;
    a = {};
    a.rank = dp->rank; a.world = dp->world; a.n = dp->n; a.n_opt = dp->n_opt;
    for (int q = 0; q < dp->world; ++q) {
        if (!dp->peer[q] || ((uintptr_t)dp->peer[q] & 15)) return dp->peer[q] ? NERF_B200_EALIGN : NERF_B200_EINVAL;
        a.peer[q] = reinterpret_cast<unsigned char *>(dp->peer[q]);
    }
    a.mc = reinterpret_cast<unsigned char *>(dp->multicast);
    a.local = reinterpret_cast<Local *>(dp->state);
    a.barriers = dp->emulate_sequential ? 0 : 1;
    return 0;
}

}  // namespace opt
}  // namespace nerfb200

using namespace nerfb200;

extern "C" {

size_t nerf_b200_dp_bytes(int64_t n) { return n > 0 ? (size_t)NERF_B200_DP_CTL_BYTES + 2 * (size_t)n * sizeof(float) : 0; }

int nerf_b200_dp_reduce(const nerf_b200_dp *dp, void *stream)
{
    opt::Args a;
    int rc = opt::fill(a, dp);
    if (rc) return rc;
    const long long shard4 = a.n / 4 / a.world;
    const int grid = (int)std::max<long long>(1, std::min<long long>(opt::kMaxBlocksA, (shard4 + opt::kThreads * 4 - 1) / (opt::kThreads * 4)));
    opt::dp_reduce_kernel<<<grid, opt::kThreads, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}

int nerf_b200_dp_adam_step(const nerf_b200_dp *dp, float *params, float *exp_avg, float *exp_avg_sq, const double *hyper,
                           float *loss_out, void *stream)
{
    opt::Args a;
    int rc = opt::fill(a, dp);
    if (rc) return rc;
    if (!params || !exp_avg || !exp_avg_sq || !hyper) return NERF_B200_EINVAL;
    if (((uintptr_t)params | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) return NERF_B200_EALIGN;
    a.p = params; a.m = exp_avg; a.v = exp_avg_sq; a.hyper = hyper; a.loss_out = loss_out;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::max<long long>(1, std::min<long long>(2 * sms, (a.n / 4 + opt::kThreads * 4 - 1) / (opt::kThreads * 4)));
    opt::dp_adam_kernel<<<grid, opt::kThreads, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}

}  // extern "C"
