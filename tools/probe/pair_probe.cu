// Probe for the next step of the render kernel: does pairing two SMs on one MMA (tcgen05 cta_group::2, M = 256,
// each CTA holding half of B in its shared memory) buy sustained throughput under the power cap?
// Both variants stream the same MMA shape the fused kernel uses (A from tensor memory, B [N=128 x K=16] from
// shared memory, fp32 accumulate, random bf16 data so the datapath toggles realistically) for about half a second on
// all 148 SMs; reported: wall-clock TFLOP/s and cycles per instruction.
//   solo: cta_group::1, M = 128, every SM reads a 4 KB B tile per MMA from its own shared memory
//   pair: cta_group::2, M = 256, issued by the even CTA of each pair; every SM reads 2 KB per MMA
// Build: make -C tools/probe pair_probe.   Run under gpurun.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../nerf_dbr_b200/csrc/ptx.cuh"
using namespace nerfb200::ptx;

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// two bf16 in (-1, 1) with full mantissas
__device__ __forceinline__ uint32_t rand_bf16x2(uint32_t seed)
{
    const uint32_t h = hash32(seed);
    const float a = (float)(int)(h & 0xffffu) * (1.0f / 65536.0f) - 0.5f, b = (float)(int)(h >> 16) * (1.0f / 65536.0f) - 0.5f;
    return pack_bf16(a, b);
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void wait_or_trap(uint32_t bar)
{
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, 0))
        if (clock64() - t0 > 6000000000LL) __trap();           // ~3 s: a protocol mistake must not hang the GPU
}
__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

constexpr uint32_t kBarOff = 64 * 1024, kTptrOff = kBarOff + 64;

// common set-up: random B tile(s) in shared memory, random A K-blocks in TMEM columns [448, 480)
__device__ void fill_operands(uint8_t *sm, uint32_t tm, uint32_t b_bytes)
{
    for (uint32_t i = threadIdx.x; i < b_bytes / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(sm)[i] = rand_bf16x2(i * 2654435761u + blockIdx.x);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = rand_bf16x2((threadIdx.x * 32 + i) * 40503u + 17u * blockIdx.x);
    tmem_st32(tm + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + 448, v);
    tmem_st_wait();
    fence_proxy_async_smem();
}

// STREAM > 0: warp 1 additionally pulls `STREAM` KB stages from a 1 MiB L2-resident buffer into shared memory with
// cp.async.bulk, back to back through a 4-slot ring (the fused kernel streams 1 MiB of weights per tile per SM, about
// 57 B/clk; unthrottled this probe pulls more) -- what does the L2 -> shared-memory weight stream cost in clocks?
template <int STREAM>
__global__ void __launch_bounds__(128, 1) solo_kernel_t(int reps, long long *cycles, const unsigned char *wbuf, unsigned long long *streamed)
{
    extern __shared__ uint8_t raw[];
    uint8_t *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    const uint32_t base = smem_u32(sm), bar = base + kBarOff, tptr = base + kTptrOff;
    volatile int *stop = reinterpret_cast<volatile int *>(sm + kTptrOff + 16);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        for (int i = 0; i < 4; ++i) mbar_init(bar + 8 + 8 * i, 1);
        *stop = 0;
        fence_mbar_init();
    }
    if (threadIdx.x < 32) tmem_alloc<512>(tptr);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = *reinterpret_cast<volatile uint32_t *>(sm + kTptrOff);
    fill_operands(sm, tm, 4 * 16384);                        // four [128 x 64] B chunks
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = idesc_bf16(128, 128);
        const uint64_t bdesc = smem_desc_sw128(base);
        const long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < reps; ++r)
#pragma unroll
            for (int k = 0; k < 16; ++k)                     // two accumulators alternate, like the kernel's halves
                mma_bf16_ts(tm + (k & 1) * 128, tm + 448 + (k & 3) * 8, bdesc + 2 * (k & 3) + 1024 * (k >> 2), idesc, (r | (k >> 1)) != 0);
        mma_commit(bar);
        wait_or_trap(bar);
        if (blockIdx.x == 0) cycles[0] = clock64() - t0;
        *stop = 1;
    } else if (STREAM > 0 && threadIdx.x == 32) {
        // ring of 4 slots behind the operands (smem offset 72 KB..), each STREAM KB
        unsigned long long n = 0;
        uint32_t phase[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            mbar_arrive_expect_tx(bar + 8 + 8 * i, STREAM * 1024);
            bulk_g2s(base + 72 * 1024 + i * STREAM * 1024, wbuf + ((n * STREAM * 1024) & ((1u << 20) - 1)), STREAM * 1024, bar + 8 + 8 * i);
            ++n;
        }
        while (!*stop) {
            const int i = (int)(n & 3);
            const long long t1 = clock64();
            while (!mbar_try_wait(bar + 8 + 8 * i, phase[i])) if (clock64() - t1 > 2000000000LL) __trap();
            phase[i] ^= 1;
            mbar_arrive_expect_tx(bar + 8 + 8 * i, STREAM * 1024);
            bulk_g2s(base + 72 * 1024 + i * STREAM * 1024, wbuf + ((n * STREAM * 1024) & ((1u << 20) - 1)), STREAM * 1024, bar + 8 + 8 * i);
            ++n;
        }
        for (int i = 0; i < 4; ++i) {                        // drain
            const long long t1 = clock64();
            while (!mbar_try_wait(bar + 8 + 8 * i, phase[i])) if (clock64() - t1 > 2000000000LL) __trap();
        }
        if (blockIdx.x == 0) streamed[0] = n * STREAM * 1024;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc<512>(tm); }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(int reps, long long *cycles)
{
    extern __shared__ uint8_t raw[];
    uint8_t *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);       // same offset in both CTAs of the pair
    const uint32_t base = smem_u32(sm), bar = base + kBarOff, tptr = base + kTptrOff;
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    cluster_sync_all();
    if (threadIdx.x < 32) {                                  // the same warp of both CTAs allocates for the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = *reinterpret_cast<volatile uint32_t *>(sm + kTptrOff);
    fill_operands(sm, tm, 4 * 8192);                         // this CTA's half of B: four [64 x 64] chunks
    tc_fence_before_sync();
    cluster_sync_all();
    tc_fence_after_sync();
    if (rank == 0 && threadIdx.x == 0) {
        constexpr uint32_t idesc = idesc_bf16(256, 128);     // M = 256 over the pair, N = 128 (64 rows of B per CTA)
        const uint64_t bdesc = smem_desc_sw128(base);
        const long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < reps; ++r)
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t d = tm + (k & 1) * 128, a = tm + 448 + (k & 3) * 8;
                const uint64_t b = bdesc + 2 * (k & 3) + 512 * (k >> 2);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                    ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"((uint32_t)((r | (k >> 1)) != 0)), "r"(0u) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"((unsigned short)3) : "memory");
        wait_or_trap(bar);
        if (blockIdx.x == 0) cycles[0] = clock64() - t0;
    } else if (rank == 1 && threadIdx.x == 0) {
        wait_or_trap(bar);                                   // the pair's MMAs have completed here too
    }
    tc_fence_before_sync();
    cluster_sync_all();
    if (threadIdx.x < 32) {
        tc_fence_after_sync();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
    }
}

template <int STREAM>
static void run_solo(const char *name, int reps, double flop, long long *d, const unsigned char *wbuf, unsigned long long *streamed)
{
    const int smem = (72 + 4 * (STREAM > 0 ? STREAM : 1) + 2) * 1024;
    cudaFuncSetAttribute(solo_kernel_t<STREAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int pass = 0; pass < 3; ++pass) {
        cudaMemset(streamed, 0, 8);
        cudaEventRecord(e0);
        solo_kernel_t<STREAM><<<148, 128, smem>>>(reps, d, wbuf, streamed);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); exit(1); }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        unsigned long long bytes; cudaMemcpy(&bytes, streamed, 8, cudaMemcpyDeviceToHost);
        if (pass) printf("%-9s grid 148 reps %8d | %8.2f ms | %7.1f TFLOP/s | %5.0f MHz | L2->smem stream %5.1f B/clk/SM (%.2f TB/s total)\n", name, reps, ms,
                         flop * reps * 148 / (ms * 1e-3) / 1e12, cyc / (ms * 1e3), (double)bytes / cyc, bytes * 148.0 / (ms * 1e-3) / 1e12);
    }
}

template <typename K>
static void run(const char *name, K kernel, int grid, int reps, double flop_per_cta_rep, long long *d)
{
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int pass = 0; pass < 3; ++pass) {                   // pass 0 warms up; 1 and 2 are reported
        cudaEventRecord(e0);
        kernel<<<grid, 128, 80 * 1024>>>(reps, d);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); exit(1); }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
        if (pass) printf("%-5s grid %3d reps %8d | %8.2f ms | %7.1f TFLOP/s | %6.1f cycles per MMA instruction | %5.0f MHz\n", name, grid, reps, ms,
                         flop_per_cta_rep * reps * grid / (ms * 1e-3) / 1e12, (double)cyc / (16.0 * reps), cyc / (ms * 1e3));
    }
}

int main(int argc, char **argv)
{
    long long *d; cudaMalloc(&d, 16);
    const int reps = argc > 1 ? atoi(argv[1]) : 700000;      // 11.2 M MMAs per issuer: ~0.5 s
    const double flop = 16.0 * 2.0 * 128 * 128 * 16;          // per CTA and repetition (the pair issuer covers two CTAs)
    unsigned char *wbuf; cudaMalloc(&wbuf, 1 << 20); cudaMemset(wbuf, 0x3c, 1 << 20);
    unsigned long long *streamed; cudaMalloc(&streamed, 8);
    if (argc > 2) {                                           // L2 -> shared-memory weight-stream experiment
        run_solo<0>("solo", reps, flop, d, wbuf, streamed);
        run_solo<8>("solo+8KB", reps, flop, d, wbuf, streamed);
        run_solo<32>("solo+32KB", reps, flop, d, wbuf, streamed);
        run_solo<0>("solo", reps, flop, d, wbuf, streamed);
        return 0;
    }
    run_solo<0>("solo", reps, flop, d, wbuf, streamed);
    run("pair", pair_kernel, 148, reps, flop, d);
    run_solo<0>("solo", reps, flop, d, wbuf, streamed);
    run("pair", pair_kernel, 148, reps, flop, d);
    return 0;
}
