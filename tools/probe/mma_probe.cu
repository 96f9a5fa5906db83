// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, operand source
// (A from shared memory = SS, from tensor memory = TS) and the number of independent accumulators
// the issue order interleaves.  One CTA per SM, one issuing thread.  Build: see tools/probe/Makefile.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../nerf_dbr_b200/csrc/ptx.cuh"
using namespace nerfb200::ptx;

struct Cfg { int n, ts, indep, reps; };

// straight-line issue: 16 K-steps x INDEP accumulators per repetition, descriptors hoisted
template <int N, int TS, int INDEP>
__global__ void __launch_bounds__(128, 1) probe(int reps, long long *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    const uint32_t base = smem_u32(sm);
    const uint32_t bar = base + 200 * 1024, tptr = bar + 64;
    for (uint32_t i = threadIdx.x; i < 50 * 1024; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(tptr);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = *reinterpret_cast<volatile uint32_t *>(sm + 200 * 1024 + 64);
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = idesc_bf16(128, N);
        const uint64_t adesc = smem_desc_sw128(base), bdesc = smem_desc_sw128(base + 65536);
        long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
#pragma unroll
                for (int a = 0; a < INDEP; ++a) {
                    const uint32_t d = tm + a * N;
                    if (TS) mma_bf16_ts(d, tm + 448 + (k & 3) * 8, bdesc + 2 * (k & 3) + 512 * (k >> 2), idesc, true);
                    else mma_bf16_ss(d, adesc + 2 * (k & 3) + 1024 * (k >> 2), bdesc + 2 * (k & 3) + 512 * (k >> 2), idesc, true);
                }
            }
        }
        long long t1 = clock64();
        mma_commit(bar);
        while (!mbar_try_wait(bar, 0)) {}
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc<512>(tm); }
}

template <int N, int TS, int INDEP>
void run(long long *d)
{
    const int reps = 64;
    cudaFuncSetAttribute(probe<N, TS, INDEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    for (int grid : {1, 148}) {
        probe<N, TS, INDEP><<<grid, 128, 210 * 1024>>>(reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        double n_mma = (double)reps * 16 * INDEP;
        printf("%5d %3d %6d | %12lld %12lld | %10.1f %12.0f  (grid %d)\n", N, TS, INDEP, h[0], h[1], h[1] / n_mma,
               128.0 * N * 16 * n_mma / h[1], grid);
    }
}

int main()
{
    long long *d; cudaMalloc(&d, 16);
    printf("%5s %3s %6s | %12s %12s | %10s %12s\n", "N", "TS", "indep", "issue_cyc", "total_cyc", "cyc/MMA", "MAC/cyc/SM");
    run<256, 0, 1>(d); run<128, 0, 1>(d); run<64, 0, 1>(d); run<32, 0, 1>(d);
    run<128, 0, 2>(d); run<64, 0, 2>(d); run<64, 0, 4>(d);
    run<256, 1, 1>(d); run<128, 1, 1>(d); run<64, 1, 1>(d);
    run<128, 1, 2>(d); run<64, 1, 2>(d); run<64, 1, 4>(d); run<192, 1, 1>(d); run<96, 1, 4>(d);
    return 0;
}
