// Microbenchmark: does the placement of the A operand (in tensor memory) relative to the accumulator D, and of concurrent
// epilogue traffic (tcgen05.ld 64 columns + tcgen05.st 32 columns per warp, like the fused kernel's trunk epilogue),
// change the rate of tcgen05.mma (M=128, N=128, K=16, bf16)?  One CTA per SM; thread 0 issues the MMAs, warps 4..11 run
// the epilogue loop on a given column base until the MMAs are done.  Columns are given as a_col (K-block base: the four
// K-steps read a_col + 8k), d_col, e_col (epilogue warps w2 = 0/1 use e_col + 64 w2).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../nerf_dbr_b200/csrc/ptx.cuh"
using namespace nerfb200::ptx;

__global__ void __launch_bounds__(384, 1) probe(int reps, int a_col, int a_col2, int d_col, int d_col2, int e_col, int epi_warps, long long *out)
{
    extern __shared__ uint8_t raw[];
    uint8_t *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    const uint32_t base = smem_u32(sm);
    const uint32_t bar = base + 200 * 1024, tptr = bar + 64;
    volatile int *done = reinterpret_cast<volatile int *>(sm + 200 * 1024 + 128);
    for (uint32_t i = threadIdx.x; i < 50 * 1024; i += blockDim.x) reinterpret_cast<uint32_t *>(sm)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); *done = 0; }
    if (threadIdx.x < 32) tmem_alloc<512>(tptr);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = *reinterpret_cast<volatile uint32_t *>(sm + 200 * 1024 + 64);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = idesc_bf16(128, 128);
        const uint64_t bdesc = smem_desc_sw128(base + 65536);
        long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < reps; ++r) {
            // one "chunk pair": 4 K-steps with (A, D) and 4 with (A2, D2), like alternating halves / K-blocks
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_bf16_ts(tm + d_col, tm + a_col + 8 * k, bdesc + 2 * k, idesc, true);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_bf16_ts(tm + d_col2, tm + a_col2 + 8 * k, bdesc + 2 * k + 512, idesc, true);
        }
        mma_commit(bar);
        while (!mbar_try_wait(bar, 0)) {}
        long long t2 = clock64();
        *done = 1;
        if (blockIdx.x == 0) out[0] = t2 - t0;
    } else if (warp >= 4 && warp < 4 + epi_warps) {
        const int ew = warp - 4, q = ew & 3, w2 = ew >> 2;
        const uint32_t t_cols = tm + ((uint32_t)(q * 32) << 16) + e_col + 64 * w2;
        long long n = 0;
        while (!*done) {
            uint32_t xa[32], xb[32], pk[32];
            tmem_ld32(t_cols, xa);
            tmem_ld32(t_cols + 32, xb);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) { pk[i] = xa[2 * i] ^ xa[2 * i + 1]; pk[16 + i] = xb[2 * i] + xb[2 * i + 1]; }
            tmem_st32(t_cols, pk);
            tmem_st_wait();
            ++n;
        }
        if (blockIdx.x == 0 && threadIdx.x == 128) out[1] = n;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc<512>(tm); }
}

int main()
{
    long long *d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    struct C { const char *name; int a, a2, d, d2, e, w; };
    const C cases[] = {
        {"no epilogue: A 256/320, D 0/128 (old map)", 256, 320, 0, 128, 384, 0},
        {"no epilogue: A 128/192 (same half as D 0), D 0/0", 128, 192, 0, 0, 384, 0},
        {"no epilogue: A 128, D 0 | A 384, D 256 (new map, same halves)", 128, 384, 0, 256, 384, 0},
        {"no epilogue: A 128, D 256 | A 384, D 0 (cross halves)", 128, 384, 256, 0, 384, 0},
        {"old map + epilogue in A's half:  A 256/320, D 0/128, E 384 x8", 256, 320, 0, 128, 384, 8},
        {"old map + epilogue in D's half:  A 256/320, D 0/128, E 128 x8 (not a real case)", 256, 320, 0, 0, 128, 8},
        {"new map h0: A 128/192, D 0/0, E 384 x8 (h1 of previous layer)", 128, 192, 0, 0, 384, 8},
        {"new map h0: A 384/448 (kb 2,3), D 0/0, E 256 x8", 384, 448, 0, 0, 256, 8},
        {"new map h1: A 128/192, D 256/256, E 384 x8", 128, 192, 256, 256, 384, 8},
        {"new map h1: A 384/448, D 256/256, E 128 x8", 384, 448, 256, 256, 128, 8},
        {"new map mix: A 128, D 0 | A 384, D 256, E 384+ x8", 128, 384, 0, 256, 448 - 64, 8},
        {"4 epilogue warps: A 256/320, D 0/128, E 384 x4", 256, 320, 0, 128, 384, 4},
        {"4 epilogue warps: A 128/192, D 0/0, E 384 x4", 128, 192, 0, 0, 384, 4},
    };
    const int reps = 256;
    printf("%-82s | %10s %10s %12s\n", "case", "cyc/MMA", "total", "epi loops");
    for (const C &c : cases) {
        for (int grid : {148}) {
            cudaMemset(d, 0, 16);
            probe<<<grid, 384, 210 * 1024>>>(reps, c.a, c.a2, c.d, c.d2, c.e, c.w, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%-82s | %10.1f %10lld %12lld\n", c.name, h[0] / (reps * 8.0), h[0], h[1]);
        }
    }
    return 0;
}
