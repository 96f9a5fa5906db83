// Tensor-memory read-back rate (tcgen05.ld) on B200: what bounds the epilogue of a fused MLP whose activations live in
// TMEM.  W warps of one CTA (warp w reads lane quadrant w % 4, a private 64-column window) issue back-to-back
// tcgen05.ld.32x32b.x32 (128 B per thread, 4 KB per warp and instruction) for `iters` rounds; printed: bytes per SM clock for
// the whole SM and per warp, for 4 warps (one per quadrant / SM sub-partition), 8 (two per quadrant, the layout of the
// render kernel's epilogue) and 16.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/ldtm_probe tools/probe/ldtm_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512, 1) ldtm_kernel(int warps, int iters, long long *cycles, uint32_t *sink)
{
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(&tslot);
    const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 64;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < warps) {
        for (int it = 0; it < iters; ++it) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(addr + (uint32_t)(it & 1) * 32) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= r[0] ^ r[31];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *cycles;
    uint32_t *sink;
    cudaMallocManaged(&cycles, 256 * sizeof(long long));
    cudaMalloc(&sink, 4);
    printf("tcgen05.ld.32x32b.x32 read-back, one CTA per SM on %d SMs, 4 KB per warp and instruction\n", sms);
    for (int warps : {1, 4, 8, 16}) {
        const int iters = 4000;
        for (int rep = 0; rep < 2; ++rep) {
            ldtm_kernel<<<sms, 512>>>(warps, iters, cycles, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
        }
        double mean = 0;
        for (int b = 0; b < sms; ++b) mean += (double)cycles[b] / sms;
        const double bytes = (double)warps * iters * 4096.0;
        printf("%2d warps: %7.1f B/clk/SM, %6.1f B/clk per warp, %6.1f cycles per instruction per warp\n", warps, bytes / mean,
               bytes / mean / warps, mean / iters);
    }
    return 0;
}
