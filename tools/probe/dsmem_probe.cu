// DSMEM hand-off rate between the CTAs of a cluster on B200 -- the number that decides whether weight-gradient operands
// can travel SM to SM instead of through HBM ("why the backward is not one on-chip kernel").
//
//   mode 0  cp.async.bulk shared::cta -> shared::cluster (TMA engine), 2-slot flow-controlled pipe per CTA: the
//           receiver waits for a full slot (mbarrier complete_tx), re-arms it and tells the sender (remote arrive)
//           that the slot is free -- what a real operand hand-off has to do.
//   mode 1  st.shared::cluster.v4.f32 from all 256 threads (no flow control): the plain remote-store rate.
//
// Every CTA of a cluster sends to its right neighbour and receives from its left one, all clusters at once (grid =
// whole GPU).  Prints bytes per SM clock per SM, sent (= received), for cluster sizes 2, 4, 8 and chunk sizes 8 / 32 KB.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/dsmem_probe tools/probe/dsmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity)
{
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 2000000000LL) __trap();
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }

constexpr int kMaxChunk = 32 * 1024;

__global__ void __launch_bounds__(256, 1) dsmem_kernel(int mode, int chunk, int iters, long long *cycles)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *src = smem, *dst = smem + kMaxChunk;                  // dst: 2 slots
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 3 * kMaxChunk);  // full[2], ready[2]
    const uint32_t me = cluster_rank(), n = cluster_size(), right = (me + 1) % n, left = (me + n - 1) % n;
    const uint32_t full0 = smem_u32(bars), ready0 = smem_u32(bars + 2);
    for (int i = threadIdx.x; i < kMaxChunk / 4; i += blockDim.x) reinterpret_cast<float *>(src)[i] = (float)i;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(ready0 + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < 2; ++s) mbar_expect_tx(full0 + 8 * s, (uint32_t)chunk);       // both slots armed before anyone sends
    }
    __syncthreads();
    cluster_sync();
    const long long t0 = clock64();
    if (mode == 0) {
        if (threadIdx.x == 0) {                                          // sender
            for (int it = 0; it < iters; ++it) {
                const int s = it & 1;
                if (it >= 2) wait(ready0 + 8 * s, ((it >> 1) - 1) & 1);  // the neighbour freed and re-armed the slot
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(mapa(smem_u32(dst + s * kMaxChunk), right)), "r"(smem_u32(src)), "r"((uint32_t)chunk),
                               "r"(mapa(full0 + 8 * s, right)) : "memory");
            }
        } else if (threadIdx.x == 32) {                                  // receiver
            for (int it = 0; it < iters; ++it) {
                const int s = it & 1;
                wait(full0 + 8 * s, (it >> 1) & 1);
                if (it + 2 < iters) {
                    mbar_expect_tx(full0 + 8 * s, (uint32_t)chunk);
                    mbar_arrive_remote(mapa(ready0 + 8 * s, left));
                }
            }
        }
    } else {
        const uint32_t base = mapa(smem_u32(dst), right);
        for (int it = 0; it < iters; ++it)
            for (int off = threadIdx.x * 16; off < chunk; off += blockDim.x * 16)
                asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(base + (it & 1) * kMaxChunk + off), "f"((float)it) : "memory");
    }
    __syncthreads();
    cluster_sync();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *cycles;
    cudaMallocManaged(&cycles, 256 * sizeof(long long));
    const size_t smem = 3 * kMaxChunk + 64;
    cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    printf("%d SMs; bytes per SM clock per SM (sent = received); clocks are per-SM clock64\n", sms);
    for (int mode = 0; mode < 2; ++mode)
        for (int cs : {2, 4, 8})
            for (int chunk : {8192, 32768}) {
                const int grid = sms / cs * cs, iters = 2000;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaError_t e = cudaLaunchKernelEx(&cfg, dsmem_kernel, mode, chunk, iters, cycles);
                    if (e == cudaSuccess) e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("mode %d cluster %d chunk %d: %s\n", mode, cs, chunk, cudaGetErrorString(e)); return 1; }
                }
                double worst = 0, mean = 0;
                for (int b = 0; b < grid; ++b) { worst = cycles[b] > worst ? cycles[b] : worst; mean += (double)cycles[b] / grid; }
                printf("%-28s cluster %d  chunk %5d B  grid %3d: %6.1f B/clk/SM (mean), %6.1f (slowest CTA)\n",
                       mode == 0 ? "cp.async.bulk smem->dsmem" : "st.shared::cluster.v4", cs, chunk, grid,
                       (double)iters * chunk / mean, (double)iters * chunk / worst);
            }
    return 0;
}
