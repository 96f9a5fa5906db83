// Microbenchmark: what bounds a store-bound kernel on this part?  Write-only bandwidth with different store flavours
// (plain 128-bit, streaming .cs, 256-bit, L2 evict-first, bulk shared->global copies = the TMA engine), and how much read
// bandwidth is left beside a write stream (the TRAIN forward / dgrad kernels write at ~3.5 TB/s while wgrad reads at ~5.9:
// could they run side by side?).  4 GiB buffers, persistent grid.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { ST_PLAIN, ST_CS, ST_V8, ST_EVICT_FIRST, ST_BULK, ST_TILE, ST_TILE_BULK, ST_TILE_LANEROW, N_MODES };
static const char *kNames[] = {"st.global.v4", "st.global.cs.v4", "st.global.v8 (256-bit)", "st.global.L2::evict_first.v8 (256-bit)", "cp.async.bulk shared->global (16 KB)",
    "workspace pattern: 4 KB per warp, 512 B per instruction", "workspace pattern: 4 KB per warp, cp.async.bulk", "workspace pattern: 4 KB per warp, lane = row (16 B per line per instruction)"};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// blocks [0, n_writers) write `wbytes` of w, blocks [n_writers, grid) read `rbytes` of r
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(uint4 *w, size_t wbytes, const uint4 *r, size_t rbytes, int n_writers, unsigned long long *sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    if ((int)blockIdx.x < n_writers) {
        if (MODE == ST_BULK) {
            // every warp owns a 16 KB staging buffer and copies it out again and again
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = 12;
            if (warp >= warps) return;
            uint4 *buf = reinterpret_cast<uint4 *>(sm + warp * 16384);
            for (int i = lane; i < 1024; i += 32) buf[i] = make_uint4(i, warp, blockIdx.x, 7);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            const size_t chunk = 16384, n_chunks = wbytes / chunk;
            if (lane == 0) {
                int inflight = 0;
                for (size_t c = (size_t)blockIdx.x * warps + warp; c < n_chunks; c += (size_t)n_writers * warps) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<unsigned char *>(w) + c * chunk), "r"(smem_u32(buf)), "r"((uint32_t)chunk) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (++inflight >= 4) { asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); }
                }
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            }
            return;
        }
        if (MODE == ST_TILE || MODE == ST_TILE_BULK || MODE == ST_TILE_LANEROW) {
            // the training workspace: [slab of 64 samples][70 blocks of 64 features] tiles of 8 KB; a CTA takes 128-sample
            // tiles (2 slabs) in turn; per "layer half" its 8 warps (warps 4..11 here: any 8) each write one 4 KB half tile
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            if (warp >= 8) return;
            const int q = warp & 3, w2 = warp >> 2;
            uint4 *buf = reinterpret_cast<uint4 *>(sm + warp * 4096);
            const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
            const size_t slab_bytes = 70 * 8192, n_tiles = wbytes / (2 * slab_bytes);
            for (size_t t = blockIdx.x; t < n_tiles; t += n_writers) {
                for (int blk = 0; blk < 70; blk += 2) {                  // 35 "layer halves" of 2 blocks (128 features)
                    unsigned char *dst = reinterpret_cast<unsigned char *>(w) + ((2 * t + (q >> 1)) * 70 + blk + w2) * 8192 + (q & 1) * 4096;
                    if (MODE == ST_TILE) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) reinterpret_cast<uint4 *>(dst)[j * 32 + lane] = v;
                    } else if (MODE == ST_TILE_LANEROW) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) reinterpret_cast<uint4 *>(dst)[lane * 8 + (j ^ (lane & 7))] = v;
                    } else {
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; ++j) buf[lane * 8 + (j ^ (lane & 7))] = v;
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(dst), "r"(smem_u32(buf)) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                }
            }
            if (MODE == ST_TILE_BULK && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            return;
        }
        const size_t n = wbytes / 16, stride = (size_t)n_writers * blockDim.x;
        const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
        if (MODE == ST_V8 || MODE == ST_EVICT_FIRST) {
            const size_t n8 = wbytes / 32;
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
                if (MODE == ST_V8)
                    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %1, %2, %3, %4};" ::"l"(reinterpret_cast<unsigned char *>(w) + i * 32), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                else
                    asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1, %2, %3, %4, %1, %2, %3, %4};" ::"l"(reinterpret_cast<unsigned char *>(w) + i * 32), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
            return;
        }
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            if (MODE == ST_PLAIN) w[i] = v;
            else asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(w + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    } else {
        const int n_readers = gridDim.x - n_writers, b = blockIdx.x - n_writers;
        const size_t n = rbytes / 16, stride = (size_t)n_readers * blockDim.x;
        unsigned long long acc = 0;
        for (size_t i = (size_t)b * blockDim.x + threadIdx.x; i < n; i += stride) {
            uint4 x;
            asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(r + i));
            acc += x.x ^ x.y ^ x.z ^ x.w;
        }
        if (acc == 0x1234567ull) *sink = acc;
    }
}

template <int MODE>
static void run(uint4 *w, const uint4 *r, size_t bytes, int writers, int readers, double read_frac, unsigned long long *sink)
{
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const size_t wb = writers ? bytes : 0, rb = readers ? (size_t)(bytes * read_frac) / 4096 * 4096 : 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        probe<MODE><<<writers + readers, 512, MODE == ST_BULK ? 12 * 16384 : 32768>>>(w, wb, r, rb, writers, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-40s writers %3d readers %3d | %7.3f ms | write %6.2f TB/s  read %6.2f TB/s  total %6.2f TB/s %s\n", kNames[MODE], writers, readers, best,
           wb / (best * 1e-3) / 1e12, rb / (best * 1e-3) / 1e12, (wb + rb) / (best * 1e-3) / 1e12, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    const size_t bytes = 4ull << 30;
    uint4 *w, *r; unsigned long long *sink;
    cudaMalloc(&w, bytes); cudaMalloc(&r, bytes); cudaMalloc(&sink, 8);
    cudaMemset(r, 1, bytes);
    run<ST_PLAIN>(w, r, bytes, 148, 0, 0, sink);
    run<ST_CS>(w, r, bytes, 148, 0, 0, sink);
    run<ST_V8>(w, r, bytes, 148, 0, 0, sink);
    run<ST_EVICT_FIRST>(w, r, bytes, 148, 0, 0, sink);
    run<ST_BULK>(w, r, bytes, 148, 0, 0, sink);
    const size_t tiled = bytes / (2 * 70 * 8192) * (2 * 70 * 8192);
    run<ST_TILE>(w, r, tiled, 148, 0, 0, sink);
    run<ST_TILE_BULK>(w, r, tiled, 148, 0, 0, sink);
    run<ST_TILE_LANEROW>(w, r, tiled, 148, 0, 0, sink);
    return 0;
}
