// tcgen05.mma.kind::f8f6f4 operand layouts on sm_100a, checked against the host before the FP8 render mode relies on them:
//   D[128 x 128] f32 (TMEM) = A[128 x 128] e4m3 (TMEM, 4 consecutive-k values per 32-bit column, written with
//   tcgen05.st) x B[128 n x 128 k]^T e4m3 (shared memory, K-major rows of 128 B, 128-byte swizzle -- the same
//   descriptor as a [128 x 64] bf16 chunk), as four K = 32 instructions (+32 B / +8 columns per step).
// Data are small integers (exact in e4m3 and in the fp32 sum), so the result must match the host bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/fp8_probe tools/probe/fp8_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_fp8.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t a)
{
    return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// .kind::f8f6f4: [4,6) D = f32 (1), [7,10) A format e4m3 (0), [10,13) B format e4m3 (0), [17,23) N >> 3, [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_e4m3(uint32_t m, uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24); }

__global__ void __launch_bounds__(128, 1) fp8_probe_kernel(const uint8_t *a_g, const uint8_t *b_g, float *d_g, int a_in_smem)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sb = sm, *sa = sm + 16384;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 32768);
    uint32_t *tslot = reinterpret_cast<uint32_t *>(sm + 32768 + 64);
    const int t = threadIdx.x, warp = t >> 5;
    // B (and A for the SS form): element (row, k) at row * 128 + (((k >> 4) ^ (row & 7)) << 4 | (k & 15))
    for (int i = t; i < 128 * 128; i += 128) {
        const int n = i >> 7, k = i & 127;
        const int off = n * 128 + ((((k >> 4) ^ (n & 7)) << 4) | (k & 15));
        sb[off] = b_g[i];
        sa[off] = a_g[i];
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(tslot);
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    // A in TMEM: thread t = row t; column c (of 32, at +128) = k 4c .. 4c+3, lowest k in the lowest byte
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) {
        uint32_t w = 0;
        for (int j = 0; j < 4; ++j) w |= (uint32_t)a_g[t * 128 + 4 * c + j] << (8 * j);
        v[c] = w;
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(t_lane + 128), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (t == 0) {
        const uint64_t bdesc = smem_desc_sw128(smem_u32(sb)), adesc = smem_desc_sw128(smem_u32(sa));
        const uint32_t idesc = idesc_e4m3(128, 128);
        for (int k = 0; k < 4; ++k) {
            const uint32_t acc = k > 0;
            if (a_in_smem)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(acc) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(tmem), "r"(tmem + 128 + 8 * k), "l"(bdesc + 2 * k), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(t_lane + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 32; ++c) d_g[t * 128 + c0 + c] = __uint_as_float(r[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main()
{
    static uint8_t a[128 * 128], b[128 * 128];
    static float fa[128 * 128], fb[128 * 128], ref[128 * 128];
    srand(7);
    const float vals[7] = {-3.f, -2.f, -1.f, 0.f, 1.f, 2.f, 0.5f};
    for (int i = 0; i < 128 * 128; ++i) {
        fa[i] = vals[rand() % 7]; fb[i] = vals[rand() % 7];
        a[i] = (uint8_t)__nv_cvt_float_to_fp8(fa[i], __NV_SATFINITE, __NV_E4M3);
        b[i] = (uint8_t)__nv_cvt_float_to_fp8(fb[i], __NV_SATFINITE, __NV_E4M3);
    }
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            float s = 0.f;
            for (int k = 0; k < 128; ++k) s += fa[m * 128 + k] * fb[n * 128 + k];
            ref[m * 128 + n] = s;
        }
    uint8_t *da, *db;
    float *dd;
    cudaMalloc(&da, sizeof(a)); cudaMalloc(&db, sizeof(b)); cudaMalloc(&dd, sizeof(ref));
    cudaMemcpy(da, a, sizeof(a), cudaMemcpyHostToDevice); cudaMemcpy(db, b, sizeof(b), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(fp8_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 36 * 1024);
    static float out[128 * 128];
    for (int ss = 0; ss < 2; ++ss) {
        cudaMemset(dd, 0, sizeof(ref));
        fp8_probe_kernel<<<1, 128, 36 * 1024>>>(da, db, dd, ss);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s form: %s\n", ss ? "SS" : "TS", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out, dd, sizeof(out), cudaMemcpyDeviceToHost);
        int bad = 0;
        double worst = 0;
        for (int i = 0; i < 128 * 128; ++i) { double d = fabs((double)out[i] - ref[i]); if (d > 0) ++bad; if (d > worst) worst = d; }
        printf("%s form (A in %s): %d of 16384 elements differ from the host, worst %.3f   D[0][0..3] = %.1f %.1f %.1f %.1f (host %.1f %.1f %.1f %.1f)\n",
               ss ? "SS" : "TS", ss ? "shared memory" : "tensor memory", bad, worst, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
    }
    return 0;
}
