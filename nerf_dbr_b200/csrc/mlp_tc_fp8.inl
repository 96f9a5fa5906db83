// FP8 mode of the fused render kernel (included by mlp_tc.cu inside namespace nerfb200::tc: it shares the front / back
// device code -- rays, depths, encoding, per-ray colour bias, colour head, compositing -- with the BF16 kernel and reads
// the same table offsets from the FP8 buffer's fp32 region, fp8_layout.h).
//
// Same persistent structure as the BF16 kernel: warp 0 streams weights, warp 1 issues the MMAs of a tile as
// straight-line code, warps 4-11 run the trunk epilogue on TMEM, warps 12-15 front / back; two 256-column TMEM
// regions alternate per layer, activations are written back in place and consumed from TMEM.  What changes:
//   * 256-wide contractions are tcgen05.mma.kind::f8f6f4 on e4m3 operands, K = 32 per instruction: a 64-wide K-block
//     of the previous layer is 16 packed TMEM columns and 2 instructions (128 tensor-pipe cycles instead of 256);
//     layer 0 and layer 4's skip part stay kind::f16 on the bf16 encoded position.
//   * the weight stream is 17 stages per tile (548 KB instead of 1 MiB); 17 is odd, so ring slot and barrier parity
//     of a stage are computed from a running stage count instead of being immediates.
//   * the epilogue is acc * m[n] + b'[n] (one packed FMA per column pair, tables in shared memory), ReLU + e4m3
//     conversion with saturation (cvt.rn.satfinite.relu.e4m3x2.f32), four values per TMEM column.
// The group schedule of a tile -- (layer, half, A K-block) in issue order -- is the BF16 kernel's chunk table.

constexpr uint32_t SM_QMUL = SM_STAGE;                     // [8][256] f32 multipliers (the TRAIN staging area is unused here)
static_assert(8 * 256 * 4 <= 16384, "multiplier table fits the staging area");

// D[tmem] (+)= A[tmem, e4m3 x 4 per column] * B[smem, e4m3]^T, K = 32
__device__ __forceinline__ void mma_e4m3_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// .kind::f8f6f4 instruction descriptor: D = f32, A = B = e4m3 (format 0), K-major
__host__ __device__ constexpr uint32_t idesc_e4m3(uint32_t m, uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24); }

// four fp32 -> one TMEM column of e4m3 (lowest k in the lowest byte), with ReLU and saturation
__device__ __forceinline__ uint32_t relu_pack_e4m3x4(float a, float b, float c, float d)
{
    uint32_t r;
    asm("{\n\t.reg .b16 lo, hi;\n\t"
        "cvt.rn.satfinite.relu.e4m3x2.f32 lo, %2, %1;\n\t"
        "cvt.rn.satfinite.relu.e4m3x2.f32 hi, %4, %3;\n\t"
        "mov.b32 %0, {lo, hi};\n\t}"
        : "=r"(r) : "f"(a), "f"(b), "f"(c), "f"(d));
    return r;
}
// (x0, x1) = (x0, x1) * (m0, m1) + (b0, b1) as one packed fp32x2 FMA
__device__ __forceinline__ void fma2(float &x0, float &x1, float m0, float m1, float b0, float b1)
{
    asm("{\n\t.reg .b64 a, m, b, d;\n\t"
        "mov.b64 a, {%2, %3};\n\tmov.b64 m, {%4, %5};\n\tmov.b64 b, {%6, %7};\n\t"
        "fma.rn.f32x2 d, a, m, b;\n\t"
        "mov.b64 {%0, %1}, d;\n\t}"
        : "=f"(x0), "=f"(x1) : "f"(x0), "f"(x1), "f"(m0), "f"(m1), "f"(b0), "f"(b1));
}

// stage of the FP8 stream that group CI reads, byte offset of its operand inside the stage, first / last use of the stage
__host__ __device__ constexpr int q_group_stage(int ci)
{
    const ChunkInfo c = kChunks.c[ci];
    if (c.layer == 0) return 0;
    if (c.layer == 8) return 16;
    const int first = q_layer_stage(c.layer);
    if (c.asrc == 4) return first;                          // layer 4's skip stage
    return first + (c.layer == 4 ? 1 : 0) + (c.asrc >> 1);
}
__host__ __device__ constexpr uint32_t q_group_offset(int ci)
{
    const ChunkInfo c = kChunks.c[ci];
    if (c.layer == 0 || c.asrc == 4) return (uint32_t)c.half * kChunkBytes;                      // bf16 chunk of the half
    if (c.layer == 8) return (uint32_t)(c.asrc >> 1) * (kC0Rows * 128) + (uint32_t)(c.asrc & 1) * 64u;
    return (uint32_t)c.half * kChunkBytes + (uint32_t)(c.asrc & 1) * 64u;                        // second K-block: +64 B in the rows
}
__host__ __device__ constexpr bool q_first_use(int ci)
{
    for (int j = 0; j < ci; ++j)
        if (q_group_stage(j) == q_group_stage(ci)) return false;
    return true;
}
__host__ __device__ constexpr bool q_last_use(int ci)
{
    for (int j = ci + 1; j < kChunksPerTile; ++j)
        if (q_group_stage(j) == q_group_stage(ci)) return false;
    return true;
}

struct IssueCtxQ {
    uint32_t bars;
    uint32_t region[2];
    uint32_t w_base;        // shared address of weight ring slot 0
    uint64_t pedesc;
    uint32_t pe_empty_bar;
    uint32_t sg0;           // running stage count at the start of this tile (17 per tile)
    int tile;
    uint16_t pair_mask;
    unsigned int *dbg;
    long long *trace;       // this tile's trace rows or nullptr (same slots as the BF16 kernel, tools/tc_trace.py)
};

template <int CI>
__device__ __forceinline__ void issue_group_fp8(const IssueCtxQ &x)
{
    constexpr ChunkInfo c = kChunks.c[CI];
    constexpr int stage = q_group_stage(CI);
    constexpr uint32_t off = q_group_offset(CI);
    constexpr bool bf16_group = c.layer == 0 || c.asrc == 4;
    const uint32_t sg = x.sg0 + (uint32_t)stage, slot = sg & 3u;
    if constexpr (q_first_use(CI)) wait_bar(x.bars + 8u * (B_WFULL + slot), (sg >> 2) & 1u, x.dbg, 4);
    if constexpr ((c.flags & 4) != 0) wait_bar(x.bars + 8u * (B_AREADY + c.asrc), (c.layer - 1) & 1, x.dbg, 3);
    if constexpr (c.layer == 1 && (CI == 0 || kChunks.c[CI > 0 ? CI - 1 : 0].layer != c.layer)) {
        if (x.tile > 0) wait_bar(x.bars + 8u * B_C0FREE, (x.tile - 1) & 1, x.dbg, 10);
    }
    tc_fence_after_sync();
    if (elect_one()) {
        const uint32_t d_tmem = x.region[c.layer & 1] + c.half * 128;
        const uint64_t bdesc = smem_desc_sw128(x.w_base + slot * (uint32_t)kStageSlotBytes + off);
        if constexpr (bf16_group) {
            constexpr uint32_t idesc = idesc_bf16(128, 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)             // 4 x (K = 16 bf16): +32 B inside the 128 B swizzle span
                mma_bf16_ss(d_tmem, x.pedesc + 2 * k, bdesc + 2 * k, idesc, !((c.flags & 1) && k == 0));
        } else {
            constexpr uint32_t idesc = c.layer == 8 ? idesc_e4m3(128, kC0Rows) : idesc_e4m3(128, 128);
            // A = K-block `asrc` of the previous layer: 64 e4m3 = 16 packed columns at the start of its 64-column range
            const uint32_t a_tmem = x.region[(c.layer & 1) ^ 1] + c.asrc * 64;
#pragma unroll
            for (int k = 0; k < 2; ++k)             // 2 x (K = 32 e4m3): +32 B in the rows, +8 columns in TMEM
                mma_e4m3_ts(d_tmem, a_tmem + 8 * k, bdesc + 2 * k, idesc, !((c.flags & 1) && k == 0));
        }
        if constexpr (CI == 0 || kChunks.c[CI > 0 ? CI - 1 : 0].layer != c.layer) { if (x.trace) x.trace[c.layer * 8 + 0] = clock64(); }
        if constexpr ((c.flags & 2) != 0) {
            mma_commit(x.bars + 8u * (c.layer == 8 ? B_ACCC0 : B_ACCFULL + c.half));
            if (x.trace) x.trace[c.layer * 8 + ((CI + 1 == kChunksPerTile || kChunks.c[CI + 1 < kChunksPerTile ? CI + 1 : CI].layer != c.layer) ? 2 : 1)] = clock64();
        }
        if constexpr (q_last_use(CI)) {
            if (x.pair_mask) mma_commit_mcast(x.bars + 8u * (B_WEMPTY + slot), x.pair_mask);
            else mma_commit(x.bars + 8u * (B_WEMPTY + slot));
        }
        // the encoded position is last read by layer 4's second skip group
        if constexpr (c.layer == 4 && c.asrc == 4 && c.half == 1) mma_commit(x.pe_empty_bar);
    }
    __syncwarp();
}
template <int... CI>
__device__ __forceinline__ void issue_tile_fp8(const IssueCtxQ &x, std::integer_sequence<int, CI...>)
{
    (issue_group_fp8<CI>(x), ...);
}

// epilogue of one N = 128 half for one warp (32 rows x 64 columns): acc * m + b', ReLU, e4m3, 16 packed columns back
// over the start of the columns just read = one 64-wide K-block of the next layer's A operand
__device__ __forceinline__ void epilogue_half_fp8(uint32_t t_cols, uint32_t mul_addr, uint32_t bias_addr, uint32_t bar_ready, int lane)
{
    uint32_t xa[32], xb[32], pk[16];
    tmem_ld32(t_cols, xa);
    tmem_ld32(t_cols + 32, xb);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 m = ld_shared_f4(mul_addr + 16 * i), b = ld_shared_f4(bias_addr + 16 * i);
        float x0 = __uint_as_float(xa[4 * i + 0]), x1 = __uint_as_float(xa[4 * i + 1]);
        float x2 = __uint_as_float(xa[4 * i + 2]), x3 = __uint_as_float(xa[4 * i + 3]);
        fma2(x0, x1, m.x, m.y, b.x, b.y);
        fma2(x2, x3, m.z, m.w, b.z, b.w);
        pk[i] = relu_pack_e4m3x4(x0, x1, x2, x3);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 m = ld_shared_f4(mul_addr + 128 + 16 * i), b = ld_shared_f4(bias_addr + 128 + 16 * i);
        float x0 = __uint_as_float(xb[4 * i + 0]), x1 = __uint_as_float(xb[4 * i + 1]);
        float x2 = __uint_as_float(xb[4 * i + 2]), x3 = __uint_as_float(xb[4 * i + 3]);
        fma2(x0, x1, m.x, m.y, b.x, b.y);
        fma2(x2, x3, m.z, m.w, b.z, b.w);
        pk[8 + i] = relu_pack_e4m3x4(x0, x1, x2, x3);
    }
    tmem_st16(t_cols, pk);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_ready);
}

template <int SRC>
__global__ void __launch_bounds__(kThreads, 1) fused_render_fp8_kernel(const Args a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_base = smem_u32(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bars = sm_base + SM_BAR;
    auto bar = [&](int i) { return bars + 8u * i; };
    const float *wf = reinterpret_cast<const float *>(a.packed);            // the FP8 buffer's fp32 region
    const unsigned char *wq = a.packed + Q_OFFSET;

    const int tile_begin = blockIdx.x * a.tiles_per_cta;
    const int tile_end = min(a.n_tiles, tile_begin + a.tiles_per_cta);
    const bool pair = a.pair > 1;
    const int my_tiles = pair ? a.tiles_per_cta : max(0, tile_end - tile_begin);
    const uint32_t cta_rank = pair ? cluster_ctarank() : 0u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), pair ? a.pair : 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(B_PEFULL + i), 4); mbar_init(bar(B_PEEMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) mbar_init(bar(B_ACCFULL + i), 1);
        mbar_init(bar(B_ACCC0), 1);
        mbar_init(bar(B_C0FREE), 4);
        for (int i = 0; i < 4; ++i) mbar_init(bar(B_AREADY + i), 4);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(sm_base + SM_TMEM);
    {
        float *bias = reinterpret_cast<float *>(sm + SM_BIAS), *mul = reinterpret_cast<float *>(sm + SM_QMUL);
        for (int i = threadIdx.x; i < 8 * 256; i += kThreads) { bias[i] = __ldg(wf + F_BIAS + i); mul[i] = __ldg(wf + Q_MUL + i); }
        float *wc1 = reinterpret_cast<float *>(sm + SM_WC1);
        for (int i = threadIdx.x; i < 384; i += kThreads) wc1[i] = __ldg(wf + F_WC1 + i);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (pair) cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + SM_TMEM);

    if (warp == 0) {
        // ================================ weight producer ===================================
        if (lane == 0) {
            uint32_t sg = 0;
            for (int t = 0; t < my_tiles; ++t) {
                for (int st = 0; st < kQStages; ++st, ++sg) {
                    const uint32_t slot = sg & 3u, round = sg >> 2;
                    if (round > 0) wait_bar(bar(B_WEMPTY + slot), (round - 1) & 1, a.dbg, 1);
                    const uint32_t bytes = q_stage_bytes(st), dst = sm_base + SM_W + slot * kStageSlotBytes;
                    mbar_arrive_expect_tx(bar(B_WFULL + slot), bytes);
                    if (pair) {
                        const uint32_t piece = bytes / (uint32_t)a.pair, off = cta_rank * piece;
                        bulk_g2s_mcast(dst + off, wq + q_stage_offset(st) + off, piece, bar(B_WFULL + slot), (uint16_t)((1u << a.pair) - 1u));
                        continue;
                    }
                    bulk_g2s(dst, wq + q_stage_offset(st), bytes, bar(B_WFULL + slot));
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer =========================================
        IssueCtxQ x;
        x.bars = bars;
        x.pair_mask = pair ? (uint16_t)((1u << a.pair) - 1u) : (uint16_t)0;
        x.w_base = sm_base + SM_W;
        x.dbg = a.dbg;
        for (int t = 0; t < my_tiles; ++t) {
            const int pb = t & 1, pe_use = t >> 1;
            wait_bar(bar(B_PEFULL + pb), pe_use & 1, a.dbg, 2);
            x.region[0] = tmem_base + (uint32_t)(t & 1) * 256;
            x.region[1] = tmem_base + (uint32_t)((t & 1) ^ 1) * 256;
            x.pedesc = smem_desc_sw128(sm_base + SM_PE + pb * 16384);
            x.pe_empty_bar = bar(B_PEEMPTY + pb);
            x.sg0 = (uint32_t)t * kQStages;
            x.tile = t;
            x.trace = (a.trace && blockIdx.x == 0 && t < kTraceTiles) ? a.trace + t * 72 : nullptr;
            issue_tile_fp8(x, std::make_integer_sequence<int, kChunksPerTile>{});
        }
    } else if (warp >= 4 && warp < 12) {
        // ================================ epilogue ===========================================
        const int ew = warp - 4, q = ew & 3, w2 = ew >> 2;
        uint32_t g = 0;
        for (int t = 0; t < my_tiles; ++t) {
            long long *tr = (a.trace && blockIdx.x == 0 && t < kTraceTiles && ew == 0 && lane == 0) ? a.trace + t * 72 : nullptr;
            for (int layer = 0; layer < 8; ++layer, ++g) {
                const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (g & 1) * 256 + 64 * w2;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    wait_bar(bar(B_ACCFULL + hh), layer & 1, a.dbg, 5);
                    tc_fence_after_sync();
                    if (tr) tr[layer * 8 + (hh == 0 ? 3 : 6)] = clock64();
                    const uint32_t tab = (uint32_t)(layer * 256 + hh * 128 + 64 * w2) * 4u;
                    epilogue_half_fp8(t_lane + hh * 128, sm_base + SM_QMUL + tab, sm_base + SM_BIAS + tab, bar(B_AREADY + 2 * hh + w2), lane);
                    if (tr && hh == 0) tr[layer * 8 + 4] = clock64();
                }
                if (tr) tr[layer * 8 + 5] = clock64();
            }
            ++g;                                    // colour layer 0 takes a region turn too
        }
    } else if (warp >= 12) {
        // ================================ front / back =======================================
        const int row = (warp - 12) * 32 + lane;
        const float step = linspace_step(a.n_samples);
        const float sig_inv = __ldg(wf + Q_SIGINV);
        auto produce = [&](int t) {
            const int pb = t & 1, pe_use = t >> 1;
            if (pe_use >= 1) wait_bar(bar(B_PEEMPTY + pb), (pe_use - 1) & 1, a.dbg, 8);
            produce_tile<SRC, false, false>(a, sm, tile_begin + t, pb, t & 1, row, step, wf);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_PEFULL + pb));
        };
        if (my_tiles > 0) produce(0);
        for (int t = 0; t < my_tiles; ++t) {
            if (t + 1 < my_tiles) produce(t + 1);
            const int pb = t & 1;
            wait_bar(bar(B_ACCC0), t & 1, a.dbg, 11);
            tc_fence_after_sync();
            const int rpt_shift = a.tiles_per_ray == 1 ? a.s_pad_log2 : 7;
            const uint32_t t_row = tmem_base + ((uint32_t)((warp - 12) * 32) << 16) + (uint32_t)(t & 1) * 256;
            float ypre[3], sig_pre;
            // colour layer 0's column scales are folded into the per-ray bias (x s) and W_c1 (/ s): same code as BF16 mode
            color_row(t_row, sm_base + SM_RAYB + (pb * kMaxRaysPerTile + (row >> rpt_shift)) * 512, sm_base + SM_WC1,
                      bar(B_C0FREE), lane, ypre[0], ypre[1], ypre[2], sig_pre);
            composite_tile<SRC>(a, sm, tile_begin + t, row, step, wf, sig_pre * sig_inv, ypre);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (pair) cluster_sync_all();
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}
