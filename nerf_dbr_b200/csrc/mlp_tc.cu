// BF16 mode: the fused render kernel on 5th-generation tensor cores.
//
//   rays -> uniform/stratified depths -> positional encoding -> 8x256 MLP (+ heads) -> alpha
//   compositing, one persistent CTA per SM, per-sample activations never leave the SM.
//
// Per 128-sample tile the CTA runs nine layers as tcgen05.mma (M=128, N=128, K=16 bf16, fp32
// accumulators in TMEM): L0 (K=64: encoded position), L1-L3 (K=256), L4 (K=64 encoded position +
// K=256 hidden -- the skip is a second accumulate, not a concat), L5-L7, C0 (N=128).
//   * ACTIVATIONS LIVE IN TENSOR MEMORY.  Each layer's 256 fp32 accumulator columns are two
//     N=128 halves.  The epilogue reads a half (tcgen05.ld), adds bias, applies ReLU, packs to
//     bf16 and writes it back IN PLACE (tcgen05.st) over the columns it has just read; the next
//     layer's MMA takes that as its A operand straight from TMEM (no shared-memory round trip,
//     which is what bounded the first version of this kernel: A-operand reads + epilogue stores
//     + weight fills exceeded the 128 B/clk shared-memory port).
//   * Two 256-column TMEM regions alternate per layer (the MMA of layer l+1 writes the other
//     region while layer l's activations are read from this one).
//   * The chunk order (packed_layout.h: h0k0 h0k1 h1k0 h0k2 h0k3 h1k1 h1k2 h1k3) lets half 0's
//     epilogue overlap the same layer's remaining MMAs and half 1's the next layer's first three
//     chunks, which hides the MMA -> epilogue -> MMA latency chain (~750 cycles of slack).
//   * The MMA issuer is straight-line code: the whole per-tile schedule (64 chunks, 256 MMAs) is
//     unrolled at compile time from the constexpr chunk table, every barrier parity, ring slot and
//     operand offset an immediate -- a table-driven loop issues ~1 MMA per 190 cycles, this one
//     keeps up with the pipe's 64 cycles per instruction.
//   * B operand: [128 x 64] bf16 weight chunks, pre-swizzled and stored in consumption order by
//     pack.cu, streamed L2 -> shared memory as 32 KB cp.async.bulk stages (the TMA engine)
//     through a 4-stage mbarrier ring.
//   * The encoded position (bf16, 128B-swizzled K-major [128 x 64] tile) is the only A operand
//     in shared memory (L0 and the skip part of L4).
//   * The view direction enters colour layer 0 as an fp32 per-ray bias (W_dir . enc(d) + b),
//     computed once per ray on CUDA cores: no per-sample direction encoding at all.
//   * The density head rides in colour layer 0's GEMM as output column 128 (N = 144; bf16 like every
//     other layer: < 0.01 dB against computing it in fp32, tests/diag/emulate_bf16.py) instead of 256 CUDA-core
//     FMAs per sample.  Colour layer 0's epilogue -- relu(acc + per-ray bias) . W_c1 -> sigmoid, sigma =
//     relu(col 128 + b) -- belongs to the back warps, which read the accumulator straight from TMEM and
//     then composite (segmented warp scan of transmittance products); the epilogue warps go from layer 7
//     directly to the next tile's layer 0.  The front/back warps also generate the next tile's rays,
//     depths and encodings.
//
// Warp roles (512 threads): 0 weight producer | 1 MMA issuer | 2 TMEM allocator | 4-11 epilogue
// (lane quadrant = warp % 4, 64-column part of a half = (warp-4)/4) | 12-15 front (encode) /
// back (composite).
//
// reference: PyTorchCPURenderer.render_image / _render_ray_chunk (src/benchmark/
// pytorch_renderers.py:127-170), NeRFModel.forward (src/models/nerf.py:92-131),
// execute_volume_rendering (pytorch_renderers.py:105-125).
#include <cstdlib>
#include <utility>
#include "common.cuh"
#include "ptx.cuh"
#include "train_layout.h"
#include "fp8_layout.h"

namespace nerfb200 {
namespace tc {

using namespace ptx;

constexpr int kThreads = 512;
constexpr int kTileM = 128;
constexpr int kMaxRaysPerTile = 8;      // S_pad >= 16

// Two instantiations share one shared-memory map:
//   SPLIT = false  bf16 operands; 4 weight slots of 36 KB; two single-plane encoded-position buffers
//   SPLIT = true   every operand is a bf16 (hi, lo) pair and every product three MMAs
//                  (hi*hi + hi*lo + lo*hi): fp32-class accuracy on the tensor cores (max-abs ~1e-5 against
//                  the fp32 reference).  2 weight slots of 72 KB (hi stage | lo stage), one two-plane
//                  encoded-position buffer; activations keep their lo halves in the 32 TMEM columns per
//                  K-block that the bf16 mode leaves unused.
template <bool SPLIT>
struct Cfg {
    static constexpr int kWSlots = SPLIT ? 2 : 4;
    static constexpr uint32_t kSlotBytes = (SPLIT ? 2u : 1u) * kStageSlotBytes;
    static constexpr int kPeBufs = SPLIT ? 1 : 2;
    static constexpr int kMmaPerStep = SPLIT ? 3 : 1;
};
static_assert(kStagesPerTile % 4 == 0, "ring slots and barrier parities are compile-time per tile");
constexpr int kMaxWSlots = 4;

// shared memory map (bytes from a 1024-aligned base)
constexpr uint32_t SM_PE = 0;                              // 32 KB: 2 x [128 x 64] bf16, or (hi, lo) planes of one tile
constexpr uint32_t SM_W = 32768;                           // 144 KB of weight stages
constexpr uint32_t SM_BIAS = SM_W + 4 * kStageSlotBytes;   // [8][256] f32
constexpr uint32_t SM_WC1 = SM_BIAS + 8192;                // [3][128] f32
constexpr uint32_t SM_RAYB = SM_WC1 + 1536;                // [2][8][128] f32  per-ray colour-0 bias
constexpr uint32_t SM_DE = SM_RAYB + 8192;                 // [8][32] f32      direction encodings
constexpr uint32_t SM_SCR = SM_DE + 1024;                  // back-warp scratch (256 B)
constexpr uint32_t SM_BAR = SM_SCR + 256;                  // mbarriers
constexpr uint32_t SM_TMEM = SM_BAR + 256;
constexpr uint32_t SM_STAGE = SM_TMEM + 256;               // TRAIN: six 4 KB store staging buffers (epilogue warps 0-5; warps 6, 7 use
                                                           // the 4 KB tails of weight slots 0 and 1, which no 36 KB stage ever lands in)
constexpr uint32_t SM_TOTAL = SM_STAGE + 24576;
static_assert((kStagesPerTile - kStagesC0) % 4 == 2, "the two 36 KB colour-layer-0 stages land in ring slots 2 and 3");
constexpr uint32_t kSmemBytes = SM_TOTAL + 1024;           // + alignment slack
static_assert(kSmemBytes <= 232448, "shared memory budget");

// barrier indices
enum { B_WFULL = 0, B_WEMPTY = B_WFULL + kMaxWSlots, B_PEFULL = B_WEMPTY + kMaxWSlots, B_PEEMPTY = B_PEFULL + 2,
       B_ACCFULL = B_PEEMPTY + 2, B_AREADY = B_ACCFULL + 2, B_ACCC0 = B_AREADY + 4, B_C0FREE = B_ACCC0 + 1,
       B_COUNT = B_C0FREE + 1 };
// Phase discipline (mbarrier parity waits are only sound while the producer is at most ONE phase ahead of
// every waiter): acc_full[h] completes once per trunk layer 0..7 and each completion needs the previous
// layer's epilogue; colour layer 0 has its own barrier because the NEXT tile's layer 0 follows it with no
// epilogue in between -- sharing acc_full[0] let the barrier run two phases ahead of a slow epilogue warp.
static_assert(B_COUNT * 8 <= 256, "barrier area");

constexpr ChunkTable kChunks = make_chunk_table();

enum { SRC_RAYS = 1, SRC_POSE = 2, SRC_POINTS = 3 };   // SRC_POINTS: query_network -- a row is one (point, direction) pair

struct Args {
    const unsigned char *packed;
    Pose pose;
    int width, row0;
    float half_w, half_h, focal;
    const float *rays_o, *rays_d, *t_rand;
    const float *points, *dirs;                // SRC_POINTS: [n_points,3] each; outputs sigma_out [n_points], rgb_out [n_points,3]
    float *sigma_out, *rgb_out;
    int n_points;
    const float *z_vals;       // optional explicit depths [n_rays, n_samples] (ascending per ray)
    float *weights;            // optional per-sample compositing weights out [n_rays, n_samples]
    int n_rays, n_samples;
    int s_pad_log2;            // S_pad = 1 << s_pad_log2 when tiles_per_ray == 1
    int tiles_per_ray;         // > 1 when S_pad > 128
    int n_tiles, tiles_per_cta;
    float near, far;
    float *rgb_map, *depth, *acc;
    float *ws;                 // TRAIN: training workspace [R_TOTAL][ws_ch] (train_layout.h); rays are chunk-local
    int ws_ch;
    unsigned int *dbg;         // optional: [0] = first timeout code
    int pair;                  // cluster size 1, 2 or 4: the CTAs of a cluster share the weight stream -- each loads 1/size of
                               // every stage and multicasts it to all of them (1/size of the L2 -> shared-memory reads per SM)
    int sm_limit;              // 0 = every SM; else at most this many CTAs (the caller runs something else beside)
    long long *trace;          // optional timeline (tools/tc_trace.py): CTA 0, first kTraceTiles tiles
};
constexpr int kTraceTiles = 6;
// trace layout: [tile][layer 0..8][slot 0..7] clock64 stamps
//   0 MMA: first chunk of the layer issued   1 MMA: half 0 committed   2 MMA: last half committed
//   3 EPI(warp 4): half 0 acc_full seen   4 EPI: half 0 a_ready arrive   5 EPI: layer done
//   6 EPI: last half acc_full seen

__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity, unsigned int *dbg, uint32_t code)
{
    mbar_wait_bounded(bar, parity, dbg, 0x80000000u, code);
}

struct RowInfo { int ray, s; bool valid; };

__device__ __forceinline__ RowInfo row_info(const Args &a, int tile, int row)
{
    RowInfo r;
    if (a.tiles_per_ray == 1) {
        int rpt_log2 = 7 - a.s_pad_log2;
        r.ray = (tile << rpt_log2) + (row >> a.s_pad_log2);
        r.s = row & ((1 << a.s_pad_log2) - 1);
    } else {
        r.ray = tile / a.tiles_per_ray;
        r.s = (tile - r.ray * a.tiles_per_ray) * kTileM + row;
    }
    r.valid = r.ray < a.n_rays && r.s < a.n_samples;
    return r;
}

// TRAIN: column of the workspace this tile row maps to (-1: padding row)
__device__ __forceinline__ int ws_col(const Args &a, const RowInfo &ri) { return ri.valid ? ri.ray * a.n_samples + ri.s : -1; }

template <int SRC>
__device__ __forceinline__ void ray_of(const Args &a, int ray, float (&o)[3], float (&d)[3])
{
    if (SRC == SRC_POSE) {
        int j = a.row0 + ray / a.width, i = ray % a.width;
        float dx, dy;
        pixel_dir(i, j, a.half_w, a.half_h, a.focal, dx, dy);
#pragma unroll
        for (int c = 0; c < 3; ++c) { d[c] = rotate_dir(a.pose, c, dx, dy); o[c] = a.pose.t[c]; }
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) { d[c] = __ldg(a.rays_d + 3 * (size_t)ray + c); o[c] = __ldg(a.rays_o + 3 * (size_t)ray + c); }
    }
}

__device__ __forceinline__ float depth_of(const Args &a, int ray, int s, float step)
{
    if (a.z_vals) return __ldg(a.z_vals + (size_t)ray * a.n_samples + s);
    return a.t_rand ? depth_jittered(s, a.n_samples, step, a.near, a.far, __ldg(a.t_rand + (size_t)ray * a.n_samples + s))
                    : depth_uniform(s, a.n_samples, step, a.near, a.far);
}

// phase of fl(pi_f * x) in units of 2^-32 turns.  All ten frequencies are exact left shifts of it:
// fl(fl(2^k pi) x) == 2^k fl(pi_f x) because scaling by 2^k commutes with rounding.
__device__ __forceinline__ uint32_t phase_of(float x)
{
    float arg = __fmul_rn(kPiF, x);
    double turns = (double)arg * 0.15915494309189535;
    return (uint32_t)(long long)__double2ll_rn(turns * 4294967296.0);
}
__device__ __forceinline__ void sincos_phase(uint32_t phase, float &s, float &c)
{
    float r = (float)(int)phase * 1.4629180792671596e-9f;    // 2 pi / 2^32, r in [-pi, pi)
    s = __sinf(r);
    c = __cosf(r);
}

// ------------------------------------------------------------------------------------------
// front: rays, depths, points, encoded-position tile (bf16, swizzled) and per-ray colour bias
__device__ __forceinline__ float bf16_hi(float v) { return __uint_as_float(__float_as_uint(__bfloat162float(__float2bfloat16_rn(v)))); }

// pe_bulk (TRAIN): the caller copies the warp's 32 rows of the encoded-position tile to the workspace with one bulk copy
// (the shared-memory operand tile IS the workspace image of G_PE), so the per-row stores are skipped
template <int SRC, bool SPLIT, bool TRAIN>
__device__ void produce_tile(const Args &a, uint8_t *sm, int tile, int pe_buf, int rb_buf, int row, float step,
                             const float *__restrict__ wf, bool pe_bulk = false)
{
    RowInfo ri = row_info(a, tile, row);
    if (SRC == SRC_POINTS) ri.valid = tile * kTileM + row < a.n_points;
    float feat[64];
#pragma unroll
    for (int f = 0; f < 64; ++f) feat[f] = 0.f;
    if (ri.valid) {
        float o[3], d[3], z = 0.f;
        if (SRC != SRC_POINTS) {
            ray_of<SRC>(a, ri.ray, o, d);
            z = depth_of(a, ri.ray, ri.s, step);
        }
        uint32_t ph[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float p = SRC == SRC_POINTS ? __ldg(a.points + 3 * ((size_t)tile * kTileM + row) + c) : point_on_ray(o[c], d[c], z);
            feat[c] = p;
            ph[c] = phase_of(p);
        }
#pragma unroll
        for (int k = 0; k < kPosFreq; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float sn, cs;
                sincos_phase(ph[c] << k, sn, cs);
                feat[3 + 6 * k + c] = sn;
                feat[3 + 6 * k + 3 + c] = cs;
            }
    }
    if (TRAIN) {                                           // the bf16 values the MMA sees, [feature][sample] for wgrad
        const int col = ws_col(a, ri);
        if (col >= 0 && !pe_bulk) {
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) pk[i] = pack_bf16(feat[2 * i], feat[2 * i + 1]);
            store_block_row(a.ws, G_PE, col, pk);
        }
    }
    const uint32_t pe_row = smem_u32(sm + SM_PE + pe_buf * 16384 + row * 128);
#pragma unroll
    for (int u = 0; u < 8; ++u)
        st_shared_v4(pe_row + ((u ^ (row & 7)) << 4),
                     pack_bf16(feat[8 * u + 0], feat[8 * u + 1]), pack_bf16(feat[8 * u + 2], feat[8 * u + 3]),
                     pack_bf16(feat[8 * u + 4], feat[8 * u + 5]), pack_bf16(feat[8 * u + 6], feat[8 * u + 7]));
    if (SPLIT) {                                           // low-order plane: feat - bf16(feat)
#pragma unroll
        for (int f = 0; f < 64; ++f) feat[f] -= bf16_hi(feat[f]);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            st_shared_v4(pe_row + 16384 + ((u ^ (row & 7)) << 4),
                         pack_bf16(feat[8 * u + 0], feat[8 * u + 1]), pack_bf16(feat[8 * u + 2], feat[8 * u + 3]),
                         pack_bf16(feat[8 * u + 4], feat[8 * u + 5]), pack_bf16(feat[8 * u + 6], feat[8 * u + 7]));
    }

    if (SRC == SRC_POINTS) return;                         // per-row directions: the back warps build the colour-0 bias
    // direction encodings of the tile's rays: thread (q*32 + f) -> feature f of ray q (fp32, full-range sinf/cosf)
    const int rpt = a.tiles_per_ray == 1 ? (kTileM >> a.s_pad_log2) : 1;
    float *de = reinterpret_cast<float *>(sm + SM_DE);
    for (int q = row >> 5, f = row & 31; q < rpt; q += 4) {
        {
            int ray = a.tiles_per_ray == 1 ? tile * rpt + q : tile / a.tiles_per_ray;
            float v = 0.f;
            if (ray < a.n_rays && f < kDirFeat) {
                float o[3], d[3];
                ray_of<SRC>(a, ray, o, d);
                if (f < 3) v = d[f];
                else {
                    int k = (f - 3) / 6, w = (f - 3) - 6 * k, c = w % 3;
                    float arg = __fmul_rn(kPiF * (float)(1 << k), d[c]);
                    v = w < 3 ? sinf(arg) : cosf(arg);
                }
            }
            de[q * 32 + f] = v;
        }
    }
    named_bar_sync(1, 128);
    float *rayb = reinterpret_cast<float *>(sm + SM_RAYB) + rb_buf * (kMaxRaysPerTile * 128);
    for (int q = 0; q < rpt; ++q) {
        float acc = __ldg(wf + F_BC0 + row);
#pragma unroll
        for (int j = 0; j < kDirFeat; ++j) acc = fmaf(de[q * 32 + j], __ldg(wf + F_WC0D + j * 128 + row), acc);
        rayb[q * 128 + row] = acc;
    }
    if (TRAIN) {                                           // encoded direction of this row's ray (dW of colour layer 0)
        const int col = ws_col(a, ri);
        if (col >= 0) {
            const int q = a.tiles_per_ray == 1 ? (row >> a.s_pad_log2) : 0;
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i)
                pk[i] = i < 16 ? pack_bf16(2 * i < kDirFeat ? de[q * 32 + 2 * i] : 0.f, 2 * i + 1 < kDirFeat ? de[q * 32 + 2 * i + 1] : 0.f) : 0u;
            store_block_row(a.ws, G_DE, col, pk, 4);        // 32 features: wgrad multiplies N = 32 of this block
        }
    }
    named_bar_sync(1, 128);          // de[] is rewritten by the next produce
}

// ------------------------------------------------------------------------------------------
// back: sigmoid / relu heads, alpha, segmented transmittance scan, weighted sums, output
template <int SRC>
__device__ void composite_tile(const Args &a, uint8_t *sm, int tile, int row, float step,
                               const float *__restrict__ wf, float sig_pre, const float (&ypre)[3])
{
    const int lane = row & 31, warp = row >> 5;

    RowInfo ri = row_info(a, tile, row);
    float sigma = fmaxf(sig_pre + __ldg(wf + F_BSIG), 0.f);
    float col[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) col[c] = 1.0f / (1.0f + expf(-(ypre[c] + __ldg(wf + F_BC1 + c))));
    float alpha = 0.f, keep = 1.f, z = 0.f;
    if (ri.valid && a.n_samples > 1) {                  // S == 1 renders black in the reference (empty dists)
        float o[3], d[3];
        ray_of<SRC>(a, ri.ray, o, d);
        float dn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
        z = depth_of(a, ri.ray, ri.s, step);
        float dz = 1e10f;
        if (ri.s + 1 < a.n_samples) dz = __fsub_rn(depth_of(a, ri.ray, ri.s + 1, step), z);
        float dist = __fmul_rn(dz, dn);
        alpha = __fsub_rn(1.0f, expf(__fmul_rn(-sigma, dist)));
        keep = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
    }
    // segmented inclusive product scan over rows of one ray
    const int seg = a.tiles_per_ray == 1 ? (1 << a.s_pad_log2) : kTileM;   // rows per ray in this tile
    const int wseg = seg < 32 ? seg : 32;
    float incl = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float n = __shfl_up_sync(0xffffffffu, incl, o);
        if (o < wseg && (lane & (wseg - 1)) >= o) incl *= n;
    }
    float prev = __shfl_up_sync(0xffffffffu, incl, 1);
    float trans = (lane & (wseg - 1)) == 0 ? 1.0f : prev;
    float *scr = reinterpret_cast<float *>(sm + SM_SCR);          // [0..3] warp products, [4..9] carry, [16..] sums
    float *carry = scr + 4;
    if (seg > 32) {
        if (lane == 31) scr[warp] = incl;
        named_bar_sync(1, 128);
        const int w0 = seg >= kTileM ? 0 : (warp & ~1);           // first warp of this ray's segment
        float pre = 1.0f;
        for (int w = w0; w < warp; ++w) pre *= scr[w];
        if (a.tiles_per_ray > 1 && (tile % a.tiles_per_ray) != 0) pre *= carry[0];
        trans *= pre;
    }
    float w = __fmul_rn(alpha, trans);
    if (a.weights && ri.valid) a.weights[(size_t)ri.ray * a.n_samples + ri.s] = w;
    float sums[5] = {w * col[0], w * col[1], w * col[2], w * z, w};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
        if (o < wseg) {
#pragma unroll
            for (int i = 0; i < 5; ++i) sums[i] += __shfl_xor_sync(0xffffffffu, sums[i], o);
        }
    if (seg <= 32) {
        if ((lane & (wseg - 1)) == 0 && ri.ray < a.n_rays) {
            a.rgb_map[3 * (size_t)ri.ray + 0] = sums[0]; a.rgb_map[3 * (size_t)ri.ray + 1] = sums[1];
            a.rgb_map[3 * (size_t)ri.ray + 2] = sums[2];
            a.depth[ri.ray] = sums[3];
            if (a.acc) a.acc[ri.ray] = sums[4];
        }
    } else {
        float *ws = scr + 16 + warp * 5;
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 5; ++i) ws[i] = sums[i];
        }
        named_bar_sync(1, 128);
        const int wpr = seg >> 5;                                 // warps per ray: 2 or 4
        if (lane == 0 && (warp & (wpr - 1)) == 0 && ri.ray < a.n_rays) {
            float t[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                t[i] = 0.f;
                for (int k = 0; k < wpr; ++k) t[i] += scr[16 + (warp + k) * 5 + i];
            }
            bool first = true, last = true;
            if (a.tiles_per_ray > 1) {
                int part = tile % a.tiles_per_ray;
                first = part == 0; last = part == a.tiles_per_ray - 1;
                if (!first) {
#pragma unroll
                    for (int i = 0; i < 5; ++i) t[i] += carry[1 + i];
                }
                float prod = scr[0] * scr[1] * scr[2] * scr[3];
                carry[0] = first ? prod : carry[0] * prod;
#pragma unroll
                for (int i = 0; i < 5; ++i) carry[1 + i] = t[i];
            }
            if (last) {
                a.rgb_map[3 * (size_t)ri.ray + 0] = t[0]; a.rgb_map[3 * (size_t)ri.ray + 1] = t[1];
                a.rgb_map[3 * (size_t)ri.ray + 2] = t[2];
                a.depth[ri.ray] = t[3];
                if (a.acc) a.acc[ri.ray] = t[4];
            }
        }
        named_bar_sync(1, 128);      // scratch reused by the next tile
    }
}

// ------------------------------------------------------------------------------------------
// epilogue of one N=128 accumulator half for one warp (32 rows x 64 columns): tcgen05.ld,
// + bias, ReLU, bf16, and tcgen05.st of the 32 packed columns back over the first half of the
// columns this warp has just read = one K-block of the next layer's A operand (K-major in TMEM).
// kSigma (layer 7 only) also accumulates the density-head dot product from the fp32 values.
__device__ __forceinline__ void bias_relu_pack(const uint32_t (&x)[32], uint32_t *pk, uint32_t bias_addr)
{
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 b = ld_shared_f4(bias_addr + 16 * i);
        float x0 = __uint_as_float(x[4 * i + 0]), x1 = __uint_as_float(x[4 * i + 1]);
        float x2 = __uint_as_float(x[4 * i + 2]), x3 = __uint_as_float(x[4 * i + 3]);
        add2(x0, x1, b.x, b.y);
        add2(x2, x3, b.z, b.w);
        pk[2 * i + 0] = relu_pack_bf16(x0, x1);
        pk[2 * i + 1] = relu_pack_bf16(x2, x3);
    }
}
// split precision: hi = bf16(relu(x + b)), lo = bf16(relu(x + b) - hi), 32 columns -> 16 + 16 packed columns
__device__ __forceinline__ void bias_relu_pack_split(const uint32_t (&x)[32], uint32_t *hi, uint32_t *lo, uint32_t bias_addr)
{
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 b = ld_shared_f4(bias_addr + 16 * i);
        float v[4] = {__uint_as_float(x[4 * i + 0]), __uint_as_float(x[4 * i + 1]), __uint_as_float(x[4 * i + 2]), __uint_as_float(x[4 * i + 3])};
        add2(v[0], v[1], b.x, b.y);
        add2(v[2], v[3], b.z, b.w);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float a0 = fmaxf(v[2 * j], 0.f), a1 = fmaxf(v[2 * j + 1], 0.f);
            const uint32_t h = pack_bf16(a0, a1);
            hi[2 * i + j] = h;
            lo[2 * i + j] = pack_bf16(a0 - __uint_as_float(h << 16), a1 - __uint_as_float(h & 0xffff0000u));
        }
    }
}

// ws_out (TRAIN): bf16 operand blocks of the workspace (or nullptr); ws_row = first feature of this warp's 64;
// col_r / stage: see store_block_rows_staged / store_block_rows_bulk (train_layout.h); mask_out = this thread's mask word or nullptr
// col0 >= 0: the warp's rows are the consecutive samples col0 .. col0 + 31 -> TMA store (store_block_rows_bulk)
template <bool SPLIT>
__device__ __forceinline__ void epilogue_half(uint32_t t_cols, uint32_t bias_addr, uint32_t bar_ready, int lane,
                                              unsigned short *ws_out = nullptr, int ws_row = 0, const int *col_r = nullptr,
                                              uint32_t stage = 0, unsigned long long *mask_out = nullptr, int col0 = -1)
{
    uint32_t xa[32], xb[32];
    tmem_ld32(t_cols, xa);
    tmem_ld32(t_cols + 32, xb);
    tmem_ld_wait();
    if (!SPLIT) {
        uint32_t pk[32];
        bias_relu_pack(xa, pk, bias_addr);
        bias_relu_pack(xb, pk + 16, bias_addr + 128);
        tmem_st32(t_cols, pk);                   // K-block: 64 bf16 in columns [0, 32) of this warp's range
        if (ws_out) {                            // TRAIN: hand the operand to the next layer first, then store
            tmem_st_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ready);
            // exactly what the next layer multiplies: the bf16-rounded values (128 contiguous bytes per sample)
            if (col0 >= 0) {
                store_block_rows_bulk<0>(ws_out, ws_row, col0, pk, stage, lane);
            } else {
                const int cr[4] = {col_r[0], col_r[1], col_r[2], col_r[3]};
                if (lane == 0) bulk_store_reads_done();        // an earlier tile's copy may still read the buffer
                __syncwarp();
                store_block_rows_staged(ws_out, ws_row, cr, pk, stage, lane);
            }
            // + the ReLU mask of these 64 activations for the dgrad chain
            if (mask_out) *mask_out = (unsigned long long)relu_mask_word(pk) | ((unsigned long long)relu_mask_word(pk + 16) << 32);
            return;
        }
    } else {
        uint32_t hi[16], lo[16];                 // hi halves in columns [0, 32), lo halves in [32, 64)
        bias_relu_pack_split(xa, hi, lo, bias_addr);
        tmem_st16(t_cols, hi);
        tmem_st16(t_cols + 32, lo);
        bias_relu_pack_split(xb, hi, lo, bias_addr + 128);
        tmem_st16(t_cols + 16, hi);
        tmem_st16(t_cols + 48, lo);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_ready);
}

// 32 columns of colour layer 0: relu(acc + per-ray bias) . W_c1 -> 3 partial sums
__device__ __forceinline__ void color_dot(const uint32_t (&x)[32], uint32_t rayb_addr, uint32_t wc1_addr,
                                          float &r0, float &r1, float &r2)
{
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 b = ld_shared_f4(rayb_addr + 16 * i);
        const float4 w0 = ld_shared_f4(wc1_addr + 16 * i);
        const float4 w1 = ld_shared_f4(wc1_addr + 512 + 16 * i);
        const float4 w2 = ld_shared_f4(wc1_addr + 1024 + 16 * i);
        float x0 = __uint_as_float(x[4 * i + 0]), x1 = __uint_as_float(x[4 * i + 1]);
        float x2 = __uint_as_float(x[4 * i + 2]), x3 = __uint_as_float(x[4 * i + 3]);
        add2(x0, x1, b.x, b.y);
        add2(x2, x3, b.z, b.w);
        x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
        r0 = fmaf(x0, w0.x, r0); r0 = fmaf(x1, w0.y, r0); r0 = fmaf(x2, w0.z, r0); r0 = fmaf(x3, w0.w, r0);
        r1 = fmaf(x0, w1.x, r1); r1 = fmaf(x1, w1.y, r1); r1 = fmaf(x2, w1.z, r1); r1 = fmaf(x3, w1.w, r1);
        r2 = fmaf(x0, w2.x, r2); r2 = fmaf(x1, w2.y, r2); r2 = fmaf(x2, w2.z, r2); r2 = fmaf(x3, w2.w, r2);
    }
}
// back warps: this thread's row of colour layer 0's accumulator (128 columns) -> 3 colour pre-activations
__device__ __forceinline__ void color_row(uint32_t t_row, uint32_t rayb_addr, uint32_t wc1_addr, uint32_t bar_c0free,
                                          int lane, float &r0, float &r1, float &r2, float &sig_pre)
{
    uint32_t xa[32], xb[32];
    r0 = r1 = r2 = 0.f;
    const uint32_t sg = tmem_ld1(t_row + 128);              // density head: accumulator column 128
    tmem_ld32(t_row, xa);
    tmem_ld32(t_row + 32, xb);
    tmem_ld_wait();
    color_dot(xa, rayb_addr, wc1_addr, r0, r1, r2);
    tmem_ld32(t_row + 64, xa);
    color_dot(xb, rayb_addr + 128, wc1_addr + 128, r0, r1, r2);
    tmem_ld32(t_row + 96, xb);
    tmem_ld_wait();
    // the accumulator is in registers: the next tile's layer 1 may overwrite it
    tc_fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_c0free);
    sig_pre = __uint_as_float(sg);
    color_dot(xa, rayb_addr + 256, wc1_addr + 256, r0, r1, r2);
    color_dot(xb, rayb_addr + 384, wc1_addr + 384, r0, r1, r2);
}

// SRC_POINTS back warps (query_network): this row's own view direction -> encoded direction (fp32, 27 registers) ->
// colour layer 0's bias for each column on the fly (W_dir [27][128] and b_c0 sit in shared memory, read as
// broadcasts), relu(acc + bias) . W_c1 -> sigmoid; density = relu(acc col 128 + b).  NeRFModel.forward's outputs.
__device__ __forceinline__ void query_row(const Args &a, int idx, uint32_t t_row, uint32_t wdir_addr, uint32_t wc1_addr,
                                          uint32_t bar_c0free, int lane, const float *__restrict__ wf, uint32_t warp_bias_addr)
{
    float de[kDirFeat];
    const bool on = idx < a.n_points;
    {
        float d[3] = {0.f, 0.f, 0.f};
        if (on)
#pragma unroll
            for (int c = 0; c < 3; ++c) d[c] = __ldg(a.dirs + 3 * (size_t)idx + c);
        // The usual caller passes a ray's direction on every one of its samples (base_renderer.py:165-188 after
        // sample_points_on_rays): when the warp's 32 rows carry ONE direction, the warp builds that direction's bias row
        // together -- lane j the j-th encoded feature, lane l columns 4l .. 4l+3, the same fmaf chain per column as the
        // per-row path below and as the fused render's per-ray bias (so the same bits) -- and the rows take the fused
        // render's colour head: ~150 instructions per row instead of 1728 packed FMAs + 864 shared-memory reads.
        const float d0x = __shfl_sync(0xffffffffu, d[0], 0), d0y = __shfl_sync(0xffffffffu, d[1], 0), d0z = __shfl_sync(0xffffffffu, d[2], 0);
        const bool same = !on || (__float_as_uint(d[0]) == __float_as_uint(d0x) && __float_as_uint(d[1]) == __float_as_uint(d0y) &&
                                  __float_as_uint(d[2]) == __float_as_uint(d0z));
        if (__all_sync(0xffffffffu, same)) {
            float feat = 0.f;
            if (lane < kDirFeat) {
                const int w = lane < 3 ? lane : (lane - 3) % 6, c = w % 3;
                const float dc = c == 0 ? d0x : c == 1 ? d0y : d0z;
                if (lane < 3) feat = dc;
                else {
                    const float arg = __fmul_rn(kPiF * (float)(1 << ((lane - 3) / 6)), dc);
                    feat = w < 3 ? sinf(arg) : cosf(arg);
                }
            }
            float4 b = ld_shared_f4(wdir_addr + (kDirFeat * 128 + 4 * lane) * 4);    // b_c0
#pragma unroll
            for (int j = 0; j < kDirFeat; ++j) {
                const float dj = __shfl_sync(0xffffffffu, feat, j);
                const float4 w = ld_shared_f4(wdir_addr + (j * 128 + 4 * lane) * 4);
                fma2_bcast(b.x, b.y, w.x, w.y, dj, b.x, b.y);
                fma2_bcast(b.z, b.w, w.z, w.w, dj, b.z, b.w);
            }
            st_shared_v4(warp_bias_addr + 16 * lane, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
            __syncwarp();
            float r0, r1, r2, sig_pre;
            color_row(t_row, warp_bias_addr, wc1_addr, bar_c0free, lane, r0, r1, r2, sig_pre);
            if (on) {
                a.sigma_out[idx] = fmaxf(sig_pre + __ldg(wf + F_BSIG), 0.f);
                a.rgb_out[3 * (size_t)idx + 0] = 1.0f / (1.0f + expf(-(r0 + __ldg(wf + F_BC1 + 0))));
                a.rgb_out[3 * (size_t)idx + 1] = 1.0f / (1.0f + expf(-(r1 + __ldg(wf + F_BC1 + 1))));
                a.rgb_out[3 * (size_t)idx + 2] = 1.0f / (1.0f + expf(-(r2 + __ldg(wf + F_BC1 + 2))));
            }
            __syncwarp();                                   // the bias row is rewritten by this warp's next tile
            return;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) de[c] = d[c];
#pragma unroll
        for (int k = 0; k < kDirFreq; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float arg = __fmul_rn(kPiF * (float)(1 << k), d[c]);
                de[3 + 6 * k + c] = sinf(arg);
                de[3 + 6 * k + 3 + c] = cosf(arg);
            }
    }
    float r[3] = {0.f, 0.f, 0.f};
    const uint32_t sg = tmem_ld1(t_row + 128);
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        uint32_t x[32];
        tmem_ld32(t_row + 32 * g, x);
        tmem_ld_wait();
        if (g == 3) {                                       // the accumulator is in registers
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_c0free);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c0 = 32 * g + 4 * i;
            float4 b = ld_shared_f4(wdir_addr + (kDirFeat * 128 + c0) * 4);          // b_c0
#pragma unroll
            for (int j = 0; j < kDirFeat; ++j) {
                const float4 w = ld_shared_f4(wdir_addr + (j * 128 + c0) * 4);
                fma2_bcast(b.x, b.y, w.x, w.y, de[j], b.x, b.y);     // packed FMAs: the same bits as four fmaf, half the instructions
                fma2_bcast(b.z, b.w, w.z, w.w, de[j], b.z, b.w);
            }
            const float4 w0 = ld_shared_f4(wc1_addr + c0 * 4);
            const float4 w1 = ld_shared_f4(wc1_addr + 512 + c0 * 4);
            const float4 w2 = ld_shared_f4(wc1_addr + 1024 + c0 * 4);
            const float v[4] = {fmaxf(__uint_as_float(x[4 * i + 0]) + b.x, 0.f), fmaxf(__uint_as_float(x[4 * i + 1]) + b.y, 0.f),
                                fmaxf(__uint_as_float(x[4 * i + 2]) + b.z, 0.f), fmaxf(__uint_as_float(x[4 * i + 3]) + b.w, 0.f)};
            r[0] = fmaf(v[0], w0.x, r[0]); r[0] = fmaf(v[1], w0.y, r[0]); r[0] = fmaf(v[2], w0.z, r[0]); r[0] = fmaf(v[3], w0.w, r[0]);
            r[1] = fmaf(v[0], w1.x, r[1]); r[1] = fmaf(v[1], w1.y, r[1]); r[1] = fmaf(v[2], w1.z, r[1]); r[1] = fmaf(v[3], w1.w, r[1]);
            r[2] = fmaf(v[0], w2.x, r[2]); r[2] = fmaf(v[1], w2.y, r[2]); r[2] = fmaf(v[2], w2.z, r[2]); r[2] = fmaf(v[3], w2.w, r[2]);
        }
    }
    if (on) {
        a.sigma_out[idx] = fmaxf(__uint_as_float(sg) + __ldg(wf + F_BSIG), 0.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) a.rgb_out[3 * (size_t)idx + c] = 1.0f / (1.0f + expf(-(r[c] + __ldg(wf + F_BC1 + c))));
    }
}

// TRAIN back warps: this row of colour layer 0's accumulator -> relu(acc + per-ray bias) stored as colour layer
// 0's activation, colour pre-activations -> sigmoid -> stored, density pre-activation stored.  No compositing:
// the training step's ray kernel does forward compositing, loss and backward in one pass.
__device__ __forceinline__ void train_heads_row(uint32_t t_row, uint32_t rayb_addr, uint32_t wc1_addr, uint32_t bar_c0free,
                                                int lane, const float *__restrict__ wf, float *ws, int ws_ch, int col)
{
    float r[3] = {0.f, 0.f, 0.f};
    const uint32_t sg = tmem_ld1(t_row + 128);
    unsigned long long *mask_c0 = reinterpret_cast<unsigned long long *>(ws + (size_t)R_MASKC0 * ws_ch);
    unsigned int mbits[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        uint32_t x[32], pk[16];
        tmem_ld32(t_row + 32 * g, x);
        tmem_ld_wait();
        if (g == 3) {                                       // the accumulator is in registers
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_c0free);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 b = ld_shared_f4(rayb_addr + (32 * g + 4 * i) * 4);
            const float4 w0 = ld_shared_f4(wc1_addr + (32 * g + 4 * i) * 4);
            const float4 w1 = ld_shared_f4(wc1_addr + 512 + (32 * g + 4 * i) * 4);
            const float4 w2 = ld_shared_f4(wc1_addr + 1024 + (32 * g + 4 * i) * 4);
            float v[4] = {fmaxf(__uint_as_float(x[4 * i + 0]) + b.x, 0.f), fmaxf(__uint_as_float(x[4 * i + 1]) + b.y, 0.f),
                          fmaxf(__uint_as_float(x[4 * i + 2]) + b.z, 0.f), fmaxf(__uint_as_float(x[4 * i + 3]) + b.w, 0.f)};
            r[0] = fmaf(v[0], w0.x, r[0]); r[0] = fmaf(v[1], w0.y, r[0]); r[0] = fmaf(v[2], w0.z, r[0]); r[0] = fmaf(v[3], w0.w, r[0]);
            r[1] = fmaf(v[0], w1.x, r[1]); r[1] = fmaf(v[1], w1.y, r[1]); r[1] = fmaf(v[2], w1.z, r[1]); r[1] = fmaf(v[3], w1.w, r[1]);
            r[2] = fmaf(v[0], w2.x, r[2]); r[2] = fmaf(v[1], w2.y, r[2]); r[2] = fmaf(v[2], w2.z, r[2]); r[2] = fmaf(v[3], w2.w, r[2]);
            pk[2 * i] = pack_bf16(v[0], v[1]);
            pk[2 * i + 1] = pack_bf16(v[2], v[3]);
        }
        mbits[g] = relu_mask_word(pk);
        if (col >= 0) {                                     // 32 features = four 16-byte units of this sample's row
            uint4 *rowp = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned short *>(ws) + big_row(G_C0H + 64 * (g >> 1), col));
#pragma unroll
            for (int u = 0; u < 4; ++u)
                rowp[((g & 1) * 4 + u) ^ (col & 7)] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
    }
    if (col >= 0) {
        mask_c0[col] = (unsigned long long)mbits[0] | ((unsigned long long)mbits[1] << 32);
        mask_c0[(size_t)ws_ch + col] = (unsigned long long)mbits[2] | ((unsigned long long)mbits[3] << 32);
        ws[(size_t)R_SIGPRE * ws_ch + col] = __uint_as_float(sg) + __ldg(wf + F_BSIG);
#pragma unroll
        for (int c = 0; c < 3; ++c) ws[(size_t)(R_RGB + c) * ws_ch + col] = 1.0f / (1.0f + expf(-(r[c] + __ldg(wf + F_BC1 + c))));
    }
}

// ------------------------------------------------------------------------------------------
// MMA issue: one tile's schedule as straight-line code.  Everything that depends on the chunk index
// is a compile-time constant; the per-tile variables live in IssueCtx.
struct IssueCtx {
    uint32_t bars;          // shared address of barrier 0
    uint32_t region[2];     // TMEM base of the accumulator region of even / odd layers of this tile
    uint64_t wdesc;         // smem descriptor of weight ring slot 0, chunk 0
    uint64_t pedesc;        // smem descriptor of this tile's encoded-position operand (hi plane)
    uint32_t pe_empty_bar;  // barrier released when layer 4 has consumed the encoded position
    int tile;               // CTA-local tile index
    uint16_t pair_mask;     // != 0: weight ring shared with the peer CTAs of the cluster: release slots in all of them
    unsigned int *dbg;
    long long *trace;       // this tile's trace rows or nullptr
};

template <int CI, bool SPLIT>
__device__ __forceinline__ void issue_chunk(const IssueCtx &x)
{
    using C = Cfg<SPLIT>;
    constexpr ChunkInfo c = kChunks.c[CI];
    constexpr int stage = CI / kStageChunks, slot = stage % C::kWSlots;
    constexpr uint32_t idesc = c.layer == 8 ? idesc_bf16(128, kC0Rows) : idesc_bf16(128, 128);
    constexpr uint32_t chunk_in_slot = (uint32_t)(chunk_offset(CI) - stage_offset(stage));
    constexpr bool last_of_layer = CI + 1 == kChunksPerTile || kChunks.c[CI + 1 < kChunksPerTile ? CI + 1 : CI].layer != c.layer;
    constexpr bool first_of_layer = CI == 0 || kChunks.c[CI > 0 ? CI - 1 : 0].layer != c.layer;
    if constexpr (CI % kStageChunks == 0)
        wait_bar(x.bars + 8u * (B_WFULL + slot), (stage / C::kWSlots) & 1, x.dbg, 4);
    if constexpr ((c.flags & 4) != 0)      // a_ready[kb]: one phase per producing layer 0..7 (8 per tile: parity restarts)
        wait_bar(x.bars + 8u * (B_AREADY + c.asrc), (c.layer - 1) & 1, x.dbg, 3);
    if constexpr (c.layer == 1 && first_of_layer) {
        // layer 1 overwrites the region colour layer 0 of the PREVIOUS tile accumulated into: wait until the
        // back warps have pulled it into registers
        if (x.tile > 0) wait_bar(x.bars + 8u * B_C0FREE, (x.tile - 1) & 1, x.dbg, 10);
    }
    tc_fence_after_sync();
    if (elect_one()) {
        const uint32_t d_tmem = x.region[c.layer & 1] + c.half * 128;
        const uint64_t bdesc = x.wdesc + (uint64_t)((slot * C::kSlotBytes + chunk_in_slot) >> 4);
        constexpr uint64_t kLoW = kStageSlotBytes >> 4;      // lo weight stage sits behind the hi stage in the slot
        if constexpr (c.asrc == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {        // 4 x (K = 16): +32 B inside the 128 B swizzle span
                mma_bf16_ss(d_tmem, x.pedesc + 2 * k, bdesc + 2 * k, idesc, !((c.flags & 1) && k == 0));
                if constexpr (SPLIT) {
                    mma_bf16_ss(d_tmem, x.pedesc + 2 * k, bdesc + kLoW + 2 * k, idesc, true);             // hi * lo
                    mma_bf16_ss(d_tmem, x.pedesc + (16384 >> 4) + 2 * k, bdesc + 2 * k, idesc, true);     // lo * hi
                }
            }
        } else {
            // A = K-block `asrc` of the previous layer: bf16 pairs written in place over the other
            // region's accumulator columns [64 asrc, 64 asrc + 32) (SPLIT: lo halves in [+32, +64))
            const uint32_t a_tmem = x.region[(c.layer & 1) ^ 1] + c.asrc * 64;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                mma_bf16_ts(d_tmem, a_tmem + 8 * k, bdesc + 2 * k, idesc, !((c.flags & 1) && k == 0));
                if constexpr (SPLIT) {
                    mma_bf16_ts(d_tmem, a_tmem + 8 * k, bdesc + kLoW + 2 * k, idesc, true);
                    mma_bf16_ts(d_tmem, a_tmem + 32 + 8 * k, bdesc + 2 * k, idesc, true);
                }
            }
        }
        if constexpr (first_of_layer) { if (x.trace) x.trace[c.layer * 8 + 0] = clock64(); }
        if constexpr ((c.flags & 2) != 0) {
            mma_commit(x.bars + 8u * (c.layer == 8 ? B_ACCC0 : B_ACCFULL + c.half));
            if (x.trace) x.trace[c.layer * 8 + (last_of_layer ? 2 : 1)] = clock64();
        }
        if constexpr (CI % kStageChunks == kStageChunks - 1) {
            if (x.pair_mask) mma_commit_mcast(x.bars + 8u * (B_WEMPTY + slot), x.pair_mask);
            else mma_commit(x.bars + 8u * (B_WEMPTY + slot));
        }
        if constexpr (c.layer == 4 && last_of_layer) mma_commit(x.pe_empty_bar);
    }
    __syncwarp();
}
template <bool SPLIT, int... CI>
__device__ __forceinline__ void issue_tile(const IssueCtx &x, std::integer_sequence<int, CI...>)
{
    (issue_chunk<CI, SPLIT>(x), ...);
}

// ------------------------------------------------------------------------------------------
template <int SRC, bool SPLIT, bool TRAIN>
__global__ void __launch_bounds__(kThreads, 1) fused_render_kernel(const Args a)
{
    using C = Cfg<SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array, so the compiler keeps the
    // shared address space (LDS/STS instead of generic LD/ST)
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_base = smem_u32(sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bars = sm_base + SM_BAR;
    auto bar = [&](int i) { return bars + 8u * i; };
    const float *wf = reinterpret_cast<const float *>(a.packed);
    const unsigned char *wb = a.packed + B_OFFSET;
    const unsigned char *wb_lo = a.packed + B_LO_OFFSET;

    const int tile_begin = blockIdx.x * a.tiles_per_cta;
    const int tile_end = min(a.n_tiles, tile_begin + a.tiles_per_cta);
    // paired CTAs consume the shared weight stream in lockstep: both run tiles_per_cta tiles, the ones past the end
    // of the work are ghosts (every row invalid, nothing written)
    const bool pair = a.pair > 1;
    const int my_tiles = pair ? a.tiles_per_cta : max(0, tile_end - tile_begin);
    const uint32_t cta_rank = pair ? cluster_ctarank() : 0u;           // 0 .. a.pair - 1

    // ---- one-time setup -------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int i = 0; i < C::kWSlots; ++i) { mbar_init(bar(B_WFULL + i), 1); mbar_init(bar(B_WEMPTY + i), pair ? a.pair : 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_PEFULL + i), 4); mbar_init(bar(B_PEEMPTY + i), 1);
        }
        for (int i = 0; i < 2; ++i) mbar_init(bar(B_ACCFULL + i), 1);
        mbar_init(bar(B_ACCC0), 1);
        mbar_init(bar(B_C0FREE), 4);
        for (int i = 0; i < 4; ++i) mbar_init(bar(B_AREADY + i), 4);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(sm_base + SM_TMEM);
    {   // small fp32 tables the epilogue reads as shared-memory broadcasts
        float *bias = reinterpret_cast<float *>(sm + SM_BIAS);
        for (int i = threadIdx.x; i < 8 * 256; i += kThreads) bias[i] = __ldg(wf + F_BIAS + i);
        float *wc1 = reinterpret_cast<float *>(sm + SM_WC1);
        for (int i = threadIdx.x; i < 384; i += kThreads) wc1[i] = __ldg(wf + F_WC1 + i);
        if (SRC == SRC_POINTS) {                           // W_dir [27][128] then b_c0 [128]: 14 KB of the staging region
            float *wd = reinterpret_cast<float *>(sm + SM_STAGE);
            for (int i = threadIdx.x; i < kDirFeat * 128; i += kThreads) wd[i] = __ldg(wf + F_WC0D + i);
            for (int i = threadIdx.x; i < 128; i += kThreads) wd[kDirFeat * 128 + i] = __ldg(wf + F_BC0 + i);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (pair) cluster_sync_all();                          // the peer's barriers exist before anything is multicast at them
    tc_fence_after_sync();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(sm + SM_TMEM);

    if (warp == 0) {
        // ================================ weight producer ===================================
        // the packed stream is already in consumption order: 32 consecutive 32 KB stages per tile
        if (lane == 0) {
            uint32_t sg = 0;
            for (int t = 0; t < my_tiles; ++t) {
                for (int st = 0; st < kStagesPerTile; ++st, ++sg) {
                    const uint32_t slot = sg % C::kWSlots, round = sg / C::kWSlots;
                    if (round > 0) wait_bar(bar(B_WEMPTY + slot), (round - 1) & 1, a.dbg, 1);
                    const uint32_t bytes = stage_bytes(st), dst = sm_base + SM_W + slot * C::kSlotBytes;
                    mbar_arrive_expect_tx(bar(B_WFULL + slot), SPLIT ? 2 * bytes : bytes);
                    if (pair) {
                        // this CTA fetches its 1/size piece of the stage for every CTA of the cluster; the other pieces
                        // arrive from the peers.  A slot is refilled only after ALL MMA issuers released it (B_WEMPTY
                        // counts one arrival per CTA), so no copy can overtake a reader.
                        const uint32_t piece = bytes / (uint32_t)a.pair, off = cta_rank * piece;
                        const uint16_t mask = (uint16_t)((1u << a.pair) - 1u);
                        bulk_g2s_mcast(dst + off, wb + stage_offset(st) + off, piece, bar(B_WFULL + slot), mask);
                        if (SPLIT) bulk_g2s_mcast(dst + kStageSlotBytes + off, wb_lo + stage_offset(st) + off, piece, bar(B_WFULL + slot), mask);
                        continue;
                    }
                    bulk_g2s(dst, wb + stage_offset(st), bytes, bar(B_WFULL + slot));
                    if (SPLIT) bulk_g2s(dst + kStageSlotBytes, wb_lo + stage_offset(st), bytes, bar(B_WFULL + slot));
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer =========================================
        // warp-uniform control flow; one elected lane issues the tcgen05 instructions
        IssueCtx x;
        x.bars = bars;
        x.pair_mask = pair ? (uint16_t)((1u << a.pair) - 1u) : (uint16_t)0;
        x.wdesc = smem_desc_sw128(sm_base + SM_W);
        x.dbg = a.dbg;
        for (int t = 0; t < my_tiles; ++t) {
            // encoded-position buffers: two alternate (bf16) or one is reused every tile (split)
            const int pb = SPLIT ? 0 : (t & 1), pe_use = SPLIT ? t : (t >> 1);
            wait_bar(bar(B_PEFULL + pb), pe_use & 1, a.dbg, 2);
            // nine layers per tile: the region parity of layer l of tile t is (t + l) & 1
            x.region[0] = tmem_base + (uint32_t)(t & 1) * 256;
            x.region[1] = tmem_base + (uint32_t)((t & 1) ^ 1) * 256;
            x.pedesc = smem_desc_sw128(sm_base + SM_PE + (SPLIT ? 0 : pb * 16384));
            x.pe_empty_bar = bar(B_PEEMPTY + pb);
            x.tile = t;
            x.trace = (a.trace && blockIdx.x == 0 && t < kTraceTiles) ? a.trace + t * 72 : nullptr;
            issue_tile<SPLIT>(x, std::make_integer_sequence<int, kChunksPerTile>{});
        }
    } else if (warp >= 4 && warp < 12) {
        // ================================ epilogue ===========================================
        const int ew = warp - 4, q = ew & 3, w2 = ew >> 2;
        const int row = q * 32 + lane;
        uint32_t g = 0;
        for (int t = 0; t < my_tiles; ++t) {
            long long *tr = (a.trace && blockIdx.x == 0 && t < kTraceTiles && ew == 0 && lane == 0) ? a.trace + t * 72 : nullptr;
            int col = -1, col_r[4] = {-1, -1, -1, -1}, col0 = -1;
            if (TRAIN) {
                col = ws_col(a, row_info(a, tile_begin + t, row));
#pragma unroll
                for (int j = 0; j < 4; ++j) col_r[j] = __shfl_sync(0xffffffffu, col, (lane >> 2) + 8 * j);
                // all 32 rows of the warp valid and consecutive from a multiple of 32 (every full tile when the sample count is
                // 16, 32, 64 or 128): store through the TMA engine
                const int c0 = __shfl_sync(0xffffffffu, col, 0);
                if (c0 >= 0 && (c0 & 31) == 0 && __all_sync(0xffffffffu, col == c0 + lane)) col0 = c0;
            }
            for (int layer = 0; layer < 8; ++layer, ++g) {
                const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (g & 1) * 256 + 64 * w2;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    wait_bar(bar(B_ACCFULL + hh), layer & 1, a.dbg, 5);      // 8 phases per tile: parity = layer
                    tc_fence_after_sync();
                    if (tr) tr[layer * 8 + (hh == 0 ? 3 : 6)] = clock64();
                    unsigned short *ws_out = nullptr;
                    unsigned long long *mask_out = nullptr;
                    if (TRAIN) {
                        ws_out = reinterpret_cast<unsigned short *>(a.ws);
                        if (col >= 0)
                            mask_out = reinterpret_cast<unsigned long long *>(a.ws + (size_t)R_MASK * a.ws_ch) +
                                       (size_t)(layer * 4 + hh * 2 + w2) * a.ws_ch + col;
                    }
                    epilogue_half<SPLIT>(t_lane + hh * 128, sm_base + SM_BIAS + (layer * 256 + hh * 128 + 64 * w2) * 4,
                                         bar(B_AREADY + 2 * hh + w2), lane, ws_out, G_H + layer * 256 + hh * 128 + 64 * w2, col_r,
                                         ew < 6 ? sm_base + SM_STAGE + ew * 4096 : sm_base + SM_W + (ew - 6) * kStageSlotBytes + kStageBytes,
                                         mask_out, col0);
                    if (tr && hh == 0) tr[layer * 8 + 4] = clock64();
                }
                if (tr) tr[layer * 8 + 5] = clock64();
            }
            ++g;                                    // colour layer 0 (the back warps' epilogue) takes a region turn too
        }
        if (TRAIN && lane == 0) bulk_store_drain();
    } else if (warp >= 12) {
        // ================================ front / back =======================================
        const int row = (warp - 12) * 32 + lane;
        const float step = linspace_step(a.n_samples);
        auto produce = [&](int t) {
            const int pb = SPLIT ? 0 : (t & 1), pe_use = SPLIT ? t : (t >> 1);
            if (pe_use >= 1) wait_bar(bar(B_PEEMPTY + pb), (pe_use - 1) & 1, a.dbg, 8);
            int c0 = -1;
            if (TRAIN) {
                // whole warp of valid, consecutive samples from a multiple of 32: its 32 rows of the operand tile (4 KB, already
                // the workspace image: unit u of a row at u ^ (row & 7), and col = row mod 8) go out as ONE bulk copy
                const int col = ws_col(a, row_info(a, tile_begin + t, row));
                c0 = __shfl_sync(0xffffffffu, col, 0);
                if (!(c0 >= 0 && (c0 & 31) == 0 && __all_sync(0xffffffffu, col == c0 + lane))) c0 = -1;
                if (lane == 0) bulk_store_reads_done();     // the copy issued from this buffer two tiles ago has read it
                __syncwarp();
            }
            produce_tile<SRC, SPLIT, TRAIN>(a, sm, tile_begin + t, pb, t & 1, row, step, wf, c0 >= 0);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar(B_PEFULL + pb));
                if (TRAIN && c0 >= 0) {
                    unsigned short *dst = reinterpret_cast<unsigned short *>(a.ws) + big_row(G_PE, c0);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;"
                                 ::"l"(dst), "r"(sm_base + SM_PE + pb * 16384 + (warp - 12) * 4096) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        };
        if (my_tiles > 0) produce(0);
        for (int t = 0; t < my_tiles; ++t) {
            if (t + 1 < my_tiles) produce(t + 1);
            const int pb = t & 1;                       // per-ray bias buffers always alternate
            // colour layer 0's epilogue: this thread's accumulator row straight from TMEM
            wait_bar(bar(B_ACCC0), t & 1, a.dbg, 11);
            tc_fence_after_sync();
            const int rpt_shift = a.tiles_per_ray == 1 ? a.s_pad_log2 : 7;
            const uint32_t t_row = tmem_base + ((uint32_t)((warp - 12) * 32) << 16) + (uint32_t)(t & 1) * 256;
            if (SRC == SRC_POINTS) {
                query_row(a, (tile_begin + t) * kTileM + row, t_row, sm_base + SM_STAGE, sm_base + SM_WC1, bar(B_C0FREE), lane, wf,
                          sm_base + SM_RAYB + (warp - 12) * 512);
            } else if (TRAIN) {
                train_heads_row(t_row, sm_base + SM_RAYB + (pb * kMaxRaysPerTile + (row >> rpt_shift)) * 512, sm_base + SM_WC1,
                                bar(B_C0FREE), lane, wf, a.ws, a.ws_ch, ws_col(a, row_info(a, tile_begin + t, row)));
            } else {
                float ypre[3], sig_pre;
                color_row(t_row, sm_base + SM_RAYB + (pb * kMaxRaysPerTile + (row >> rpt_shift)) * 512, sm_base + SM_WC1,
                          bar(B_C0FREE), lane, ypre[0], ypre[1], ypre[2], sig_pre);
                composite_tile<SRC>(a, sm, tile_begin + t, row, step, wf, sig_pre, ypre);
            }
        }
        if (TRAIN && lane == 0) bulk_store_drain();
    }

    // ---- teardown ---------------------------------------------------------------------------
    tc_fence_before_sync();
    __syncthreads();
    if (pair) cluster_sync_all();                          // no CTA leaves while its peer may still signal into it
    if (warp == 2) {
        tc_fence_after_sync();
        tmem_dealloc<512>(tmem_base);
    }
}

static int plan(Args &a)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (a.sm_limit > 0 && a.sm_limit < sms) sms = a.sm_limit;
    if (a.n_points > 0) {                             // SRC_POINTS: a tile is 128 (point, direction) rows
        a.n_samples = kTileM;
        a.n_rays = (a.n_points + kTileM - 1) / kTileM;
    }
    const int S = a.n_samples;
    if (S < 1 || S > 32768) return NERF_B200_EUNSUPPORTED;
    if (S <= kTileM) {
        int lg = 4;                                   // S_pad >= 16 (at most 8 rays per tile)
        while ((1 << lg) < S) ++lg;
        a.s_pad_log2 = lg;
        a.tiles_per_ray = 1;
        int rpt = kTileM >> lg;
        a.n_tiles = (a.n_rays + rpt - 1) / rpt;
    } else {
        a.s_pad_log2 = 7;
        a.tiles_per_ray = (S + kTileM - 1) / kTileM;
        a.n_tiles = a.n_rays * a.tiles_per_ray;
    }
    int per = (a.n_tiles + sms - 1) / sms;
    per = ((per + a.tiles_per_ray - 1) / a.tiles_per_ray) * a.tiles_per_ray;   // a ray never straddles CTAs
    a.tiles_per_cta = per;
    int grid = (a.n_tiles + per - 1) / per;
    if (a.pair > 1) grid = (grid + a.pair - 1) / a.pair * a.pair;   // whole clusters (trailing CTAs may run ghost tiles only)
    return grid;
}

// Cluster size of the shared weight stream: 2 unless NERF_B200_CLUSTER=1|2|4 in the environment says otherwise (A/B
// measurements; 1 = every CTA streams all the weights itself).  Measured at 800x600x128 on B200: 2 = +1.6..2.1 %
// (SM clock 1522 -> 1552 MHz under the power cap); 4 = 0.62x, because 37 clusters of four one-CTA-per-SM blocks do
// not all fit the GPCs at once and the launch runs in two waves.
static int cluster_default()
{
    static const int v = [] {                         // read once per process (documented in include/nerf_b200.h)
        const char *e = getenv("NERF_B200_CLUSTER");
        const int x = e ? atoi(e) : 2;
        return (x == 1 || x == 2 || x == 4) ? x : 2;
    }();
    return v;
}

template <int SRC, bool SPLIT, bool TRAIN = false>
static int launch(Args &a, cudaStream_t stream)
{
    a.pair = TRAIN ? 1 : cluster_default();           // TRAIN: the paired weight stream measured 3-4 % SLOWER per training step (r2)
    int grid = plan(a);
    if (a.pair > 1 && a.n_tiles < 4 * grid) { a.pair = 1; grid = plan(a); }  // small launches: ghost tiles would dominate
    if (grid < 0) return grid;
    cudaError_t e = cudaFuncSetAttribute(fused_render_kernel<SRC, SPLIT, TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    if (a.pair > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = a.pair; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, fused_render_kernel<SRC, SPLIT, TRAIN>, a);
        if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
        return launch_status();
    }
    fused_render_kernel<SRC, SPLIT, TRAIN><<<grid, kThreads, kSmemBytes, stream>>>(a);
    return launch_status();
}

#include "mlp_tc_fp8.inl"

template <int SRC>
static int launch_fp8(Args &a, cudaStream_t stream)
{
    a.pair = cluster_default();
    int grid = plan(a);
    if (a.pair > 1 && a.n_tiles < 4 * grid) { a.pair = 1; grid = plan(a); }
    if (grid < 0) return grid;
    cudaError_t e = cudaFuncSetAttribute(fused_render_fp8_kernel<SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    if (a.pair > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = a.pair; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, fused_render_fp8_kernel<SRC>, a);
        if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
        return launch_status();
    }
    fused_render_fp8_kernel<SRC><<<grid, kThreads, kSmemBytes, stream>>>(a);
    return launch_status();
}

}  // namespace tc


int tc_render_pose(const void *packed, const float *c2w, int width, int height, float focal, float near,
                   float far, int n_samples, int row0, int n_rows, bool split, float *rgb_out, float *depth_out,
                   unsigned int *dbg, cudaStream_t stream)
{
    tc::Args a = {};
    a.trace = trace_buffer();
    a.packed = reinterpret_cast<const unsigned char *>(packed);
    a.pose = pose_from_c2w(c2w);
    a.width = width; a.row0 = row0;
    a.half_w = (float)((double)width * 0.5); a.half_h = (float)((double)height * 0.5); a.focal = focal;
    a.n_rays = n_rows * width; a.n_samples = n_samples;
    a.near = near; a.far = far;
    a.rgb_map = rgb_out; a.depth = depth_out; a.acc = nullptr; a.dbg = dbg;
    return split ? tc::launch<tc::SRC_POSE, true>(a, stream) : tc::launch<tc::SRC_POSE, false>(a, stream);
}

int tc_render_rays(const void *packed, const float *rays_o, const float *rays_d, int n_rays, int n_samples,
                   float near, float far, const float *t_rand, const float *z_vals, bool split, float *rgb_out,
                   float *depth_out, float *acc_out, float *weights_out, unsigned int *dbg, cudaStream_t stream)
{
    tc::Args a = {};
    a.z_vals = z_vals; a.weights = weights_out;
    a.packed = reinterpret_cast<const unsigned char *>(packed);
    a.rays_o = rays_o; a.rays_d = rays_d; a.t_rand = t_rand;
    a.n_rays = n_rays; a.n_samples = n_samples;
    a.near = near; a.far = far;
    a.rgb_map = rgb_out; a.depth = depth_out; a.acc = acc_out; a.dbg = dbg;
    return split ? tc::launch<tc::SRC_RAYS, true>(a, stream) : tc::launch<tc::SRC_RAYS, false>(a, stream);
}

// FP8 mode (fp8_layout.h): `packed_fp8` = the buffer nerf_b200_pack_weights_fp8 filled
int tc_render_pose_fp8(const void *packed_fp8, const float *c2w, int width, int height, float focal, float near, float far,
                       int n_samples, int row0, int n_rows, float *rgb_out, float *depth_out, unsigned int *dbg, cudaStream_t stream)
{
    tc::Args a = {};
    a.packed = reinterpret_cast<const unsigned char *>(packed_fp8);
    a.pose = pose_from_c2w(c2w);
    a.width = width; a.row0 = row0;
    a.half_w = (float)((double)width * 0.5); a.half_h = (float)((double)height * 0.5); a.focal = focal;
    a.n_rays = n_rows * width; a.n_samples = n_samples;
    a.near = near; a.far = far;
    a.rgb_map = rgb_out; a.depth = depth_out; a.acc = nullptr; a.dbg = dbg;
    a.trace = trace_buffer();
    return tc::launch_fp8<tc::SRC_POSE>(a, stream);
}

int tc_render_rays_fp8(const void *packed_fp8, const float *rays_o, const float *rays_d, int n_rays, int n_samples, float near,
                       float far, const float *t_rand, const float *z_vals, float *rgb_out, float *depth_out, float *acc_out,
                       float *weights_out, unsigned int *dbg, cudaStream_t stream)
{
    tc::Args a = {};
    a.z_vals = z_vals; a.weights = weights_out;
    a.packed = reinterpret_cast<const unsigned char *>(packed_fp8);
    a.rays_o = rays_o; a.rays_d = rays_d; a.t_rand = t_rand;
    a.n_rays = n_rays; a.n_samples = n_samples;
    a.near = near; a.far = far;
    a.rgb_map = rgb_out; a.depth = depth_out; a.acc = acc_out; a.dbg = dbg;
    return tc::launch_fp8<tc::SRC_RAYS>(a, stream);
}

// NeRFModel.forward on (point, direction) rows: query_nerf_networks in BF16 mode
int tc_query_points(const void *packed, const float *points, const float *dirs, long long n, float *sigma, float *rgb,
                    unsigned int *dbg, cudaStream_t stream)
{
    if (n > (1ll << 30)) return NERF_B200_EUNSUPPORTED;
    tc::Args a = {};
    a.packed = reinterpret_cast<const unsigned char *>(packed);
    a.points = points; a.dirs = dirs; a.n_points = (int)n;
    a.sigma_out = sigma; a.rgb_out = rgb; a.dbg = dbg;
    a.near = 0.f; a.far = 1.f;
    return tc::launch<tc::SRC_POINTS, false>(a, stream);
}

// Training forward on the tensor cores: rays [0, n_rays) of the given (chunk-local) arrays; activations, head
// outputs and encodings go to the workspace (train_layout.h) for the ray kernel, the dgrad chain and wgrad.
int tc_train_forward(const void *packed, const float *rays_o, const float *rays_d, int n_rays, int n_samples, float near,
                     float far, const float *t_rand, float *ws, int ws_ch, unsigned int *dbg, int sm_limit, cudaStream_t stream)
{
    tc::Args a = {};
    a.sm_limit = sm_limit;
    a.packed = reinterpret_cast<const unsigned char *>(packed);
    a.rays_o = rays_o; a.rays_d = rays_d; a.t_rand = t_rand;
    a.n_rays = n_rays; a.n_samples = n_samples;
    a.near = near; a.far = far;
    a.ws = ws; a.ws_ch = ws_ch; a.dbg = dbg;
    a.trace = trace_buffer();
    return tc::launch<tc::SRC_RAYS, false, true>(a, stream);
}

}  // namespace nerfb200
