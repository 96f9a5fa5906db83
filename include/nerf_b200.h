/*
 * nerf_b200.h -- C ABI of the B200-native NeRF render/train hot path.
 *
 * This is the drop-in boundary for nerf-dbr's renderer plug-in interface
 * (reference: src/benchmark/base_renderer.py:90-281).  The reference has no FFI
 * (it is pure Python), so each entry point below names the reference *method* it
 * replaces; nerf_dbr_b200/host/ binds them with ctypes (see INTEGRATION.md for the
 * stub a nerf-dbr maintainer adds).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     parameter name ends in _host;  all tensors are dense row-major fp32 unless stated.
 *   - the caller owns every buffer (inputs, outputs, packed weights, workspaces); nothing here
 *     calls cudaMalloc/cudaFree and no data survives a call -> re-entrant across streams and
 *     devices (the launch counter is atomic, the two debug hooks at the end of this header are per
 *     device).  One exception: train_fwd_bwd in BF16 mode creates two auxiliary streams and three
 *     events per (device, caller stream) on first use and reuses them (its weight-gradient launches fork
 *     onto them and join back into the caller's stream before the call returns; at most 8 caller
 *     streams per device), so training calls for ONE device must come from one host thread at a time.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     all work is enqueued asynchronously on it.
 *   - return value: 0 = ok, negative = NERF_B200_E* argument error (nothing was
 *     enqueued), positive = cudaError_t from the launch.  Never throws.
 *   - there is no CPU fallback: on a machine without an sm_100 device the launch
 *     fails and the error code is returned.
 */
#ifndef NERF_B200_H
#define NERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NERF_B200_API __attribute__((visibility("default")))
#else
#define NERF_B200_API
#endif

#define NERF_B200_ABI_VERSION 1

enum {
    NERF_B200_OK = 0,
    NERF_B200_EINVAL = -1,      /* null pointer / non-positive size */
    NERF_B200_EUNSUPPORTED = -2,/* shape the selected precision mode cannot run */
    NERF_B200_EALIGN = -3       /* pointer not 16-byte aligned where required */
};

/* Arithmetic of the MLP contraction. */
enum {
    NERF_B200_FP32 = 0,   /* fp32 FFMA on CUDA cores: the <=1e-4 max-abs parity mode */
    NERF_B200_BF16 = 1,   /* bf16 operands, fp32 accumulate in TMEM (tcgen05): the throughput mode */
    NERF_B200_BF16X3 = 2  /* tensor cores with every operand split into bf16 (hi, lo) and three MMAs per product
                             (hi*hi + hi*lo + lo*hi): fp32-class accuracy (max-abs ~1e-5), ~1/3 of the BF16 rate.
                             render_image / render_rays only */
};

/* The 22 parameter tensors of one NeRFModel (reference src/models/nerf.py:72-90), as
 * they sit in its state_dict: nn.Linear weights are [out, in] row-major. */
typedef struct nerf_b200_params {
    const float *layer_w[8];   /* [256,63], 3x[256,256], [256,319] (in = [h256, pe63]), 3x[256,256] */
    const float *layer_b[8];   /* [256] each */
    const float *density_w;    /* [1,256] */
    const float *density_b;    /* [1] */
    const float *color0_w;     /* [128,283] (in = [h256, dir_pe27]) */
    const float *color0_b;     /* [128] */
    const float *color1_w;     /* [3,128] */
    const float *color1_b;     /* [3] */
} nerf_b200_params;

NERF_B200_API int nerf_b200_abi_version(void);
NERF_B200_API const char *nerf_b200_error_string(int code);

/* ---- weights -------------------------------------------------------------------------
 * Replaces the per-renderer weight conversion done in setup()
 * (pattern: src/benchmark/cpu_optimized_renderer.py:31-52).  `packed` is a caller-owned
 * device buffer of nerf_b200_packed_bytes() bytes, 1024-byte aligned; it holds the fp32
 * K-major copies used by the FP32 mode and the bf16 128B-swizzled K-chunks the tcgen05
 * kernel streams with bulk async copies. */
NERF_B200_API size_t nerf_b200_packed_bytes(void);
NERF_B200_API int nerf_b200_pack_weights(const nerf_b200_params *params_host, void *packed, void *stream);
/* pack_weights_ex: the same, writing only the parts a caller's modes read (a training loop re-packs every step).  The
 * bf16 operand stream, the biases and the head weights are always written; `what` adds the fp32 K-major matrices
 * (FP32 mode), the low-order bf16 stream (BF16X3) and the transposed stream of the backward chain (BF16 training).
 * Parts not selected keep whatever the buffer held. */
enum { NERF_B200_PACK_FP32_MATRICES = 1, NERF_B200_PACK_BF16_LO = 2, NERF_B200_PACK_DGRAD = 4, NERF_B200_PACK_ALL = 7 };
NERF_B200_API int nerf_b200_pack_weights_ex(const nerf_b200_params *params_host, void *packed, int what, void *stream);

/* ---- rays and samples ------------------------------------------------------------------
 * generate_rays: BaseUnifiedRenderer.generate_rays (base_renderer.py:223-258; also
 * NeRFTrainer._get_rays, trainer.py:271-292).  c2w_host = 16 floats, row-major 4x4, read
 * at call time.  Rows [row0, row0+n_rows) of the H x W image -> rays_o, rays_d
 * [n_rows*W,3].  Bit-exact with the reference's CPU result. */
NERF_B200_API int nerf_b200_generate_rays(const float *c2w_host, int width, int height, float focal,
                            int row0, int n_rows, float *rays_o, float *rays_d, void *stream);

/* sample_points: BaseUnifiedRenderer.sample_points_on_rays (base_renderer.py:260-281);
 * with t_rand != NULL ([n_rays,n_samples] uniforms) the stratified jitter of
 * VolumeRenderer.sample_points_on_rays(perturb=True) (src/utils/rendering.py:42-47).
 * -> points [n_rays,n_samples,3], z_vals [n_rays,n_samples].  Bit-exact. */
NERF_B200_API int nerf_b200_sample_points(const float *rays_o, const float *rays_d, int n_rays,
                            int n_samples, float near, float far, const float *t_rand,
                            float *points, float *z_vals, void *stream);

/* importance_sample: VolumeRenderer.importance_sample (src/utils/rendering.py:54-100) with
 * the shape fix it needs to run (oracle/nerf_oracle.py:importance_sample).  u [n_rays,n_new]
 * replaces the internal torch.rand.  -> indices [n_rays,n_new] int64 (searchsorted, right),
 * z_new [n_rays,n_new], points [n_rays,n_new,3].  n_samples must be a multiple of 32 and
 * <= 1024.  Bit-exact (indices, depths and points). */
NERF_B200_API int nerf_b200_importance_sample(const float *rays_o, const float *rays_d, const float *z_vals,
                                const float *weights, const float *u, int n_rays,
                                int n_samples, int n_new, int64_t *indices, float *z_new,
                                float *points, void *stream);

/* ---- network -------------------------------------------------------------------------
 * positional_encoding: PositionalEncoding.encode (src/models/nerf.py:24-45).
 * x [n,3] -> out [n, 3+6*n_freq]. */
NERF_B200_API int nerf_b200_positional_encoding(const float *x, int64_t n, int n_freq, float *out, void *stream);

/* query_network: BaseUnifiedRenderer.query_nerf_networks -> NeRFModel.forward
 * (base_renderer.py:165-188, src/models/nerf.py:92-131).  positions, directions [n,3] ->
 * sigma [n,1], rgb [n,3].  mode = NERF_B200_FP32 (CUDA cores) | NERF_B200_BF16 (the fused tcgen05 kernel with one
 * (point, direction) pair per row; `packed` 1024-byte aligned).  BF16X3: NERF_B200_EUNSUPPORTED. */
NERF_B200_API int nerf_b200_query_network(const void *packed, const float *positions, const float *directions,
                            int64_t n, int mode, float *sigma, float *rgb, void *stream);

/* ---- compositing -----------------------------------------------------------------------
 * composite: <Renderer>.execute_volume_rendering (src/benchmark/pytorch_renderers.py:105-125)
 * and VolumeRenderer.volume_render (src/utils/rendering.py:102-143).  sigma [R,S] (the
 * trailing 1 of [R,S,1] is implicit), rgb [R,S,3], z_vals [R,S], rays_d [R,3] ->
 * rgb_map [R,3], depth [R]; acc [R] and weights [R,S] when non-NULL. */
NERF_B200_API int nerf_b200_composite(const float *sigma, const float *rgb, const float *z_vals,
                        const float *rays_d, int n_rays, int n_samples, float *rgb_map,
                        float *depth, float *acc, float *weights, void *stream);

/* ---- fused render ------------------------------------------------------------------------
 * render_image: PyTorchCPURenderer.render_image (src/benchmark/pytorch_renderers.py:127-170):
 * ray generation, uniform sampling, positional encoding, the fine network and compositing in
 * ONE kernel; per-sample activations never reach HBM.  Renders rows [row0,row0+n_rows) of the
 * H x W image (the multi-GPU shard) -> rgb_out [n_rows*W,3], depth_out [n_rows*W].
 * Limits: n_samples <= 32768 in the BF16 / BF16X3 modes (128-sample tiles, transmittance carried across the tiles
 * of a ray), <= 2048 in FP32 mode; above: NERF_B200_EUNSUPPORTED.  n_samples == 1 renders black, as the reference
 * does (its distance tensor is empty). */
NERF_B200_API int nerf_b200_render_image(const void *packed, const float *c2w_host, int width, int height,
                           float focal, float near, float far, int n_samples, int row0,
                           int n_rows, int mode, float *rgb_out, float *depth_out, void *stream);

/* render_rays: the same fused kernel fed from ray arrays -- the forward of
 * NeRFTrainer._render_rays (src/training/trainer.py:294-316) for one network, and
 * _render_ray_chunk of the renderers.  t_rand NULL = uniform depths.  acc_out may be NULL. */
NERF_B200_API int nerf_b200_render_rays(const void *packed, const float *rays_o, const float *rays_d,
                          int n_rays, int n_samples, float near, float far, const float *t_rand,
                          int mode, float *rgb_out, float *depth_out, float *acc_out, void *stream);

/* render_rays_ex: render_rays with explicit per-ray depths in (z_vals [n_rays,n_samples], ascending; overrides
 * near/far/t_rand) and the per-sample compositing weights out (weights_out [n_rays,n_samples]) -- the two
 * hooks a hierarchical (coarse -> importance -> fine) render needs: the `weights` output of
 * VolumeRenderer.volume_render (src/utils/rendering.py:131) feeds importance_sample (:54-100). Either may be NULL. */
NERF_B200_API int nerf_b200_render_rays_ex(const void *packed, const float *rays_o, const float *rays_d, int n_rays,
                             int n_samples, float near, float far, const float *t_rand, const float *z_vals,
                             int mode, float *rgb_out, float *depth_out, float *acc_out, float *weights_out,
                             void *stream);

/* merge_samples: sorted union of z_sorted [n_rays,n_sorted] (ascending) and z_new [n_rays,n_new] (any order) ->
 * z_out [n_rays, n_sorted+n_new], equal to torch.sort(torch.cat([z, z_new], -1)).values bit for bit.  The
 * reference stops at importance_sample (its result is never consumed); the sorted union is the original-NeRF
 * recipe, needed because volume_render assumes ascending depths. */
NERF_B200_API int nerf_b200_merge_samples(const float *z_sorted, const float *z_new, int n_rays, int n_sorted, int n_new,
                            float *z_out, void *stream);

/* ---- FP8 mode (quantised weights and activations on the tensor cores) -----------------------------
 * The B200 counterpart of the reference's CompressedNeRFRenderer (src/benchmark/compressed_renderer.py:89-211:
 * per-tensor int8 weights dequantised to fp16 on the CPU): the fused render kernel with e4m3 operands for every
 * 256-wide contraction (tcgen05.mma.kind::f8f6f4, twice the bf16 rate), bf16 for the encoded-position inputs, fp32
 * accumulation, heads and compositing.  A lossy mode: judged against the compressed renderer, not against the 1e-4 /
 * 0.05 dB gates.
 *
 * pack_weights_fp8: `packed` = the network's regular packed buffer (already filled by nerf_b200_pack_weights: the
 * calibration pass runs the FP32 kernel on it), calib_positions / calib_directions [n_calib,3] = sample points of the
 * scene (the activation scales are 2^floor(log2(240 / max)) of what the fp32 network produces on them),
 * packed_fp8 = caller-owned device buffer of nerf_b200_packed_fp8_bytes() bytes, 1024-byte aligned.
 * render_image_fp8 / render_rays_fp8: nerf_b200_render_image / nerf_b200_render_rays_ex on that buffer. */
NERF_B200_API size_t nerf_b200_packed_fp8_bytes(void);
NERF_B200_API int nerf_b200_pack_weights_fp8(const nerf_b200_params *params_host, const void *packed,
                               const float *calib_positions, const float *calib_directions, int64_t n_calib,
                               void *packed_fp8, void *stream);
NERF_B200_API int nerf_b200_render_image_fp8(const void *packed_fp8, const float *c2w_host, int width, int height, float focal,
                               float near, float far, int n_samples, int row0, int n_rows, float *rgb_out,
                               float *depth_out, void *stream);
NERF_B200_API int nerf_b200_render_rays_fp8(const void *packed_fp8, const float *rays_o, const float *rays_d, int n_rays,
                              int n_samples, float near, float far, const float *t_rand, const float *z_vals,
                              float *rgb_out, float *depth_out, float *acc_out, float *weights_out, void *stream);

/* hierarchical_samples: the sampling side of a coarse -> importance -> fine render (BASELINE.json configs[4]) as ONE
 * kernel: VolumeRenderer.sample_points_on_rays (the coarse depths: uniform, or stratified with t_rand [n_rays,n_samples])
 * + VolumeRenderer.importance_sample (src/utils/rendering.py:54-100, with the shape fix; `weights` [n_rays,n_samples]
 * = the coarse pass's compositing weights, e.g. weights_out of render_rays_ex) + the sorted union.  Only
 * z_out [n_rays, n_samples + n_new] is written: bit for bit merge_samples(z, importance_sample(z, weights, u).z_new)
 * with z = sample_points(..).z_vals.  u [n_rays,n_new] = the uniforms of rendering.py:79; u == NULL draws them in the
 * kernel (Philox4x32-10, key `seed`, 24-bit mantissas like torch.rand).  n_samples % 32 == 0, both counts <= 1024. */
NERF_B200_API int nerf_b200_hierarchical_samples(const float *weights, int n_rays, int n_samples, int n_new, float near,
                                   float far, const float *t_rand, const float *u, uint64_t seed, float *z_out,
                                   void *stream);

/* ---- data path ---------------------------------------------------------------------------
 * composite_white: the per-image conversion of SyntheticDataset.__init__ (src/data/loader.py:46-54) for RGBA8 pixels
 * already resized on the host: img/255 in float64, rgb*alpha + (1-alpha), stored as fp32 -- bit-exact with the
 * reference's numpy float64 arithmetic followed by torch.FloatTensor.  rgba [n_pixels,4] uint8 (device) ->
 * rgb_out [n_pixels,3] fp32. */
NERF_B200_API int nerf_b200_composite_white(const unsigned char *rgba, int64_t n_pixels, float *rgb_out, void *stream);

/* ray_batch: the ray selection of NeRFTrainer.train_step (src/training/trainer.py:100-118): for pixel_index [n]
 * (int64, row-major pixel = row * width + col, device) -> rays_o, rays_d [n,3] with the bits of
 * _get_rays(pose, (H, W), focal)[index] (trainer.py:271-292; SyntheticDataset.get_rays, loader.py:78-108) and, when
 * `target` != NULL, target [n,3] = image[index] gathered from image [H*W,3] (device).  The full-image ray tensors the
 * reference builds every step never exist. */
NERF_B200_API int nerf_b200_ray_batch(const float *c2w_host, int width, int height, float focal, const int64_t *pixel_index,
                        int n, const float *image, float *rays_o, float *rays_d, float *target, void *stream);

/* ---- training ----------------------------------------------------------------------------
 * train_fwd_bwd: forward + backward of ONE network's term of the photometric loss in
 * NeRFTrainer.train_step (src/training/trainer.py:117-126):
 *   loss_term = mean_{R x 3} (C - target)^2,  C = render_rays(...)
 * d loss/d params is ACCUMULATED (+=) into `grads` (same 22 tensors / shapes as params,
 * device pointers), scaled by grad_scale / (3 * n_rays_global) * 2 so that data-parallel ranks
 * can pass their global ray count and all-reduce-sum.  loss_sum (device, 1 float) += sum of
 * squared errors over this call's rays (divide by 3*R for the mean).  rgb_out [R,3] optional.
 * workspace: caller-owned, nerf_b200_train_workspace_bytes(n_rays, n_samples) bytes. */
NERF_B200_API size_t nerf_b200_train_workspace_bytes(int n_rays, int n_samples);
NERF_B200_API int nerf_b200_train_fwd_bwd(const void *packed, const nerf_b200_params *params_dev_ptrs_host,
                            const nerf_b200_params *grads_dev_ptrs_host, const float *rays_o,
                            const float *rays_d, const float *target, int n_rays, int n_samples,
                            float near, float far, const float *t_rand, int n_rays_global,
                            int mode, void *workspace, float *loss_sum, float *rgb_out,
                            void *stream);

/* Split form for callers that overlap the two networks of a step (B200TrainStep): the activation phase of one
 * network (forward, compositing backward, dgrad chain -- bound by HBM *writes*) can run beside the weight-gradient
 * phase of the other (bound by HBM *reads*) on disjoint SMs.  `phases`: ACTIVATIONS leaves everything the
 * weight-gradient phase needs in `workspace`; WEIGHT_GRADS consumes it (same arguments, same workspace, later on
 * any stream ordered after the first call) and accumulates into `grads`; ALL = both = nerf_b200_train_fwd_bwd.
 * Split phases need mode BF16 and a batch that fits one workspace chunk (524288 samples), else
 * NERF_B200_EUNSUPPORTED.  `sm_limit` > 0 caps the CTAs of every launch of this call (0 = all SMs). */
enum { NERF_B200_TRAIN_ACTIVATIONS = 1, NERF_B200_TRAIN_WEIGHT_GRADS = 2, NERF_B200_TRAIN_ALL = 3 };
NERF_B200_API int nerf_b200_train_fwd_bwd_ex(const void *packed, const nerf_b200_params *params_dev_ptrs_host,
                               const nerf_b200_params *grads_dev_ptrs_host, const float *rays_o,
                               const float *rays_d, const float *target, int n_rays, int n_samples,
                               float near, float far, const float *t_rand, int n_rays_global,
                               int mode, void *workspace, float *loss_sum, float *rgb_out,
                               int phases, int sm_limit, void *stream);

/* ---- optimizer step + data-parallel gradient exchange -------------------------------------------
 * Replaces the tail of NeRFTrainer.train_step (src/training/trainer.py:125-136): clip_grad_norm_ over both networks,
 * Adam.step() (weight_decay = L2 in the gradient), and -- new, the reference is single-device -- the sum of the
 * gradients over data-parallel ranks, done by this library's own kernels over NVLink peer memory instead of a
 * collective library call.
 *
 * All parameters of a step live in ONE flat fp32 bucket; gradients, exp_avg and exp_avg_sq likewise (same offsets).
 * The gradient bucket G and the reduced bucket Gsum sit in a SYMMETRIC allocation of nerf_b200_dp_bytes(n) bytes per
 * rank, laid out [ctl: NERF_B200_DP_CTL_BYTES | G: n floats | Gsum: n floats], zero-initialised, which every rank maps
 * from every other rank (the caller obtains the mappings, e.g. torch symmetric memory; `peer[q]` is rank q's
 * allocation as mapped in THIS process, `multicast` optionally the NVSwitch multicast mapping of the same
 * allocation).  n must be a multiple of 4 * world; floats [0, n_opt) are parameters' gradients (zero in the pads),
 * floats [n_opt, n) ride along un-optimised (slot n_opt carries the loss).  `state`: NERF_B200_DP_STATE_BYTES of
 * zeroed private device memory per rank.
 *
 *   nerf_b200_dp_reduce     after the backward kernels on the same stream: waits until every rank's G is final, sums
 *                           this rank's 1/world shard over the ranks in rank order (or with one in-switch
 *                           multimem.ld_reduce when `multicast` is set) and stores the sums into every rank's Gsum.
 *   nerf_b200_dp_adam_step  waits until every shard arrived; gradient norm, clip (max_norm <= 0: off), Adam on the
 *                           whole bucket, G zeroed for the next step.  hyper (device, 8 DOUBLES, static): lr0, gamma,
 *                           beta1, beta2, eps, weight_decay, max_norm, loss_scale.  The step count t lives on the
 *                           device (`state`, uint32 at byte offset NERF_B200_DP_STATE_OPT_STEP; the kernel increments
 *                           it; the host sets it on resume): lr_t = lr0 gamma^t (ExponentialLR stepped once per
 *                           iteration, trainer.py:62-64,136) and Adam's 1 - beta^(t+1) are formed in the kernel, so a
 *                           step costs no host-to-device traffic and the launch pair can sit in a CUDA graph.
 *                           loss_out (device, 3 floats, optional): [0] = summed slot n_opt * loss_scale, [1] = gradient
 *                           norm before clipping, [2] = the learning rate used.
 *
 * Every rank computes the same bits (fixed summation order, replicated update).  world == 1 needs no peer: pass
 * peer[0] = the local allocation.  Every rank must call both functions once per step, in this order; a rank that
 * never arrives makes the others trap after ~20 s instead of hanging.  `emulate_sequential` (tests): no flag waits --
 * several "ranks" whose allocations live on one GPU are run as reduce(0..W-1) then adam_step(0..W-1) on one stream. */
#define NERF_B200_DP_MAX_WORLD 16
#define NERF_B200_DP_CTL_BYTES 1024
#define NERF_B200_DP_STATE_BYTES 2048
#define NERF_B200_DP_STATE_OPT_STEP 12
typedef struct nerf_b200_dp {
    int rank, world;
    int64_t n, n_opt;
    void *peer[NERF_B200_DP_MAX_WORLD];
    void *multicast;
    void *state;
    int emulate_sequential;
} nerf_b200_dp;
NERF_B200_API size_t nerf_b200_dp_bytes(int64_t n);
NERF_B200_API int nerf_b200_dp_reduce(const nerf_b200_dp *dp_host, void *stream);
NERF_B200_API int nerf_b200_dp_adam_step(const nerf_b200_dp *dp_host, float *params, float *exp_avg, float *exp_avg_sq,
                           const double *hyper, float *loss_out, void *stream);

/* Environment (read ONCE per process, at the first tensor-core render launch; for A/B measurements only):
 * NERF_B200_CLUSTER=1 makes every CTA stream the whole weight set itself instead of sharing the stream inside 2-CTA
 * clusters (the default, 2). */

/* ---- introspection (tests / bench) -------------------------------------------------------
 * Number of kernel launches this library has enqueued since load (bench.py's gpu_launches). */
NERF_B200_API uint64_t nerf_b200_launch_count(void);
/* Optional watchdog word (device, 4 bytes, zero it first), kept PER DEVICE: the word is filed under the device that
 * owns the pointer and only handed to launches on that device; NULL detaches the current device's word.  While a word
 * is attached the tensor-core kernels' barrier waits are bounded (~2 s of SM clocks): on a timeout the kernel stores
 * tag | code<<16 | block here (tag 0x8 render/forward, 0x9 dgrad chain, 0xA wgrad) and traps, so a protocol bug fails
 * the launch instead of hanging the device.  With no word attached (production) a wait never traps -- preemption,
 * a debugger or throttled clocks may stretch it -- it backs off with nanosleep and keeps waiting. */
NERF_B200_API void nerf_b200_set_watchdog_word(unsigned int *device_word);
/* Optional timeline buffer (device, 6*9*8 int64, zeroed): CTA 0 of render_image(BF16) stores clock64
 * stamps of its MMA issuer and one epilogue warp for its first 6 tiles (tools/tc_trace.py).  Per device like the
 * watchdog word; NULL detaches the current device's buffer. */
NERF_B200_API void nerf_b200_set_trace_buffer(long long *device_buf);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H */
