/*
 * scalar_oracle.c -- scalar C restatement of the bit-exact pieces of nerf-dbr's
 * render path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): loaded by
 * tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg through
 * oracle/scalar.py; never linked into or called from the product library.
 *
 * Each function states, operation by operation and rounding by rounding, what
 * the reference's torch-CPU expression evaluates to ("fl" = round to nearest
 * fp32).  The recipes were pinned against the reference itself
 * (tests/golden/make_golden.py) and are what the CUDA kernels mirror with
 * __fmul_rn/__fadd_rn/__fdiv_rn/fmaf.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).  Contraction
 * must stay off: every product and sum below is a separate rounding unless it is
 * written as fmaf().
 *
 * Reference lines (paths relative to the nerf-dbr root):
 *   so_linspace / so_z_vals   src/benchmark/base_renderer.py:274-275 (torch.linspace, affine)
 *   so_camera_rays            src/benchmark/base_renderer.py:238-256
 *   so_points                 src/benchmark/base_renderer.py:279
 *   so_stratified             src/utils/rendering.py:42-47
 *   so_encode                 src/models/nerf.py:40-45
 *   so_mlp                    src/models/nerf.py:108-129
 *   so_composite              src/benchmark/pytorch_renderers.py:105-125, src/utils/rendering.py:117-141
 *   so_importance             src/utils/rendering.py:73-95 (with the z_vals-gather shape fix)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SO_API __attribute__((visibility("default")))

/* torch.linspace(start, end, n), fp32, as ATen's vectorised CPU kernel computes it:
 * step = fl((end-start)/fl(n-1)); the lower half counts up from start and the upper
 * half counts down from end, each with ONE rounding (fused multiply-add). */
SO_API void so_linspace(float start, float end, int n, float *out)
{
    if (n <= 0) return;
    if (n == 1) { out[0] = start; return; }
    float step = (end - start) / (float)(n - 1);
    int half = n / 2;
    for (int i = 0; i < n; ++i)
        out[i] = (i < half) ? fmaf(step, (float)i, start)
                            : fmaf(-step, (float)(n - 1 - i), end);
}

/* z_i = fl( fl(near*fl(1-t_i)) + fl(far*t_i) ), t = linspace(0,1,S). */
SO_API void so_z_vals(int n_samples, float near, float far, float *z)
{
    so_linspace(0.0f, 1.0f, n_samples, z);
    for (int i = 0; i < n_samples; ++i) {
        float t = z[i];
        float a = near * (1.0f - t);
        float b = far * t;
        z[i] = a + b;
    }
}

/* rays_o, rays_d: [H*W,3], pixel (row j, col i) at index j*W+i.  Directions are not
 * normalised.  dx = fl(fl(i - fl(W/2))/f), dy = fl(-(fl(j - fl(H/2))/f)), dz = -1;
 * rays_d[c] = fl(fl(fl(+0 + fl(dx*R[c][0])) + fl(dy*R[c][1])) + fl(dz*R[c][2])): torch.sum
 * starts from a +0 accumulator, which only shows when all three products are -0
 * (the result is then +0, e.g. the y component of the centre row of an axis-aligned
 * camera). */
SO_API void so_camera_rays(const float *c2w, int width, int height, float focal,
                           float *rays_o, float *rays_d)
{
    float half_w = (float)((double)width * 0.5);
    float half_h = (float)((double)height * 0.5);
    for (int j = 0; j < height; ++j) {
        for (int i = 0; i < width; ++i) {
            float dx = ((float)i - half_w) / focal;
            float dy = -(((float)j - half_h) / focal);
            float dz = -1.0f;
            float *o = rays_o + 3 * ((size_t)j * width + i);
            float *d = rays_d + 3 * ((size_t)j * width + i);
            for (int c = 0; c < 3; ++c) {
                float p0 = dx * c2w[4 * c + 0];
                float p1 = dy * c2w[4 * c + 1];
                float p2 = dz * c2w[4 * c + 2];
                float s = 0.0f + p0;
                s = s + p1;
                d[c] = s + p2;
                o[c] = c2w[4 * c + 3];
            }
        }
    }
}

/* points[r][s][c] = fl(o[r][c] + fl(d[r][c] * z)), z = z_vals[r*z_stride + s]
 * (z_stride = 0 when every ray shares one depth row). */
SO_API void so_points(const float *rays_o, const float *rays_d, const float *z_vals,
                      int n_rays, int n_samples, int z_stride, float *points)
{
    for (int r = 0; r < n_rays; ++r)
        for (int s = 0; s < n_samples; ++s) {
            float z = z_vals[(size_t)r * z_stride + s];
            for (int c = 0; c < 3; ++c) {
                float m = rays_d[3 * r + c] * z;
                points[((size_t)r * n_samples + s) * 3 + c] = rays_o[3 * r + c] + m;
            }
        }
}

/* Stratified jitter given the uniform depths z[S] and t_rand[R][S]:
 * mid_i = fl(0.5*fl(z_{i+1}+z_i)); lower = [z_0, mid...], upper = [mid..., z_{S-1}];
 * z' = fl(lower + fl(fl(upper-lower)*t)). */
SO_API void so_stratified(const float *z, const float *t_rand, int n_rays, int n_samples,
                          float *z_out)
{
    for (int r = 0; r < n_rays; ++r)
        for (int s = 0; s < n_samples; ++s) {
            float lo = (s == 0) ? z[0] : 0.5f * (z[s] + z[s - 1]);
            float hi = (s == n_samples - 1) ? z[n_samples - 1] : 0.5f * (z[s + 1] + z[s]);
            float span = hi - lo;
            float j = span * t_rand[(size_t)r * n_samples + s];
            z_out[(size_t)r * n_samples + s] = lo + j;
        }
}

/* out[n][3 + 6L]: [x | sin(c_0 x) | cos(c_0 x) | sin(c_1 x) | ...], c_k = fl(2^k * pi)
 * = 2^k * fl32(pi), argument = fl(c_k * x) (one rounding).  sinf/cosf are libm's
 * (<= 1 ulp; torch uses SLEEF, also <= 1 ulp -- last-bit differences are possible,
 * which is why encoded features are a tolerance gate, the *arguments* a bit gate). */
SO_API void so_encode(const float *x, int n, int n_freq, float *out)
{
    const float pi_f = 3.14159274101257324f; /* fl32(pi) */
    int width = 3 + 6 * n_freq;
    for (int i = 0; i < n; ++i) {
        float *o = out + (size_t)i * width;
        for (int c = 0; c < 3; ++c) o[c] = x[3 * i + c];
        float ck = pi_f;
        for (int k = 0; k < n_freq; ++k) {
            for (int c = 0; c < 3; ++c) {
                float arg = ck * x[3 * i + c];
                o[3 + 6 * k + c] = sinf(arg);
                o[3 + 6 * k + 3 + c] = cosf(arg);
            }
            ck = ck * 2.0f;
        }
    }
}

SO_API void so_encode_args(const float *x, int n, int n_freq, float *args)
{
    const float pi_f = 3.14159274101257324f;
    for (int i = 0; i < n; ++i) {
        float ck = pi_f;
        for (int k = 0; k < n_freq; ++k) {
            for (int c = 0; c < 3; ++c)
                args[((size_t)i * n_freq + k) * 3 + c] = ck * x[3 * i + c];
            ck = ck * 2.0f;
        }
    }
}

static void dense(const float *w, const float *b, const float *x, int n_out, int n_in,
                  float *y, int relu)
{
    for (int o = 0; o < n_out; ++o) {
        float acc = 0.0f;
        for (int k = 0; k < n_in; ++k) acc = fmaf(w[(size_t)o * n_in + k], x[k], acc);
        acc = acc + b[o];
        y[o] = (relu && acc < 0.0f) ? 0.0f : acc;
    }
}

/* params: 22 pointers in state-dict order (layers.0.weight, layers.0.bias, ...,
 * layers.7.bias, density_head.weight, .bias, color_layers.0.weight, .bias,
 * color_layers.1.weight, .bias).  Naive fp32 dot products (sequential fmaf): agrees with
 * the MKL-backed reference to fp32 rounding noise, not bit-for-bit. */
SO_API void so_mlp(const float *const *params, const float *pos, const float *dir, int n,
                   float *sigma, float *rgb)
{
    float pe[63], de[27], h[256 + 63], t[256 + 63], c[128], y[3];
    for (int i = 0; i < n; ++i) {
        so_encode(pos + 3 * i, 1, 10, pe);
        so_encode(dir + 3 * i, 1, 4, de);
        dense(params[0], params[1], pe, 256, 63, h, 1);
        for (int l = 1; l < 8; ++l) {
            int n_in = 256;
            if (l == 4) { memcpy(h + 256, pe, sizeof(pe)); n_in = 319; }
            dense(params[2 * l], params[2 * l + 1], h, 256, n_in, t, 1);
            memcpy(h, t, 256 * sizeof(float));
        }
        dense(params[16], params[17], h, 1, 256, sigma + i, 1);
        memcpy(t, h, 256 * sizeof(float));
        memcpy(t + 256, de, sizeof(de));
        dense(params[18], params[19], t, 128, 283, c, 1);
        dense(params[20], params[21], c, 3, 128, y, 0);
        for (int k = 0; k < 3; ++k) rgb[3 * i + k] = 1.0f / (1.0f + expf(-y[k]));
    }
}

/* Alpha compositing.  The running transmittance product is kept in double and rounded
 * to fp32 per element, as ATen's CPU cumprod does for float inputs.
 * outputs: rgb_map[R][3], depth[R], acc[R] (nullable), weights[R][S] (nullable). */
SO_API void so_composite(const float *sigma, const float *rgb, const float *z_vals,
                         const float *rays_d, int n_rays, int n_samples,
                         float *rgb_map, float *depth, float *acc, float *weights)
{
    for (int r = 0; r < n_rays; ++r) {
        const float *d = rays_d + 3 * r;
        float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        double run = 1.0;
        float cr = 0.0f, cg = 0.0f, cb = 0.0f, cd = 0.0f, ca = 0.0f;
        for (int s = 0; s < n_samples; ++s) {
            size_t i = (size_t)r * n_samples + s;
            float dz = (s + 1 < n_samples) ? z_vals[i + 1] - z_vals[i] : 1e10f;
            float dist = dz * nrm;
            float sg = sigma[i] > 0.0f ? sigma[i] : 0.0f;
            /* n_samples == 1: the reference's dists tensor is empty ([R,0]) and every output is 0 */
            float a = n_samples > 1 ? 1.0f - expf((-sg) * dist) : 0.0f;
            float trans = (float)run;               /* exclusive product */
            float keep = (1.0f - a) + 1e-10f;
            run *= (double)keep;
            float w = a * trans;
            cr += w * rgb[3 * i + 0];
            cg += w * rgb[3 * i + 1];
            cb += w * rgb[3 * i + 2];
            cd += w * z_vals[i];
            ca += w;
            if (weights) weights[i] = w;
        }
        rgb_map[3 * r + 0] = cr; rgb_map[3 * r + 1] = cg; rgb_map[3 * r + 2] = cb;
        depth[r] = cd;
        if (acc) acc[r] = ca;
    }
}

/* torch.sum(x, dim=-1) of one contiguous fp32 row whose length is a multiple of 32: four
 * 8-lane accumulators over 32-element chunks, combined ((a0+a1)+a2)+a3, lanes added 0..7. */
static float row_sum_vec(const float *x, int n)
{
    float acc[4][8];
    memset(acc, 0, sizeof(acc));
    for (int c = 0; c + 32 <= n; c += 32)
        for (int a = 0; a < 4; ++a)
            for (int l = 0; l < 8; ++l) acc[a][l] += x[c + 8 * a + l];
    float lane[8];
    for (int l = 0; l < 8; ++l) lane[l] = ((acc[0][l] + acc[1][l]) + acc[2][l]) + acc[3][l];
    float s = lane[0];
    for (int l = 1; l < 8; ++l) s += lane[l];
    return s;
}

/* Inverse-CDF sampling for n_samples in {32,64,128,256} (the row-sum recipe above is
 * verified for those).  z_vals[R][S], weights[R][S], u[R][n] -> idx[R][n] (int64, the
 * searchsorted(right=True) result), z_new[R][n].  Returns only the new samples, in the
 * order of u (unsorted, not merged) exactly like the reference function. */
SO_API int so_importance(const float *z_vals, const float *weights, const float *u,
                         int n_rays, int n_samples, int n_new, int64_t *idx, float *z_new)
{
    if (n_samples % 32 != 0) return -1;
    float *w = (float *)malloc(sizeof(float) * n_samples);
    float *cdf = (float *)malloc(sizeof(float) * (n_samples + 1));
    for (int r = 0; r < n_rays; ++r) {
        const float *z = z_vals + (size_t)r * n_samples;
        for (int s = 0; s < n_samples; ++s) w[s] = weights[(size_t)r * n_samples + s] + 1e-5f;
        float total = row_sum_vec(w, n_samples);
        double run = 0.0;
        cdf[0] = 0.0f;
        for (int s = 0; s < n_samples; ++s) {
            float pdf = w[s] / total;
            run += (double)pdf;
            cdf[s + 1] = (float)run;
        }
        for (int k = 0; k < n_new; ++k) {
            float uk = u[(size_t)r * n_new + k];
            int lo = 0, hi = n_samples + 1;            /* first index with cdf > u */
            while (lo < hi) { int mid = (lo + hi) / 2; if (cdf[mid] <= uk) lo = mid + 1; else hi = mid; }
            int id = lo;
            int below = id - 1; if (below < 0) below = 0; if (below > n_samples - 1) below = n_samples - 1;
            int above = id;     if (above > n_samples - 1) above = n_samples - 1;
            float den = cdf[above] - cdf[below];
            if (den < 1e-5f) den = 1.0f;
            float t = (uk - cdf[below]) / den;
            float span = z[above] - z[below];
            float step = t * span;
            z_new[(size_t)r * n_new + k] = z[below] + step;
            idx[(size_t)r * n_new + k] = id;
        }
    }
    free(w); free(cdf);
    return 0;
}
