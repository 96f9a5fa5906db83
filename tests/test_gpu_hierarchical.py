"""GPU: the hierarchical path of BASELINE.json configs[4] -- weights out of the fused kernel, inverse-CDF
samples, sorted union, fine pass at explicit depths.  Stages are checked against the oracle one by one
(bit-exact where they are index/position work); the end-to-end composition is checked in the fp32 mode."""
import numpy as np
import pytest
import torch

from gpu_util import Watchdog, packed_net
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


def rays(poses, w=31, h=17):
    ro, rd = O.camera_rays(poses["generic"], w, h)
    return ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()


def test_merge_is_bit_exact_sorted_union():
    from nerf_dbr_b200.host import ops
    g = torch.Generator().manual_seed(0)
    for na, nb in ((128, 128), (64, 128), (32, 7), (96, 1)):
        a = torch.sort(torch.rand(333, na, generator=g) * 4 + 2, dim=-1).values
        b = torch.rand(333, nb, generator=g) * 4 + 2
        b[:, : min(5, nb)] = a[:, : min(5, nb)]                       # exact ties between the two sets
        if nb > 8:
            b[:, 6] = b[:, 7]                                          # ties inside the new set
        ref = torch.sort(torch.cat([a, b], -1), -1).values
        out = ops.merge_samples(a.cuda(), b.cuda())
        assert torch.equal(out.cpu(), ref)


# (mode, weights tolerance, rgb/depth/acc tolerance): fp32 and bf16x3 are held to the 1e-4 gate
@pytest.mark.parametrize("mode,tol,tol_img", [(0, 5e-6, 1e-4), (2, 1e-4, 1e-4), (1, 3e-2, 1e-1)])
def test_weights_output_and_explicit_depths(mode, tol, tol_img, checkpoints, poses):
    from nerf_dbr_b200.host import ops
    w = checkpoints["lego"]["fine_model"]
    net = packed_net(w)
    ro, rd = rays(poses)
    with Watchdog() as wd:
        for S in (32, 128):
            _, z = O.sample_along_rays(ro, rd, S)
            ref = O.render_rays_at(w, ro, rd, z.contiguous())
            rgb, dep, acc, wts = ops.render_rays(net, ro.cuda(), rd.cuda(), S, mode=mode, want_acc=True, want_weights=True)
            assert (wts.cpu() - ref[3]).abs().max() <= tol and (acc.cpu() - ref[2]).abs().max() <= tol_img
        # explicit, non-uniform, ascending depths (256 = two tiles per ray in the tensor-core kernel)
        g = torch.Generator().manual_seed(4)
        z = torch.sort(torch.rand(ro.shape[0], 256, generator=g) * 4 + 2, dim=-1).values
        ref = O.render_rays_at(w, ro, rd, z)
        rgb, dep, wts = ops.render_rays(net, ro.cuda(), rd.cuda(), 256, mode=mode, z_vals=z.cuda(), want_weights=True)
        torch.cuda.synchronize()
        assert int(wd.word.item()) == 0
        assert (wts.cpu() - ref[3]).abs().max() <= tol
        assert (rgb.cpu() - ref[0]).abs().max() <= tol_img and (dep.cpu() - ref[1]).abs().max() <= tol_img


def test_hierarchical_render_end_to_end_fp32(checkpoints, poses):
    """128 coarse + 128 importance samples.  fp32 mode: the coarse weights agree with the oracle to ~1e-6, so
    the inverse-CDF indices agree except where u falls within that distance of a cdf step."""
    from nerf_dbr_b200.host import ops
    cw, fw = checkpoints["lego"]["coarse_model"], checkpoints["semi30"]["fine_model"]
    ro, rd = rays(poses, 23, 13)
    u = torch.rand(ro.shape[0], 128, generator=torch.Generator().manual_seed(9))
    ref_f, ref_d, ref_c, z_ref, _ = O.render_hierarchical(cw, fw, ro, rd, 128, u)
    rgb_f, dep_f, rgb_c, z_all = ops.render_hierarchical(packed_net(cw), packed_net(fw), ro.cuda(), rd.cuda(), 128, 128,
                                                         mode=0, u=u.cuda())
    assert z_all.shape == (ro.shape[0], 256)
    assert bool((z_all[:, 1:] >= z_all[:, :-1]).all())
    assert (rgb_c.cpu() - ref_c).abs().max() <= 1e-4
    # the inverse CDF is continuous in the weights: 1e-6 differences in the coarse weights move a new depth by
    # ulps (or, inside a near-empty bin, by a visible fraction of that bin) -- never to another region
    dz = (z_all.cpu() - z_ref).abs()
    assert (dz == 0).float().mean().item() >= 0.98 and (dz <= 1e-4).float().mean().item() >= 0.995, dz.max()
    assert (rgb_f.cpu() - ref_f).abs().max() <= 2e-3 and (dep_f.cpu() - ref_d).abs().max() <= 2e-2


def test_hierarchical_render_bf16_runs_at_config5_shape(checkpoints, poses):
    """Tensor-core mode at the config-5 sample counts: finite, sorted, close to the fp32-mode result."""
    from nerf_dbr_b200.host import ops
    cw = fw = checkpoints["lego"]["fine_model"]
    ro, rd = rays(poses, 40, 30)
    u = torch.rand(ro.shape[0], 128, generator=torch.Generator().manual_seed(2)).cuda()
    c, f = packed_net(cw), packed_net(fw)
    with Watchdog() as wd:
        rgb16, dep16, _, z16 = ops.render_hierarchical(c, f, ro.cuda(), rd.cuda(), 128, 128, mode=1, u=u)
        rgb32, dep32, _, z32 = ops.render_hierarchical(c, f, ro.cuda(), rd.cuda(), 128, 128, mode=0, u=u)
        torch.cuda.synchronize()
        assert int(wd.word.item()) == 0
    assert torch.isfinite(rgb16).all() and bool((z16[:, 1:] >= z16[:, :-1]).all())
    mse = float(((rgb16 - rgb32) ** 2).mean())
    assert -10 * np.log10(mse) >= 45.0


@pytest.mark.parametrize("S,n_new,jitter", [(128, 128, False), (64, 128, True), (32, 7, False), (96, 100, True), (128, 1, False)])
def test_fused_sampling_kernel_equals_the_three_stage_path(S, n_new, jitter, poses):
    """nerf_b200_hierarchical_samples (coarse depths + inverse CDF + sorted union, one launch, only z_all written) against
    sample_points -> importance_sample -> merge_samples, which the goldens pin to the reference: bit for bit, including
    degenerate rays (all-zero weights), u = 0 and u just below 1, and stratified coarse depths."""
    from nerf_dbr_b200.host import ops
    ro, rd = rays(poses, 37, 21)
    ro, rd = ro.cuda(), rd.cuda()
    n = ro.shape[0]
    g = torch.Generator().manual_seed(S + n_new)
    w = (torch.rand(n, S, generator=g) ** 6).cuda()
    w[:9] = 0.0
    u = torch.rand(n, n_new, generator=g)
    u[:, 0] = 0.0
    if n_new > 1:
        u[:, 1] = 0.99999994
    u = u.cuda()
    tr = torch.rand(n, S, generator=g).cuda() if jitter else None
    _, z = ops.sample_points(ro, rd, S, t_rand=tr)
    _, z_new, _ = ops.importance_sample(ro, rd, z, w, u)
    ref = ops.merge_samples(z, z_new)
    out = ops.hierarchical_samples(w, n_new, t_rand=tr, u=u)
    assert torch.equal(out, ref)
    assert torch.equal(out.cpu(), torch.sort(torch.cat([z, z_new], -1), -1).values.cpu())


@pytest.mark.parametrize("S_", [32, 64, 128])
def test_fused_sampling_kernel_against_the_reference_golden(S_):
    """The golden importance fixture (reference source + the documented one-line fix, tests/golden/make_golden.py): uniform
    coarse depths from near = 2, far = 6, the reference's z_new -> the fused kernel's union must be their sorted cat."""
    from conftest import load_npz
    from nerf_dbr_b200.host import ops
    g = load_npz("golden_importance.npz")
    z, w, u, z_new = (torch.from_numpy(g[f"S{S_}|{k}"]) for k in ("z", "w", "u", "z_new"))
    out = ops.hierarchical_samples(w.cuda(), u.shape[1], 2.0, 6.0, u=u.cuda())
    assert torch.equal(out.cpu(), torch.sort(torch.cat([z, z_new], -1), -1).values)


def test_fused_sampling_kernel_draws_its_own_uniforms():
    """u = None: Philox in the kernel.  Deterministic per seed, different across seeds, the union is sorted and
    contains every coarse depth, and the new samples follow the piecewise-constant pdf (all the mass in a quarter of the
    bins -> nearly all samples land in [z16, z32], uniformly)."""
    from nerf_dbr_b200.host import ops
    n, S, k = 4096, 64, 128
    w = torch.zeros(n, S)
    w[:, 16:32] = 1.0                                       # all the mass in bins 16..31 of 64 (plus the 1e-5 floor)
    w = w.cuda()
    a, b, c = (ops.hierarchical_samples(w, k, seed=s) for s in (7, 7, 8))
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert bool((a[:, 1:] >= a[:, :-1]).all())
    _, z = ops.sample_points(torch.zeros(n, 3).cuda(), torch.ones(n, 3).cuda(), S)
    lo, hi = float(z[0, 15]), float(z[0, 32])
    inside = ((a > lo) & (a < hi)).float().sum(-1) - 16     # minus the 16 coarse depths strictly inside
    frac = float(inside.mean()) / k
    assert 0.97 <= frac <= 1.0, frac
    # uniformity inside the mass: bin i of the pdf spans [z_i, z_i+1], so the new samples' mean sits at the middle of [z16, z32]
    new_mean = float(((a * ((a > lo) & (a < hi))).sum() - z[:, 16:32].sum()) / inside.sum())
    assert abs(new_mean - 0.5 * (float(z[0, 16]) + float(z[0, 32]))) <= 0.02 * (hi - lo), new_mean


def test_hierarchical_against_the_oracle_in_every_mode(checkpoints, poses):
    """configs[4] composition against the oracle's hierarchical render (same u): BF16X3 tight, BF16 by PSNR.  The
    importance samples depend on the coarse weights, so bf16 coarse weights move them slightly -- a different, equally
    valid quadrature of the same integral; the gate is on the image."""
    from nerf_dbr_b200.host import ops
    from gpu_util import psnr
    cw, fw = checkpoints["lego"]["coarse_model"], checkpoints["lego"]["fine_model"]
    ro, rd = rays(poses, 48, 36)
    u = torch.rand(ro.shape[0], 128, generator=torch.Generator().manual_seed(3))
    ref_f, ref_d, ref_c, _, _ = O.render_hierarchical(cw, fw, ro, rd, 128, u)
    c, f = packed_net(cw), packed_net(fw)
    with Watchdog() as wd:
        out = {m: ops.render_hierarchical(c, f, ro.cuda(), rd.cuda(), 128, 128, mode=m, u=u.cuda()) for m in (2, 1)}
        torch.cuda.synchronize()
        assert int(wd.word.item()) == 0
    e3 = (out[2][0].cpu() - ref_f).abs().max().item()
    p16 = psnr(out[1][0].cpu().numpy(), ref_f.numpy())
    print(f"hierarchical 128+128 vs oracle: BF16X3 max-abs {e3:.2e}; BF16 PSNR {p16:.1f} dB")
    assert e3 <= 1e-4              # measured 2.0e-5
    assert p16 >= 50.0             # measured 63.5 dB
