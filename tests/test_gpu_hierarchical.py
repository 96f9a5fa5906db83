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
