"""GPU, FP32 mode (CUDA-core FFMA): rendered RGB/depth within max-abs 1e-4 of the reference's
PyTorchCPURenderer (golden fixtures) -- north_star tolerance for the fp32 mode."""
import numpy as np
import pytest
import torch

from conftest import load_npz
from gpu_util import packed_net
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4          # BASELINE.json north_star: "max-abs 1e-4 in an fp32 mode"


def test_query_network_fp32_matches_golden(checkpoints):
    from nerf_dbr_b200.host import ops
    g = load_npz("golden_network.npz")
    pos, dirs = torch.from_numpy(g["pos"]).cuda(), torch.from_numpy(g["dirs"]).cuda()
    for cname in ("trained11", "lego", "semi30"):
        net = packed_net(checkpoints[cname]["fine_model"])
        sigma, rgb = ops.query_network(net, pos, dirs)
        ref_s, ref_c = g[f"{cname}|sigma"], g[f"{cname}|rgb"]
        assert sigma.shape == (pos.shape[0], 1) and rgb.shape == (pos.shape[0], 3)
        # sigma is unbounded (hundreds for trained weights): relative to its scale
        assert np.abs(sigma.cpu().numpy() - ref_s).max() <= 2e-5 * max(1.0, np.abs(ref_s).max())
        assert np.abs(rgb.cpu().numpy() - ref_c).max() <= 2e-5


def test_render_image_fp32_matches_golden(checkpoints, poses):
    from nerf_dbr_b200.host import ops
    g = load_npz("golden_render.npz")
    keys = sorted({k.rsplit("|", 1)[0] for k in g.files})
    worst = 0.0
    for k in keys:
        cname, pname, dims = k.split("|")
        w, h, s = (int(x) for x in dims.split("x"))
        net = packed_net(checkpoints[cname]["fine_model"])
        rgb, dep = ops.render_image(net, poses[pname], w, h, s, mode=0)
        assert rgb.shape == (h, w, 3) and dep.shape == (h, w)
        e_rgb = np.abs(rgb.cpu().numpy() - g[k + "|rgb"]).max()
        e_dep = np.abs(dep.cpu().numpy() - g[k + "|depth"]).max()
        worst = max(worst, e_rgb, e_dep)
        assert e_rgb <= TOL and e_dep <= TOL, (k, e_rgb, e_dep)
    print("fp32 mode worst max-abs vs reference:", worst)


def test_render_rays_fp32_jitter_and_ragged(checkpoints, poses):
    """Ray-array entry (training forward): stratified jitter, ragged ray count, odd sample counts."""
    from nerf_dbr_b200.host import ops
    w = checkpoints["trained11"]["coarse_model"]
    net = packed_net(w)
    ro, rd = O.camera_rays(poses["generic"], 31, 17)
    ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
    for S in (5, 48, 64, 200):
        t = torch.rand(ro.shape[0], S, generator=torch.Generator().manual_seed(S))
        for tr in (None, t):
            ref = O.render_rays(w, ro, rd, S, t_rand=tr)
            rgb, dep, acc = ops.render_rays(net, ro.cuda(), rd.cuda(), S, mode=0,
                                            t_rand=None if tr is None else tr.cuda(), want_acc=True)
            assert (rgb.cpu() - ref[0]).abs().max() <= TOL
            assert (dep.cpu() - ref[1]).abs().max() <= TOL
            assert (acc.cpu() - ref[2]).abs().max() <= TOL


def test_renderer_interface_like_reference_integration_test(checkpoints, tmp_path):
    """The reference's own integration check (test_system.py:290-333): fake checkpoint -> setup ->
    render_image at 64x48x16 -> shapes; plus the other interface methods and the registration rule."""
    import nerf_dbr_b200 as nb
    path = str(tmp_path / "fake.pth")
    torch.save(checkpoints["semi30"], path)
    r = nb.B200Renderer("fp32")
    assert r.name == "B200 FP32" and r.device == "cuda"
    r.setup(path)
    pose = torch.eye(4)
    pose[2, 3] = 4.0
    with r.performance_monitor():
        rgb, depth = r.render_image(pose, resolution=(64, 48), samples_per_ray=16)
    assert rgb.shape == (48, 64, 3) and depth.shape == (48, 64)
    assert r.last_render_time > 0 and r.peak_memory_mb > 0 and "CUDA" in r.get_device_info()
    ref_rgb, ref_dep = O.render_image(checkpoints["semi30"]["fine_model"], pose, 64, 48, 16)
    assert (rgb.cpu() - ref_rgb).abs().max() <= TOL and (depth.cpu() - ref_dep).abs().max() <= TOL
    # the decomposed calls the reference renderers make in _render_ray_chunk (pytorch_renderers.py:156-170)
    ro, rd = r.generate_rays(pose, 64, 48)
    ro, rd = ro.reshape(-1, 3)[:512], rd.reshape(-1, 3)[:512]
    pts, z = r.sample_points_on_rays(ro, rd, 16)
    dens, col = r.query_nerf_networks(pts.reshape(-1, 3), rd[:, None, :].expand_as(pts).reshape(-1, 3), use_fine=True)
    rgb2, dep2 = r.execute_volume_rendering(dens.reshape(512, 16, 1), col.reshape(512, 16, 3), z, rd)
    assert (rgb2.cpu() - ref_rgb.reshape(-1, 3)[:512]).abs().max() <= TOL
    assert (dep2.cpu() - ref_dep.reshape(-1)[:512]).abs().max() <= TOL
    with pytest.raises(RuntimeError):
        nb.B200Renderer("fp32").render_image(pose, (8, 8), 16)       # setup() not called
