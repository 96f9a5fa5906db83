"""GPU: rays, depths, points, importance indices are BIT-EXACT with the oracle and the golden
digests; positional encoding and standalone compositing within fp32 rounding."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_json, load_npz
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


def sha(t) -> str:
    a = t.detach().cpu().contiguous().numpy() if isinstance(t, torch.Tensor) else np.ascontiguousarray(t)
    return hashlib.sha256(a.tobytes()).hexdigest()


@pytest.mark.parametrize("key", sorted(load_json("golden_geometry.json").keys()))
def test_rays_and_samples_match_golden_digests(key, poses):
    from nerf_dbr_b200.host import ops
    gold = load_json("golden_geometry.json")[key]
    pname, dims = key.split("_")
    w, h, s = (int(x) for x in dims.split("x"))
    ro, rd = ops.generate_rays(poses[pname], w, h)
    assert sha(ro) == gold["rays_o"] and sha(rd) == gold["rays_d"]
    pts, z = ops.sample_points(ro.reshape(-1, 3), rd.reshape(-1, 3), s)
    assert sha(z[0]) == gold["z_row"]
    assert torch.equal(z, z[0:1].expand_as(z))
    assert sha(pts) == gold["points"]


def test_row_shards_equal_whole_image(poses):
    from nerf_dbr_b200.host import ops
    ro, rd = ops.generate_rays(poses["generic"], 200, 150)
    for row0, n in ((0, 19), (19, 75), (94, 56)):
        so, sd = ops.generate_rays(poses["generic"], 200, 150, row0=row0, n_rows=n)
        assert torch.equal(so, ro[row0:row0 + n]) and torch.equal(sd, rd[row0:row0 + n])


@pytest.mark.parametrize("w,h", [(31, 17), (37, 23), (40, 30), (64, 48), (5, 3), (1, 1), (4, 1)])
def test_generate_rays_both_kernels_bit_exact(w, h, poses):
    """generate_rays against the oracle for image sizes whose float count is and is not a multiple of four (the float4
    kernel and the one-float-per-thread kernel), whole images and row bands that start at odd rows."""
    from nerf_dbr_b200.host import ops
    for pname in ("generic", "bench0"):
        ro_ref, rd_ref = O.camera_rays(poses[pname], w, h)
        ro, rd = ops.generate_rays(poses[pname], w, h)
        assert torch.equal(ro.cpu(), ro_ref) and torch.equal(rd.cpu(), rd_ref)
        for row0, n in ((0, h), (1, h - 1), (h // 2, h - h // 2), (h - 1, 1)):
            if n <= 0:
                continue
            so, sd = ops.generate_rays(poses[pname], w, h, row0=row0, n_rows=n)
            assert torch.equal(so.cpu(), ro_ref[row0:row0 + n]) and torch.equal(sd.cpu(), rd_ref[row0:row0 + n])


@pytest.mark.parametrize("n_samples", [1, 2, 7, 16, 33, 64, 100, 128, 192, 256])
def test_sample_points_bit_exact_any_count(n_samples, poses):
    from nerf_dbr_b200.host import ops
    ro, rd = O.camera_rays(poses["generic"], 37, 23)          # ragged: 851 rays
    ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
    pts, z = O.sample_along_rays(ro, rd, n_samples)
    g_pts, g_z = ops.sample_points(ro.cuda(), rd.cuda(), n_samples)
    assert torch.equal(g_z.cpu(), z) and torch.equal(g_pts.cpu(), pts)
    if n_samples > 1:
        t = torch.rand(ro.shape[0], n_samples, generator=torch.Generator().manual_seed(n_samples))
        pts, z = O.sample_along_rays(ro, rd, n_samples, t_rand=t)
        g_pts, g_z = ops.sample_points(ro.cuda(), rd.cuda(), n_samples, t_rand=t.cuda())
        assert torch.equal(g_z.cpu(), z) and torch.equal(g_pts.cpu(), pts)


def test_stratified_matches_golden():
    from nerf_dbr_b200.host import ops
    g = load_npz("golden_stratified.npz")
    ro, rd = ops.generate_rays(torch.from_numpy(g["pose"]), 64, 48)
    ro, rd = ro.reshape(-1, 3)[:1024], rd.reshape(-1, 3)[:1024]
    pts, z = ops.sample_points(ro, rd, 64, t_rand=torch.from_numpy(g["t_rand"]).cuda())
    assert np.array_equal(z.cpu().numpy(), g["z"]) and sha(pts) == str(g["points_sha"])


@pytest.mark.parametrize("S_", [32, 64, 128])
def test_importance_sampling_bit_exact(S_):
    from nerf_dbr_b200.host import ops
    g = load_npz("golden_importance.npz")
    t = {k: torch.from_numpy(g[f"S{S_}|{k}"]).cuda() for k in ("rays_o", "rays_d", "z", "w", "u")}
    pts, z_new, idx = ops.importance_sample(t["rays_o"], t["rays_d"], t["z"], t["w"], t["u"])
    assert np.array_equal(z_new.cpu().numpy(), g[f"S{S_}|z_new"])
    assert np.array_equal(pts.cpu().numpy(), g[f"S{S_}|points"])
    _, _, idx_ref = O.importance_sample(t["rays_o"].cpu(), t["rays_d"].cpu(), t["z"].cpu(), t["w"].cpu(), t["u"].cpu())
    assert torch.equal(idx.cpu(), idx_ref)


def test_importance_sampling_large_random():
    """Property at scale: bit-exact against the oracle on 20k rays x 128 -> 128 (config 5 shape)."""
    from nerf_dbr_b200.host import ops
    g = torch.Generator().manual_seed(3)
    R, S, N = 20000, 128, 128
    ro = torch.randn(R, 3, generator=g)
    rd = torch.randn(R, 3, generator=g)
    _, z = O.sample_along_rays(ro, rd, S)
    z = z.contiguous()
    w = torch.rand(R, S, generator=g) ** 8
    u = torch.rand(R, N, generator=g)
    pts_ref, z_ref, idx_ref = O.importance_sample(ro, rd, z, w, u)
    pts, zn, idx = ops.importance_sample(ro.cuda(), rd.cuda(), z.cuda(), w.cuda(), u.cuda())
    assert torch.equal(idx.cpu(), idx_ref) and torch.equal(zn.cpu(), z_ref) and torch.equal(pts.cpu(), pts_ref)
    with pytest.raises(Exception):
        ops.importance_sample(ro.cuda(), rd.cuda(), z[:, :100].contiguous().cuda(), w[:, :100].contiguous().cuda(), u.cuda())


@pytest.mark.parametrize("S,N", [(128, 128), (64, 100), (32, 7), (256, 256)])
def test_importance_sampling_serial_cdf_fallback(S, N):
    """The warp-per-ray kernels sum the cdf with a parallel scan when every quotient is >= 2^-28 (all partial sums are then
    exact in double, any order gives the reference's serial result) and fall back to the serial double loop otherwise.
    Rays with a 3e5 spike have quotients ~3e-11: both branches, interleaved ray by ray, bit-exact against the oracle --
    for the importance kernel and for the fused sampling kernel (sorted union)."""
    from nerf_dbr_b200.host import ops
    g = torch.Generator().manual_seed(S * 1000 + N)
    R = 3000
    ro = torch.randn(R, 3, generator=g)
    rd = torch.randn(R, 3, generator=g)
    _, z = O.sample_along_rays(ro, rd, S)
    z = z.contiguous()
    w = torch.rand(R, S, generator=g) ** 8
    w[::2, 5] = 3e5
    w[1::7] = 0.0
    u = torch.rand(R, N, generator=g)
    pts_ref, z_ref, idx_ref = O.importance_sample(ro, rd, z, w, u)
    pts, zn, idx = ops.importance_sample(ro.cuda(), rd.cuda(), z.cuda(), w.cuda(), u.cuda())
    assert torch.equal(idx.cpu(), idx_ref) and torch.equal(zn.cpu(), z_ref) and torch.equal(pts.cpu(), pts_ref)
    union = ops.hierarchical_samples(w.cuda(), N, u=u.cuda())
    assert torch.equal(union.cpu(), torch.sort(torch.cat([z, z_ref], -1), -1).values)


def test_positional_encoding():
    from nerf_dbr_b200.host import ops
    g = load_npz("golden_network.npz")
    pe = ops.positional_encoding(torch.from_numpy(g["pos"][:128]).cuda(), 10).cpu().numpy()
    de = ops.positional_encoding(torch.from_numpy(g["dirs"][:128]).cuda(), 4).cpu().numpy()
    assert pe.shape == (128, 63) and de.shape == (128, 27)
    assert np.array_equal(pe[:, :3], g["pe_pos"][:, :3])                # identity columns: exact
    assert np.abs(pe - g["pe_pos"]).max() <= 3e-7                        # sinf/cosf: <= 2 ulp
    assert np.abs(de - g["pe_dir"]).max() <= 3e-7


@pytest.mark.parametrize("n_freq", [10, 4, 6])
def test_positional_encoding_wide_range(n_freq):
    """Against the exact sin / cos (float64) of the reference's fp32 argument fl(fl(2^k pi) x): scene-sized, large and
    huge coordinates (the phase-shift path of the 10- and 4-frequency kernels hands |x| > 16384 to sincosf), ragged row
    counts around the 84-row tile, and the generic kernel (6 frequencies)."""
    from nerf_dbr_b200.host import ops
    g = torch.Generator().manual_seed(n_freq)
    x = torch.cat([torch.rand(5003, 3, generator=g) * 8 - 4, torch.randn(997, 3, generator=g) * 300,
                   torch.randn(85, 3, generator=g) * 1e6, torch.tensor([[0.0, 1.0, -1.0], [0.5, 2.0, 4.0], [1e-8, -1e-8, 16384.0]])])
    got = ops.positional_encoding(x.cuda(), n_freq).cpu()
    assert got.shape == (x.shape[0], 3 + 6 * n_freq) and torch.equal(got[:, :3], x)
    for k in range(n_freq):
        arg = (torch.tensor(2.0 ** k) * torch.pi * x).double()              # fp32 product, as nerf.py:42-43
        assert (got[:, 3 + 6 * k: 6 + 6 * k].double() - torch.sin(arg)).abs().max() <= 2.5e-7
        assert (got[:, 6 + 6 * k: 9 + 6 * k].double() - torch.cos(arg)).abs().max() <= 2.5e-7
    for n in (1, 83, 84, 85, 169):
        assert torch.equal(ops.positional_encoding(x[:n].cuda(), n_freq).cpu(), got[:n])


def test_composite_matches_golden():
    from nerf_dbr_b200.host import ops
    g = load_npz("golden_composite.npz")
    t = {k: torch.from_numpy(g[k]).cuda() for k in ("sigma", "rgb", "z", "rays_d")}
    rgb, dep, acc, w = ops.composite(t["sigma"], t["rgb"], t["z"], t["rays_d"], want_aux=True)
    assert np.abs(rgb.cpu().numpy() - g["rgb_map"]).max() <= 2e-6
    assert np.abs(dep.cpu().numpy() - g["depth"]).max() <= 1e-5
    assert np.abs(acc.cpu().numpy() - g["acc"]).max() <= 2e-6
    assert np.abs(w.cpu().numpy() - g["weights"]).max() <= 3e-7
    rgb2, dep2 = ops.composite(t["sigma"], t["rgb"], t["z"], t["rays_d"])
    assert torch.equal(rgb2, rgb) and torch.equal(dep2, dep)


def test_composite_edge_cases():
    """One sample per ray, ragged S, all-zero density (reference: no background term -> black)."""
    from nerf_dbr_b200.host import ops
    g = torch.Generator().manual_seed(5)
    for S in (1, 3, 31, 45, 200):
        R = 77
        sig = torch.rand(R, S, 1, generator=g) * 20 - 2
        col = torch.rand(R, S, 3, generator=g)
        rd = torch.randn(R, 3, generator=g)
        _, z = O.sample_along_rays(torch.zeros(R, 3), rd, S)
        ref = O.composite(sig, col, z.contiguous(), rd)
        rgb, dep, acc, w = ops.composite(sig[..., 0].cuda(), col.cuda(), z.contiguous().cuda(), rd.cuda(), want_aux=True)
        assert (rgb.cpu() - ref[0]).abs().max() <= 3e-6 and (dep.cpu() - ref[1]).abs().max() <= 2e-5
        if S > 1:                      # S == 1: the reference weights tensor is [R,0]
            assert (w.cpu() - ref[3]).abs().max() <= 3e-7
    z0 = torch.zeros(5, 8).cuda()
    rgb, dep = ops.composite(z0, torch.rand(5, 8, 3).cuda(), z0 + 3, torch.ones(5, 3).cuda())
    assert float(rgb.abs().max()) == 0.0 and float(dep.abs().max()) == 0.0
