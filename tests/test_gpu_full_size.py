"""GPU: BASELINE.json's full size (configs[2]: 800x600, 128 samples per ray), where the CPU oracle needs minutes per
view -- parity through size-independent properties instead:
  * decomposition: any split into row bands (the multi-GPU sharding) gives the bits of the single launch;
  * determinism: a second launch gives the same bits;
  * cross-mode agreement: BF16X3 (max-abs gate of the fp32 mode) and BF16 (PSNR gate) against the FP32 CUDA-core
    mode, which small-size tests pin to the reference within 1.6e-6;
  * composition: the staged reference pipeline (generate_rays -> sample_points -> query_network -> composite) on a
    band equals the fused kernel's band."""
import numpy as np
import pytest
import torch

from gpu_util import Watchdog, packed_net, psnr

pytestmark = pytest.mark.gpu
W, H, S = 800, 600, 128


def test_full_size_properties(checkpoints, poses):
    from nerf_dbr_b200.host import lib as L, ops
    from nerf_dbr_b200.host.parallel import row_band
    from nerf_dbr_b200.host.synthetic import orbit_pose
    net = packed_net(checkpoints["lego"]["fine_model"])
    pose = orbit_pose(7, 40)
    with Watchdog() as wd:
        rgb, dep = ops.render_image(net, pose, W, H, S, mode=L.BF16)
        # decomposition into the 8-GPU bands and into ragged bands; determinism
        for world in (8, 7):
            for rank in range(world):
                row0, n = row_band(rank, world, H)
                r2, d2 = ops.render_image(net, pose, W, H, S, mode=L.BF16, row0=row0, n_rows=n)
                assert torch.equal(r2, rgb[row0:row0 + n]) and torch.equal(d2, dep[row0:row0 + n]), (world, rank)
        r3, d3 = ops.render_image(net, pose, W, H, S, mode=L.BF16)
        assert torch.equal(r3, rgb) and torch.equal(d3, dep)
        torch.cuda.synchronize()
        assert int(wd.word.item()) == 0
        # cross-mode agreement
        r32, d32 = ops.render_image(net, pose, W, H, S, mode=L.FP32)
        rx3, dx3 = ops.render_image(net, pose, W, H, S, mode=L.BF16X3)
        torch.cuda.synchronize()
        assert int(wd.word.item()) == 0
    assert torch.isfinite(rgb).all() and torch.isfinite(dep).all()
    e3 = (rx3 - r32).abs().max().item()
    p16 = psnr(rgb.cpu().numpy(), r32.cpu().numpy())
    print(f"800x600x128: BF16X3 vs FP32 max-abs {e3:.2e}; BF16 vs FP32 PSNR {p16:.1f} dB, max-abs {(rgb - r32).abs().max().item():.2e}")
    assert e3 <= 1e-4
    assert p16 >= 50.0
    # composition: the staged pipeline on one band (BF16 query_network) against the fused band
    row0, n = 300, 16
    ro, rd = ops.generate_rays(pose, W, H, row0=row0, n_rows=n)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    pts, z = ops.sample_points(ro, rd, S)
    sigma, col = ops.query_network(net, pts.reshape(-1, 3), rd[:, None, :].expand(-1, S, -1).reshape(-1, 3).contiguous(), mode=L.BF16)
    rgb_s, dep_s = ops.composite(sigma.reshape(-1, S, 1), col.reshape(-1, S, 3), z, rd)[:2]
    assert (rgb_s.reshape(n, W, 3) - rgb[row0:row0 + n]).abs().max().item() <= 2e-5
    assert (dep_s.reshape(n, W) - dep[row0:row0 + n]).abs().max().item() <= 2e-4
