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


def test_full_size_training_step_properties():
    """BASELINE.json configs[3]: 4096 rays, 64 coarse + 128 fine samples.  Properties that need no oracle:
      * additivity (the data-parallel contract): the gradients and loss of the batch equal the accumulated gradients
        and summed losses of its shards computed separately with the global ray count -- here 2 and 3 ragged shards;
      * cross-mode agreement: the tensor-core mode against the FP32 CUDA-core mode, which small-size tests pin to the
        reference's autograd (loss within 0.5 %, every gradient tensor within 15 %: the mode's bf16 precision)."""
    from nerf_dbr_b200.host import lib as L
    from nerf_dbr_b200.host.synthetic import seeded_models
    from nerf_dbr_b200.host.trainer import B200TrainStep
    n = 4096
    g = torch.Generator().manual_seed(11)
    ro = (torch.zeros(n, 3) + torch.tensor([0.0, 0.0, 4.0])).cuda()
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.25 + torch.tensor([0.0, 0.0, -1.0]), dim=-1).cuda()
    tgt, tr = torch.rand(n, 3, generator=g).cuda(), torch.rand(n, 64, generator=g).cuda()

    def run(mode, shards):
        coarse, fine = seeded_models(5, 30.0, "cuda")
        step = B200TrainStep(coarse, fine, 64, 128, mode=mode)
        total, acc, first = 0.0, None, 0
        for count in shards:
            sl = slice(first, first + count)
            loss, _, _ = step(ro[sl], rd[sl], tgt[sl], t_rand=tr[sl].contiguous(), n_rays_global=n)   # zeroes, then accumulates
            grads = [p.grad.double().clone() for p in step.parameters()]
            acc = grads if acc is None else [a + b for a, b in zip(acc, grads)]
            total += float(loss)
            first += count
        return total, acc, [k for m in (coarse, fine) for k, _ in m.named_parameters()]

    with Watchdog() as wd:
        l1, g1, names = run(L.BF16, [n])
        l2, g2, _ = run(L.BF16, [2048, 2048])
        l3, g3, _ = run(L.BF16, [1000, 2000, 1096])
        l32, g32, _ = run(L.FP32, [n])
        torch.cuda.synchronize()
        assert int(wd.word.item()) == 0
    for lx, gx in ((l2, g2), (l3, g3)):
        assert abs(lx - l1) <= 1e-5 * l1
        worst = max(float((a - b).norm()) / max(float(a.norm()), 1e-30) for a, b in zip(g1, gx))
        assert worst <= 1e-3, worst
    rel = sorted(((float((a - b).norm()) / max(float(b.norm()), 1e-30), k) for a, b, k in zip(g1, g32, names)), reverse=True)
    print(f"4096 rays: loss bf16 {l1:.6f} fp32 {l32:.6f}; worst gradient tensors bf16 vs fp32:", [(f"{e:.1e}", k) for e, k in rel[:3]])
    assert abs(l1 - l32) <= 5e-3 * l32
    assert rel[0][0] <= 0.15, rel[0]


def test_full_size_sampling_kernels_properties():
    """configs[4]'s ray count (1600 x 1200 = 1.92 M rays, 128 coarse + 128 importance samples), where the CPU oracle needs
    minutes: the fused sampling kernel equals the three-stage path (sample_points -> importance_sample -> merge_samples, each
    pinned to the reference at small sizes) bit for bit; the union is sorted and holds every coarse depth; the staged
    compositing of the 800x600x128 shape is deterministic and additive over ray shards; the in-kernel uniforms are
    deterministic per seed."""
    from nerf_dbr_b200.host import ops
    dev = torch.device("cuda")
    R, Sc, Nn = 1600 * 1200, 128, 128
    g = torch.Generator(device=dev).manual_seed(11)
    w = torch.rand(R, Sc, device=dev, generator=g) ** 6
    w[::97] = 0.0                                              # empty rays: uniform pdf from the 1e-5 floor
    w[5::1013, 17] = 4e5                                       # rays that take the serial cdf path (quotients < 2^-28)
    u = torch.rand(R, Nn, device=dev, generator=g)
    ro = torch.zeros(R, 3, device=dev)
    rd = torch.randn(R, 3, device=dev, generator=g)
    _, z = ops.sample_points(ro, rd, Sc)
    assert torch.equal(z, z[:1].expand_as(z))
    _, z_new, idx = ops.importance_sample(ro, rd, z, w, u)
    assert int(idx.min()) >= 1 and int(idx.max()) <= Sc + 1
    ref = ops.merge_samples(z, z_new)
    del idx
    out = ops.hierarchical_samples(w, Nn, u=u)
    assert torch.equal(out, ref)
    assert bool((out[:, 1:] >= out[:, :-1]).all())
    # every coarse depth is in the union: removing the new samples' multiset leaves z (checked through sums of exact membership)
    pos = torch.searchsorted(out, z)                           # first position of each coarse depth in its union row
    assert bool((torch.gather(out, 1, pos.clamp(max=Sc + Nn - 1)) == z).all())
    del ref, pos, z_new
    a, b, c = (ops.hierarchical_samples(w, Nn, seed=s) for s in (3, 3, 4))
    assert torch.equal(a, b) and not torch.equal(a, c) and bool((a[:, 1:] >= a[:, :-1]).all())
    del a, b, c, out, u
    # staged compositing at the headline shape: deterministic, and a shard of the rays gives the bits of the whole
    R2 = 800 * 600
    sigma = torch.rand(R2, S, device=dev, generator=g) * 8 - 1
    col = torch.rand(R2, S, 3, device=dev, generator=g)
    full = ops.composite(sigma, col, z[:R2], rd[:R2], want_aux=True)
    again = ops.composite(sigma, col, z[:R2], rd[:R2], want_aux=True)
    assert all(torch.equal(x, y) for x, y in zip(full, again))
    lo, hi = 123457, 345679
    part = ops.composite(sigma[lo:hi], col[lo:hi], z[lo:hi], rd[lo:hi], want_aux=True)
    assert all(torch.equal(x[lo:hi], y) for x, y in zip(full, part))
    assert bool(((full[2] >= 0) & (full[2] <= 1 + 1e-5)).all())   # accumulated opacity in [0, 1]
