"""GPU: FP8 mode -- the fused render kernel with e4m3 weights and activations (tcgen05.mma.kind::f8f6f4), SURVEY 8 f4.

A lossy mode, so the gates are its own:
  * the kernel computes what the mode DEFINES: tests/diag/emulate_fp8.py is that definition on the CPU (same scales, e4m3
    rounding with saturation, bf16 encoded-position inputs, fp32 accumulation) -- kernel vs model tight;
  * the scales the pack kernels derive (calibration with the fp32 CUDA kernel, per-row weight maxima) equal the model's;
  * against the reference it is judged like the reference's own lossy renderer: CompressedNeRFRenderer
    (src/benchmark/compressed_renderer.py, default config of the suite: 8-bit weights, 10 % pruning) run on the CPU on the
    same view -- FP8 mode must be in its class (PSNR against PyTorchCPURenderer within 6 dB of the compressed renderer's);
  * structural properties as in the other modes: every tile shape, row bands bit-identical, deterministic."""
import io
import contextlib

import numpy as np
import pytest
import torch

from conftest import load_npz
from gpu_util import Watchdog, packed_net, psnr
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
FP8 = 3
Q_MUL = 531208                       # fp8_layout.h (floats): F_WO; Q_SA = Q_MUL + 2048, Q_SW = Q_SA + 16
Q_SA, Q_SW = Q_MUL + 2048, Q_MUL + 2048 + 16


def fp8_net(state_dict):
    from nerf_dbr_b200.host import ops
    return ops.pack_weights_fp8({k: v.cuda() for k, v in state_dict.items()}, torch.device("cuda", 0))


def test_scales_equal_the_numerical_model(checkpoints):
    from diag.emulate_fp8 import activation_maxima, calibration_points, scales
    for cname in ("lego", "semi30"):
        w = checkpoints[cname]["fine_model"]
        q = fp8_net(w).view(torch.float32).cpu()
        with torch.no_grad():
            amax = activation_maxima(w, calibration_points()[0])
            sa, sw = scales(w, amax)
        got_amax = q[Q_SA + 8:Q_SA + 16]
        assert torch.allclose(got_amax, torch.tensor(amax), rtol=1e-4), (got_amax, amax)
        assert q[Q_SA:Q_SA + 8].tolist() == sa
        for l in range(1, 8):
            assert torch.equal(q[Q_SW + 256 * l:Q_SW + 256 * l + 256], sw[l]), l
        assert torch.equal(q[Q_SW + 256 * 8:Q_SW + 256 * 8 + 129], sw[8])


def test_kernel_computes_what_the_mode_defines(checkpoints, poses):
    from diag.emulate_fp8 import render_image as model_render
    from nerf_dbr_b200.host import ops
    g = load_npz("golden_render.npz")
    with Watchdog() as wd, torch.no_grad():
        for cname, key in (("lego", "lego|generic|96x64x64"), ("lego", "lego|bench1|64x48x16"), ("semi30", "semi30|generic|64x48x16")):
            w = checkpoints[cname]["fine_model"]
            net = fp8_net(w)
            _, pname, dims = key.split("|")
            wd_, ht, s = (int(x) for x in dims.split("x"))
            rgb, dep = ops.render_image(net, poses[pname], wd_, ht, s, mode=FP8)
            torch.cuda.synchronize()
            assert int(wd.word.item()) == 0, hex(int(wd.word.item()) & 0xffffffff)
            assert torch.isfinite(rgb).all() and torch.isfinite(dep).all()
            m_rgb, m_dep = model_render(w, poses[pname], wd_, ht, s)
            p_km, p_mr, p_kr = psnr(rgb.cpu().numpy(), m_rgb.numpy()), psnr(m_rgb.numpy(), g[key + "|rgb"]), psnr(rgb.cpu().numpy(), g[key + "|rgb"])
            print(f"{key}: PSNR kernel vs fp8 model {p_km:.1f} dB; model vs reference {p_mr:.1f} dB; kernel vs reference {p_kr:.1f} dB")
            assert p_km >= min(p_mr + 6.0, 60.0), (key, p_km, p_mr)        # tight: well inside the mode's own error
            assert p_kr >= 28.0, (key, p_kr)


def test_fp8_is_in_the_class_of_the_reference_compressed_renderer(checkpoints, poses, tmp_path):
    from oracle import refload
    from nerf_dbr_b200.host import ops
    if refload.reference_root() is None:
        pytest.skip("reference sources not vendored on this box (tools/vendor_reference.sh)")
    refload.import_reference()
    from src.benchmark.compressed_renderer import CompressedNeRFRenderer
    g = load_npz("golden_render.npz")
    ck = str(tmp_path / "lego.pth")
    torch.save(checkpoints["lego"], ck)
    with contextlib.redirect_stdout(io.StringIO()):
        comp = CompressedNeRFRenderer({"quantization_bits": 8, "pruning_ratio": 0.1, "use_mixed_precision": True,
                                       "compress_activations": True})             # the suite's default (benchmark_suite.py:69-75)
        comp.setup(ck)
    net = fp8_net(checkpoints["lego"]["fine_model"])
    for pname in ("bench1", "generic"):
        ref = g[f"lego|{pname}|64x48x16|rgb"]
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            c_rgb, _ = comp.render_image(poses[pname], (64, 48), 16)
        rgb, _ = ops.render_image(net, poses[pname], 64, 48, 16, mode=FP8)
        p_comp, p_fp8 = psnr(c_rgb.float().numpy(), ref), psnr(rgb.cpu().numpy(), ref)
        print(f"{pname}: PSNR against PyTorchCPURenderer -- reference CompressedNeRFRenderer {p_comp:.1f} dB, B200 FP8 mode {p_fp8:.1f} dB")
        assert p_fp8 >= p_comp - 6.0, (pname, p_fp8, p_comp)


@pytest.mark.parametrize("S", [16, 24, 64, 100, 128, 256])
def test_render_rays_fp8_tile_shapes(S, checkpoints, poses):
    """8/4/2/1 rays per tile, padded sample counts, multi-tile rays, ragged ray count, jitter, acc and weights out --
    against the BF16 mode (PSNR: the mode's own error level), finite, deterministic."""
    from nerf_dbr_b200.host import ops
    w = checkpoints["lego"]["fine_model"]
    net8, net16 = fp8_net(w), packed_net(w)
    ro, rd = O.camera_rays(poses["generic"], 41, 27)
    ro, rd = ro.reshape(-1, 3).contiguous().cuda(), rd.reshape(-1, 3).contiguous().cuda()
    tr = torch.rand(ro.shape[0], S, generator=torch.Generator().manual_seed(S)).cuda()
    with Watchdog() as wd:
        for t in (None, tr):
            a = ops.render_rays(net8, ro, rd, S, mode=FP8, t_rand=t, want_acc=True, want_weights=True)
            b = ops.render_rays(net16, ro, rd, S, mode=1, t_rand=t, want_acc=True, want_weights=True)
            a2 = ops.render_rays(net8, ro, rd, S, mode=FP8, t_rand=t, want_acc=True, want_weights=True)
            torch.cuda.synchronize()
            assert int(wd.word.item()) == 0
            assert all(torch.isfinite(x).all() for x in a)
            assert all(torch.equal(x, y) for x, y in zip(a, a2))
            p = psnr(a[0].cpu().numpy(), b[0].cpu().numpy())
            assert p >= 27.0, (S, p)
            assert (a[2] - b[2]).abs().max().item() <= 0.5 and (a[3].sum(-1) - a[2]).abs().max().item() <= 1e-4


def test_fp8_row_bands_and_renderer(checkpoints, poses, tmp_path):
    import nerf_dbr_b200 as nb
    from nerf_dbr_b200.host import ops
    net = fp8_net(checkpoints["lego"]["fine_model"])
    with Watchdog() as wd:
        rgb, dep = ops.render_image(net, poses["bench1"], 200, 150, 32, mode=FP8)
        for row0, n in ((0, 19), (19, 75), (94, 56)):
            r2, d2 = ops.render_image(net, poses["bench1"], 200, 150, 32, mode=FP8, row0=row0, n_rows=n)
            assert torch.equal(r2, rgb[row0:row0 + n]) and torch.equal(d2, dep[row0:row0 + n])
        first = None
        for i in range(24):                                  # soak: many tiles and launches back to back
            r3, _ = ops.render_image(net, poses["generic"], 400, 300, 64, mode=FP8)
            first = r3.clone() if first is None else first
        torch.cuda.synchronize()
        assert int(wd.word.item()) == 0 and torch.equal(r3, first)
    path = str(tmp_path / "ck.pth")
    torch.save(checkpoints["lego"], path)
    r = nb.B200Renderer("fp8")
    assert r.name == "B200 FP8"
    r.setup(path)
    with r.performance_monitor():
        img, depth = r.render_image(poses["bench1"], (64, 48), 16)
    g = load_npz("golden_render.npz")
    assert img.shape == (48, 64, 3) and psnr(img.cpu().numpy(), g["lego|bench1|64x48x16|rgb"]) >= 28.0
