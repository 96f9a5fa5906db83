"""GPU: BASELINE.json's own sizes pinned DIRECTLY to the reference -- pixels written by the unmodified
``PyTorchCPURenderer`` (tests/golden/make_golden_configs.py):

* configs[1]: the full 400x300 frame, 64 samples per ray (``lego`` fixture, view 5 of the 40-view orbit);
* configs[2]: rows 0-7, 296-303 and 592-599 of the 800x600 frame, 128 samples per ray (``lego`` and ``semi30``,
  view 7 of the orbit), rendered here as row bands of the full-size image (``row0`` / ``n_rows`` of
  ``nerf_b200_render_image``: the multi-GPU shard entry).

Gates (north_star): FP32 and BF16X3 modes max-abs <= 1e-4 on rgb and depth; BF16 mode <= 0.05 dB PSNR difference
against a target at 32 dB.  The BF16 numbers are also REPORTED (not gated) with the target at 40 dB, together with the
raw PSNR(bf16, reference) -- the target level is the one free knob of that gate, so both ends are on record
(gpurun_out/bf16_gate.json, copied to profiles/)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import REPO, load_npz
from gpu_util import Watchdog, packed_net, psnr

pytestmark = pytest.mark.gpu
TOL = 1e-4
BANDS = ((0, 8), (296, 8), (592, 8))


def _poses(g):
    return torch.from_numpy(g["pose_c2"]), torch.from_numpy(g["pose_c3"])


def _cases(g, checkpoints):
    """(tag, network, pose, W, H, S, row0, n_rows, ref_rgb, ref_depth)"""
    p2, p3 = _poses(g)
    k = "lego|view5of40|400x300x64"
    yield k, "lego", p2, 400, 300, 64, 0, 300, g[k + "|rgb"], g[k + "|depth"]
    for cname in ("lego", "semi30"):
        for row0, n in BANDS:
            k = f"{cname}|view7of40|800x600x128|rows{row0}+{n}"
            yield k, cname, p3, 800, 600, 128, row0, n, g[k + "|rgb"], g[k + "|depth"]


@pytest.mark.parametrize("mode,name", [(0, "fp32"), (2, "bf16x3")])
def test_configs_max_abs_gate(mode, name, checkpoints):
    from nerf_dbr_b200.host import ops
    g = load_npz("golden_configs.npz")
    nets = {c: packed_net(checkpoints[c]["fine_model"]) for c in ("lego", "semi30")}
    worst = 0.0
    with Watchdog() as wd:
        for tag, cname, pose, w, h, s, row0, n, ref_rgb, ref_dep in _cases(g, checkpoints):
            rgb, dep = ops.render_image(nets[cname], pose, w, h, s, mode=mode, row0=row0, n_rows=n)
            torch.cuda.synchronize()
            assert int(wd.word.item()) == 0, hex(int(wd.word.item()) & 0xffffffff)
            e_rgb = np.abs(rgb.cpu().numpy() - ref_rgb).max()
            e_dep = np.abs(dep.cpu().numpy() - ref_dep).max()
            print(f"{name} {tag}: max|rgb| {e_rgb:.2e} max|depth| {e_dep:.2e}")
            worst = max(worst, e_rgb, e_dep)
            assert e_rgb <= TOL and e_dep <= TOL, (tag, e_rgb, e_dep)
    print(f"{name}: worst max-abs vs the reference at configs[1]/[2] sizes: {worst:.2e}")


def test_configs_bf16_psnr_gate(checkpoints):
    from nerf_dbr_b200.host import ops
    g = load_npz("golden_configs.npz")
    nets = {c: packed_net(checkpoints[c]["fine_model"]) for c in ("lego", "semi30")}
    report = []
    with Watchdog() as wd:
        for tag, cname, pose, w, h, s, row0, n, ref_rgb, ref_dep in _cases(g, checkpoints):
            rgb, dep = ops.render_image(nets[cname], pose, w, h, s, mode=1, row0=row0, n_rows=n)
            torch.cuda.synchronize()
            assert int(wd.word.item()) == 0
            rgb = rgb.cpu().numpy()
            row = {"case": tag, "psnr_bf16_vs_reference_db": psnr(rgb, ref_rgb), "max_abs_rgb": float(np.abs(rgb - ref_rgb).max()),
                   "max_abs_depth": float(np.abs(dep.cpu().numpy() - ref_dep).max())}
            for level_db, sigma in ((32, 0.0251), (40, 0.01)):
                target = ref_rgb.astype(np.float64) + np.random.default_rng(0).normal(0.0, sigma, ref_rgb.shape)
                row[f"dpsnr_target_{level_db}db"] = abs(psnr(rgb, target) - psnr(ref_rgb, target))
            print(row)
            report.append(row)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "bf16_gate.json"), "w") as fh:
        json.dump({"gate": "north_star: <= 0.05 dB PSNR difference (target at 32 dB gated; 40 dB reported)",
                   "min_psnr_bf16_vs_reference_db": min(r["psnr_bf16_vs_reference_db"] for r in report),
                   "max_dpsnr_target_32db": max(r["dpsnr_target_32db"] for r in report),
                   "max_dpsnr_target_40db": max(r["dpsnr_target_40db"] for r in report), "cases": report}, fh, indent=1)
    for row in report:
        assert row["dpsnr_target_32db"] <= 0.05, row
        assert row["psnr_bf16_vs_reference_db"] >= 50.0, row


def test_trained11_bf16_kernel_against_its_numerical_model(checkpoints, poses):
    """``trained11`` (i.i.d. Gaussian weights at trained magnitudes) amplifies ANY rounding of its activations, so the
    reference-level gates of the other fixtures say nothing about a kernel on it.  What can be pinned: the kernel
    computes what BF16 mode DEFINES -- tests/diag/emulate_bf16.py is that definition on the CPU (bf16-rounded operands,
    fp32 accumulation, fp32 direction bias, phase-shift encoding) -- tightly, and that definition sits where bf16
    arithmetic puts it against the reference, loosely.  Differences kernel-vs-model are accumulation order and
    sin/cos ulps (fp32-level), which this fixture also amplifies: hence PSNR, not max-abs."""
    from diag.emulate_bf16 import render_image as model_render
    from nerf_dbr_b200.host import ops
    g = load_npz("golden_render.npz")
    w = checkpoints["trained11"]["fine_model"]
    net = packed_net(w)
    with Watchdog() as wd, torch.no_grad():
        for key in ("trained11|generic|96x64x64", "trained11|bench0|64x48x16"):
            _, pname, dims = key.split("|")
            wd_, ht, s = (int(x) for x in dims.split("x"))
            rgb, dep = ops.render_image(net, poses[pname], wd_, ht, s, mode=1)
            torch.cuda.synchronize()
            assert int(wd.word.item()) == 0
            m_rgb, m_dep = model_render(w, poses[pname], wd_, ht, s)
            p_kernel_model = psnr(rgb.cpu().numpy(), m_rgb.numpy())
            p_model_ref = psnr(m_rgb.numpy(), g[key + "|rgb"])
            p_kernel_ref = psnr(rgb.cpu().numpy(), g[key + "|rgb"])
            print(f"{key}: PSNR kernel vs bf16 model {p_kernel_model:.1f} dB; model vs reference {p_model_ref:.1f} dB; "
                  f"kernel vs reference {p_kernel_ref:.1f} dB")
            assert p_kernel_model >= p_model_ref + 6.0, (key, p_kernel_model, p_model_ref)   # tight: well inside the mode's own error
            assert p_model_ref >= 30.0 and p_kernel_ref >= 30.0, (key, p_model_ref, p_kernel_ref)   # loose
