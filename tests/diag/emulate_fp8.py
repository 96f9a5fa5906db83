"""Numerical model of FP8 mode (csrc/mlp_tc_fp8.inl, csrc/pack_fp8.cu, csrc/fp8_layout.h) in torch on the CPU: the same scales
(calibrated on the same points), e4m3 weights and activations with saturation, bf16 encoded-position inputs and layer 0 /
skip weights, fp32 accumulation, fp32 heads and compositing.  Predicts the mode's error against the oracle before GPU time
is spent, and pins the kernel in tests/test_gpu_fp8.py (kernel vs model tight, model vs reference loose).  Not product code."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import nerf_oracle as O
from diag.emulate_bf16 import bf, encode_phase


def e4m3(x):
    """cvt.rn.satfinite.e4m3: saturate to +-448, round to nearest even"""
    return x.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).to(torch.float32)


def pow2_floor_scale(target, maxabs):
    maxabs = torch.as_tensor(maxabs, dtype=torch.float32)
    s = torch.exp2(torch.floor(torch.log2(target / maxabs)))
    return torch.where(maxabs > 0, s, torch.ones_like(s))


def calibration_points(n_views=4, width=40, height=30, n_samples=32):
    pos, dirs = [], []
    for i in range(n_views):
        ro, rd = O.camera_rays(O.benchmark_pose(i, n_views), width, height)
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        pts, _ = O.sample_along_rays(ro, rd, n_samples)
        pos.append(pts.reshape(-1, 3))
        dirs.append(rd[:, None, :].expand(-1, n_samples, -1).reshape(-1, 3))
    return torch.cat(pos), torch.cat(dirs)


def activation_maxima(w, pos):
    """max of every trunk layer's fp32 activations over the calibration points (what simt_calibrate records)"""
    pe = O.encode(pos, 10)
    h, out = pe, []
    for i in range(8):
        x = torch.cat([h, pe], -1) if i == 4 else h
        h = torch.relu(x @ w[f"layers.{i}.weight"].T + w[f"layers.{i}.bias"])
        out.append(float(h.max()))
    return out


def scales(w, amax):
    sa = [float(pow2_floor_scale(240.0, a)) for a in amax]
    sw = {l: pow2_floor_scale(448.0, w[f"layers.{l}.weight"][:, :256].abs().amax(dim=1)) for l in range(1, 8)}
    sw[8] = pow2_floor_scale(448.0, torch.cat([w["color_layers.0.weight"][:, :256].abs().amax(dim=1), w["density_head.weight"].abs().amax(dim=1)]))
    return sa, sw


def mlp_fp8(w, sa, sw, pts, dirs):
    pe = bf(encode_phase(pts))
    acc = pe @ bf(w["layers.0.weight"]).T
    hq = e4m3(torch.relu(acc * sa[0] + w["layers.0.bias"] * sa[0]))
    for l in range(1, 8):
        W = w[f"layers.{l}.weight"]
        acc = hq @ e4m3(W[:, :256] * sw[l][:, None]).T
        if l == 4:
            acc = acc + pe[:, :63] @ bf(W[:, 256:] * (sa[3] * sw[4])[:, None]).T
        m = sa[l] / (sa[l - 1] * sw[l])
        hq = e4m3(torch.relu(acc * m + w[f"layers.{l}.bias"] * sa[l]))
    Wc = w["color_layers.0.weight"]
    s = sa[7] * sw[8][:128]
    acc = hq @ e4m3(Wc[:, :256] * sw[8][:128, None]).T
    sig_acc = hq @ e4m3(w["density_head.weight"] * sw[8][128]).T
    sigma = torch.relu(sig_acc * (1.0 / (sa[7] * sw[8][128])) + w["density_head.bias"])
    de = O.encode(dirs, 4)
    rayb = de @ (Wc[:, 256:] * s[:, None]).T + w["color_layers.0.bias"] * s
    c = torch.relu(acc + rayb)
    y = c @ (w["color_layers.1.weight"] / s[None, :]).T + w["color_layers.1.bias"]
    return sigma, torch.sigmoid(y)


def render_image(w, pose, W_, H_, S, sa=None, sw=None):
    if sa is None:
        sa, sw = scales(w, activation_maxima(w, calibration_points()[0]))
    ro, rd = O.camera_rays(pose, W_, H_)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    pts, z = O.sample_along_rays(ro, rd, S)
    d = rd[:, None, :].expand_as(pts).reshape(-1, 3)
    sg, col = mlp_fp8(w, sa, sw, pts.reshape(-1, 3), d)
    rgb, dep, _, _ = O.composite(sg.reshape(-1, S, 1), col.reshape(-1, S, 3), z, rd)
    return rgb.reshape(H_, W_, 3), dep.reshape(H_, W_)


if __name__ == "__main__":
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "golden", "golden_render.npz"))
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "golden", "ckpt_lego_stuffed_fp16.npz"))
    cks = {"rand2": O.seeded_checkpoint(2), "semi30": O.seeded_checkpoint(2, 30.0),
           "lego": {"fine_model": {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}}}
    poses = {"bench0": O.benchmark_pose(0, 3), "bench1": O.benchmark_pose(1, 3), "generic": O.generic_pose()}
    ps = lambda a, b: -10 * np.log10(np.mean((np.asarray(a, np.float64) - b) ** 2))
    with torch.no_grad():
        for cname in ("lego", "semi30", "rand2"):
            w = cks[cname]["fine_model"]
            amax = activation_maxima(w, calibration_points()[0])
            sa, sw = scales(w, amax)
            print(cname, "activation maxima", ["%.2f" % a for a in amax], "sa", sa)
            for k in sorted({k.rsplit("|", 1)[0] for k in g.files if k.startswith(cname)}):
                _, pname, dims = k.split("|")
                w_, h_, s_ = (int(v) for v in dims.split("x"))
                rgb, dep = render_image(w, poses[pname], w_, h_, s_, sa, sw)
                ref, refd = g[k + "|rgb"], g[k + "|depth"]
                print(f"  {k:34s} PSNR(fp8 model, ref)={ps(rgb.numpy(), ref.astype(np.float64)):6.1f} dB  max|rgb|={np.abs(rgb.numpy() - ref).max():.2e}"
                      f"  max|depth|={np.abs(dep.numpy() - refd).max():.2e}")
