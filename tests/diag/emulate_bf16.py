"""Numerical model of the BF16 tensor-core kernel (mlp_tc.cu) in torch on the CPU: bf16-rounded
operands, fp32 accumulation, fp32 per-ray direction bias, phase-shift positional encoding.
Used to predict the kernel's error against the oracle before spending GPU time, and to localise
a discrepancy (layout bug vs. numerics) when a GPU parity test fails.  Not part of the product."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import math
import numpy as np
import torch
from oracle import nerf_oracle as O


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def encode_phase(p):
    """the kernel's encoding: a = fl(pi_f * p); phase = round(a/(2 pi) * 2^32) mod 2^32; shifts per octave"""
    pi_f = torch.tensor(math.pi, dtype=torch.float32)
    a = (pi_f * p).to(torch.float64)
    ph = torch.round(a * 0.15915494309189535 * 4294967296.0).to(torch.int64) & 0xffffffff
    feats = [p]
    for k in range(10):
        phk = (ph << k) & 0xffffffff
        phk = torch.where(phk >= 2 ** 31, phk - 2 ** 32, phk)
        r = (phk.to(torch.float32) * np.float32(1.4629180792671596e-9))
        feats += [torch.sin(r), torch.cos(r)]
    return torch.cat(feats, -1)


def mlp_bf16(w, pts, rays_d_per_sample):
    pe = bf(encode_phase(pts))
    h = pe
    for i in range(8):
        W = bf(w[f"layers.{i}.weight"])
        if i == 4:
            acc = h @ W[:, :256].T + pe @ W[:, 256:].T
        else:
            acc = h @ W.T
        x = torch.relu(acc + w[f"layers.{i}.bias"])
        h = bf(x)
    # the density head is column 128 of colour layer 0's GEMM: bf16 h7 x bf16 w_sigma, fp32 accumulate
    sigma = torch.relu(h @ bf(w["density_head.weight"]).T + w["density_head.bias"])
    de = O.encode(rays_d_per_sample, 4)
    Wc = w["color_layers.0.weight"]
    bias = de @ Wc[:, 256:].T + w["color_layers.0.bias"]
    c = torch.relu(h @ bf(Wc[:, :256]).T + bias)
    y = c @ w["color_layers.1.weight"].T + w["color_layers.1.bias"]
    return sigma, torch.sigmoid(y)


def render_image(w, pose, W_, H_, S):
    ro, rd = O.camera_rays(pose, W_, H_)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    pts, z = O.sample_along_rays(ro, rd, S)
    d = rd[:, None, :].expand_as(pts).reshape(-1, 3)
    sg, col = mlp_bf16(w, pts.reshape(-1, 3), d)
    rgb, dep, _, _ = O.composite(sg.reshape(-1, S, 1), col.reshape(-1, S, 3), z, rd)
    return rgb.reshape(H_, W_, 3), dep.reshape(H_, W_)


if __name__ == "__main__":
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "golden", "golden_render.npz"))
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "golden", "ckpt_lego_stuffed_fp16.npz"))
    cks = {"rand2": O.seeded_checkpoint(2), "semi30": O.seeded_checkpoint(2, 30.0), "trained11": O.trained_like_checkpoint(11),
           "lego": {"fine_model": {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}}}
    poses = {"bench0": O.benchmark_pose(0, 3), "bench1": O.benchmark_pose(1, 3), "generic": O.generic_pose()}
    x = torch.rand(4096, 3) * 8 - 4
    print("encoding max err vs reference encode:", (encode_phase(x) - O.encode(x, 10)).abs().max().item())
    with torch.no_grad():
        for k in sorted({k.rsplit("|", 1)[0] for k in g.files}):
            cname, pname, dims = k.split("|")
            w_, h_, s_ = (int(v) for v in dims.split("x"))
            if w_ > 100:
                continue
            rgb, dep = render_image(cks[cname]["fine_model"], poses[pname], w_, h_, s_)
            ref, refd = g[k + "|rgb"], g[k + "|depth"]
            noise = np.random.default_rng(0).normal(0, 0.01, ref.shape)
            T = ref.astype(np.float64) + noise
            ps = lambda a, b: -10 * np.log10(np.mean((np.asarray(a, np.float64) - b) ** 2))
            print(f"{k:34s} PSNR(bf16,ref)={ps(rgb.numpy(), ref.astype(np.float64)):6.1f}  dPSNR={abs(ps(rgb.numpy(), T) - ps(ref, T)):.4f}"
                  f"  max|rgb|={np.abs(rgb.numpy() - ref).max():.2e}  max|depth|={np.abs(dep.numpy() - refd).max():.2e}")
