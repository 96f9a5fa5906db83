"""HBM roofline of the standalone geometry / encoding / compositing kernels (the staged API; the fused kernel needs
none of them): algorithmic bytes / measured time, at the 800x600x128 shape.  Run under gpurun."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_dbr_b200.host import ops  # noqa: E402

dev = torch.device("cuda", 0)
W, H, S = 800, 600, 128
R = W * H
pose = torch.eye(4)
pose[2, 3] = 4.0


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
ro, rd = ops.generate_rays(pose, W, H)
ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
out["generate_rays"] = (timed(lambda: ops.generate_rays(pose, W, H)), 24 * R)
pts, z = ops.sample_points(ro, rd, S)
out["sample_points"] = (timed(lambda: ops.sample_points(ro, rd, S)), 16 * R * S + 24 * R)
tr = torch.rand(R, S, device=dev)
out["sample_points_jitter"] = (timed(lambda: ops.sample_points(ro, rd, S, t_rand=tr)), 20 * R * S + 24 * R)
n_enc = 1 << 23
x = torch.rand(n_enc, 3, device=dev)
out["positional_encoding_L10"] = (timed(lambda: ops.positional_encoding(x, 10)), (12 + 4 * 63) * n_enc)
out["positional_encoding_L4"] = (timed(lambda: ops.positional_encoding(x, 4)), (12 + 4 * 27) * n_enc)
sigma = torch.rand(R, S, device=dev)
col = torch.rand(R, S, 3, device=dev)
out["composite"] = (timed(lambda: ops.composite(sigma, col, z, rd)), 20 * R * S + 12 * R + 16 * R)
out["composite_with_weights"] = (timed(lambda: ops.composite(sigma, col, z, rd, want_aux=True)), 24 * R * S + 12 * R + 20 * R)
u = torch.rand(R, S, device=dev)
wts = torch.rand(R, S, device=dev)
out["importance_sample"] = (timed(lambda: ops.importance_sample(ro, rd, z, wts, u)), (12 + 24) * R * S + 24 * R)
_, z_new, _ = ops.importance_sample(ro, rd, z, wts, u)
out["merge_samples"] = (timed(lambda: ops.merge_samples(z, z_new)), 16 * R * S)
# the fused sampling kernel of the hierarchical path: weights (+ uniforms) in, sorted union out
out["hierarchical_samples_u_given"] = (timed(lambda: ops.hierarchical_samples(wts, S, u=u)), (4 + 4 + 8) * R * S)
out["hierarchical_samples_philox"] = (timed(lambda: ops.hierarchical_samples(wts, S, seed=1)), (4 + 8) * R * S)
# data path: RGBA8 -> fp32 RGB on white (100 images of 800x800), ray batch of 4096 pixels
rgba = torch.randint(0, 256, (100, 800, 800, 4), dtype=torch.uint8, device=dev)
out["composite_white"] = (timed(lambda: ops.composite_white(rgba)), 16 * rgba.numel() // 4)
img = torch.rand(H, W, 3, device=dev)
sel = torch.randperm(R, device=dev)[:4096]
out["ray_batch_4096"] = (timed(lambda: ops.ray_batch(pose, W, H, 800.0, sel, img)), (8 + 12 + 36) * 4096)
print(json.dumps({k: {"ms": round(ms, 4), "algorithmic_GB": round(b / 1e9, 3), "GB_per_s": round(b / ms / 1e6, 1)} for k, (ms, b) in out.items()}, indent=1))
