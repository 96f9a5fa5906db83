"""One warm launch + one measured launch of every standalone HBM kernel at 800x600x128, for an `ncu --set full` capture
(tools/bench_geometry.py launches each kernel 13 times, which makes a full-set capture needlessly long).
usage (gpurun): ncu --set full --clock-control none -k regex:'rays_kernel|composite4|encode_rows|warp_kernel|white4' \
    -o gpurun_out/geometry_full python tools/ncu_geometry_once.py"""
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
ro, rd = ops.generate_rays(pose, W, H)
ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
tr = torch.rand(R, S, device=dev)
x = torch.rand(1 << 23, 3, device=dev)
sigma = torch.rand(R, S, device=dev)
col = torch.rand(R, S, 3, device=dev)
u = torch.rand(R, S, device=dev)
wts = torch.rand(R, S, device=dev)
rgba = torch.randint(0, 256, (100, 800, 800, 4), dtype=torch.uint8, device=dev)
for _ in range(2):
    pts, z = ops.sample_points(ro, rd, S)
    ops.sample_points(ro, rd, S, t_rand=tr)
    ops.positional_encoding(x, 10)
    ops.positional_encoding(x, 4)
    ops.composite(sigma, col, z, rd)
    ops.composite(sigma, col, z, rd, want_aux=True)
    _, z_new, _ = ops.importance_sample(ro, rd, z, wts, u)
    ops.merge_samples(z, z_new)
    ops.hierarchical_samples(wts, S, u=u)
    ops.hierarchical_samples(wts, S, seed=1)
    ops.composite_white(rgba)
    torch.cuda.synchronize()
print("ok")
