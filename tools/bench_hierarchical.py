"""BASELINE.json configs[4] shape on one GPU: 1600x1200, 128 coarse + 128 importance samples (inverse-CDF fine pass).
Extra evidence, not the headline metric.  Prints stage timings (CUDA events) and Mrays/s."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nerf_dbr_b200.host import ops
from nerf_dbr_b200.host.synthetic import orbit_pose

dev = torch.device("cuda", 0)
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ckpt_lego_stuffed_fp16.npz"))
net = ops.pack_weights({k: torch.from_numpy(z[k].astype(np.float32)).to(dev) for k in z.files}, dev)
W, H, NC, NI = 1600, 1200, 128, 128
pose = orbit_pose(1, 64)
ro, rd = ops.generate_rays(pose, W, H)
ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
u = torch.rand(ro.shape[0], NI, device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


def run(mode):
    marks = [ev() for _ in range(6)]
    marks[0].record()
    rgb_c, _, wts = ops.render_rays(net, ro, rd, NC, mode, want_weights=True); marks[1].record()
    _, z_c = ops.sample_points(ro, rd, NC); marks[2].record()
    _, z_new, _ = ops.importance_sample(ro, rd, z_c, wts, u); marks[3].record()
    z_all = ops.merge_samples(z_c, z_new); marks[4].record()
    rgb_f, dep_f = ops.render_rays(net, ro, rd, NC + NI, mode, z_vals=z_all); marks[5].record()
    torch.cuda.synchronize()
    return [marks[i].elapsed_time(marks[i + 1]) for i in range(5)], rgb_f


for mode, name in ((1, "bf16"),):
    run(mode)
    t, rgb = run(mode)
    total = sum(t)
    print(json.dumps({"config": "1600x1200, 128 coarse + 128 importance (fine pass on the 256-sample sorted union)", "mode": name,
                      "ms": {"coarse_render": t[0], "sample_points": t[1], "importance_sample": t[2], "merge": t[3], "fine_render": t[4]},
                      "ms_total": total, "mrays_per_s": W * H / total / 1e3,
                      "tflops_mlp": (NC + NC + NI) * W * H * 1.055744e6 / (t[0] + t[4]) / 1e9, "finite": bool(torch.isfinite(rgb).all())}))
