"""Small end-to-end pass of every product kernel, meant for `compute-sanitizer --tool memcheck`
(compute-sanitizer is closed on the round-1 GPU pool, so this has only been run plain, as an edge-shape pass):
    compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py
Renders a 32x24x32 view in the three modes, a 100-sample (ragged: padded tiles) view, the hierarchical path, and
one training step in both modes with a ray count that leaves a partial slab.  No oracle, no checks: the sanitizer
is the checker."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nerf_dbr_b200 as nb  # noqa: E402
from nerf_dbr_b200.host import ops  # noqa: E402
from nerf_dbr_b200.host.trainer import B200TrainStep  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    z = np.load(os.path.join(ROOT, "tests", "golden", "ckpt_lego_stuffed_fp16.npz"))
    weights = {k: torch.from_numpy(z[k].astype(np.float32)).to(dev) for k in z.files}
    net = ops.pack_weights(weights, dev)
    pose = torch.eye(4)
    pose[2, 3] = 4.0
    for mode in (0, 1, 2):
        rgb, depth = ops.render_image(net, pose, 32, 24, 32, mode=mode)
        assert torch.isfinite(rgb).all()
    ops.render_image(net, pose, 16, 8, 100, mode=1)           # S not a power of two
    ops.render_image(net, pose, 8, 4, 300, mode=1)            # several tiles per ray
    ro_h, rd_h = ops.generate_rays(pose, 16, 8)
    ops.render_hierarchical(net, net, ro_h.reshape(-1, 3), rd_h.reshape(-1, 3), 32, 32, mode=1)
    torch.cuda.synchronize()
    g = torch.Generator().manual_seed(0)
    n = 37                                                    # 37 x 64 and 37 x 128 samples: partial last slab
    ro = torch.zeros(n, 3) + torch.tensor([0.0, 0.0, 4.0])
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
    tgt = torch.rand(n, 3, generator=g)
    for mode in (0, 1):
        coarse, fine = nb.NeRFModel().to(dev), nb.NeRFModel().to(dev)
        step = B200TrainStep(coarse, fine, 64, 128, mode=mode)
        loss, _, _ = step(ro.to(dev), rd.to(dev), tgt.to(dev))
        assert torch.isfinite(loss).all()
    torch.cuda.synchronize()
    print("sanitize_small: done,", ops.launch_count(), "launches")


if __name__ == "__main__":
    main()
