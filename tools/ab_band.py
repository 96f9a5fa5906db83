"""A/B helper: device time of the fused render for a row band of 800x600x128 (what one rank of an N-GPU run renders),
e.g.  NERF_B200_CLUSTER=1 python tools/ab_band.py 75   vs   NERF_B200_CLUSTER=2 python tools/ab_band.py 75"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_dbr_b200.host import lib as L, ops  # noqa: E402
from nerf_dbr_b200.host.synthetic import orbit_pose  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 75
dev = torch.device("cuda", 0)
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ckpt_lego_stuffed_fp16.npz"))
net = ops.pack_weights({k: torch.from_numpy(z[k].astype(np.float32)).to(dev) for k in z.files}, dev)
n = max(20, 4800 // rows)
for _ in range(10):
    ops.render_image(net, orbit_pose(1, 40), 800, 600, 128, mode=L.BF16, row0=0, n_rows=rows)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    ops.render_image(net, orbit_pose(i % 40, 40), 800, 600, 128, mode=L.BF16, row0=0, n_rows=rows)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"cluster={os.environ.get('NERF_B200_CLUSTER', 'default')} rows={rows}: {ms:.3f} ms per band, {800 * rows / ms / 1e3:.3f} Mrays/s")
