"""A/B of two builds in the regime the power cap does not reach: single 800x75-row bands of the 800x600x128 view (what one
GPU of eight renders), 30 ms of idle between launches so the SM clock stays at its maximum.  Median ms per launch."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import os, sys, json, time, torch
sys.path.insert(0, %r)
import numpy as np
from nerf_dbr_b200.host import ops, lib as L
from nerf_dbr_b200.host.synthetic import orbit_pose
_lib = L.load_library()
z = np.load(os.path.join(%r, "tests", "golden", "ckpt_lego_stuffed_fp16.npz"))
dev = torch.device("cuda", 0)
net = ops.pack_weights({k: torch.from_numpy(z[k].astype(np.float32)).to(dev) for k in z.files}, dev)
rgb, dep = torch.empty(75, 800, 3, device=dev), torch.empty(75, 800, device=dev)
ms = []
for i in range(45):
    time.sleep(0.03)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.render_image(net, orbit_pose(i %% 40, 40), 800, 600, 128, 1, row0=300, n_rows=75, out_rgb=rgb, out_depth=dep)
    e1.record(); torch.cuda.synchronize()
    if i >= 5: ms.append(e0.elapsed_time(e1))
ms.sort()
print(json.dumps({"lib": os.environ.get("NERF_B200_LIB", "in-tree"), "median_ms": ms[len(ms) // 2], "min_ms": ms[0], "mrays_per_s": 60000 / ms[len(ms) // 2] / 1e3}))
''' % (ROOT, ROOT)
libs = [None, os.path.join(ROOT, "tools", "ab", sys.argv[1] if len(sys.argv) > 1 else "libnerf_b200_head.so")]
for rep in range(3):
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["NERF_B200_LIB"] = lib
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
