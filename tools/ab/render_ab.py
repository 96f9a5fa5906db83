"""A/B of two builds of libnerf_b200.so on ONE box: 800x600x128 BF16 render, alternating processes (each build gets a
fresh process; NERF_B200_LIB selects the library).  Prints Mrays/s per run."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
import numpy as np
from nerf_dbr_b200.host import ops, lib as L
from nerf_dbr_b200.host.synthetic import orbit_pose
_lib = L.load_library()
if not hasattr(_lib, "nerf_b200_pack_weights_ex"):        # an older build: same packing through its only entry point
    _lib.nerf_b200_pack_weights_ex = lambda p, v, what, s: _lib.nerf_b200_pack_weights(p, v, s)
z = np.load(os.path.join(%r, "tests", "golden", "ckpt_lego_stuffed_fp16.npz"))
dev = torch.device("cuda", 0)
net = ops.pack_weights({k: torch.from_numpy(z[k].astype(np.float32)).to(dev) for k in z.files}, dev)
rgb, dep = torch.empty(600, 800, 3, device=dev), torch.empty(600, 800, device=dev)
for i in range(5):
    ops.render_image(net, orbit_pose(i, 40), 800, 600, 128, 1, out_rgb=rgb, out_depth=dep)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30):
    ops.render_image(net, orbit_pose(i %% 40, 40), 800, 600, 128, 1, out_rgb=rgb, out_depth=dep)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"lib": os.environ.get("NERF_B200_LIB", "in-tree"), "mrays_per_s": 480000 * 30 / e0.elapsed_time(e1) / 1e3}))
''' % (ROOT, ROOT)
libs = [None, os.path.join(ROOT, "tools", "ab", sys.argv[1] if len(sys.argv) > 1 else "libnerf_b200_r1.so")]
for rep in range(3):
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["NERF_B200_LIB"] = lib
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
