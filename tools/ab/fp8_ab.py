"""A/B of builds of libnerf_b200.so on ONE box, FP8 mode, 800x600x128: in-tree build first, then every library named on the
command line (files under tools/ab/), each in a fresh process.  Prints Mrays/s per run."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
import numpy as np
from nerf_dbr_b200.host import ops, lib as L
from nerf_dbr_b200.host.synthetic import orbit_pose
z = np.load(os.path.join(%r, "tests", "golden", "ckpt_lego_stuffed_fp16.npz"))
dev = torch.device("cuda", 0)
sd = {k: torch.from_numpy(z[k].astype(np.float32)).to(dev) for k in z.files}
net = ops.pack_weights_fp8(sd, dev)
rgb, dep = torch.empty(600, 800, 3, device=dev), torch.empty(600, 800, device=dev)
for i in range(5):
    ops.render_image(net, orbit_pose(i, 40), 800, 600, 128, L.FP8, out_rgb=rgb, out_depth=dep)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30):
    ops.render_image(net, orbit_pose(i %% 40, 40), 800, 600, 128, L.FP8, out_rgb=rgb, out_depth=dep)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"lib": os.path.basename(os.environ.get("NERF_B200_LIB", "in-tree")), "mrays_per_s": 480000 * 30 / e0.elapsed_time(e1) / 1e3}))
''' % (ROOT, ROOT)
libs = [None] + [os.path.join(ROOT, "tools", "ab", a) for a in sys.argv[1:]]
for rep in range(2):
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["NERF_B200_LIB"] = lib
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
