"""Throughput of query_network (NeRFModel.forward on (point, direction) rows) in BF16 and FP32 mode.  Run under gpurun."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_dbr_b200.host import lib as L, ops  # noqa: E402

dev = torch.device("cuda", 0)
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ckpt_lego_stuffed_fp16.npz"))
net = ops.pack_weights({k: torch.from_numpy(z[k].astype(np.float32)).to(dev) for k in z.files}, dev)
out = {}
for mode, name, n in ((L.BF16, "bf16", 1 << 24), (L.FP32, "fp32", 1 << 21)):
    g = torch.Generator(device=dev).manual_seed(0)
    pos = (torch.rand(n, 3, device=dev, generator=g) - 0.5) * 4
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, device=dev, generator=g), dim=-1)
    for _ in range(3):
        ops.query_network(net, pos, dirs, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.query_network(net, pos, dirs, mode=mode)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[name] = {"rows": n, "ms": ms, "msamples_per_s": n / ms / 1e3, "tflops": 1055744 * n / (ms * 1e-3) / 1e12}
    if mode == L.BF16:          # the reference's usual call: a ray's direction repeated on its 128 samples
        dirs = dirs[:: 128].repeat_interleave(128, dim=0).contiguous()
        for _ in range(3):
            ops.query_network(net, pos, dirs, mode=mode)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            ops.query_network(net, pos, dirs, mode=mode)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[name + "_ray_directions"] = {"rows": n, "ms": ms, "msamples_per_s": n / ms / 1e3, "tflops": 1055744 * n / (ms * 1e-3) / 1e12}
print(json.dumps(out))
