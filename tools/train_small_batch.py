"""What one GPU does in an 8-way strong-scaled training step: TrainEngine with 512 rays of a 4096-ray global batch
(the exchange degenerates to local at world 1, everything else is the per-GPU work at N = 8).  Prints ms per step for
the graph and for eager launches; with --ncu-friendly runs only a few eager steps (for an ncu launch list)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerf_dbr_b200.host import lib as L
from nerf_dbr_b200.host.engine import TrainEngine
from nerf_dbr_b200.host.synthetic import seeded_models

rays = int(sys.argv[1]) if len(sys.argv) > 1 else 512
few = "--ncu-friendly" in sys.argv
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
ro = torch.zeros(rays, 3) + torch.tensor([0.0, 0.0, 4.0])
rd = torch.nn.functional.normalize(torch.randn(rays, 3, generator=g) * 0.25 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
b = [t.to(dev) for t in (ro, rd, torch.rand(rays, 3, generator=g), torch.rand(rays, 64, generator=g))]
out = {"rays_per_gpu": rays, "global_rays": 4096}
conc = None if "--auto" in sys.argv else ("--concurrent" in sys.argv)
out["concurrent_passes"] = conc
for use_graph in ((False,) if few else (True, False)):
    c, f = seeded_models(5, 30.0, dev)
    eng = TrainEngine(c, f, rays, 64, 128, mode=L.BF16, lr=5e-4, max_norm=1.0, n_rays_global=4096, use_graph=use_graph, data_parallel=False,
                      concurrent_passes=conc)
    n_warm, n = (2, 3) if few else (10, 100)
    for _ in range(n_warm):
        eng.step(*b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.step(*b)
    e1.record()
    torch.cuda.synchronize()
    out["ms_per_step_graph" if use_graph else "ms_per_step_eager"] = e0.elapsed_time(e1) / n
    out["kernels_per_step"] = eng.launches_per_step
print(json.dumps(out))
