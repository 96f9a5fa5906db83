"""Diagnostic: TrainEngine per-step loss / gradient norm, graph vs eager launches, bf16 (not a test)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from nerf_dbr_b200.host.engine import TrainEngine
from nerf_dbr_b200.host.synthetic import seeded_models
from nerf_dbr_b200.host import ops

def batch(n, seed):
    g = torch.Generator().manual_seed(seed)
    ro = (torch.zeros(n, 3) + torch.tensor([0.0, 0.0, 4.0])).cuda()
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.25 + torch.tensor([0.0, 0.0, -1.0]), dim=-1).cuda()
    tg = (0.5 + 0.5 * torch.sin(3.0 * rd)).cuda()
    return ro, rd, tg, torch.rand(n, 16, generator=g).cuda()

for mode in (1, 0):
    for use_graph in (True, False):
        torch.manual_seed(3)
        c, f = seeded_models(3, 1.0, "cuda")
        eng = TrainEngine(c, f, 256, 16, 32, mode=mode, lr=5e-4, gamma=0.1 ** (1 / 250000), max_norm=1.0, use_graph=use_graph)
        out = []
        for i in range(40):
            eng.step(*batch(256, i % 2))
            st = eng.last_stats()
            out.append("%.4f/%.3f" % (st["loss"], st["grad_norm"]))
        print("mode", mode, "graph" if use_graph else "eager", " ".join(out), flush=True)

# the same problem through the unfused sequence (B200TrainStep + torch clip / Adam / ExponentialLR), fp32 mode (deterministic)
from nerf_dbr_b200.host.trainer import B200TrainStep
for mode in (0, 1):
    c, f = seeded_models(3, 1.0, "cuda")
    step = B200TrainStep(c, f, 16, 32, mode=mode)
    opt = torch.optim.Adam(step.parameters(), lr=5e-4)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.1 ** (1 / 250000))
    out = []
    for i in range(40):
        ro, rd, tg, tr = batch(256, i % 2)
        loss, _, _ = step(ro, rd, tg, t_rand=tr)
        norm = torch.nn.utils.clip_grad_norm_(step.parameters(), 1.0)
        opt.step(); sched.step()
        out.append("%.4f/%.3f" % (float(loss), float(norm)))
    print("mode", mode, "unfused-torch", " ".join(out), flush=True)
