"""Diagnostic: B200Trainer loss curves, fused engine vs unfused path, both precisions (not a test)."""
import os, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import nerf_dbr_b200 as nb
from test_gpu_trainer_loop import TinyDataset, _config
from pathlib import Path

ds = TinyDataset()
for precision in ("fp32", "bf16"):
    for fused in (False, True):
        with tempfile.TemporaryDirectory() as tmp:
            cfg = dict(_config(Path(tmp), precision), fused_step=fused)
            tr = nb.B200Trainer(cfg)
            losses = [tr.train_step(ds[i % 2]) for i in range(60)]
            print(precision, "fused" if fused else "unfused", "first10 %.4f last10 %.4f" % (sum(losses[:10]), sum(losses[-10:])),
                  ["%.4f" % x for x in losses[::6]], flush=True)
            if fused:
                print("   stats", tr.engine.last_stats(), "graph", bool(tr.engine.graph), flush=True)
