"""CPU: host-side logic that needs no GPU -- the flat bucket layout of the training engine, the reference loader of the
test infrastructure, and bench.py's reference arm (the JSON line the driver parses; it runs on the host cores by design)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import REPO


def test_flat_layout_alignment_and_world_divisibility():
    from nerf_dbr_b200.host.engine import FlatLayout
    from nerf_dbr_b200.host.model import NeRFModel
    params = list(NeRFModel().parameters()) + list(NeRFModel().parameters())
    assert len(params) == 44 and sum(p.numel() for p in params) == 2 * 530052
    for world in (1, 2, 3, 4, 8, 16):
        lay = FlatLayout([p.shape for p in params], world)
        assert all(o % 4 == 0 for o in lay.offsets)                       # every tensor starts 16-byte aligned
        assert lay.n_opt % 4 == 0 and lay.n_opt >= 2 * 530052
        assert lay.n % (4 * world) == 0 and lay.n >= lay.n_opt + 4        # room for the tail slots (loss)
        flat = torch.arange(lay.n, dtype=torch.float32)
        views = lay.views(flat)
        for p, v, o in zip(params, views, lay.offsets):
            assert v.shape == p.shape and v.data_ptr() == flat.data_ptr() + 4 * o
        # views tile the bucket without overlap
        ends = [o + p.numel() for o, p in zip(lay.offsets, params)]
        assert all(e <= nxt for e, nxt in zip(ends, lay.offsets[1:] + [lay.n_opt]))


def test_reference_loader_prefers_real_checkout_and_stubs_matplotlib():
    from oracle import refload
    root = refload.reference_root()
    if root is None:
        pytest.skip("no reference sources here (tools/vendor_reference.sh)")
    src = refload.import_reference()
    from src.benchmark.pytorch_renderers import PyTorchCPURenderer
    assert os.path.dirname(os.path.dirname(src.__file__)) == root or src.__name__ == "src"
    assert PyTorchCPURenderer.__name__ == "PyTorchCPURenderer"
    import matplotlib.pyplot as plt
    assert plt is not None


def test_bench_reference_arm_emits_the_contract_line():
    """`bench.py --impl reference`: ONE JSON line on stdout with impl, metric, value, cpu_baseline(kind, cores, sample) and
    an e2e object repeating the value; every host thread is used even when the launcher exported OMP_NUM_THREADS=1."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s at 800x600, 128 samples/ray" and d["unit"] == "Mrays/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    ncores = len(os.sched_getaffinity(0))
    assert cb["cores"] == ncores, (cb["cores"], ncores)                    # not the single thread torchrun would leave it with
    from oracle import refload
    assert cb["kind"] == ("reference" if refload.reference_root() else "port")
