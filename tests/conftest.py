import json
import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def checkpoints():
    """The four checkpoint fixtures the goldens were generated with (tests/golden/make_golden.py)."""
    from oracle import nerf_oracle as O
    z = load_npz("ckpt_lego_stuffed_fp16.npz")
    lego = {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}
    return {
        "rand2": O.seeded_checkpoint(2),
        "semi30": O.seeded_checkpoint(2, 30.0),
        "trained11": O.trained_like_checkpoint(11),
        "lego": {"coarse_model": lego, "fine_model": lego},
    }


@pytest.fixture(scope="session")
def poses():
    from oracle import nerf_oracle as O
    return {"bench0": O.benchmark_pose(0, 3), "bench1": O.benchmark_pose(1, 3),
            "generic": O.generic_pose()}
