"""Synthetic inputs for benchmarks and tools (no dataset or trained checkpoint ships with the reference): the
suite's camera orbit and seeded networks.  Product-side helpers: nothing here touches ``oracle/``."""
from __future__ import annotations

import math
from typing import Tuple

import torch

from .model import NeRFModel


def orbit_pose(i: int, n_views: int) -> torch.Tensor:
    """Camera-to-world matrix of view ``i`` of ``n_views``: the camera sits at z = 4 and the scene turns about the y
    axis by ``2 pi i / n_views`` (``UnifiedBenchmarkSuite.generate_test_poses``, src/benchmark/benchmark_suite.py:132-149)."""
    th = 2.0 * math.pi * i / n_views
    return torch.tensor([[math.cos(th), 0.0, math.sin(th), 0.0],
                         [0.0, 1.0, 0.0, 0.0],
                         [-math.sin(th), 0.0, math.cos(th), 4.0],
                         [0.0, 0.0, 0.0, 1.0]], dtype=torch.float32)


def seeded_models(seed: int, density_gain: float = 1.0, device="cpu") -> Tuple[NeRFModel, NeRFModel]:
    """(coarse, fine) with the default initialisation under ``torch.manual_seed(seed)``, coarse constructed first (the
    reference's fixture order, test_system.py:197-201); ``density_gain`` scales both density heads (30 gives a
    semi-opaque volume whose weights spread along the ray: a gradient-flow workload)."""
    torch.manual_seed(seed)
    coarse, fine = NeRFModel(), NeRFModel()
    if density_gain != 1.0:
        with torch.no_grad():
            for m in (coarse, fine):
                m.density_head.weight.mul_(density_gain)
                m.density_head.bias.mul_(density_gain)
    return coarse.to(device), fine.to(device)
