"""Host-side mirror of the reference's renderer plug-in interface
(src/benchmark/base_renderer.py:16-281): ``SharedNeRFModel`` (per-device checkpoint cache) and
``BaseUnifiedRenderer`` (name/device, setup, performance_monitor, get_device_info,
query_nerf_networks, generate_rays, sample_points_on_rays and the two abstract methods).

When nerf-dbr itself is importable (``src.benchmark.base_renderer``), B200Renderer derives from
the reference's own base class instead, so the suite's isinstance/duck-typing sees a native
renderer.  This mirror is what runs standalone.
"""
from __future__ import annotations

import os
import threading
import time
from abc import ABC, abstractmethod
from contextlib import contextmanager
from typing import Dict, Optional, Tuple

import torch

from .model import NeRFModel


# device -> (checkpoint path the pair was read from or None, coarse, fine); every SharedNeRFModel() is a
# stateless view of this one table, which is what makes it "shared" between renderers
_MODEL_CACHE: Dict[str, Tuple[Optional[str], NeRFModel, NeRFModel]] = {}


class SharedNeRFModel:
    """One (coarse, fine) model pair per device, read from a reference-format ``.pth`` (keys 'coarse_model' /
    'fine_model', base_renderer.py:42-48).  Like the reference, a missing checkpoint file leaves the randomly
    initialised pair in place and says so (base_renderer.py:62-76)."""

    def load_models(self, checkpoint_path: str, device: str = "cpu") -> None:
        cached = _MODEL_CACHE.get(device)
        if cached is not None and cached[0] == checkpoint_path:
            return
        pair = [NeRFModel().to(device).eval() for _ in range(2)]        # coarse first, fine second
        source: Optional[str] = None
        if os.path.exists(checkpoint_path):
            state = torch.load(checkpoint_path, map_location=device, weights_only=False)
            for model, key in zip(pair, ("coarse_model", "fine_model")):
                model.load_state_dict(state[key])
            source = checkpoint_path
        else:
            print(f"checkpoint {checkpoint_path!r} does not exist: rendering with randomly initialised networks")
        _MODEL_CACHE[device] = (source, pair[0], pair[1])

    def get_models(self, device: str = "cpu") -> Tuple[NeRFModel, NeRFModel]:
        if device not in _MODEL_CACHE:
            raise RuntimeError(f"no models for device {device!r}: call load_models(checkpoint_path, device) first")
        _, coarse, fine = _MODEL_CACHE[device]
        return coarse, fine


class _RssSampler(threading.Thread):
    """Polls the process's resident set size every 10 ms until stopped; ``peak_mb`` is the largest value seen."""

    def __init__(self):
        super().__init__(daemon=True)
        import psutil
        self._proc, self._done = psutil.Process(), threading.Event()
        self.peak_mb = self._proc.memory_info().rss / 2 ** 20

    def run(self):
        while not self._done.wait(0.01):
            self.peak_mb = max(self.peak_mb, self._proc.memory_info().rss / 2 ** 20)

    def finish(self) -> float:
        self._done.set()
        self.join()
        return max(self.peak_mb, self._proc.memory_info().rss / 2 ** 20)


class BaseUnifiedRenderer(ABC):
    """Same surface as the reference base class (base_renderer.py:90-281)."""

    def __init__(self, name: str, device: str = "cpu"):
        self.name = name
        self.device = device
        self.shared_model = SharedNeRFModel()
        self.last_render_time = 0.0
        self.peak_memory_mb = 0.0
        self.near = 2.0
        self.far = 6.0

    def setup(self, checkpoint_path: str):
        self.shared_model.load_models(checkpoint_path, self.device)

    @contextmanager
    def performance_monitor(self):
        """Sets ``last_render_time`` (wall clock, device-synchronised on both sides) and ``peak_memory_mb`` (peak host
        RSS) for the body, the two attributes the suite reads (base_renderer.py:118-147, benchmark_suite.py:207-208)."""
        on_gpu = self.device.startswith("cuda")
        sampler = _RssSampler()
        sampler.start()
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            yield
        finally:
            if on_gpu:
                torch.cuda.synchronize()
            self.last_render_time = time.perf_counter() - t0
            self.peak_memory_mb = sampler.finish()

    def get_device_info(self) -> str:
        if self.device.startswith("cuda"):
            return f"CUDA - {torch.cuda.get_device_name()}"
        import psutil
        return f"CPU - {psutil.cpu_count()} cores"

    @abstractmethod
    def query_nerf_networks(self, positions, directions, use_fine: bool = True):
        ...

    @abstractmethod
    def execute_volume_rendering(self, densities, colors, z_vals, ray_directions) -> Tuple[torch.Tensor, torch.Tensor]:
        ...

    @abstractmethod
    def render_image(self, camera_pose, resolution: Tuple[int, int], samples_per_ray: int = 64):
        ...
