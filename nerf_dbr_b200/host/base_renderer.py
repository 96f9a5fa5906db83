"""Host-side mirror of the reference's renderer plug-in interface
(src/benchmark/base_renderer.py:16-281): ``SharedNeRFModel`` (per-device checkpoint cache) and
``BaseUnifiedRenderer`` (name/device, setup, performance_monitor, get_device_info,
query_nerf_networks, generate_rays, sample_points_on_rays and the two abstract methods).

When nerf-dbr itself is importable (``src.benchmark.base_renderer``), B200Renderer derives from
the reference's own base class instead, so the suite's isinstance/duck-typing sees a native
renderer.  This mirror is what runs standalone.
"""
from __future__ import annotations

import threading
import time
from abc import ABC, abstractmethod
from contextlib import contextmanager
from typing import Tuple

import torch

from .model import NeRFModel


class SharedNeRFModel:
    """One (coarse, fine) model pair per device, loaded from a reference-format ``.pth``
    (keys 'coarse_model' / 'fine_model', base_renderer.py:42-48).  Like the reference, a missing
    checkpoint yields randomly initialised models (base_renderer.py:62-76)."""

    _instance = None
    _models_by_device: dict = {}
    _loaded_checkpoint = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
        return cls._instance

    def load_models(self, checkpoint_path: str, device: str = "cpu"):
        if device in self._models_by_device and self._loaded_checkpoint == checkpoint_path:
            return
        coarse, fine = NeRFModel().to(device), NeRFModel().to(device)
        try:
            ckpt = torch.load(checkpoint_path, map_location=device, weights_only=False)
            coarse.load_state_dict(ckpt["coarse_model"])
            fine.load_state_dict(ckpt["fine_model"])
            self._loaded_checkpoint = checkpoint_path
        except FileNotFoundError:
            print("Checkpoint not found, using randomly initialized models")
        coarse.eval()
        fine.eval()
        self._models_by_device[device] = {"coarse": coarse, "fine": fine}

    def get_models(self, device: str = "cpu"):
        if device not in self._models_by_device:
            raise RuntimeError(f"Models not loaded for device {device}. Call load_models() first.")
        m = self._models_by_device[device]
        return m["coarse"], m["fine"]


class BaseUnifiedRenderer(ABC):
    """Same surface as the reference base class (base_renderer.py:90-281)."""

    def __init__(self, name: str, device: str = "cpu"):
        self.name = name
        self.device = device
        self.shared_model = SharedNeRFModel()
        self.last_render_time = 0.0
        self.peak_memory_mb = 0.0
        self._monitoring = False
        self.near = 2.0
        self.far = 6.0

    def setup(self, checkpoint_path: str):
        self.shared_model.load_models(checkpoint_path, self.device)

    @contextmanager
    def performance_monitor(self):
        """Wall time with device sync on both sides and peak host RSS, like the reference
        (base_renderer.py:118-147)."""
        import psutil
        proc = psutil.Process()
        self.peak_memory_mb = proc.memory_info().rss / 1024 / 1024
        self._monitoring = True

        def poll():
            while self._monitoring:
                self.peak_memory_mb = max(self.peak_memory_mb, proc.memory_info().rss / 1024 / 1024)
                time.sleep(0.01)
        th = threading.Thread(target=poll)
        th.start()
        if self.device.startswith("cuda"):
            torch.cuda.synchronize()
        start = time.time()
        try:
            yield
        finally:
            if self.device.startswith("cuda"):
                torch.cuda.synchronize()
            self.last_render_time = time.time() - start
            self._monitoring = False
            th.join()

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
