"""nerf_dbr_b200 -- B200-native render/train hot path for nerf-dbr's renderer plug-in interface.

The product is ``libnerf_b200.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/nerf_b200.h``); this package is the thin Python host that mirrors the reference's
renderer interface (``src/benchmark/base_renderer.py``) on top of it.  There is no CPU path:
every entry point raises if the CUDA library or a CUDA device is missing.
"""
from .host.lib import NerfB200Error, build_library, load_library  # noqa: F401
from .host.model import NeRFModel, PositionalEncoding  # noqa: F401
from .host.base_renderer import BaseUnifiedRenderer, SharedNeRFModel  # noqa: F401
from .host.b200_renderer import B200Renderer  # noqa: F401
from .host.trainer import B200TrainStep, load_checkpoint, save_checkpoint  # noqa: F401
from .host.b200_trainer import B200Trainer  # noqa: F401
from .host.engine import TrainEngine  # noqa: F401
from .host.data import SyntheticDataset, load_synthetic_data, write_standin_dataset  # noqa: F401
from .host.autograd import render_rays_autograd  # noqa: F401

__all__ = ["B200Renderer", "B200Trainer", "B200TrainStep", "TrainEngine", "SyntheticDataset", "load_synthetic_data", "write_standin_dataset", "save_checkpoint", "load_checkpoint", "render_rays_autograd", "BaseUnifiedRenderer", "SharedNeRFModel", "NeRFModel", "PositionalEncoding",
           "NerfB200Error", "build_library", "load_library"]
