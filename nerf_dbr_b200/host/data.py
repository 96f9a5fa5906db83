"""Data path with the reference loader's interface (``SyntheticDataset`` / ``load_synthetic_data``,
src/data/loader.py:13-129) and its numbers, device-resident.

Same constructor, same items (``{'image': [H,W,3] fp32, 'pose': [4,4] fp32, 'focal': float}``), same ``get_rays``.
What moves: the PNGs are decoded and LANCZOS-resized on the host exactly as the reference does (PIL), uploaded as
RGBA8 and composited onto white by ``nerf_b200_composite_white`` with the reference's float64 arithmetic -- bit-exact
images at a sixth of the upload; rays for a training step come from ``nerf_b200_ray_batch`` (pixel indices in, rays
and target colours out).  ``write_standin_dataset`` renders a small Blender-format dataset from a checkpoint so the
training loop can be exercised from files: the reference ships no dataset (``data/.gitkeep``).
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import lib as L
from . import ops


class SyntheticDataset:
    """Mirror of the reference class (loader.py:13-108).  ``device`` must be a CUDA device: there is no CPU path."""

    def __init__(self, data_dir: str, split: str = "train", img_wh: Tuple[int, int] = (800, 800), device: str = "cuda"):
        from PIL import Image
        L.load_library()
        self.data_dir, self.split = data_dir, split
        self.img_w, self.img_h = img_wh
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.NerfB200Error("SyntheticDataset", -101, "device must be a CUDA device (nerf_dbr_b200 has no CPU path)")
        with open(os.path.join(data_dir, f"transforms_{split}.json"), "r") as fh:      # FileNotFoundError as in the reference
            self.meta = json.load(fh)
        self.focal = 0.5 * self.img_w / np.tan(0.5 * self.meta["camera_angle_x"])      # loader.py:37
        frames = self.meta["frames"]
        staged = torch.empty(len(frames), self.img_h, self.img_w, 4, dtype=torch.uint8).pin_memory()
        poses = []
        for k, frame in enumerate(frames):
            img = Image.open(os.path.join(data_dir, frame["file_path"] + ".png")).convert("RGBA")
            img = img.resize((self.img_w, self.img_h), Image.LANCZOS)                  # loader.py:45-46: host work, as there
            staged[k] = torch.from_numpy(np.array(img))
            poses.append(np.array(frame["transform_matrix"]))
        with torch.cuda.device(self.device):
            rgba = staged.to(self.device, non_blocking=True)
            self.images = ops.composite_white(rgba)                                    # [N,H,W,3] fp32, loader.py:49-54 on the device
            self.poses = torch.from_numpy(np.stack(poses)).to(torch.float32).to(self.device) if poses else \
                torch.empty(0, 4, 4, device=self.device)
            torch.cuda.synchronize(self.device)
        print(f"Loaded {len(self.images)} images from {split} split")

    def __len__(self):
        return len(self.images)

    def __getitem__(self, idx):
        return {"image": self.images[idx], "pose": self.poses[idx], "focal": self.focal}

    def get_rays(self, pose: torch.Tensor):
        """(rays_o, rays_d) [H,W,3] for a pose (loader.py:78-108), bit-exact."""
        return ops.generate_rays(pose, self.img_w, self.img_h, float(self.focal), device=self.device)

    def ray_batch(self, idx: int, pixel_index: torch.Tensor):
        """(rays_o, rays_d, target) [n,3] for pixels ``pixel_index`` (int64, row * W + col) of image ``idx``."""
        return ops.ray_batch(self.poses[idx], self.img_w, self.img_h, float(self.focal), pixel_index.to(self.device),
                             self.images[idx])


def load_synthetic_data(data_dir: str, device: str = "cuda", img_wh: Tuple[int, int] = (800, 800)) -> Dict[str, SyntheticDataset]:
    """All splits that exist (loader.py:111-129)."""
    datasets = {}
    for split in ("train", "val", "test"):
        try:
            datasets[split] = SyntheticDataset(data_dir, split, img_wh=img_wh, device=device)
        except FileNotFoundError:
            print(f"Warning: {split} split not found in {data_dir}")
    return datasets


def write_standin_dataset(data_dir: str, renderer, n_views: Dict[str, int], img_wh: Tuple[int, int] = (100, 75),
                          samples_per_ray: int = 64, camera_angle_x: Optional[float] = None) -> None:
    """Write a NeRF-synthetic (Blender) format dataset -- ``transforms_{split}.json`` + RGBA PNGs -- rendered from the
    networks loaded in ``renderer`` (a set-up ``B200Renderer``): colour = the fine network's render, alpha = its
    accumulated weight, poses = the benchmark orbit.  ``camera_angle_x`` defaults to the angle whose focal length is the
    renderers' fixed 800 at this width (base_renderer.py:224)."""
    from PIL import Image
    from .synthetic import orbit_pose
    width, height = img_wh
    if camera_angle_x is None:
        camera_angle_x = 2.0 * math.atan(0.5 * width / 800.0)
    os.makedirs(data_dir, exist_ok=True)
    for split, count in n_views.items():
        os.makedirs(os.path.join(data_dir, split), exist_ok=True)
        frames = []
        for i in range(count):
            pose = orbit_pose(i, max(count, 1))
            ro, rd = renderer.generate_rays(pose, width, height)
            rgb, _, acc = ops.render_rays(renderer._net(True, render=True), ro.reshape(-1, 3), rd.reshape(-1, 3), samples_per_ray, renderer.mode,
                                          renderer.near, renderer.far, want_acc=True)
            rgba = torch.cat([rgb, acc[:, None]], dim=-1).clamp(0.0, 1.0).reshape(height, width, 4)
            Image.fromarray((rgba.cpu().numpy() * 255.0 + 0.5).astype(np.uint8), mode="RGBA").save(
                os.path.join(data_dir, split, f"r_{i}.png"))
            frames.append({"file_path": f"./{split}/r_{i}", "transform_matrix": pose.tolist()})
        with open(os.path.join(data_dir, f"transforms_{split}.json"), "w") as fh:
            json.dump({"camera_angle_x": camera_angle_x, "frames": frames}, fh, indent=1)
