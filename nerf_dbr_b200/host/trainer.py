"""Training step on the B200 kernels with the reference trainer's semantics
(src/training/trainer.py:83-138, 294-316): coarse network on ``n_coarse`` stratified samples, fine
network on ``n_fine`` uniform samples, loss = mse(coarse) + mse(fine); the optimizer, gradient clipping
and LR schedule stay plain PyTorch on the same ``nn.Parameter``s (trainer.py:125-136)."""
from __future__ import annotations

from typing import Optional

import torch

from . import lib as L
from . import ops
from .model import NeRFModel
from .parallel import allreduce_sum_


class B200TrainStep:
    """Replaces NeRFTrainer._render_rays + loss + backward.  Data parallel: every rank passes its
    shard of the ray batch and the GLOBAL ray count; gradients are all-reduce-summed over NCCL."""

    def __init__(self, coarse: NeRFModel, fine: NeRFModel, n_coarse: int = 64, n_fine: int = 128,
                 near: float = 2.0, far: float = 6.0, mode: int = L.FP32):
        self.coarse, self.fine = coarse, fine
        self.n_coarse, self.n_fine, self.near, self.far, self.mode = n_coarse, n_fine, near, far, mode

    def parameters(self):
        return list(self.coarse.parameters()) + list(self.fine.parameters())

    def __call__(self, rays_o, rays_d, target, t_rand: Optional[torch.Tensor] = None,
                 n_rays_global: Optional[int] = None, allreduce: bool = True):
        """Zeroes the gradients, runs forward+backward of both networks, all-reduces when a process
        group is initialised, returns (loss, rgb_coarse, rgb_fine)."""
        grads = [p.grad for p in self.parameters() if p.grad is not None]
        if grads:
            torch._foreach_zero_(grads)              # one multi-tensor launch instead of 48 fills
        if t_rand is None:                       # the reference jitters the coarse samples (rendering.py:46)
            t_rand = torch.rand(rays_o.shape[0], self.n_coarse, device=rays_o.device)
        lc, rgb_c = ops.train_fwd_bwd(self.coarse, rays_o, rays_d, target, self.n_coarse, t_rand, n_rays_global,
                                      self.near, self.far, self.mode)
        lf, rgb_f = ops.train_fwd_bwd(self.fine, rays_o, rays_d, target, self.n_fine, None, n_rays_global,
                                      self.near, self.far, self.mode)
        loss = lc + lf
        if allreduce:
            loss = allreduce_sum_([p.grad for p in self.parameters()], extra=loss)   # one 4.24 MB sum over NVLink
        return loss, rgb_c, rgb_f


def save_checkpoint(path: str, step: B200TrainStep, optimizer, scheduler=None, config=None, train_losses=(),
                    val_losses=()) -> None:
    """Write a checkpoint in the reference's format (NeRFTrainer.save_checkpoint, src/training/trainer.py:374-388):
    the renderers read 'coarse_model' / 'fine_model' (base_renderer.py:47-48), the reference trainer resumes from
    the rest."""
    torch.save({"coarse_model": {k: v.detach().cpu() for k, v in step.coarse.state_dict().items()},
                "fine_model": {k: v.detach().cpu() for k, v in step.fine.state_dict().items()},
                "optimizer": optimizer.state_dict(),
                "scheduler": scheduler.state_dict() if scheduler is not None else {},
                "config": dict(config or {}), "train_losses": list(train_losses), "val_losses": list(val_losses)}, path)


def load_checkpoint(path: str, step: B200TrainStep, optimizer=None, scheduler=None):
    """Counterpart of NeRFTrainer.load_checkpoint (trainer.py:390-402); returns (train_losses, val_losses)."""
    ck = torch.load(path, map_location=next(step.coarse.parameters()).device, weights_only=False)
    step.coarse.load_state_dict(ck["coarse_model"])
    step.fine.load_state_dict(ck["fine_model"])
    if optimizer is not None and ck.get("optimizer"):
        optimizer.load_state_dict(ck["optimizer"])
    if scheduler is not None and ck.get("scheduler"):
        scheduler.load_state_dict(ck["scheduler"])
    return ck.get("train_losses", []), ck.get("val_losses", [])
