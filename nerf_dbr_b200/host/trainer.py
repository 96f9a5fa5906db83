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


def _world_size() -> int:
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class B200TrainStep:
    """Replaces NeRFTrainer._render_rays + loss + backward.  Data parallel: every rank passes its
    shard of the ray batch and the GLOBAL ray count; gradients are all-reduce-summed over NCCL.

    (A software-pipelined schedule over the split-phase entry point -- one pass's weight gradients on a second stream
    and a few SMs beside the next pass's activation phase -- was measured at 4.30-4.80 ms per step against 4.20 ms back
    to back and removed; ``ops.TrainPass`` keeps the split phases for callers that want to overlap a collective.)"""

    def __init__(self, coarse: NeRFModel, fine: NeRFModel, n_coarse: int = 64, n_fine: int = 128,
                 near: float = 2.0, far: float = 6.0, mode: int = L.FP32):
        self.coarse, self.fine = coarse, fine
        self.n_coarse, self.n_fine, self.near, self.far, self.mode = n_coarse, n_fine, near, far, mode
        self._flat = None                            # one bucket: every gradient is a view of it, + 1 slot for the loss

    def _attach_flat_grads(self):
        """Gradients live in ONE flat fp32 buffer (each ``p.grad`` a view of it, like a DDP bucket): zeroing is one
        memset, the data-parallel all-reduce runs in place on the bucket with no gather / scatter copies.  Re-attached
        if someone replaced a ``.grad`` (e.g. ``optimizer.zero_grad(set_to_none=True)``)."""
        params = self.parameters()
        if not params[0].is_cuda:
            return None
        n = sum((p.numel() + 3) // 4 * 4 for p in params)      # every view starts 16-byte aligned
        if self._flat is None or self._flat.device != params[0].device or self._flat.numel() != n + 4:
            self._flat = torch.zeros(n + 4, device=params[0].device, dtype=torch.float32)
        off = 0
        for p in params:
            view = self._flat[off:off + p.numel()].view_as(p)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr() or p.grad.shape != p.shape:
                p.grad = view
            off += (p.numel() + 3) // 4 * 4
        return self._flat

    def parameters(self):
        return list(self.coarse.parameters()) + list(self.fine.parameters())

    def __call__(self, rays_o, rays_d, target, t_rand: Optional[torch.Tensor] = None,
                 n_rays_global: Optional[int] = None, allreduce: bool = True):
        """Zeroes the gradients, runs forward+backward of both networks, all-reduces when a process
        group is initialised, returns (loss, rgb_coarse, rgb_fine)."""
        flat = self._attach_flat_grads()
        if flat is not None:
            flat.zero_()                             # one memset for the 48 gradient tensors
        else:
            for p in self.parameters():
                if p.grad is not None:
                    p.grad.zero_()
        if t_rand is None:                       # the reference jitters the coarse samples (rendering.py:46)
            t_rand = torch.rand(rays_o.shape[0], self.n_coarse, device=rays_o.device)
        n = rays_o.shape[0]
        kw = dict(n_rays_global=n if n_rays_global is None else n_rays_global, near=self.near, far=self.far, mode=self.mode)
        lc, rgb_c = ops.train_fwd_bwd(self.coarse, rays_o, rays_d, target, self.n_coarse, t_rand, **kw)
        lf, rgb_f = ops.train_fwd_bwd(self.fine, rays_o, rays_d, target, self.n_fine, None, **kw)
        loss = lc + lf
        if allreduce and _world_size() > 1:
            if flat is not None:                     # in place on the bucket (4.24 MB over NVLink); the loss rides in its last slot
                flat[-1] = loss
                allreduce_sum_([flat])
                loss = flat[-1].clone()
            else:
                loss = allreduce_sum_([p.grad for p in self.parameters()], extra=loss)
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
