"""Training loop on the B200 kernels with the reference trainer's interface and semantics
(``NeRFTrainer``, src/training/trainer.py): same config keys, same batch format
(``{'image': [H,W,3], 'pose': [4,4], 'focal': float}``), same per-step order -- random ray selection, coarse
(stratified) + fine (uniform) render, ``mse + mse``, optional ``clip_grad_norm_``, Adam, exponential LR decay --
same checkpoint format and resume rule.  What differs is where the work runs: rays, both networks' forward and
backward and the full-image validation render are ``nerf_b200_*`` calls.  By default (``fused_step``) the gradient
exchange, clipping, Adam and the learning-rate decay are this library's kernels too and the whole iteration is one CUDA
graph (``TrainEngine``); with ``fused_step: False`` optimizer, clipping and scheduler are plain PyTorch on the same
``nn.Parameter``s (trainer.py:125-136) and the gradients are all-reduced with NCCL.  Data parallel when
``torch.distributed`` is initialised: every rank draws the same ray selection and trains on its shard."""
from __future__ import annotations

import math
import os
import re
from typing import Dict, Optional

import torch

from . import lib as L
from . import ops
from .model import NeRFModel
from .parallel import broadcast_parameters_, ray_shard
from .engine import TrainEngine
from .trainer import B200TrainStep, load_checkpoint, save_checkpoint


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class B200Trainer:
    """Mirror of ``NeRFTrainer`` (trainer.py:18-81).  ``config`` keys read: hidden_dim / position_encoding_levels /
    direction_encoding_levels (must be the reference's 256 / 10 / 4: the kernels are specialised for that network),
    lr, weight_decay, lr_decay, decay_steps, n_coarse, n_fine, n_rays, near, far, gradient_clipping,
    checkpoint_frequency, plus ``fused_step`` (default True), ``dp_transport`` ('auto' | 'p2p' | 'multimem' | 'nccl'),
    ``precision`` ('bf16' | 'fp32'), ``device_index`` (default: ``$LOCAL_RANK``, else the current device), ``seed``,
    ``checkpoint_dir`` (default 'checkpoints', as the reference)."""

    def __init__(self, config: Dict):
        if not torch.cuda.is_available():
            raise RuntimeError("B200Trainer needs a CUDA device (nerf_dbr_b200 has no CPU path)")
        L.load_library()
        self.config = config
        if (config.get("hidden_dim", 256), config.get("position_encoding_levels", 10),
                config.get("direction_encoding_levels", 4)) != (256, 10, 4):
            raise ValueError("the B200 kernels implement the reference's 8x256 network with 10 / 4 encoding levels")
        # one process per GPU under torchrun: LOCAL_RANK picks the device unless the config names one
        default_index = int(os.environ.get("LOCAL_RANK", torch.cuda.current_device()))
        self.device = torch.device("cuda", int(config.get("device_index", default_index)))
        seed = config.get("seed")
        if seed is not None:
            torch.manual_seed(int(seed))
        self.coarse_model = NeRFModel().to(self.device)
        self.fine_model = NeRFModel().to(self.device)
        params = list(self.coarse_model.parameters()) + list(self.fine_model.parameters())
        # data parallel: every replica starts from rank 0's weights (an unseeded config would otherwise give every rank
        # its own random initialisation, and the all-reduced gradients would belong to no single model)
        broadcast_parameters_(params)
        self.optimizer = torch.optim.Adam(params, lr=config.get("lr", 5e-4), weight_decay=config.get("weight_decay", 0.0))
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(
            self.optimizer, gamma=config.get("lr_decay", 0.1) ** (1 / config.get("decay_steps", 250000)))
        self.n_coarse, self.n_fine = config.get("n_coarse", 64), config.get("n_fine", 128)
        self.near, self.far = config.get("near", 2.0), config.get("far", 6.0)
        self.gradient_clipping = config.get("gradient_clipping", None)
        self.checkpoint_frequency = config.get("checkpoint_frequency", 50)
        self.checkpoint_dir = config.get("checkpoint_dir", "checkpoints")
        self.mode = {"bf16": L.BF16, "fp32": L.FP32}[config.get("precision", "bf16")]
        self.step_fn = B200TrainStep(self.coarse_model, self.fine_model, self.n_coarse, self.n_fine, self.near, self.far,
                                     mode=self.mode)
        self.train_losses, self.val_losses = [], []
        self.fused_step = bool(config.get("fused_step", True))
        self._loaded_state = False
        self.engine: Optional[TrainEngine] = None          # built at the first step, when the batch size is known
        # ray selection and jitter: one generator, identical on every rank (each rank then takes its shard)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(seed) if seed is not None else 0)

    # ------------------------------------------------------------------ one step (trainer.py:83-138)
    def _engine_for(self, n_rays: int) -> TrainEngine:
        rank, world = _dist()
        count = ray_shard(rank, world, n_rays)[1]
        if self.engine is None or self.engine.n_rays != count or self.engine.n_rays_global != n_rays:
            if world > 1 and any(ray_shard(r, world, n_rays)[1] != count for r in range(world)):
                raise ValueError(f"fused_step needs n_rays ({n_rays}) divisible by the world size ({world})")
            cfg, t = self.config, (self.engine.opt_step if self.engine is not None else None)
            if self.engine is not None:
                self.engine.sync_holders()
            previous = (self.optimizer, self.scheduler) if self.engine is not None or self._loaded_state else None
            self.engine = TrainEngine(self.coarse_model, self.fine_model, count, self.n_coarse, self.n_fine, self.near, self.far,
                                      mode=self.mode, lr=cfg.get("lr", 5e-4),
                                      gamma=cfg.get("lr_decay", 0.1) ** (1 / cfg.get("decay_steps", 250000)),
                                      weight_decay=cfg.get("weight_decay", 0.0), max_norm=self.gradient_clipping,
                                      n_rays_global=n_rays, transport=cfg.get("dp_transport", "auto"))
            if previous is not None:                         # carry Adam moments / step count over (resume, or a new batch size)
                self.engine.optimizer.load_state_dict(previous[0].state_dict())
                self.engine.scheduler.load_state_dict(previous[1].state_dict())
                self.engine.adopt_holders()
                if t is not None:
                    self.engine.opt_step = t
            self.optimizer, self.scheduler = self.engine.optimizer, self.engine.scheduler
            self.step_fn.coarse, self.step_fn.fine = self.coarse_model, self.fine_model
        return self.engine

    def train_step(self, batch: Dict, sync: bool = True):
        """One iteration (trainer.py:83-138).  Returns the loss as a float like the reference; ``sync=False`` returns a
        0-dim device tensor instead and does not wait for the GPU."""
        self.coarse_model.train()
        self.fine_model.train()
        image = batch["image"].to(self.device, torch.float32)
        height, width = image.shape[:2]
        n_pixels = height * width
        n_rays = min(int(self.config.get("n_rays", 1024)), n_pixels)
        select = torch.randperm(n_pixels, device=self.device, generator=self._gen)[:n_rays]
        t_rand = torch.rand(n_rays, self.n_coarse, device=self.device, generator=self._gen)
        rank, world = _dist()
        first, count = ray_shard(rank, world, n_rays)
        # this rank's rays and target colours straight from the pixel indices (the bits of _get_rays(pose)[select] and
        # image.reshape(-1, 3)[select], trainer.py:100-118, without the full-image ray tensors)
        rays_o, rays_d, target = ops.ray_batch(batch["pose"], width, height, float(batch["focal"]), select[first:first + count], image)
        if self.fused_step:
            eng = self._engine_for(n_rays)
            eng.step(rays_o, rays_d, target, t_rand[first:first + count])
            loss = eng.out[0].clone()
            return float(loss) if sync else loss
        loss, _, _ = self.step_fn(rays_o, rays_d, target, t_rand=t_rand[first:first + count].contiguous(), n_rays_global=n_rays)
        if self.gradient_clipping is not None:
            torch.nn.utils.clip_grad_norm_(self.step_fn.parameters(), self.gradient_clipping)
        self.optimizer.step()
        self.scheduler.step()
        return float(loss) if sync else loss.detach()

    # ------------------------------------------------------------------ validation (trainer.py:140-170, 355-372)
    def render_image(self, pose: torch.Tensor, img_shape, focal: float) -> torch.Tensor:
        """The fine network's image, ``n_fine`` uniform samples per ray: what ``NeRFTrainer._render_image`` returns."""
        height, width = img_shape
        net = ops.pack_weights({k: v.detach() for k, v in self.fine_model.state_dict().items()}, self.device)
        return ops.render_image(net, pose, width, height, self.n_fine, self.mode, float(focal), self.near, self.far)[0]

    def validate(self, val_dataset) -> float:
        self.coarse_model.eval()
        self.fine_model.eval()
        losses = []
        with torch.no_grad():
            for i in range(min(5, len(val_dataset))):
                batch = val_dataset[i]
                pred = self.render_image(batch["pose"], batch["image"].shape[:2], batch["focal"])
                losses.append(float(torch.mean((pred - batch["image"].to(self.device, torch.float32)) ** 2)))
        return float(sum(losses) / max(len(losses), 1))

    # ------------------------------------------------------------------ epochs, checkpoints (trainer.py:172-268, 374-402)
    def train(self, train_dataset, val_dataset=None, n_epochs: int = 100, verbose: bool = True) -> None:
        say = print if verbose and _dist()[0] == 0 else (lambda *a, **k: None)
        latest = self._find_latest_checkpoint()
        if latest:
            self.load_checkpoint(latest)
            say(f"Resuming from {latest}: {len(self.train_losses)}/{n_epochs} epochs done")
        start = len(self.train_losses)
        for epoch in range(start, n_epochs):
            losses = [self.train_step(train_dataset[i], sync=False) for i in range(len(train_dataset))]      # no read-back per step
            self.train_losses.append(float(torch.stack(losses).mean()) if losses else 0.0)                  # one per epoch
            if val_dataset is not None and (epoch + 1) % 10 == 0:
                self.val_losses.append(self.validate(val_dataset))
                say(f"Epoch {epoch + 1}: train {self.train_losses[-1]:.4f}, val {self.val_losses[-1]:.4f}")
            else:
                say(f"Epoch {epoch + 1}: train {self.train_losses[-1]:.4f}")
            if (epoch + 1) % self.checkpoint_frequency == 0 and _dist()[0] == 0:
                self.save_checkpoint(f"checkpoint_epoch_{epoch + 1}.pth")

    def _find_latest_checkpoint(self) -> Optional[str]:
        if not os.path.isdir(self.checkpoint_dir):
            return None
        best = None
        for name in os.listdir(self.checkpoint_dir):
            m = re.fullmatch(r"checkpoint_epoch_(\d+)\.pth", name)
            if m and (best is None or int(m.group(1)) > best[0]):
                best = (int(m.group(1)), os.path.join(self.checkpoint_dir, name))
        return best[1] if best else None

    def save_checkpoint(self, filename: str) -> str:
        if self.engine is not None:
            self.engine.sync_holders()                      # step count / lr of the device-resident schedule into the torch objects
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        path = os.path.join(self.checkpoint_dir, filename)
        save_checkpoint(path, self.step_fn, self.optimizer, self.scheduler, self.config, self.train_losses, self.val_losses)
        return path

    def load_checkpoint(self, path: str) -> None:
        self.train_losses, self.val_losses = (list(x) for x in load_checkpoint(path, self.step_fn, self.optimizer, self.scheduler))
        broadcast_parameters_(self.step_fn.parameters())       # replicas restart identical even if only rank 0's file is current
        self._loaded_state = True
        if self.engine is not None:
            self.engine.adopt_holders()

    @staticmethod
    def psnr(mse: float) -> float:
        return float("inf") if mse <= 0 else -10.0 * math.log10(mse)
