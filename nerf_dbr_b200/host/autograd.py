"""``NeRFTrainer._render_rays`` as a ``torch.autograd.Function`` (training boundary): the reference's
unchanged ``train_step`` -- ``F.mse_loss`` (or any other loss), ``loss.backward()``, ``clip_grad_norm_``, Adam
(src/training/trainer.py:117-136) -- keeps working when its ``_render_rays`` (trainer.py:294-316) is this call.

forward: the fused render kernels (no activations kept).  backward: the fused training kernels recompute the
forward and run the backward; they form dL/dC = 2 (C - target) / (3 R) from a target themselves, so an arbitrary
upstream gradient g is fed as the pseudo-target C - 1.5 R g (for the reference's MSE that is its own target again, to
rounding).  For the MSE loss ``B200TrainStep`` does the same work in one pass instead of two."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import lib as L
from . import ops
from .model import STATE_ORDER


class _RenderRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, t_rand, cfg, *params):
        n_coarse, n_fine, near, far, mode = cfg
        k = len(STATE_ORDER)
        coarse = dict(zip(STATE_ORDER, (p.detach() for p in params[:k])))
        fine = dict(zip(STATE_ORDER, (p.detach() for p in params[k:])))
        dev = rays_o.device
        rgb_c = ops.render_rays(ops.pack_weights(coarse, dev), rays_o, rays_d, n_coarse, mode, near, far, t_rand)[0]
        rgb_f = ops.render_rays(ops.pack_weights(fine, dev), rays_o, rays_d, n_fine, mode, near, far)[0]
        ctx.cfg = cfg
        ctx.save_for_backward(rays_o, rays_d, t_rand, rgb_c, rgb_f, *params)
        return rgb_c, rgb_f

    @staticmethod
    def backward(ctx, g_c, g_f):
        n_coarse, n_fine, near, far, mode = ctx.cfg
        rays_o, rays_d, t_rand, rgb_c, rgb_f, *params = ctx.saved_tensors
        k, n = len(STATE_ORDER), rays_o.shape[0]
        grads = []
        for which, (net, g, rgb, s, tr) in enumerate(((params[:k], g_c, rgb_c, n_coarse, t_rand), (params[k:], g_f, rgb_f, n_fine, None))):
            named = dict(zip(STATE_ORDER, net))
            out = {name: torch.zeros_like(p) for name, p in named.items()}
            if g is not None:
                target = rgb - (1.5 * n) * g.to(torch.float32)
                ops.TrainPass(named, rays_o, rays_d, target.contiguous(), s, tr, n_rays_global=n, near=near, far=far, mode=mode,
                              want_rgb=False, slot=which, grad_out=out).run()
            grads.extend(out[name] for name in STATE_ORDER)
        return (None, None, None, None, *grads)


def render_rays_autograd(coarse_model, fine_model, rays_o: torch.Tensor, rays_d: torch.Tensor, n_coarse: int = 64,
                         n_fine: int = 128, near: float = 2.0, far: float = 6.0, mode: int = L.FP32,
                         t_rand: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(rgb_coarse [R,3], rgb_fine [R,3]), differentiable with respect to both models' parameters.  ``t_rand``
    [R,n_coarse]: the coarse pass's stratified jitter (drawn here when None, as rendering.py:46 does)."""
    dev = rays_o.device
    if t_rand is None:
        t_rand = torch.rand(rays_o.shape[0], n_coarse, device=dev)
    params = []
    for m in (coarse_model, fine_model):
        named = dict(m.named_parameters())
        params.extend(named[k] for k in STATE_ORDER)
    return _RenderRays.apply(rays_o.contiguous(), rays_d.contiguous(), t_rand.contiguous(), (n_coarse, n_fine, near, far, mode), *params)
