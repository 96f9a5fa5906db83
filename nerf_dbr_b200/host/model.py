"""Host-side mirror of the reference network definition (src/models/nerf.py): same module and
parameter names, so reference ``.pth`` checkpoints (src/training/trainer.py:374-388) load with
``load_state_dict`` and gradients land in the same 22 ``nn.Parameter``s per network.  The modules
hold parameters only -- the arithmetic runs in libnerf_b200.so."""
from __future__ import annotations

import torch
import torch.nn as nn

STATE_ORDER = [f"layers.{i}.{p}" for i in range(8) for p in ("weight", "bias")] + [
    "density_head.weight", "density_head.bias",
    "color_layers.0.weight", "color_layers.0.bias",
    "color_layers.1.weight", "color_layers.1.bias"]


class PositionalEncoding:
    """[x, sin(2^k pi x), cos(2^k pi x)] for k < L (reference src/models/nerf.py:13-45), on the GPU
    through nerf_b200_positional_encoding."""

    def __init__(self, L: int = 10):
        self.L = L

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        from . import ops
        shape = x.shape
        out = ops.positional_encoding(x.reshape(-1, 3), self.L)
        return out.reshape(*shape[:-1], 3 + 6 * self.L)


class NeRFModel(nn.Module):
    """Parameter container with the reference's names and shapes (src/models/nerf.py:48-90):
    8 x 256 trunk with the encoded position entering layer index 4 after the hidden state,
    density head 256->1, colour head 283->128->3.  Construction order matches the reference so a
    seeded random init reproduces its weights."""

    def __init__(self, pos_L: int = 10, dir_L: int = 4, hidden_dim: int = 256):
        super().__init__()
        if (pos_L, dir_L, hidden_dim) != (10, 4, 256):
            raise ValueError("the B200 kernels are specialised for pos_L=10, dir_L=4, hidden_dim=256 "
                             "(the reference's only configuration, main.py:25-62)")
        self.pos_L, self.dir_L = pos_L, dir_L
        self.pos_dim, self.dir_dim = 3 + 6 * pos_L, 3 + 6 * dir_L
        self.pos_encoder, self.dir_encoder = PositionalEncoding(pos_L), PositionalEncoding(dir_L)
        h = hidden_dim
        self.layers = nn.ModuleList([nn.Linear(self.pos_dim, h), nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h),
                                     nn.Linear(h + self.pos_dim, h), nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)])
        self.density_head = nn.Linear(h, 1)
        self.color_layers = nn.ModuleList([nn.Linear(h + self.dir_dim, h // 2), nn.Linear(h // 2, 3)])

    def forward(self, positions: torch.Tensor, directions: torch.Tensor):
        """(density [N,1], rgb [N,3]) like the reference forward (nerf.py:92-131); inference only
        (training goes through ops.train_fwd_bwd).  Packs the weights on every call -- renderers keep
        a packed copy instead (B200Renderer.setup)."""
        from . import ops
        packed = ops.pack_weights(self)
        return ops.query_network(packed, positions, directions)
