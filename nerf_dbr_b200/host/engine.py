"""The whole training iteration on the device: ``NeRFTrainer.train_step`` (src/training/trainer.py:83-138) from the
ray batch to the updated weights as one CUDA graph of this library's kernels.

    forward + backward of both networks (nerf_b200_train_fwd_bwd x 2, gradients into one flat bucket)
    -> gradient sum over the data-parallel ranks over NVLink peer memory (nerf_b200_dp_reduce)
    -> gradient norm, clip, Adam with the exponential learning-rate decay (nerf_b200_dp_adam_step)
    -> bf16 re-pack of both networks for the next step (nerf_b200_pack_weights x 2)

What stays on the host is choosing the rays.  Parameters, gradients and Adam moments are flat fp32 buckets (every
``nn.Parameter`` / ``.grad`` / optimizer-state tensor is a view), the step count lives on the device, nothing is read
back unless the caller asks for the loss.  ``torch.optim.Adam`` and ``ExponentialLR`` objects are kept as *state
holders* so that checkpoints keep the reference's format (trainer.py:374-402); they are never stepped.

Data parallel (one process per GPU, ``torch.distributed`` initialised): the gradient bucket sits in torch symmetric
memory, each rank sums its 1/world shard straight out of the peers' buckets and writes the sums into every rank --
one launch, no collective library on the path; replicas stay bit-identical (fixed summation order).  When symmetric
memory cannot be set up the bucket is all-reduced with NCCL and the same two kernels run locally.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, List, Optional

import torch

from . import lib as L
from . import ops
from .model import STATE_ORDER, NeRFModel


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


_SIDE_STREAMS: Dict[torch.device, "torch.cuda.Stream"] = {}


class FlatLayout:
    """Offsets of a list of tensors inside one flat fp32 bucket; every tensor starts 16-byte aligned."""

    def __init__(self, shapes: List[torch.Size], world: int = 1, tail: int = 4):
        self.shapes, self.offsets, off = list(shapes), [], 0
        for s in self.shapes:
            self.offsets.append(off)
            off += (math.prod(s) + 3) // 4 * 4
        self.n_opt = off                                              # parameters (and zero pads)
        quantum = 4 * max(1, world)
        self.n = (off + tail + quantum - 1) // quantum * quantum      # + the tail slots (slot 0: loss), a multiple of 4 * world

    def views(self, flat: torch.Tensor) -> List[torch.Tensor]:
        return [flat[o:o + math.prod(s)].view(s) for o, s in zip(self.offsets, self.shapes)]


class TrainEngine:
    """One object = one training configuration on one GPU (static ray-batch size).  ``step(rays_o, rays_d, target,
    t_rand)`` runs a whole iteration asynchronously; ``loss()`` / ``last_stats()`` read back when asked."""

    def __init__(self, coarse: NeRFModel, fine: NeRFModel, n_rays: int, n_coarse: int = 64, n_fine: int = 128,
                 near: float = 2.0, far: float = 6.0, mode: int = L.BF16, lr: float = 5e-4, gamma: float = 1.0,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, max_norm: Optional[float] = None,
                 n_rays_global: Optional[int] = None, use_graph: bool = True, transport: str = "auto",
                 data_parallel: bool = True, concurrent_passes: Optional[bool] = None):
        """``n_rays``: this rank's share of the batch (static).  ``transport``: 'auto' (NVLink peer memory when the
        process group supports symmetric memory -- through the NVSwitch's in-network reduction when it offers a
        multicast mapping -- else NCCL), 'p2p' (peer loads / stores), 'multimem' (in-switch reduction), 'nccl'.
        ``data_parallel=False`` ignores an initialised process group (a single-replica engine).
        ``concurrent_passes``: run the coarse and the fine network's chains side by side on two streams and disjoint SM
        sets (a third / two thirds).  Default: only for small per-GPU batches (BF16 mode, under 200k samples), where the
        fixed cost of every kernel of a chain -- ramp, tail, the small kernels between the big ones -- is what is left;
        at full batch the passes are bound by HBM and gain nothing from sharing it."""
        self.lib = L.load_library()
        self.coarse, self.fine = coarse, fine
        self.params: List[torch.nn.Parameter] = list(coarse.parameters()) + list(fine.parameters())
        self.dev = self.params[0].device
        if self.dev.type != "cuda":
            raise L.NerfB200Error("TrainEngine", -101, "parameters must live on a CUDA device (nerf_dbr_b200 has no CPU path)")
        dist, self.rank, self.world = _dist() if data_parallel else (None, 0, 1)
        self.n_rays, self.n_coarse, self.n_fine, self.mode = int(n_rays), n_coarse, n_fine, mode
        self.n_rays_global = int(n_rays_global) if n_rays_global is not None else self.n_rays * self.world
        self.hyper_host = dict(lr=lr, gamma=gamma, beta1=betas[0], beta2=betas[1], eps=eps, weight_decay=weight_decay,
                               max_norm=(max_norm if max_norm else 0.0))
        lay = self.layout = FlatLayout([p.shape for p in self.params], self.world)
        with torch.cuda.device(self.dev):
            # ---- flat parameter bucket: every nn.Parameter becomes a view of it
            self.P = torch.zeros(lay.n, device=self.dev)
            for p, view in zip(self.params, lay.views(self.P)):
                view.copy_(p.data)
                p.data = view
            self.M, self.V = torch.zeros(lay.n, device=self.dev), torch.zeros(lay.n, device=self.dev)
            # ---- symmetric [ctl | G | Gsum] and the peers' mappings
            self.transport, self._symm = "local" if self.world == 1 else transport, None
            nbytes = self.lib.nerf_b200_dp_bytes(lay.n)
            self.dp = L.DP()
            self.dp.n, self.dp.n_opt = lay.n, lay.n_opt
            self.block = None
            if self.world > 1 and transport != "nccl":
                try:
                    self.block = self._symmetric_block(nbytes, want_multicast=transport)
                except Exception as e:  # noqa: BLE001 -- fall back to NCCL, say why
                    if transport in ("p2p", "multimem"):
                        raise
                    print(f"[nerf_dbr_b200] symmetric memory unavailable ({type(e).__name__}: {e}); gradients go through NCCL")
            if self.block is None:
                self.transport = "local" if self.world == 1 else "nccl"
                self.block = torch.zeros(nbytes, dtype=torch.uint8, device=self.dev)
                self.dp.rank, self.dp.world = 0, 1
                self.dp.peer[0] = self.block.data_ptr()
            self.state = torch.zeros(L.DP_STATE_BYTES, dtype=torch.uint8, device=self.dev)
            self.dp.state = self.state.data_ptr()
            fl = self.block[L.DP_CTL_BYTES:].view(torch.float32)
            self.G, self.Gsum = fl[:lay.n], fl[lay.n:2 * lay.n]
            for p, view in zip(self.params, lay.views(self.G)):
                p.grad = view
            self.loss_slot = self.G[lay.n_opt:lay.n_opt + 1]
            self.out = torch.zeros(4, device=self.dev)                 # loss, gradient norm, lr
            self.hyper = torch.tensor([lr, gamma, betas[0], betas[1], eps, weight_decay, self.hyper_host["max_norm"],
                                       1.0 / (3.0 * self.n_rays_global)], dtype=torch.float64, device=self.dev)
            # ---- static inputs and the two passes
            self.rays_o, self.rays_d, self.target = (torch.zeros(self.n_rays, 3, device=self.dev) for _ in range(3))
            self.t_rand = torch.zeros(self.n_rays, n_coarse, device=self.dev)
            self.packed = [torch.empty(self.lib.nerf_b200_packed_bytes() + 1024, dtype=torch.uint8, device=self.dev) for _ in range(2)]
            # per-step re-pack: only what this mode's kernels read (the first pack below writes everything once)
            self._pack_what = L.PACK_DGRAD if mode == L.BF16 else L.PACK_ALL
            self._named = []
            for model in (coarse, fine):
                self._named.append({k: dict(model.named_parameters())[k] for k in STATE_ORDER})
            self._packed_views = [ops.pack_weights({k: v.detach() for k, v in nm.items()}, self.dev, out=buf)
                                  for nm, buf in zip(self._named, self.packed)]
            self.passes = []
            for which, (nm, samples, tr) in enumerate(((self._named[0], n_coarse, self.t_rand), (self._named[1], n_fine, None))):
                grads = {k: nm[k].grad for k in STATE_ORDER}
                self.passes.append(ops.TrainPass({k: v.detach() for k, v in nm.items()}, self.rays_o, self.rays_d, self.target, samples, tr,
                                                 n_rays_global=self.n_rays_global, near=near, far=far, mode=mode, want_rgb=False,
                                                 slot=which, packed=self._packed_views[which], grad_out=grads, loss_sum=self.loss_slot))
            torch.cuda.synchronize(self.dev)
        if concurrent_passes is None:
            concurrent_passes = mode == L.BF16 and self.n_rays * (n_coarse + n_fine) < 200_000
        self.concurrent_passes = bool(concurrent_passes)
        # one side stream per device for every engine of the process (the library keeps a set of auxiliary streams per
        # caller stream, include/nerf_b200.h)
        if self.concurrent_passes and self.dev not in _SIDE_STREAMS:
            _SIDE_STREAMS[self.dev] = torch.cuda.Stream(device=self.dev)
        self._side = _SIDE_STREAMS.get(self.dev) if self.concurrent_passes else None
        self.graph = None
        self.use_graph = use_graph
        self.steps_enqueued = 0
        self.launches_per_step = 0
        self._t_host = 0                                            # optimizer steps taken, tracked on the host too (no read-back)
        self._holders()

    # ------------------------------------------------------------------ symmetric memory (torch plumbing)
    def _symmetric_block(self, nbytes: int, want_multicast: str) -> torch.Tensor:
        """``want_multicast``: 'multimem' = required, 'auto' = used when the group has an NVSwitch multicast mapping
        (the in-switch reduction measured 0.679 against 0.697 ms per 4096-ray step on 8 B200), 'p2p' = never."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        block = symm.empty(nbytes, dtype=torch.uint8, device=self.dev)
        block.zero_()
        handle = symm.rendezvous(block, dist.group.WORLD)
        ptrs = list(handle.buffer_ptrs)
        if len(ptrs) != self.world or self.world > L.DP_MAX_WORLD:
            raise RuntimeError(f"symmetric memory gave {len(ptrs)} peer mappings for world {self.world}")
        self.dp.rank, self.dp.world = self.rank, self.world
        for q, ptr in enumerate(ptrs):
            self.dp.peer[q] = ptr
        self.transport = "p2p"
        if want_multicast in ("multimem", "auto"):
            mc = int(getattr(handle, "multicast_ptr", 0) or 0)
            if mc:
                self.dp.multicast = mc
                self.transport = "multimem"
            elif want_multicast == "multimem":
                raise RuntimeError("this process group has no NVSwitch multicast mapping (multicast_ptr == 0)")
        self._symm = handle
        torch.cuda.synchronize(self.dev)
        dist.barrier()                                               # every rank's block is zeroed before anyone signals
        return block

    # ------------------------------------------------------------------ the step
    def _enqueue(self) -> None:
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if self.concurrent_passes:
            main = torch.cuda.current_stream()
            sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
            share = max(1, round(sms * self.n_coarse / (self.n_coarse + self.n_fine)))
            self._side.wait_stream(main)                                 # fork: the coarse chain on the side stream
            with torch.cuda.stream(self._side):
                self.passes[0].run(sm_limit=share)
            self.passes[1].run(sm_limit=sms - share)
            main.wait_stream(self._side)                                 # join before the gradient exchange
        else:
            for tp in self.passes:
                tp.run()
        if self.transport == "nccl":
            import torch.distributed as dist
            dist.all_reduce(self.G)
        L.check("nerf_b200_dp_reduce", self.lib.nerf_b200_dp_reduce(ctypes.byref(self.dp), stream))
        L.check("nerf_b200_dp_adam_step", self.lib.nerf_b200_dp_adam_step(
            ctypes.byref(self.dp), ctypes.c_void_p(self.P.data_ptr()), ctypes.c_void_p(self.M.data_ptr()),
            ctypes.c_void_p(self.V.data_ptr()), ctypes.c_void_p(self.hyper.data_ptr()), ctypes.c_void_p(self.out.data_ptr()), stream))
        for nm, buf in zip(self._named, self.packed):                  # next step's operand streams from the new weights
            ops.pack_weights({k: v.detach() for k, v in nm.items()}, self.dev, out=buf, what=self._pack_what)

    def _capture(self) -> None:
        """Warm up eagerly on a side stream (library-owned streams/events and function attributes get created outside
        the capture), then capture one step.  The warm-up steps are real steps."""
        self.graph = False
        if not self.use_graph or self.transport == "nccl":
            return
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self.graph = g
        except Exception as e:  # noqa: BLE001
            print(f"[nerf_dbr_b200] CUDA graph capture failed ({type(e).__name__}: {e}); launching eagerly")
            torch.cuda.synchronize(self.dev)
            self.graph = False

    def load_batch(self, rays_o, rays_d, target, t_rand=None) -> None:
        """Copy this rank's share of the batch into the static input buffers (device-to-device or pinned host-to-device)."""
        self.rays_o.copy_(rays_o, non_blocking=True)
        self.rays_d.copy_(rays_d, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        if t_rand is None:                                          # the reference jitters the coarse samples (rendering.py:46)
            self.t_rand.uniform_()
        else:
            self.t_rand.copy_(t_rand, non_blocking=True)

    def step(self, rays_o=None, rays_d=None, target=None, t_rand=None) -> None:
        """One iteration, asynchronous.  Without arguments the static input buffers are used as they are."""
        with torch.cuda.device(self.dev):
            if rays_o is not None:
                self.load_batch(rays_o, rays_d, target, t_rand)
            if self.graph is None and self.steps_enqueued > 0:
                self._capture()                                      # (the capture itself is not a step)
            if self.graph:
                self.graph.replay()
            else:
                n0 = ops.launch_count()
                self._enqueue()                                      # the first step runs eagerly: creates the library's lazy state
                self.launches_per_step = ops.launch_count() - n0     # kernels of this library per iteration (= graph kernel nodes)
            self.steps_enqueued += 1
            self._t_host += 1
            for grp in self.optimizer.param_groups:                  # what the reference's scheduler would show
                grp["lr"] = self.hyper_host["lr"] * self.hyper_host["gamma"] ** self._t_host

    # ------------------------------------------------------------------ read-back (synchronises)
    def last_stats(self) -> Dict[str, float]:
        loss, norm, lr = self.out[:3].tolist()
        return {"loss": loss, "grad_norm": norm, "lr": lr}

    def loss(self) -> float:
        return float(self.out[0])

    @property
    def opt_step(self) -> int:
        """Optimizer steps taken (Adam's ``step``, the scheduler's ``last_epoch``); the device holds the same count."""
        return self._t_host

    @opt_step.setter
    def opt_step(self, t: int) -> None:
        self._t_host = int(t)
        self.state[L.DP_STATE_OPT_STEP:L.DP_STATE_OPT_STEP + 4].view(torch.int32).fill_(int(t))

    def device_opt_step(self) -> int:
        return int(self.state[L.DP_STATE_OPT_STEP:L.DP_STATE_OPT_STEP + 4].view(torch.int32).item())

    # ------------------------------------------------------------------ reference-format optimizer / scheduler state
    def _holders(self) -> None:
        h = self.hyper_host
        self.optimizer = torch.optim.Adam(self.params, lr=h["lr"], betas=(h["beta1"], h["beta2"]), eps=h["eps"],
                                          weight_decay=h["weight_decay"])
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, gamma=h["gamma"])
        for p, m, v in zip(self.params, self.layout.views(self.M), self.layout.views(self.V)):
            self.optimizer.state[p] = {"step": torch.tensor(0.0), "exp_avg": m, "exp_avg_sq": v}

    def sync_holders(self) -> None:
        """Bring the torch optimizer / scheduler objects up to date with the device state (before a checkpoint)."""
        t = self.opt_step
        for p in self.params:
            self.optimizer.state[p]["step"] = torch.tensor(float(t))
        lr_t = self.hyper_host["lr"] * self.hyper_host["gamma"] ** t
        for grp in self.optimizer.param_groups:
            grp["lr"] = lr_t
        self.scheduler.last_epoch = t
        self.scheduler._step_count = t + 1
        self.scheduler._last_lr = [lr_t]

    def adopt_holders(self) -> None:
        """After ``optimizer.load_state_dict`` / ``scheduler.load_state_dict`` / ``model.load_state_dict``: move the loaded
        state into the flat buckets, re-point the views, re-pack the weights and take over the step count."""
        with torch.cuda.device(self.dev), torch.no_grad():
            steps = []
            for p, pv, m, v in zip(self.params, self.layout.views(self.P), self.layout.views(self.M), self.layout.views(self.V)):
                if p.data.data_ptr() != pv.data_ptr():
                    pv.copy_(p.data)
                    p.data = pv
                st = self.optimizer.state.get(p, {})
                if "exp_avg" in st and st["exp_avg"].data_ptr() != m.data_ptr():
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                if "step" in st:
                    steps.append(int(float(st["step"])))
                self.optimizer.state[p] = {"step": torch.tensor(float(steps[-1] if steps else 0)), "exp_avg": m, "exp_avg_sq": v}
            t = max(steps) if steps else int(getattr(self.scheduler, "last_epoch", 0))
            self.opt_step = t
            base = getattr(self.scheduler, "base_lrs", [self.hyper_host["lr"]])[0]
            self.hyper_host["lr"] = float(base)
            self.hyper[0] = float(base)
            for p, g in zip(self.params, self.layout.views(self.G)):
                p.grad = g
            for nm, buf in zip(self._named, self.packed):
                ops.pack_weights({k: v.detach() for k, v in nm.items()}, self.dev, out=buf)
            torch.cuda.synchronize(self.dev)
