"""Tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers and the
current CUDA stream.  Every function requires CUDA tensors (fp32, contiguous) -- no CPU path."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import lib as L
from .model import STATE_ORDER


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise L.NerfB200Error(name, -101, "expected a CUDA tensor (nerf_dbr_b200 has no CPU path)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _c2w(pose: torch.Tensor):
    flat = pose.detach().to("cpu", torch.float32).reshape(-1)
    if flat.numel() != 16:
        raise ValueError("camera_pose must be 4x4")
    return (ctypes.c_float * 16)(*flat.tolist())


def params_struct(tensors: dict) -> L.Params:
    """nerf_b200_params from a name -> CUDA tensor mapping (state-dict names)."""
    p = L.Params()
    for i in range(8):
        p.layer_w[i] = tensors[f"layers.{i}.weight"].data_ptr()
        p.layer_b[i] = tensors[f"layers.{i}.bias"].data_ptr()
    p.density_w = tensors["density_head.weight"].data_ptr()
    p.density_b = tensors["density_head.bias"].data_ptr()
    p.color0_w = tensors["color_layers.0.weight"].data_ptr()
    p.color0_b = tensors["color_layers.0.bias"].data_ptr()
    p.color1_w = tensors["color_layers.1.weight"].data_ptr()
    p.color1_b = tensors["color_layers.1.bias"].data_ptr()
    return p


def pack_weights(model_or_state, device: Optional[torch.device] = None,
                 out: Optional[torch.Tensor] = None, what: int = L.PACK_ALL) -> torch.Tensor:
    """Pack one network's 22 tensors (nn.Module or state dict) into the kernel layout.  ``what``: the optional parts
    (``L.PACK_*``; default all) -- a training loop that re-packs every step writes only what its mode reads."""
    lib = L.load_library()
    sd = model_or_state.state_dict() if hasattr(model_or_state, "state_dict") else model_or_state
    if device is None:
        device = next(iter(sd.values())).device
        if device.type != "cuda":
            device = torch.device("cuda", torch.cuda.current_device())
    keep = {k: sd[k].detach().to(device, torch.float32).contiguous() for k in STATE_ORDER}
    nbytes = lib.nerf_b200_packed_bytes()
    if out is None:
        out = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-out.data_ptr()) % 1024
    view = out[off:off + nbytes]
    p = params_struct(keep)
    with torch.cuda.device(device):
        L.check("nerf_b200_pack_weights_ex", lib.nerf_b200_pack_weights_ex(ctypes.byref(p), _ptr(view), what, _stream()))
    view._keepalive = (out, keep)      # the pack kernel is asynchronous
    return view


def calibration_points(n_views: int = 4, width: int = 40, height: int = 30, n_samples: int = 32, near: float = 2.0,
                       far: float = 6.0, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(positions, directions) [n,3] of sample points along the rays of a few orbit views: what the FP8 mode's
    activation scales are calibrated on when the caller has no better sample of the scene."""
    from .synthetic import orbit_pose
    device = device or torch.device("cuda", torch.cuda.current_device())
    pos, dirs = [], []
    for i in range(n_views):
        ro, rd = generate_rays(orbit_pose(i, n_views), width, height, device=device)
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        pts, _ = sample_points(ro, rd, n_samples, near, far)
        pos.append(pts.reshape(-1, 3))
        dirs.append(rd[:, None, :].expand(-1, n_samples, -1).reshape(-1, 3))
    return torch.cat(pos).contiguous(), torch.cat(dirs).contiguous()


def pack_weights_fp8(model_or_state, device: Optional[torch.device] = None, calib=None, packed: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Quantise one network for FP8 mode (nerf_b200_pack_weights_fp8): e4m3 weights with per-row power-of-two scales,
    activation scales calibrated by running the fp32 network on ``calib`` = (positions, directions) [n,3] (default:
    ``calibration_points``).  Returns the 1024-aligned FP8 buffer (its fp32 head holds the scales, fp8_layout.h)."""
    lib = L.load_library()
    sd = model_or_state.state_dict() if hasattr(model_or_state, "state_dict") else model_or_state
    if device is None:
        device = next(iter(sd.values())).device
        if device.type != "cuda":
            device = torch.device("cuda", torch.cuda.current_device())
    keep = {k: sd[k].detach().to(device, torch.float32).contiguous() for k in STATE_ORDER}
    if packed is None:
        packed = pack_weights(keep, device)
    if calib is None:
        calib = calibration_points(device=device)
    pos, dirs = _dev(calib[0], "pack_weights_fp8"), _dev(calib[1], "pack_weights_fp8")
    nbytes = lib.nerf_b200_packed_fp8_bytes()
    buf = torch.zeros(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 1024
    view = buf[off:off + nbytes]
    p = params_struct(keep)
    with torch.cuda.device(device):
        L.check("nerf_b200_pack_weights_fp8", lib.nerf_b200_pack_weights_fp8(ctypes.byref(p), _ptr(packed), _ptr(pos), _ptr(dirs),
                                                                             pos.shape[0], _ptr(view), _stream()))
    view._keepalive = (buf, keep, packed, pos, dirs)
    return view


def generate_rays(pose: torch.Tensor, width: int, height: int, focal: float = 800.0, row0: int = 0,
                  n_rows: Optional[int] = None, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = L.load_library()
    n_rows = height - row0 if n_rows is None else n_rows
    device = device or torch.device("cuda", torch.cuda.current_device())
    ro = torch.empty(n_rows, width, 3, device=device)
    rd = torch.empty(n_rows, width, 3, device=device)
    with torch.cuda.device(device):
        L.check("nerf_b200_generate_rays", lib.nerf_b200_generate_rays(
            _c2w(pose), width, height, focal, row0, n_rows, _ptr(ro), _ptr(rd), _stream()))
    return ro, rd


def sample_points(rays_o: torch.Tensor, rays_d: torch.Tensor, n_samples: int, near: float = 2.0,
                  far: float = 6.0, t_rand: Optional[torch.Tensor] = None):
    lib = L.load_library()
    ro, rd = _dev(rays_o, "sample_points"), _dev(rays_d, "sample_points")
    n = ro.shape[0]
    pts = torch.empty(n, n_samples, 3, device=ro.device)
    z = torch.empty(n, n_samples, device=ro.device)
    tr = None if t_rand is None else _dev(t_rand, "sample_points")
    with torch.cuda.device(ro.device):
        L.check("nerf_b200_sample_points", lib.nerf_b200_sample_points(
            _ptr(ro), _ptr(rd), n, n_samples, near, far, _ptr(tr), _ptr(pts), _ptr(z), _stream()))
    return pts, z


def importance_sample(rays_o, rays_d, z_vals, weights, u):
    lib = L.load_library()
    ro, rd, z, w, u = (_dev(t, "importance_sample") for t in (rays_o, rays_d, z_vals, weights, u))
    n, s = z.shape
    k = u.shape[1]
    idx = torch.empty(n, k, dtype=torch.int64, device=z.device)
    zn = torch.empty(n, k, device=z.device)
    pts = torch.empty(n, k, 3, device=z.device)
    with torch.cuda.device(z.device):
        L.check("nerf_b200_importance_sample", lib.nerf_b200_importance_sample(
            _ptr(ro), _ptr(rd), _ptr(z), _ptr(w), _ptr(u), n, s, k, _ptr(idx), _ptr(zn), _ptr(pts), _stream()))
    return pts, zn, idx


def positional_encoding(x: torch.Tensor, n_freq: int) -> torch.Tensor:
    lib = L.load_library()
    x = _dev(x, "positional_encoding")
    out = torch.empty(x.shape[0], 3 + 6 * n_freq, device=x.device)
    if x.shape[0] == 0:
        return out
    with torch.cuda.device(x.device):
        L.check("nerf_b200_positional_encoding", lib.nerf_b200_positional_encoding(
            _ptr(x), x.shape[0], n_freq, _ptr(out), _stream()))
    return out


def query_network(packed: torch.Tensor, positions: torch.Tensor, directions: torch.Tensor,
                  mode: int = L.FP32):
    lib = L.load_library()
    pos, dirs = _dev(positions, "query_network"), _dev(directions, "query_network")
    n = pos.shape[0]
    sigma = torch.empty(n, 1, device=pos.device)
    rgb = torch.empty(n, 3, device=pos.device)
    with torch.cuda.device(pos.device):
        L.check("nerf_b200_query_network", lib.nerf_b200_query_network(
            _ptr(packed), _ptr(pos), _ptr(dirs), n, mode, _ptr(sigma), _ptr(rgb), _stream()))
    return sigma, rgb


def composite(sigma, rgb, z_vals, rays_d, want_aux: bool = False):
    lib = L.load_library()
    sg, col, z, rd = (_dev(t, "composite") for t in (sigma, rgb, z_vals, rays_d))
    n, s = z.shape
    rgb_map = torch.empty(n, 3, device=z.device)
    depth = torch.empty(n, device=z.device)
    acc = torch.empty(n, device=z.device) if want_aux else None
    wts = torch.empty(n, s, device=z.device) if want_aux else None
    with torch.cuda.device(z.device):
        L.check("nerf_b200_composite", lib.nerf_b200_composite(
            _ptr(sg), _ptr(col), _ptr(z), _ptr(rd), n, s, _ptr(rgb_map), _ptr(depth), _ptr(acc), _ptr(wts), _stream()))
    return (rgb_map, depth, acc, wts) if want_aux else (rgb_map, depth)


def render_image(packed: torch.Tensor, pose: torch.Tensor, width: int, height: int, n_samples: int,
                 mode: int = L.BF16, focal: float = 800.0, near: float = 2.0, far: float = 6.0,
                 row0: int = 0, n_rows: Optional[int] = None, out_rgb=None, out_depth=None):
    lib = L.load_library()
    n_rows = height - row0 if n_rows is None else n_rows
    dev = packed.device
    rgb = out_rgb if out_rgb is not None else torch.empty(n_rows, width, 3, device=dev)
    depth = out_depth if out_depth is not None else torch.empty(n_rows, width, device=dev)
    with torch.cuda.device(dev):
        if mode == L.FP8:                            # `packed` is the FP8 buffer (pack_weights_fp8)
            L.check("nerf_b200_render_image_fp8", lib.nerf_b200_render_image_fp8(
                _ptr(packed), _c2w(pose), width, height, focal, near, far, n_samples, row0, n_rows, _ptr(rgb), _ptr(depth), _stream()))
            return rgb, depth
        L.check("nerf_b200_render_image", lib.nerf_b200_render_image(
            _ptr(packed), _c2w(pose), width, height, focal, near, far, n_samples, row0, n_rows, mode,
            _ptr(rgb), _ptr(depth), _stream()))
    return rgb, depth


def render_rays(packed: torch.Tensor, rays_o, rays_d, n_samples: int, mode: int = L.BF16,
                near: float = 2.0, far: float = 6.0, t_rand=None, want_acc: bool = False, z_vals=None,
                want_weights: bool = False):
    """Fused render of a ray batch.  ``z_vals`` [R,S] (ascending) overrides the uniform/stratified depths;
    ``want_weights`` also returns the per-sample compositing weights [R,S]."""
    lib = L.load_library()
    ro, rd = _dev(rays_o, "render_rays"), _dev(rays_d, "render_rays")
    n = ro.shape[0]
    rgb = torch.empty(n, 3, device=ro.device)
    depth = torch.empty(n, device=ro.device)
    acc = torch.empty(n, device=ro.device) if want_acc else None
    tr = None if t_rand is None else _dev(t_rand, "render_rays")
    zv = None if z_vals is None else _dev(z_vals, "render_rays")
    if zv is not None and tuple(zv.shape) != (n, n_samples):
        raise ValueError("z_vals must be [n_rays, n_samples]")
    wts = torch.empty(n, n_samples, device=ro.device) if want_weights else None
    with torch.cuda.device(ro.device):
        if mode == L.FP8:                            # `packed` is the FP8 buffer (pack_weights_fp8)
            L.check("nerf_b200_render_rays_fp8", lib.nerf_b200_render_rays_fp8(
                _ptr(packed), _ptr(ro), _ptr(rd), n, n_samples, near, far, _ptr(tr), _ptr(zv), _ptr(rgb), _ptr(depth), _ptr(acc),
                _ptr(wts), _stream()))
        else:
            L.check("nerf_b200_render_rays_ex", lib.nerf_b200_render_rays_ex(
                _ptr(packed), _ptr(ro), _ptr(rd), n, n_samples, near, far, _ptr(tr), _ptr(zv), mode,
                _ptr(rgb), _ptr(depth), _ptr(acc), _ptr(wts), _stream()))
    out = (rgb, depth) + ((acc,) if want_acc else ()) + ((wts,) if want_weights else ())
    return out


def merge_samples(z_sorted, z_new) -> torch.Tensor:
    lib = L.load_library()
    a, b = _dev(z_sorted, "merge_samples"), _dev(z_new, "merge_samples")
    out = torch.empty(a.shape[0], a.shape[1] + b.shape[1], device=a.device)
    with torch.cuda.device(a.device):
        L.check("nerf_b200_merge_samples", lib.nerf_b200_merge_samples(
            _ptr(a), _ptr(b), a.shape[0], a.shape[1], b.shape[1], _ptr(out), _stream()))
    return out


def hierarchical_samples(weights, n_new: int, near: float = 2.0, far: float = 6.0, t_rand=None, u=None, seed: int = 0) -> torch.Tensor:
    """Sorted union [R, S + n_new] of the coarse depths and the inverse-CDF samples drawn from ``weights`` [R,S] -- one
    kernel (nerf_b200_hierarchical_samples).  ``u`` [R,n_new]: the uniforms (None: drawn in the kernel from ``seed``)."""
    lib = L.load_library()
    w = _dev(weights, "hierarchical_samples")
    n, s = w.shape
    tr = None if t_rand is None else _dev(t_rand, "hierarchical_samples")
    uu = None if u is None else _dev(u, "hierarchical_samples")
    if uu is not None and tuple(uu.shape) != (n, n_new):
        raise ValueError("u must be [n_rays, n_new]")
    out = torch.empty(n, s + n_new, device=w.device)
    with torch.cuda.device(w.device):
        L.check("nerf_b200_hierarchical_samples", lib.nerf_b200_hierarchical_samples(
            _ptr(w), n, s, n_new, near, far, _ptr(tr), _ptr(uu), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(out), _stream()))
    return out


def render_hierarchical(coarse_packed, fine_packed, rays_o, rays_d, n_coarse: int, n_importance: int,
                        mode: int = L.BF16, near: float = 2.0, far: float = 6.0, u=None, t_rand=None, seed: int = 0):
    """Coarse pass -> inverse-CDF importance samples -> fine pass on the sorted union (BASELINE.json configs[4]:
    128 coarse + 128 importance): three launches -- fused render (weights out), fused sampling, fused render.
    ``u`` [R,n_importance] uniforms (None: drawn inside the sampling kernel from ``seed``).  Returns
    (rgb_fine, depth_fine, rgb_coarse, z_union)."""
    ro, rd = _dev(rays_o, "render_hierarchical"), _dev(rays_d, "render_hierarchical")
    rgb_c, _, wts = render_rays(coarse_packed, ro, rd, n_coarse, mode, near, far, t_rand, want_weights=True)
    z_all = hierarchical_samples(wts, n_importance, near, far, t_rand, u, seed)
    rgb_f, depth_f = render_rays(fine_packed, ro, rd, n_coarse + n_importance, mode, near, far, z_vals=z_all)
    return rgb_f, depth_f, rgb_c, z_all


_workspaces = {}


class TrainPass:
    """One network's forward + backward over one ray batch, prepared once and run in one or two phases
    (``nerf_b200_train_fwd_bwd_ex``): ``run(L.TRAIN_ALL)``, or ``run(L.TRAIN_ACTIVATIONS)`` followed -- possibly
    on another stream, ordered after it -- by ``run(L.TRAIN_WEIGHT_GRADS)``.  ``slot`` selects the cached
    workspace: passes whose phases overlap in time need different slots."""

    def __init__(self, model, rays_o, rays_d, target, n_samples: int, t_rand=None, n_rays_global: Optional[int] = None,
                 near: float = 2.0, far: float = 6.0, mode: int = L.FP32, want_rgb: bool = True, slot: int = 0, packed=None,
                 grad_out: Optional[dict] = None, loss_sum: Optional[torch.Tensor] = None):
        """``model``: an ``nn.Module`` with the reference's parameter names (gradients accumulate into ``p.grad``), or
        a name -> tensor dict together with ``grad_out`` (name -> tensor the gradients accumulate into)."""
        self.lib = L.load_library()
        self.ro, self.rd, self.tg = (_dev(x, "train_fwd_bwd") for x in (rays_o, rays_d, target))
        self.n, self.n_samples, self.near, self.far, self.mode = self.ro.shape[0], n_samples, near, far, mode
        self.n_global = self.n if n_rays_global is None else n_rays_global
        dev = self.dev = self.ro.device
        named = model if isinstance(model, dict) else dict(model.named_parameters())
        if isinstance(model, dict) and grad_out is None:
            raise ValueError("a parameter dict needs grad_out")
        params, grads = {}, {}
        for k in STATE_ORDER:
            p = named[k]
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise L.NerfB200Error("train_fwd_bwd", -101, f"parameter {k} must be a contiguous fp32 CUDA tensor")
            if grad_out is not None:
                gk = grad_out[k]
                if not gk.is_cuda or gk.dtype != torch.float32 or not gk.is_contiguous() or gk.shape != p.shape:
                    raise L.NerfB200Error("train_fwd_bwd", -101, f"grad_out[{k}] must be a contiguous fp32 CUDA tensor like the parameter")
                params[k], grads[k] = p.detach(), gk
                continue
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            params[k], grads[k] = p.detach(), p.grad
        self.packed = pack_weights(params, dev) if packed is None else packed
        nbytes = self.lib.nerf_b200_train_workspace_bytes(self.n, n_samples)
        ws = _workspaces.get((dev, slot))
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _workspaces[(dev, slot)] = ws
        self.ws = ws
        # the kernels ACCUMULATE the sum of squared errors here: a caller-provided slot (e.g. the tail of a gradient
        # bucket, so that the loss rides the gradient exchange) or a fresh zero
        self.loss_sum = torch.zeros(1, device=dev) if loss_sum is None else loss_sum
        self.rgb = torch.empty(self.n, 3, device=dev) if want_rgb else None
        self.tr = None if t_rand is None else _dev(t_rand, "train_fwd_bwd")
        self.ps, self.gs = params_struct(params), params_struct(grads)
        self._keep = (params, grads)

    def run(self, phases: int = L.TRAIN_ALL, sm_limit: int = 0) -> "TrainPass":
        with torch.cuda.device(self.dev):
            L.check("nerf_b200_train_fwd_bwd_ex", self.lib.nerf_b200_train_fwd_bwd_ex(
                _ptr(self.packed), ctypes.byref(self.ps), ctypes.byref(self.gs), _ptr(self.ro), _ptr(self.rd), _ptr(self.tg),
                self.n, self.n_samples, self.near, self.far, _ptr(self.tr), self.n_global, self.mode, _ptr(self.ws),
                _ptr(self.loss_sum), _ptr(self.rgb), phases, sm_limit, _stream()))
        return self

    @property
    def loss(self):
        """This pass's loss term: sum of squared errors over its rays with the global normaliser."""
        return self.loss_sum[0] / (3.0 * self.n_global)


def train_fwd_bwd(model, rays_o, rays_d, target, n_samples: int, t_rand=None, n_rays_global: Optional[int] = None,
                  near: float = 2.0, far: float = 6.0, mode: int = L.FP32, want_rgb: bool = True):
    """Forward + backward of one network's loss term mean((C - target)^2) (reference
    NeRFTrainer.train_step, trainer.py:117-126).  ``model`` is a CUDA ``NeRFModel`` (or any module with the
    reference's parameter names); d loss/d params is accumulated into ``p.grad`` (created as zeros when
    None) -- scaled for ``n_rays_global`` rays so data-parallel ranks can all-reduce-sum.  Returns
    (loss_term as a 0-dim tensor computed over this call's rays with the global normaliser, rgb [R,3])."""
    tp = TrainPass(model, rays_o, rays_d, target, n_samples, t_rand, n_rays_global, near, far, mode, want_rgb).run()
    return tp.loss, tp.rgb


def composite_white(rgba: torch.Tensor) -> torch.Tensor:
    """[..., 4] uint8 RGBA (CUDA) -> [..., 3] fp32 RGB on a white background, the reference loader's float64 arithmetic
    (src/data/loader.py:46-54) bit for bit."""
    lib = L.load_library()
    if not rgba.is_cuda or rgba.dtype != torch.uint8 or rgba.shape[-1] != 4:
        raise L.NerfB200Error("composite_white", -101, "expected a CUDA uint8 tensor [..., 4] (nerf_dbr_b200 has no CPU path)")
    rgba = rgba.contiguous()
    out = torch.empty(*rgba.shape[:-1], 3, device=rgba.device)
    if rgba.numel():
        with torch.cuda.device(rgba.device):
            L.check("nerf_b200_composite_white", lib.nerf_b200_composite_white(_ptr(rgba), rgba.numel() // 4, _ptr(out), _stream()))
    return out


def ray_batch(pose: torch.Tensor, width: int, height: int, focal: float, pixel_index: torch.Tensor,
              image: Optional[torch.Tensor] = None):
    """(rays_o [n,3], rays_d [n,3], target [n,3] or None) for the selected pixels: the bits of
    ``_get_rays(pose)[index]`` and ``image.reshape(-1, 3)[index]`` (trainer.py:100-118) without the full-image tensors."""
    lib = L.load_library()
    if not pixel_index.is_cuda or pixel_index.dtype != torch.int64:
        raise L.NerfB200Error("ray_batch", -101, "pixel_index must be a CUDA int64 tensor (nerf_dbr_b200 has no CPU path)")
    idx = pixel_index.contiguous()
    n, dev = idx.numel(), idx.device
    ro, rd = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
    tg = None
    if image is not None:
        image = _dev(image, "ray_batch")
        if image.numel() != width * height * 3:
            raise ValueError("image must hold height * width * 3 values")
        tg = torch.empty(n, 3, device=dev)
    if n:
        with torch.cuda.device(dev):
            L.check("nerf_b200_ray_batch", lib.nerf_b200_ray_batch(_c2w(pose), width, height, focal, _ptr(idx), n, _ptr(image),
                                                                   _ptr(ro), _ptr(rd), _ptr(tg), _stream()))
    return ro, rd, tg


def launch_count() -> int:
    return int(L.load_library().nerf_b200_launch_count())
