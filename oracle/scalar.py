"""ctypes binding of oracle/scalar_oracle.c (TEST INFRASTRUCTURE; see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libscalar_oracle.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_STATE_ORDER = [f"layers.{i}.{p}" for i in range(8) for p in ("weight", "bias")] + [
    "density_head.weight", "density_head.bias",
    "color_layers.0.weight", "color_layers.0.bias",
    "color_layers.1.weight", "color_layers.1.bias"]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "scalar_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(_f)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def linspace(start: float, end: float, n: int) -> np.ndarray:
    out = np.empty(n, np.float32)
    lib().so_linspace(ctypes.c_float(start), ctypes.c_float(end), n, _p(out))
    return out


def z_vals(n_samples: int, near: float = 2.0, far: float = 6.0) -> np.ndarray:
    out = np.empty(n_samples, np.float32)
    lib().so_z_vals(n_samples, ctypes.c_float(near), ctypes.c_float(far), _p(out))
    return out


def camera_rays(c2w, width: int, height: int, focal: float = 800.0):
    c = _f32(c2w).reshape(16)
    ro = np.empty((height * width, 3), np.float32)
    rd = np.empty((height * width, 3), np.float32)
    lib().so_camera_rays(_p(c), width, height, ctypes.c_float(focal), _p(ro), _p(rd))
    return ro, rd


def points(rays_o, rays_d, z, per_ray: bool = False) -> np.ndarray:
    ro, rd, z = _f32(rays_o), _f32(rays_d), _f32(z)
    n_rays = ro.shape[0]
    n_samples = z.shape[-1]
    out = np.empty((n_rays, n_samples, 3), np.float32)
    lib().so_points(_p(ro), _p(rd), _p(z), n_rays, n_samples, n_samples if per_ray else 0, _p(out))
    return out


def stratified(z, t_rand) -> np.ndarray:
    z, t = _f32(z), _f32(t_rand)
    out = np.empty_like(t)
    lib().so_stratified(_p(z), _p(t), t.shape[0], t.shape[1], _p(out))
    return out


def encode(x, n_freq: int) -> np.ndarray:
    x = _f32(x)
    out = np.empty((x.shape[0], 3 + 6 * n_freq), np.float32)
    lib().so_encode(_p(x), x.shape[0], n_freq, _p(out))
    return out


def encode_args(x, n_freq: int) -> np.ndarray:
    x = _f32(x)
    out = np.empty((x.shape[0], n_freq, 3), np.float32)
    lib().so_encode_args(_p(x), x.shape[0], n_freq, _p(out))
    return out


def mlp(weights, pos, dirs):
    """weights: mapping name -> array (state-dict names)."""
    arrs = [_f32(weights[k].detach().numpy() if hasattr(weights[k], "detach") else weights[k])
            for k in _STATE_ORDER]
    ptrs = (_f * len(arrs))(*[_p(a) for a in arrs])
    pos, dirs = _f32(pos), _f32(dirs)
    n = pos.shape[0]
    sigma = np.empty((n, 1), np.float32)
    rgb = np.empty((n, 3), np.float32)
    lib().so_mlp(ptrs, _p(pos), _p(dirs), n, _p(sigma), _p(rgb))
    return sigma, rgb


def composite(sigma, rgb, z, rays_d):
    sigma, rgb, z, rays_d = _f32(sigma), _f32(rgb), _f32(z), _f32(rays_d)
    n_rays, n_samples = z.shape
    rgb_map = np.empty((n_rays, 3), np.float32)
    depth = np.empty(n_rays, np.float32)
    acc = np.empty(n_rays, np.float32)
    wts = np.empty((n_rays, n_samples), np.float32)
    lib().so_composite(_p(sigma.reshape(-1)), _p(rgb.reshape(-1, 3)), _p(z), _p(rays_d),
                       n_rays, n_samples, _p(rgb_map), _p(depth), _p(acc), _p(wts))
    return rgb_map, depth, acc, wts


def importance(z, weights, u):
    z, w, u = _f32(z), _f32(weights), _f32(u)
    n_rays, n_samples = z.shape
    n_new = u.shape[1]
    idx = np.empty((n_rays, n_new), np.int64)
    z_new = np.empty((n_rays, n_new), np.float32)
    rc = lib().so_importance(_p(z), _p(w), _p(u), n_rays, n_samples, n_new,
                             idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _p(z_new))
    if rc != 0:
        raise ValueError("so_importance: n_samples must be a multiple of 32")
    return idx, z_new
