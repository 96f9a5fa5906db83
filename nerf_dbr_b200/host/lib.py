"""ctypes binding of libnerf_b200.so (include/nerf_b200.h).  Pointers and sizes only: tensors are
passed as ``tensor.data_ptr()``, the stream as ``torch.cuda.current_stream().cuda_stream``."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
_CSRC = os.path.join(_PKG, "csrc")
# NERF_B200_LIB: load another build of the same ABI instead (A/B measurements of one kernel change on one box)
LIB_PATH = os.environ.get("NERF_B200_LIB") or os.path.join(_PKG, "libnerf_b200.so")

FP32, BF16, BF16X3 = 0, 1, 2
FP8 = 3                         # host-side selector only: FP8 mode has its own entry points and weight buffer
TRAIN_ACTIVATIONS, TRAIN_WEIGHT_GRADS, TRAIN_ALL = 1, 2, 3      # nerf_b200_train_fwd_bwd_ex phases
PACK_FP32_MATRICES, PACK_BF16_LO, PACK_DGRAD, PACK_ALL = 1, 2, 4, 7   # nerf_b200_pack_weights_ex parts

_lib = None


class NerfB200Error(RuntimeError):
    """A nerf_b200_* call returned non-zero (argument error < 0, cudaError_t > 0)."""

    def __init__(self, fn: str, code: int, text: str):
        super().__init__(f"{fn} failed with code {code}: {text}")
        self.code = code


class Params(ctypes.Structure):
    """nerf_b200_params: the 22 device pointers of one network, state-dict order."""
    _fields_ = [("layer_w", c_void_p * 8), ("layer_b", c_void_p * 8),
                ("density_w", c_void_p), ("density_b", c_void_p),
                ("color0_w", c_void_p), ("color0_b", c_void_p),
                ("color1_w", c_void_p), ("color1_b", c_void_p)]


DP_MAX_WORLD, DP_CTL_BYTES, DP_STATE_BYTES, DP_STATE_OPT_STEP = 16, 1024, 2048, 12


class DP(ctypes.Structure):
    """nerf_b200_dp: the data-parallel exchange's addresses (include/nerf_b200.h)."""
    _fields_ = [("rank", c_int), ("world", c_int), ("n", c_int64), ("n_opt", c_int64),
                ("peer", c_void_p * DP_MAX_WORLD), ("multicast", c_void_p), ("state", c_void_p),
                ("emulate_sequential", c_int)]


# every symbol include/nerf_b200.h declares: name -> (restype, argtypes)
_F = ctypes.POINTER(c_float)
PROTOTYPES = {
    "nerf_b200_abi_version": (c_int, []),
    "nerf_b200_error_string": (ctypes.c_char_p, [c_int]),
    "nerf_b200_packed_bytes": (c_size_t, []),
    "nerf_b200_pack_weights": (c_int, [ctypes.POINTER(Params), c_void_p, c_void_p]),
    "nerf_b200_pack_weights_ex": (c_int, [ctypes.POINTER(Params), c_void_p, c_int, c_void_p]),
    "nerf_b200_generate_rays": (c_int, [_F, c_int, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "nerf_b200_sample_points": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "nerf_b200_importance_sample": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                            c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nerf_b200_positional_encoding": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "nerf_b200_query_network": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "nerf_b200_composite": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "nerf_b200_render_image": (c_int, [c_void_p, _F, c_int, c_int, c_float, c_float, c_float, c_int, c_int,
                                       c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "nerf_b200_render_rays": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p,
                                      c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nerf_b200_render_rays_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p,
                                         c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nerf_b200_merge_samples": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "nerf_b200_packed_fp8_bytes": (c_size_t, []),
    "nerf_b200_pack_weights_fp8": (c_int, [ctypes.POINTER(Params), c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "nerf_b200_render_image_fp8": (c_int, [c_void_p, _F, c_int, c_int, c_float, c_float, c_float, c_int, c_int, c_int, c_void_p,
                                           c_void_p, c_void_p]),
    "nerf_b200_render_rays_fp8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nerf_b200_hierarchical_samples": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_uint64,
                                               c_void_p, c_void_p]),
    "nerf_b200_composite_white": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "nerf_b200_ray_batch": (c_int, [_F, c_int, c_int, c_float, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nerf_b200_train_workspace_bytes": (c_size_t, [c_int, c_int]),
    "nerf_b200_train_fwd_bwd": (c_int, [c_void_p, ctypes.POINTER(Params), ctypes.POINTER(Params), c_void_p,
                                        c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_int,
                                        c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nerf_b200_train_fwd_bwd_ex": (c_int, [c_void_p, ctypes.POINTER(Params), ctypes.POINTER(Params), c_void_p,
                                           c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_int,
                                           c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "nerf_b200_dp_bytes": (c_size_t, [c_int64]),
    "nerf_b200_dp_reduce": (c_int, [ctypes.POINTER(DP), c_void_p]),
    "nerf_b200_dp_adam_step": (c_int, [ctypes.POINTER(DP), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "nerf_b200_launch_count": (c_uint64, []),
    "nerf_b200_set_watchdog_word": (None, [c_void_p]),
    "nerf_b200_set_trace_buffer": (None, [c_void_p]),
}


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", _CSRC, "-j8"] + (["-B"] if force else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libnerf_b200.so failed")
    return LIB_PATH


def load_library() -> ctypes.CDLL:
    """Load the in-tree CUDA library; fails loudly when it is missing (there is no CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NerfB200Error("load_library", -100,
                            f"{LIB_PATH} not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a); nerf_dbr_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        if os.environ.get("NERF_B200_LIB") and not hasattr(lib, name):
            continue                     # an older build loaded for an A/B run: its missing entry points stay unbound
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(fn: str, code: int) -> None:
    if code != 0:
        text = load_library().nerf_b200_error_string(code).decode()
        raise NerfB200Error(fn, code, text)
