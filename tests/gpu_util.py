import numpy as np
import torch


def dev():
    return torch.device("cuda", 0)


def packed_net(state_dict):
    from nerf_dbr_b200.host import ops
    return ops.pack_weights({k: v.to(dev()) for k, v in state_dict.items()}, dev())


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else -10.0 * np.log10(mse)


class Watchdog:
    """Arms the tensor-core kernel's bounded-wait word so a barrier bug fails instead of hanging."""

    def __init__(self):
        from nerf_dbr_b200.host import lib as L
        self.word = torch.zeros(1, dtype=torch.int32, device=dev())
        self.lib = L.load_library()

    def __enter__(self):
        import ctypes
        self.lib.nerf_b200_set_watchdog_word(ctypes.c_void_p(self.word.data_ptr()))
        return self

    def __exit__(self, *exc):
        self.lib.nerf_b200_set_watchdog_word(None)
        return False
