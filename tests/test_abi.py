"""CPU: the C-ABI library builds, loads and exports every symbol include/nerf_b200.h declares; the
host package fails loudly without CUDA (no CPU fallback).  No compute calls here."""
import os
import re

import pytest
import torch

from conftest import REPO


def header_symbols():
    text = open(os.path.join(REPO, "include", "nerf_b200.h")).read()
    return sorted(set(re.findall(r"NERF_B200_API[^;]*?\b(nerf_b200_\w+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from nerf_dbr_b200.host import lib as L
    L.build_library()
    import ctypes
    so = ctypes.CDLL(L.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(so, s), f"{s} declared in include/nerf_b200.h but not exported"
    assert sorted(L.PROTOTYPES) == syms, "ctypes prototypes out of sync with the header"
    lib = L.load_library()
    assert lib.nerf_b200_abi_version() == 1
    assert lib.nerf_b200_packed_bytes() % 1024 == 0
    assert lib.nerf_b200_error_string(-2).decode().startswith("shape not supported")


def test_sass_is_blackwell_native():
    """tcgen05.mma / tcgen05.ld / bulk-copy (TMA engine) must be in the binary (B200_PROFILING.md)."""
    import shutil
    import subprocess
    from nerf_dbr_b200.host import lib as L
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    L.build_library()
    sass = subprocess.run(["cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True).stdout
    # tcgen05 MMA, TMEM loads, bulk (TMA-engine) copies, and the cluster forms of the paired weight stream
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP", "UBLKCP.S.G.MULTICAST", "UTCBAR.MULTICAST", "UCGABAR_ARV"):
        assert mnemonic in sass, mnemonic
    assert "HMMA." not in sass.replace("UTCHMMA", "")      # no legacy mma.sync path


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import nerf_dbr_b200 as nb
    from nerf_dbr_b200.host import ops
    with pytest.raises(RuntimeError):
        nb.B200Renderer()                      # same behaviour as PyTorchCUDARenderer (pytorch_renderers.py:176-178)
    with pytest.raises(nb.NerfB200Error):
        ops.positional_encoding(torch.zeros(4, 3), 10)
    with pytest.raises(nb.NerfB200Error):
        ops.composite(torch.zeros(2, 4), torch.zeros(2, 4, 3), torch.zeros(2, 4), torch.zeros(2, 3))


def test_model_mirror_matches_reference_names(checkpoints):
    """The host NeRFModel mirror loads reference-format state dicts and reproduces the seeded init."""
    import nerf_dbr_b200 as nb
    torch.manual_seed(2)
    coarse, fine = nb.NeRFModel(), nb.NeRFModel()
    for k, v in fine.state_dict().items():
        assert torch.equal(v, checkpoints["rand2"]["fine_model"][k]), k
    for k, v in coarse.state_dict().items():
        assert torch.equal(v, checkpoints["rand2"]["coarse_model"][k]), k
    m = nb.NeRFModel()
    missing = m.load_state_dict(checkpoints["lego"]["fine_model"])
    assert not missing.missing_keys and not missing.unexpected_keys
    assert sum(p.numel() for p in m.parameters()) == 530052


def test_argument_errors_return_codes_and_enqueue_nothing():
    """include/nerf_b200.h: argument errors return negative NERF_B200_E* codes before anything is enqueued -- so they
    can be exercised without a GPU.  (Pointers are never dereferenced on the host.)"""
    import ctypes
    from nerf_dbr_b200.host import lib as L
    lib = L.load_library()
    EINVAL, launches = -1, lib.nerf_b200_launch_count()
    null = None
    fake = ctypes.c_void_p(0x1000)
    assert lib.nerf_b200_query_network(null, fake, fake, 8, L.BF16, fake, fake, null) == EINVAL
    assert lib.nerf_b200_query_network(fake, fake, fake, 0, L.FP32, fake, fake, null) == EINVAL
    assert lib.nerf_b200_positional_encoding(null, 4, 10, fake, null) == EINVAL
    assert lib.nerf_b200_train_workspace_bytes(0, 64) == 0
    ps = L.Params()
    assert lib.nerf_b200_train_fwd_bwd_ex(fake, ctypes.byref(ps), ctypes.byref(ps), fake, fake, fake, 4, 64, 2.0, 6.0, null, 4,
                                          L.BF16, fake, fake, null, 7, 0, null) == EINVAL          # phases not in {1, 2, 3}
    assert lib.nerf_b200_train_fwd_bwd_ex(fake, ctypes.byref(ps), ctypes.byref(ps), fake, fake, fake, 4, 64, 2.0, 6.0, null, 4,
                                          L.BF16, fake, fake, null, L.TRAIN_ALL, -1, null) == EINVAL   # negative SM limit
    assert lib.nerf_b200_train_fwd_bwd_ex(fake, ctypes.byref(ps), ctypes.byref(ps), fake, fake, fake, 4, 64, 2.0, 6.0, null, 4,
                                          L.FP32, fake, fake, null, L.TRAIN_ACTIVATIONS, 0, null) == -2  # split phases: BF16 only
    # round-2 entry points
    assert lib.nerf_b200_pack_weights_ex(ctypes.byref(ps), fake, 8, null) == EINVAL                    # unknown part flag
    assert lib.nerf_b200_hierarchical_samples(null, 4, 64, 8, 2.0, 6.0, null, null, 0, fake, null) == EINVAL
    assert lib.nerf_b200_hierarchical_samples(fake, 4, 48, 8, 2.0, 6.0, null, null, 0, fake, null) == -2  # S % 32 != 0
    assert lib.nerf_b200_hierarchical_samples(fake, 4, 64, 2048, 2.0, 6.0, null, null, 0, fake, null) == -2
    assert lib.nerf_b200_composite_white(null, 16, fake, null) == EINVAL
    assert lib.nerf_b200_composite_white(ctypes.c_void_p(0x1002), 16, fake, null) == -3                # RGBA pixels are 4-byte units
    c2w = (ctypes.c_float * 16)(*([0.0] * 16))
    assert lib.nerf_b200_ray_batch(c2w, 8, 8, 800.0, fake, 4, null, fake, fake, fake, null) == EINVAL   # target without an image
    dp = L.DP()
    dp.rank, dp.world, dp.n, dp.n_opt = 0, 2, 1026, 1000                                                # n not a multiple of 4 * world
    dp.peer[0], dp.peer[1], dp.state = 0x1000, 0x2000, 0x3000
    assert lib.nerf_b200_dp_reduce(ctypes.byref(dp), null) == EINVAL
    dp.n, dp.rank = 1024, 2                                                                             # rank outside the world
    assert lib.nerf_b200_dp_reduce(ctypes.byref(dp), null) == EINVAL
    dp.rank, dp.peer[1] = 1, 0                                                                          # a peer mapping is missing
    assert lib.nerf_b200_dp_adam_step(ctypes.byref(dp), fake, fake, fake, fake, null, null) == EINVAL
    assert lib.nerf_b200_dp_bytes(1024) == L.DP_CTL_BYTES + 2 * 1024 * 4 and lib.nerf_b200_dp_bytes(0) == 0
    for code in (-1, -2, -3):
        assert lib.nerf_b200_error_string(code)
    assert lib.nerf_b200_launch_count() == launches
