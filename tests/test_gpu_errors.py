"""GPU: error behaviour of the C ABI with real device pointers (include/nerf_b200.h: negative codes, nothing enqueued);
the Python layer raises NerfB200Error, which the reference suite's per-view `except Exception`
(benchmark_suite.py:212-214) reports as FAILED like any other renderer."""
import pytest
import torch

from gpu_util import packed_net

pytestmark = pytest.mark.gpu


def test_unsupported_and_misaligned_arguments(checkpoints):
    import nerf_dbr_b200 as nb
    from nerf_dbr_b200.host import lib as L, ops
    net = packed_net(checkpoints["lego"]["fine_model"])
    pos, dirs = torch.zeros(8, 3, device="cuda"), torch.ones(8, 3, device="cuda")
    n0 = ops.launch_count()
    with pytest.raises(nb.NerfB200Error) as e:
        ops.query_network(net, pos, dirs, mode=L.BF16X3)                 # no split-precision variant of this entry point
    assert e.value.code == -2
    with pytest.raises(nb.NerfB200Error) as e:
        ops.query_network(net.view(torch.uint8)[16:], pos, dirs, mode=L.BF16)   # the packed network must be 1024-byte aligned
    assert e.value.code == -3
    with pytest.raises(nb.NerfB200Error) as e:
        ops.render_image(net, torch.eye(4), 8, 8, 40000, mode=L.BF16)    # more samples per ray than the kernel tiles
    assert e.value.code == -2
    with pytest.raises(nb.NerfB200Error):
        ops.render_image(net, torch.eye(4), 8, 8, 0, mode=L.BF16)
    with pytest.raises(nb.NerfB200Error):
        ops.query_network(net, pos.cpu(), dirs, mode=L.FP32)             # host tensor where a device tensor is required
    assert ops.launch_count() == n0                                       # nothing was enqueued by any of them
    sigma, rgb = ops.query_network(net, pos, dirs, mode=L.BF16)          # and the library is still usable
    assert torch.isfinite(sigma).all() and torch.isfinite(rgb).all()
