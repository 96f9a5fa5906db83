"""CPU, world_size 2, gloo: the host-side multi-GPU logic -- row-band partition, the one-bucket gradient
all-reduce used by the training step, and band gathering."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_dbr_b200.host.parallel import allreduce_sum_, broadcast_parameters_, gather_rows, ray_shard, row_band


def test_row_bands_tile_the_image_exactly():
    for height in (1, 7, 150, 600, 1200):
        for world in (1, 2, 3, 4, 8):
            bands = [row_band(r, world, height) for r in range(world)]
            assert bands[0][0] == 0 and sum(n for _, n in bands) == height
            for (a0, an), (b0, _) in zip(bands, bands[1:]):
                assert a0 + an == b0
            sizes = [n for _, n in bands]
            assert max(sizes) - min(sizes) <= 1
    assert ray_shard(1, 2, 4096) == (2048, 2048)
    with pytest.raises(ValueError):
        row_band(2, 2, 10)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        grads = [torch.randn(256, 63, generator=g), torch.randn(256, generator=g), torch.randn(3, 128, generator=g)]
        ref = [torch.randn(t.shape, generator=torch.Generator().manual_seed(100)) * 0 for t in grads]
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            for i, t in enumerate(grads):
                ref[i] += torch.randn(t.shape, generator=gr)
        loss = allreduce_sum_(grads, extra=torch.tensor(float(rank + 1)))
        ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(grads, ref)) and float(loss) == sum(range(1, world + 1))
        # bucket form (B200TrainStep: every gradient is a view of one flat buffer, the loss rides in its last slot)
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
        views = [flat[:6].view(2, 3), flat[6:9]]
        assert allreduce_sum_([flat]) is None
        tot = sum(range(1, world + 1))
        ok = ok and torch.equal(views[0], (torch.arange(6, dtype=torch.float32) * tot).view(2, 3)) and float(flat[-1]) == 9.0 * tot
        # render shard: every rank fills its band with its rank id; gather on rank 0
        height, width = 7, 5
        row0, n = row_band(rank, world, height)
        band = torch.full((n, width, 3), float(rank))
        img = gather_rows(band, height)
        if rank == 0:
            exp = torch.cat([torch.full((row_band(r, world, height)[1], width, 3), float(r)) for r in range(world)])
            ok = ok and torch.equal(img, exp)
        else:
            ok = ok and img is None
        # replica synchronisation (B200Trainer.__init__ / load_checkpoint): unseeded ranks construct different networks;
        # after the broadcast they hold rank 0's, and one "step" (same all-reduced gradient, same update) keeps them equal
        from nerf_dbr_b200.host.model import NeRFModel
        torch.manual_seed(1000 + rank)
        model = NeRFModel()
        params = list(model.parameters())
        broadcast_parameters_(params)
        torch.manual_seed(1000)
        ref_model = NeRFModel()
        ok = ok and all(torch.equal(a, b) for a, b in zip(params, ref_model.parameters()))
        grads = [torch.full_like(p, float(rank + 1)) for p in params]
        allreduce_sum_(grads)
        with torch.no_grad():
            for p, gr in zip(params, grads):
                p.add_(gr, alpha=-1e-3)
        digest = torch.stack([p.detach().double().sum() for p in params]).sum().reshape(1)
        both = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(both, digest)
        ok = ok and all(torch.equal(b, both[0]) for b in both)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_allreduce_and_gather_world2_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_single_process_is_a_noop():
    t = [torch.ones(3)]
    assert allreduce_sum_(t, extra=torch.tensor(2.0)).item() == 2.0 and torch.equal(t[0], torch.ones(3))
