"""Multi-GPU check of the fused gradient exchange (csrc/optim.cu) -- run under gpurun with N >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dp_check.py [p2p|multimem|nccl|auto]

Every rank trains on its shard of a common 4096-ray batch through TrainEngine; checked: the replicas' weights are
bit-identical after every step, and losses / weights follow a single-replica engine fed the whole batch (summation
order differs, so to tolerance).  Prints one JSON line on rank 0 (also timing: ms per step, device events)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_dbr_b200.host import lib as L                      # noqa: E402
from nerf_dbr_b200.host.engine import TrainEngine            # noqa: E402
from nerf_dbr_b200.host.parallel import ray_shard            # noqa: E402
from nerf_dbr_b200.host.synthetic import seeded_models       # noqa: E402


def batch(n, seed, dev):
    g = torch.Generator().manual_seed(seed)
    ro = torch.zeros(n, 3) + torch.tensor([0.0, 0.0, 4.0])
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.25 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
    return [t.to(dev) for t in (ro, rd, torch.rand(n, 3, generator=g), torch.rand(n, 64, generator=g))]


def main():
    transport = sys.argv[1] if len(sys.argv) > 1 else "auto"
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, steps = 4096, 6
    first, count = ray_shard(rank, world, n)
    c, f = seeded_models(5, 30.0, dev)
    eng = TrainEngine(c, f, count, 64, 128, mode=L.BF16, lr=5e-4, gamma=0.999, weight_decay=1e-6, max_norm=1.0,
                      n_rays_global=n, transport=transport)
    ref = None
    if rank == 0:
        c1, f1 = seeded_models(5, 30.0, dev)
        ref = TrainEngine(c1, f1, n, 64, 128, mode=L.BF16, lr=5e-4, gamma=0.999, weight_decay=1e-6, max_norm=1.0,
                          data_parallel=False)
    identical, losses, ref_losses = True, [], []
    for i in range(steps):
        b = batch(n, 50 + i, dev)
        eng.step(*(t[first:first + count] for t in b))
        losses.append(eng.loss())
        digest = torch.stack([eng.P.double().sum(), eng.P.double().abs().sum(), eng.M.double().sum(), eng.V.double().sum()])
        every = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(every, digest)
        identical = identical and all(torch.equal(e, every[0]) for e in every)
        if ref is not None:
            ref.step(*b)
            ref_losses.append(ref.loss())
    diff = None
    if rank == 0:
        diff = float((eng.P[:eng.layout.n_opt] - ref.P[:ref.layout.n_opt]).norm()) / float(ref.P[:ref.layout.n_opt].norm())
    # timing: steps on static inputs, events on this rank's stream, max over ranks
    for _ in range(5):
        eng.step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        eng.step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 50], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ok = identical and all(abs(a - b) <= 3e-3 * abs(b) for a, b in zip(losses, ref_losses)) and diff <= 1e-3
        print(json.dumps({"ok": bool(ok), "world": world, "transport": eng.transport, "graph": bool(eng.graph),
                          "replicas_bit_identical": bool(identical), "losses": losses, "single_replica_losses": ref_losses,
                          "weights_rel_diff_vs_single_replica": diff, "ms_per_step_strong_4096": float(t.item())}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
