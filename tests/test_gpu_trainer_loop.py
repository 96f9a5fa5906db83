"""GPU: B200Trainer, the mirror of the reference's NeRFTrainer loop (src/training/trainer.py): train_step on the
reference's batch format, validation render against the oracle, epoch loop with checkpoint + resume."""
import math

import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
W, H, FOCAL = 32, 24, 40.0


class TinyDataset:
    """Two views of a procedural target: {'image': [H,W,3], 'pose': [4,4], 'focal': float} like SyntheticDataset."""

    def __init__(self, n=2):
        ys, xs = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
        self.items = []
        for i in range(n):
            img = torch.stack([xs, ys, 0.5 + 0.5 * torch.sin(6.0 * xs + i)], dim=-1)
            self.items.append({"image": img.contiguous(), "pose": O.benchmark_pose(i, 8), "focal": FOCAL})

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def _config(tmp_path, precision):
    return {"lr": 5e-4, "n_rays": 256, "n_coarse": 16, "n_fine": 32, "precision": precision, "seed": 3,
            "gradient_clipping": 1.0, "checkpoint_frequency": 1, "checkpoint_dir": str(tmp_path / "checkpoints")}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_train_steps_reduce_the_loss(precision, tmp_path):
    import nerf_dbr_b200 as nb
    ds = TinyDataset()
    tr = nb.B200Trainer(_config(tmp_path, precision))
    losses = [tr.train_step(ds[i % 2]) for i in range(60)]
    assert all(math.isfinite(x) for x in losses)
    # The steps reduce the loss.  This tiny problem (256 rays, lr 5e-4) sits at the edge of the usual NeRF collapse
    # (density dies, loss returns to its start): tests/diag/engine_curves.py shows the reference's own sequence --
    # torch clip + torch Adam -- collapsing at step 22 for one seed, bit for bit like the fused step.  Whether a run
    # collapses after its descent is chaos of the optimisation, not a property of the kernels, so the gate is on the
    # descent (the fused step is pinned to torch's sequence step by step in tests/test_gpu_engine.py).
    assert min(losses) < 0.6 * losses[0], (losses[:10], min(losses))
    # the schedule of the reference: lr * (lr_decay ** (1 / decay_steps)) ** steps
    assert abs(tr.optimizer.param_groups[0]["lr"] - 5e-4 * (0.1 ** (1 / 250000)) ** 60) < 1e-12


def test_validate_matches_oracle_render(tmp_path):
    import nerf_dbr_b200 as nb
    ds = TinyDataset()
    tr = nb.B200Trainer(_config(tmp_path, "fp32"))
    ck = O.seeded_checkpoint(4, 30.0)
    tr.fine_model.load_state_dict(ck["fine_model"])
    ref = []
    for b in ds.items:
        rgb, _ = O.render_image(ck["fine_model"], b["pose"], W, H, 32, focal=FOCAL)      # the checker
        ref.append(float(torch.mean((rgb - b["image"]) ** 2)))
    assert abs(tr.validate(ds) - sum(ref) / len(ref)) <= 1e-6


def test_epoch_loop_checkpoints_and_resumes(tmp_path):
    import nerf_dbr_b200 as nb
    ds = TinyDataset()
    cfg = _config(tmp_path, "bf16")
    tr = nb.B200Trainer(cfg)
    tr.train(ds, n_epochs=2, verbose=False)
    assert len(tr.train_losses) == 2
    saved = torch.load(str(tmp_path / "checkpoints" / "checkpoint_epoch_2.pth"), weights_only=False)
    assert set(saved) == {"coarse_model", "fine_model", "optimizer", "scheduler", "config", "train_losses", "val_losses"}
    # a fresh trainer finds the checkpoint, restores weights / optimizer / history and has nothing left to do
    tr2 = nb.B200Trainer(cfg)
    tr2.train(ds, n_epochs=2, verbose=False)
    assert len(tr2.train_losses) == 2 and tr2.train_losses == tr.train_losses
    for a, b in zip(tr.step_fn.parameters(), tr2.step_fn.parameters()):
        assert torch.equal(a, b)
    assert tr2.optimizer.state_dict()["state"][0]["step"] == tr.optimizer.state_dict()["state"][0]["step"]
    tr2.train(ds, n_epochs=3, verbose=False)                       # and continues from there
    assert len(tr2.train_losses) == 3
