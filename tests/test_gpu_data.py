"""GPU: the device data path against the reference's own loader run on the CPU.

A small Blender-format dataset is written to disk (random RGBA PNGs at one size, loaded at another so that the
LANCZOS resize of loader.py:45-46 is exercised; partly transparent pixels so that the white-background composite
matters).  ``nerf_dbr_b200.SyntheticDataset`` must then hold the bits of the reference ``SyntheticDataset``
(src/data/loader.py:13-108): images, poses, focal, ``get_rays``; ray batches from pixel indices must equal
``NeRFTrainer._get_rays(...)[select]``; and ``B200Trainer.train()`` runs an epoch on a stand-in dataset rendered
from the lego fixture."""
import json
import math
import os

import numpy as np
import pytest
import torch

from conftest import load_npz

pytestmark = pytest.mark.gpu


def _write_random_dataset(root, n=3, wh=(40, 30)):
    from PIL import Image
    rng = np.random.default_rng(5)
    os.makedirs(os.path.join(root, "train"), exist_ok=True)
    frames = []
    for i in range(n):
        rgba = rng.integers(0, 256, size=(wh[1], wh[0], 4), dtype=np.uint8)
        rgba[: wh[1] // 3, :, 3] = 255                     # opaque band
        rgba[-wh[1] // 3:, :, 3] = 0                       # transparent band -> white
        Image.fromarray(rgba, mode="RGBA").save(os.path.join(root, "train", f"r_{i}.png"))
        th = 0.7 * i
        pose = [[math.cos(th), 0.0, math.sin(th), 0.3 * i], [0.1 * i, 1.0, 0.0, -0.2], [-math.sin(th), 0.0, math.cos(th), 4.0], [0.0, 0.0, 0.0, 1.0]]
        frames.append({"file_path": f"./train/r_{i}", "transform_matrix": pose})
    with open(os.path.join(root, "transforms_train.json"), "w") as fh:
        json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames}, fh)


def test_dataset_bits_equal_the_reference_loader(tmp_path):
    from oracle import refload
    if refload.reference_root() is None:
        pytest.skip("reference sources not vendored on this box (tools/vendor_reference.sh)")
    refload.import_reference()
    from src.data.loader import SyntheticDataset as RefDataset
    import nerf_dbr_b200 as nb
    root = str(tmp_path / "scene")
    _write_random_dataset(root)
    for wh in ((40, 30), (24, 18), (50, 37)):               # native size, LANCZOS down, LANCZOS up
        ref = RefDataset(root, "train", img_wh=wh, device="cpu")
        ds = nb.SyntheticDataset(root, "train", img_wh=wh, device="cuda")
        assert len(ds) == len(ref) == 3 and ds.focal == ref.focal
        assert ds.images.dtype == torch.float32 and tuple(ds.images.shape) == tuple(ref.images.shape)
        assert torch.equal(ds.images.cpu(), ref.images), wh             # bit-exact white-background composite
        assert torch.equal(ds.poses.cpu(), ref.poses)
        item, ritem = ds[1], ref[1]
        assert set(item) == {"image", "pose", "focal"} and torch.equal(item["image"].cpu(), ritem["image"])
        ro, rd = ds.get_rays(ds.poses[2])
        rro, rrd = ref.get_rays(ref.poses[2])
        assert torch.equal(ro.cpu(), rro.contiguous()) and torch.equal(rd.cpu(), rrd)
        # a training step's ray batch: indices in, rays + targets out
        sel = torch.randperm(wh[0] * wh[1], generator=torch.Generator().manual_seed(1))[:97]
        bo, bd, bt = ds.ray_batch(2, sel.cuda())
        assert torch.equal(bo.cpu(), rro.reshape(-1, 3)[sel]) and torch.equal(bd.cpu(), rrd.reshape(-1, 3)[sel])
        assert torch.equal(bt.cpu(), ref.images[2].reshape(-1, 3)[sel])
    assert float(ds.images[:, -1].min()) == 1.0             # transparent band composited onto white
    both = nb.load_synthetic_data(root, device="cuda", img_wh=(24, 18))
    assert set(both) == {"train"}                            # missing splits are skipped with a warning, as in the reference
    with pytest.raises(nb.NerfB200Error):
        nb.SyntheticDataset(root, "train", img_wh=(24, 18), device="cpu")       # no CPU path


def test_ray_batch_equals_trainer_get_rays(tmp_path):
    """The reference trainer's own ray construction (trainer.py:271-292) for a generic pose and focal."""
    from oracle import nerf_oracle as O
    from nerf_dbr_b200.host import ops
    pose = O.generic_pose()
    w, h, focal = 53, 31, 71.25
    ro, rd = O.camera_rays(pose, w, h, focal)
    sel = torch.randperm(w * h, generator=torch.Generator().manual_seed(2))[:500]
    img = torch.rand(h, w, 3, generator=torch.Generator().manual_seed(3))
    bo, bd, bt = ops.ray_batch(pose, w, h, focal, sel.cuda(), img.cuda())
    assert torch.equal(bo.cpu(), ro.reshape(-1, 3)[sel]) and torch.equal(bd.cpu(), rd.reshape(-1, 3)[sel])
    assert torch.equal(bt.cpu(), img.reshape(-1, 3)[sel])


def test_trainer_runs_an_epoch_from_files(tmp_path):
    """Stand-in dataset rendered from the lego fixture (RGBA PNG + transforms json) -> SyntheticDataset ->
    B200Trainer.train(): the reference's batch format from files, end to end on the device."""
    import nerf_dbr_b200 as nb
    z = load_npz("ckpt_lego_stuffed_fp16.npz")
    lego = {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}
    ck = str(tmp_path / "lego.pth")
    torch.save({"coarse_model": lego, "fine_model": lego}, ck)
    r = nb.B200Renderer("bf16")
    r.setup(ck)
    root = str(tmp_path / "standin")
    nb.write_standin_dataset(root, r, {"train": 4, "val": 2}, img_wh=(48, 36), samples_per_ray=32)
    data = nb.load_synthetic_data(root, device="cuda", img_wh=(48, 36))
    assert set(data) == {"train", "val"} and len(data["train"]) == 4
    assert abs(data["train"].focal - 800.0) <= 1e-3           # camera_angle_x was chosen for the renderers' fixed focal
    img = data["train"][0]["image"]
    assert img.shape == (36, 48, 3) and 0.0 <= float(img.min()) and float(img.max()) <= 1.0 and float(img.std()) > 0.01
    cfg = {"lr": 5e-4, "n_rays": 512, "n_coarse": 16, "n_fine": 32, "precision": "bf16", "seed": 1, "gradient_clipping": 1.0,
           "checkpoint_frequency": 1, "checkpoint_dir": str(tmp_path / "checkpoints")}
    tr = nb.B200Trainer(cfg)
    tr.train(data["train"], data["val"], n_epochs=2, verbose=False)
    assert len(tr.train_losses) == 2 and all(math.isfinite(x) for x in tr.train_losses)
    assert os.path.exists(str(tmp_path / "checkpoints" / "checkpoint_epoch_2.pth"))
    assert math.isfinite(tr.validate(data["val"]))


def test_composite_white_every_value_pair_and_ragged_sizes():
    """composite_white against the loader's float64 arithmetic (loader.py:46-54) for EVERY (colour, alpha) byte pair -- the
    vector kernel forms c / 255 from a product and two fused corrections instead of a division -- and for pixel counts that
    leave a tail for the scalar kernel."""
    import numpy as np
    from nerf_dbr_b200.host import ops
    c, a = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    rgba = np.stack([c, c[::-1], (c.astype(np.int32) * 7 % 256).astype(np.uint8), a], -1).reshape(-1, 4)      # 65536 pixels
    img = rgba.astype(np.float64) / 255.0
    ref = (img[:, :3] * img[:, 3:] + (1.0 - img[:, 3:])).astype(np.float32)
    for n in (65536, 65535, 4099, 5, 3, 1):
        out = ops.composite_white(torch.from_numpy(rgba[:n].copy()).cuda())
        assert out.shape == (n, 3) and np.array_equal(out.cpu().numpy(), ref[:n]), n
