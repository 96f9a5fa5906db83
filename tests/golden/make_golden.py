#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

The reference's tests hold no expected values (shape/range checks only), so these
fixtures -- outputs of the unmodified reference classes PyTorchCPURenderer,
NeRFModel, VolumeRenderer and NeRFTrainer on seeded inputs -- are what pins
oracle/ to the reference.  Big arrays whose gate is bit-exactness are stored as
sha256 digests of their raw little-endian fp32 bytes.

Environment it was generated with is recorded in golden_meta.json (torch version,
CPU capability: torch.linspace bits differ between FMA and non-FMA builds).
"""
import hashlib
import json
import os
import re
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("NERF_DBR_REFERENCE", "/root/reference")

# the reference imports matplotlib at module import (benchmark_suite.py:8, trainer.py:13);
# it is not installed here and not needed on this path
for _m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(_m, types.ModuleType(_m))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)
os.chdir(tempfile.mkdtemp())  # the reference creates outputs/ and checkpoints/ in cwd

from src.benchmark.pytorch_renderers import PyTorchCPURenderer  # noqa: E402
from src.benchmark.benchmark_suite import UnifiedBenchmarkSuite  # noqa: E402
from src.models.nerf import NeRFModel  # noqa: E402
from src.training.trainer import NeRFTrainer  # noqa: E402
from src.utils.rendering import VolumeRenderer  # noqa: E402

from oracle import nerf_oracle as O  # noqa: E402  (only for the fixture *inputs*)


def sha(t) -> str:
    a = t.detach().contiguous().numpy() if isinstance(t, torch.Tensor) else np.ascontiguousarray(t)
    return hashlib.sha256(a.tobytes()).hexdigest()


def lego_stuffed():
    """SURVEY 8c fixture 2: pour the bmild Keras arrays into the reference's shapes.  A
    value-distribution fixture (trained magnitudes, saturated alphas), not a lego image.
    Rounded to fp16 so the committed file stays ~1 MB; the goldens are generated from the
    rounded weights."""
    arr = np.load(os.path.join(REF, "data/lego_example_weights/model_fine_200000.npy"),
                  allow_pickle=True)
    sd = {}
    plain = {0: 0, 1: 2, 2: 4, 3: 6, 5: 8, 6: 12, 7: 14}
    for layer, a in plain.items():
        sd[f"layers.{layer}.weight"] = arr[a].T
        sd[f"layers.{layer}.bias"] = arr[a + 1]
    w4 = arr[10]                                  # rows [pts63, h256] -> columns [h256, pe63]
    sd["layers.4.weight"] = np.concatenate([w4[63:], w4[:63]], 0).T
    sd["layers.4.bias"] = arr[11]
    sd["density_head.weight"] = arr[22].T
    sd["density_head.bias"] = arr[23]
    sd["color_layers.0.weight"] = arr[18].T
    sd["color_layers.0.bias"] = arr[19]
    sd["color_layers.1.weight"] = arr[20].T
    sd["color_layers.1.bias"] = arr[21]
    return {k: np.ascontiguousarray(v).astype(np.float16) for k, v in sd.items()}


def checkpoints():
    np.savez_compressed(os.path.join(HERE, "ckpt_lego_stuffed_fp16.npz"), **lego_stuffed())
    z = np.load(os.path.join(HERE, "ckpt_lego_stuffed_fp16.npz"))
    lego = {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files}
    return {
        "rand2": O.seeded_checkpoint(2),
        "semi30": O.seeded_checkpoint(2, 30.0),
        "trained11": O.trained_like_checkpoint(11),
        "lego": {"coarse_model": lego, "fine_model": lego},
    }


def ref_renderer(ckpt):
    path = os.path.join(os.getcwd(), f"ck_{np.random.randint(1 << 30)}.pth")
    torch.save(ckpt, path)
    r = PyTorchCPURenderer()
    r.setup(path)
    NeRFModel().load_state_dict(ckpt["fine_model"])  # "All keys matched"
    return r


def main():
    meta = {"torch": torch.__version__, "cpu_capability": torch.backends.cpu.get_cpu_capability(),
            "threads": torch.get_num_threads(), "reference": REF}
    assert meta["cpu_capability"] != "DEFAULT"
    ckpts = checkpoints()
    poses = {"bench0": O.benchmark_pose(0, 3), "bench1": O.benchmark_pose(1, 3),
             "generic": O.generic_pose()}
    for i, p in enumerate(UnifiedBenchmarkSuite.generate_test_poses(None, 3)):
        assert torch.equal(p, O.benchmark_pose(i, 3))
    meta["seeded_init_sha"] = {
        "rand2.fine": sha(torch.cat([v.reshape(-1) for v in ckpts["rand2"]["fine_model"].values()])),
        "rand2.coarse": sha(torch.cat([v.reshape(-1) for v in ckpts["rand2"]["coarse_model"].values()])),
        "trained11.fine": sha(torch.cat([v.reshape(-1) for v in ckpts["trained11"]["fine_model"].values()])),
    }
    # seeded init must equal the reference fixture recipe (test_system.py:197-201)
    torch.manual_seed(2)
    c, f = NeRFModel(), NeRFModel()
    for k, v in f.state_dict().items():
        assert torch.equal(v, ckpts["rand2"]["fine_model"][k])
    for k, v in c.state_dict().items():
        assert torch.equal(v, ckpts["rand2"]["coarse_model"][k])

    # ---- 1. rays / samples: bit-exact gates, stored as digests --------------------------
    r0 = ref_renderer(ckpts["rand2"])
    geom = {}
    for pname, pose in poses.items():
        for (w, h, s) in [(64, 48, 16), (200, 150, 32), (400, 300, 64), (800, 600, 128)]:
            if (w, h) == (800, 600) and pname != "bench1":
                continue
            ro, rd = r0.generate_rays(pose, w, h)
            ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
            pts, z = r0.sample_points_on_rays(ro, rd, s)
            geom[f"{pname}_{w}x{h}x{s}"] = {
                "rays_o": sha(ro), "rays_d": sha(rd), "z_row": sha(z[0]), "points": sha(pts),
                "points_first": pts[0, :2].tolist(), "points_last": pts[-1, -2:].tolist()}
    # stratified jitter (rendering.py:42-47) with captured t_rand
    vr = VolumeRenderer("cpu")
    ro, rd = r0.generate_rays(poses["generic"], 64, 48)
    ro, rd = ro.reshape(-1, 3)[:1024], rd.reshape(-1, 3)[:1024]
    captured = {}
    orig_rand_like = torch.rand_like

    def rec_rand_like(x, *a, **k):
        captured["t"] = orig_rand_like(x, *a, **k)
        return captured["t"]
    torch.manual_seed(123)
    torch.rand_like = rec_rand_like
    try:
        pts_j, z_j = vr.sample_points_on_rays(ro, rd, 2.0, 6.0, 64, perturb=True)
    finally:
        torch.rand_like = orig_rand_like
    np.savez_compressed(os.path.join(HERE, "golden_stratified.npz"),
                        t_rand=captured["t"].numpy(), z=z_j.numpy(),
                        points_sha=np.array(sha(pts_j)), pose=poses["generic"].numpy())
    with open(os.path.join(HERE, "golden_geometry.json"), "w") as fh:
        json.dump(geom, fh, indent=1)

    # ---- 2. rendered images through the unmodified PyTorchCPURenderer --------------------
    out = {}
    for cname, ck in ckpts.items():
        r = ref_renderer(ck)
        for pname, pose in poses.items():
            rgb, dep = r.render_image(pose, (64, 48), 16)
            out[f"{cname}|{pname}|64x48x16|rgb"] = rgb.numpy()
            out[f"{cname}|{pname}|64x48x16|depth"] = dep.numpy()
        if cname in ("trained11", "lego"):
            rgb, dep = r.render_image(poses["generic"], (96, 64), 64)
            out[f"{cname}|generic|96x64x64|rgb"] = rgb.numpy()
            out[f"{cname}|generic|96x64x64|depth"] = dep.numpy()
    # config 1 of BASELINE.json (200x150x32) on one fixture
    r = ref_renderer(ckpts["trained11"])
    rgb, dep = r.render_image(poses["bench1"], (200, 150), 32)
    out["trained11|bench1|200x150x32|rgb"] = rgb.numpy()
    out["trained11|bench1|200x150x32|depth"] = dep.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_render.npz"), **out)

    # ---- 3. network level: NeRFModel.forward on real sample points -----------------------
    net = {}
    ro, rd = r0.generate_rays(poses["generic"], 64, 48)
    pts, _ = r0.sample_points_on_rays(ro.reshape(-1, 3)[::24], rd.reshape(-1, 3)[::24], 16)
    pos = pts.reshape(-1, 3).contiguous()
    dirs = rd.reshape(-1, 3)[::24][:, None, :].expand_as(pts).reshape(-1, 3).contiguous()
    net["pos"], net["dirs"] = pos.numpy(), dirs.numpy()
    for cname in ("trained11", "lego", "semi30"):
        m = NeRFModel()
        m.load_state_dict(ckpts[cname]["fine_model"])
        m.eval()
        with torch.no_grad():
            sg, col = m(pos, dirs)
            net[f"{cname}|sigma"], net[f"{cname}|rgb"] = sg.numpy(), col.numpy()
            if cname == "trained11":
                pe = m.pos_encoder.encode(pos[:128])
                de = m.dir_encoder.encode(dirs[:128])
                net["pe_pos"], net["pe_dir"] = pe.numpy(), de.numpy()
                x = m.pos_encoder.encode(pos[:16])
                pe16 = x
                for i, layer in enumerate(m.layers):
                    if i == 4:
                        x = torch.cat([x, pe16], -1)
                    x = torch.relu(layer(x))
                    if i in (0, 4, 7):
                        net[f"trained11|hidden{i}"] = x.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_network.npz"), **net)

    # ---- 4. compositing alone (PyTorchCPURenderer.execute_volume_rendering + volume_render)
    g = torch.Generator().manual_seed(7)
    R, S = 300, 64
    sig = torch.rand(R, S, 1, generator=g) * 8 - 1      # includes negatives (relu inside)
    col = torch.rand(R, S, 3, generator=g)
    ro, rd = r0.generate_rays(poses["generic"], 20, 15)
    rd = rd.reshape(-1, 3).contiguous()
    _, z = r0.sample_points_on_rays(ro.reshape(-1, 3), rd, S)
    z = z.contiguous()
    rgb_map, depth = r0.execute_volume_rendering(sig, col, z, rd)
    rgb4, depth4, acc4, w4 = vr.volume_render(sig, col, z, rd)
    assert torch.equal(rgb_map, rgb4) and torch.equal(depth, depth4)
    np.savez_compressed(os.path.join(HERE, "golden_composite.npz"), sigma=sig.numpy(),
                        rgb=col.numpy(), z=z.numpy(), rays_d=rd.numpy(), rgb_map=rgb_map.numpy(),
                        depth=depth.numpy(), acc=acc4.numpy(), weights=w4.numpy())

    # ---- 5. inverse-CDF sampling: reference source + the one-line shape fix ---------------
    src = open(os.path.join(REF, "src/utils/rendering.py")).read()
    broken = "z_vals_g = torch.gather(z_vals.unsqueeze(-2).expand(matched_shape), dim=-1,"
    assert broken in src
    fixed = src.replace(
        broken,
        "z_vals_g = torch.gather(z_vals.unsqueeze(-2).expand(list(indices_g.shape[:-1]) + [z_vals.shape[-1]]), dim=-1,")
    ns = {}
    exec(compile(fixed, "rendering_fixed.py", "exec"), ns)
    vr_fixed = ns["VolumeRenderer"]("cpu")
    # the shipped function really is dead code that raises
    try:
        vr.importance_sample(torch.zeros(4, 3), torch.ones(4, 3), torch.rand(4, 64),
                             torch.rand(4, 64), 8)
        raised = False
    except RuntimeError:
        raised = True
    meta["importance_sample_shipped_raises"] = raised
    imp = {}
    for S in (32, 64, 128):
        g = torch.Generator().manual_seed(100 + S)
        R, n_imp = 96, 128
        ro, rd = r0.generate_rays(poses["generic"], 12, 8)
        ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
        _, z = r0.sample_points_on_rays(ro, rd, S)
        w = torch.rand(R, S, generator=g) ** 6            # peaky, many near-zero bins
        w[:8] = 0.0                                        # degenerate rays: pdf uniform
        u_fix = torch.rand(R, n_imp, generator=g)
        u_fix[:, 0] = 0.0
        u_fix[:, 1] = 0.99999994
        orig_rand = torch.rand
        torch.rand = lambda *a, **k: u_fix
        try:
            pts, zs = vr_fixed.importance_sample(ro, rd, z.contiguous(), w, n_imp)
        finally:
            torch.rand = orig_rand
        imp[f"S{S}|z"], imp[f"S{S}|w"], imp[f"S{S}|u"] = z.contiguous().numpy(), w.numpy(), u_fix.numpy()
        imp[f"S{S}|z_new"], imp[f"S{S}|points"] = zs.numpy(), pts.numpy()
        imp[f"S{S}|rays_o"], imp[f"S{S}|rays_d"] = ro.numpy(), rd.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_importance.npz"), **imp)

    # ---- 6. one NeRFTrainer.train_step: loss and the 44 gradients -------------------------
    H, W, n_rays = 150, 200, 512
    torch.manual_seed(5)
    trainer = NeRFTrainer({"device": "cpu", "n_rays": n_rays, "n_coarse": 64, "n_fine": 128,
                           "lr": 5e-4})
    with torch.no_grad():
        for m in (trainer.coarse_model, trainer.fine_model):
            m.density_head.weight.mul_(30.0)
            m.density_head.bias.mul_(30.0)
    ref_ck = O.seeded_checkpoint(5, 30.0)
    for k, v in trainer.fine_model.state_dict().items():
        assert torch.equal(v, ref_ck["fine_model"][k]), k
    image = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(0))
    pose = torch.eye(4)
    pose[2, 3] = 4.0
    rec = {}
    orig_perm, orig_rand_like = torch.randperm, torch.rand_like

    def rec_perm(*a, **k):
        rec["perm"] = orig_perm(*a, **k)
        return rec["perm"]

    def rec_rl(x, *a, **k):
        rec["t_rand"] = orig_rand_like(x, *a, **k)
        return rec["t_rand"]
    torch.manual_seed(0)
    torch.randperm, torch.rand_like = rec_perm, rec_rl
    try:
        loss = trainer.train_step({"image": image, "pose": pose, "focal": 800.0})
    finally:
        torch.randperm, torch.rand_like = orig_perm, orig_rand_like
    tr = {"loss": np.float32(loss), "select": rec["perm"][:n_rays].numpy(),
          "t_rand": rec["t_rand"].numpy(), "H": H, "W": W, "seed": 5, "density_gain": 30.0}
    for tag, m in (("coarse", trainer.coarse_model), ("fine", trainer.fine_model)):
        for name, p in m.named_parameters():
            gflat = p.grad.reshape(-1)
            tr[f"{tag}|{name}|norm"] = np.float64(gflat.double().norm().item())
            tr[f"{tag}|{name}|strided"] = gflat[::37].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "golden_train.npz"), **tr)

    with open(os.path.join(HERE, "golden_meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print("golden fixtures written to", HERE)
    for fn in sorted(os.listdir(HERE)):
        print(f"  {fn:36s} {os.path.getsize(os.path.join(HERE, fn)) / 1024:8.1f} KB")


if __name__ == "__main__":
    main()
