"""GPU: the training step's forward+backward against the reference's autograd (golden fixture from
NeRFTrainer.train_step) and against the oracle on small ragged cases."""
import numpy as np
import pytest
import torch

from conftest import load_npz
from gpu_util import Watchdog
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
REL = 2e-3          # per-tensor relative L2 error of a gradient vs the reference's fp32 autograd on the CPU
                    # (fp32 summation-order noise: the oracle itself moves by ~5e-4 on layers.0.weight between
                    # fp32 and fp64; measured worst case here 4.8e-4)


def models_from(ck):
    import nerf_dbr_b200 as nb
    c, f = nb.NeRFModel().cuda(), nb.NeRFModel().cuda()
    c.load_state_dict(ck["coarse_model"])
    f.load_state_dict(ck["fine_model"])
    return c, f


def test_train_step_matches_reference_golden():
    from nerf_dbr_b200.host.trainer import B200TrainStep
    g = load_npz("golden_train.npz")
    ck = O.seeded_checkpoint(int(g["seed"]), float(g["density_gain"]))
    coarse, fine = models_from(ck)
    H, W = int(g["H"]), int(g["W"])
    image = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(0))
    pose = torch.eye(4)
    pose[2, 3] = 4.0
    ro, rd = O.camera_rays(pose, W, H)
    sel = torch.from_numpy(g["select"])
    ro, rd, tgt = ro.reshape(-1, 3)[sel].cuda(), rd.reshape(-1, 3)[sel].cuda(), image.reshape(-1, 3)[sel].cuda()
    step = B200TrainStep(coarse, fine, 64, 128)
    loss, rgb_c, rgb_f = step(ro, rd, tgt, t_rand=torch.from_numpy(g["t_rand"]).cuda())
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * float(g["loss"])
    worst, rows = 0.0, []
    for tag, m in (("coarse", coarse), ("fine", fine)):
        for name, p in m.named_parameters():
            ref_norm = float(g[f"{tag}|{name}|norm"])
            got = p.grad.reshape(-1).cpu()
            ref = torch.from_numpy(g[f"{tag}|{name}|strided"])
            err = float((got[::37] - ref).double().norm()) / max(float(ref.double().norm()), 1e-30)
            nerr = abs(float(got.double().norm()) - ref_norm) / max(ref_norm, 1e-30)
            worst = max(worst, err, nerr)
            rows.append((err, nerr, tag, name))
    rows.sort(reverse=True)
    print("worst relative gradient errors vs reference autograd:", [(f"{e:.1e}", f"{n:.1e}", t, k) for e, n, t, k in rows[:6]])
    assert worst <= REL, rows[0]


@pytest.mark.parametrize("n_rays,S,jitter", [(37, 16, True), (130, 64, False), (65, 100, True), (9, 200, False)])
def test_train_fwd_bwd_vs_oracle_autograd(n_rays, S, jitter, checkpoints, poses):
    """Ragged ray counts and sample counts, trained-magnitude weights (saturated alphas), accumulation
    into existing grads, global-count scaling (the data-parallel contract)."""
    import nerf_dbr_b200 as nb
    from nerf_dbr_b200.host import ops
    w = checkpoints["semi30"]["fine_model"]
    ro, rd = O.camera_rays(poses["generic"], 23, 11)
    gen = torch.Generator().manual_seed(n_rays)
    idx = torch.randperm(ro.reshape(-1, 3).shape[0], generator=gen)[:n_rays]
    ro, rd = ro.reshape(-1, 3)[idx].contiguous(), rd.reshape(-1, 3)[idx].contiguous()
    tgt = torch.rand(n_rays, 3, generator=gen)
    tr = torch.rand(n_rays, S, generator=gen) if jitter else None
    # oracle: one network's term, autograd
    wt = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    pts, z = O.sample_along_rays(ro, rd, S, t_rand=tr)
    sg, col = O.mlp(wt, pts.reshape(-1, 3), rd[:, None, :].expand_as(pts).reshape(-1, 3))
    rgb_ref = O.composite(sg.reshape(n_rays, S, 1), col.reshape(n_rays, S, 3), z, rd)[0]
    n_global = 3 * n_rays
    loss_ref = ((rgb_ref - tgt) ** 2).sum() / (3 * n_global)
    loss_ref.backward()
    m = nb.NeRFModel().cuda()
    m.load_state_dict(w)
    loss, rgb = ops.train_fwd_bwd(m, ro.cuda(), rd.cuda(), tgt.cuda(), S, None if tr is None else tr.cuda(),
                                  n_rays_global=n_global)
    assert (rgb.cpu() - rgb_ref.detach()).abs().max() <= 1e-4
    assert abs(float(loss) - float(loss_ref.detach())) <= 1e-5 * max(float(loss_ref.detach()), 1e-12)
    first = {}
    for name, p in m.named_parameters():
        got, ref = p.grad.cpu().double(), wt[name].grad.double()
        assert float((got - ref).norm()) <= REL * max(float(ref.norm()), 1e-12), name
        first[name] = p.grad.clone()
    # accumulation contract: a second call adds onto the existing grads
    ops.train_fwd_bwd(m, ro.cuda(), rd.cuda(), tgt.cuda(), S, None if tr is None else tr.cuda(), n_rays_global=n_global)
    for name, p in m.named_parameters():
        assert float((p.grad - 2 * first[name]).norm()) <= 1e-5 * max(float(first[name].norm()), 1e-20), name


def test_adam_step_follows_reference_trainer(checkpoints):
    """Three optimizer steps with the unchanged torch Adam on top of the CUDA gradients track the
    oracle's CPU training loop (same jitter, same batches)."""
    from nerf_dbr_b200.host.trainer import B200TrainStep
    ck = O.seeded_checkpoint(5, 30.0)
    coarse, fine = models_from(ck)
    step = B200TrainStep(coarse, fine, 32, 64)
    opt = torch.optim.Adam(step.parameters(), lr=5e-4)
    cw = {k: v.clone() for k, v in ck["coarse_model"].items()}
    fw = {k: v.clone() for k, v in ck["fine_model"].items()}
    cpu_params = [torch.nn.Parameter(v) for v in list(cw.values()) + list(fw.values())]
    cpu_opt = torch.optim.Adam(cpu_params, lr=5e-4)
    pose = torch.eye(4)
    pose[2, 3] = 4.0
    ro_all, rd_all = O.camera_rays(pose, 40, 30)
    gen = torch.Generator().manual_seed(1)
    for it in range(3):
        sel = torch.randperm(1200, generator=gen)[:96]
        ro, rd = ro_all.reshape(-1, 3)[sel].contiguous(), rd_all.reshape(-1, 3)[sel].contiguous()
        tgt = torch.rand(96, 3, generator=gen)
        tr = torch.rand(96, 32, generator=gen)
        loss, _, _ = step(ro.cuda(), rd.cuda(), tgt.cuda(), t_rand=tr.cuda())
        opt.step()
        keys = list(cw.keys())
        cwd = {k: p for k, p in zip(keys, cpu_params[:len(keys)])}
        fwd = {k: p for k, p in zip(keys, cpu_params[len(keys):])}
        l_ref, _, _, gc, gf = O.train_loss_and_grads(cwd, fwd, ro, rd, tgt, 32, 64, tr)
        cpu_opt.zero_grad()
        for k in keys:
            cwd[k].grad, fwd[k].grad = gc[k], gf[k]
        cpu_opt.step()
        assert abs(float(loss) - float(l_ref)) <= 1e-4 * float(l_ref)
    # Adam normalises every element's update to ~lr, so elements whose gradient is rounding noise can move
    # differently; the check is on the aggregate parameter update
    num = den = 0.0
    for models, off in ((coarse, 0), (fine, len(cw))):
        for i, (k, p) in enumerate(models.named_parameters()):
            start = (ck["coarse_model"] if off == 0 else ck["fine_model"])[k]
            d_gpu = p.detach().cpu().double() - start.double()
            d_cpu = cpu_params[off + i].detach().double() - start.double()
            num += float((d_gpu - d_cpu).norm() ** 2)
            den += float(d_cpu.norm() ** 2)
    assert (num / den) ** 0.5 <= 0.05, (num / den) ** 0.5


def _bf16_forward_autograd(w, ro, rd, tgt, S, tr):
    """fp32 autograd through a forward whose operands are rounded to bf16 exactly where the tensor-core kernel
    rounds them (straight-through): the gradient a bf16 forward SHOULD give -- separates kernel bugs from the
    precision of the mode."""
    import torch.nn.functional as F

    def bf(x):
        return x.to(torch.bfloat16).to(torch.float32)

    def st(x):
        return x + (bf(x) - x).detach()
    wt = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    pts, z = O.sample_along_rays(ro, rd, S, t_rand=tr)
    p = pts.reshape(-1, 3)
    d = rd[:, None, :].expand_as(pts).reshape(-1, 3)
    pe = st(O.encode(p, 10))
    h = pe
    for i in range(8):
        W = st(wt[f"layers.{i}.weight"])
        acc = h @ W[:, :256].T + pe @ W[:, 256:].T if i == 4 else h @ W.T
        h = st(torch.relu(acc + wt[f"layers.{i}.bias"]))
    sigma = torch.relu(h @ st(wt["density_head.weight"]).T + wt["density_head.bias"])
    Wc = wt["color_layers.0.weight"]
    c = torch.relu(h @ st(Wc[:, :256]).T + O.encode(d, 4) @ Wc[:, 256:].T + wt["color_layers.0.bias"])
    col = torch.sigmoid(c @ wt["color_layers.1.weight"].T + wt["color_layers.1.bias"])
    rgb = O.composite(sigma.reshape(-1, S, 1), col.reshape(-1, S, 3), z, rd)[0]
    loss = F.mse_loss(rgb, tgt)
    loss.backward()
    return float(loss), {k: v.grad for k, v in wt.items()}


def test_train_step_bf16_mode_tensor_cores():
    """BF16 mode (forward and weight gradients on the tensor cores; bf16 operands, fp32 accumulation).
    Two gates on the golden train step:
      * correctness: per-tensor gradients within 2 % (relative L2) of fp32 autograd through a bf16-rounded forward
        (what this mode is supposed to compute);
      * precision: loss within 0.5 % and gradients within 15 % of the reference's fp32 autograd -- bf16 activations
        move the first layers' gradients (sums over sign-alternating high-frequency encodings) by up to 9 % on this
        fixture, in the emulation as on the device."""
    from nerf_dbr_b200.host.trainer import B200TrainStep
    g = load_npz("golden_train.npz")
    ck = O.seeded_checkpoint(int(g["seed"]), float(g["density_gain"]))
    coarse, fine = models_from(ck)
    H, W = int(g["H"]), int(g["W"])
    image = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(0))
    pose = torch.eye(4)
    pose[2, 3] = 4.0
    ro, rd = O.camera_rays(pose, W, H)
    sel = torch.from_numpy(g["select"])
    ro, rd, tgt = ro.reshape(-1, 3)[sel], rd.reshape(-1, 3)[sel], image.reshape(-1, 3)[sel]
    t_rand = torch.from_numpy(g["t_rand"])
    step = B200TrainStep(coarse, fine, 64, 128, mode=1)
    loss, _, _ = step(ro.cuda(), rd.cuda(), tgt.cuda(), t_rand=t_rand.cuda())
    assert abs(float(loss) - float(g["loss"])) <= 5e-3 * float(g["loss"])
    lc, gc = _bf16_forward_autograd(ck["coarse_model"], ro, rd, tgt, 64, t_rand)
    lf, gf = _bf16_forward_autograd(ck["fine_model"], ro, rd, tgt, 128, None)
    assert abs(float(loss) - (lc + lf)) <= 1e-3 * (lc + lf)
    rows_emu, rows_ref = [], []
    for tag, m, emu in (("coarse", coarse, gc), ("fine", fine, gf)):
        for name, p in m.named_parameters():
            got = p.grad.cpu().double()
            e = emu[name].double()
            rows_emu.append((float((got - e).norm()) / max(float(e.norm()), 1e-30), tag, name))
            ref = torch.from_numpy(g[f"{tag}|{name}|strided"]).double()
            rows_ref.append((float((got.reshape(-1)[::37] - ref).norm()) / max(float(ref.norm()), 1e-30), tag, name))
    rows_emu.sort(reverse=True)
    rows_ref.sort(reverse=True)
    print("bf16 mode vs bf16-forward autograd:", [(f"{e:.1e}", t, k) for e, t, k in rows_emu[:4]])
    print("bf16 mode vs reference fp32 autograd:", [(f"{e:.1e}", t, k) for e, t, k in rows_ref[:4]])
    assert rows_emu[0][0] <= 2e-2, rows_emu[0]
    assert rows_ref[0][0] <= 0.15, rows_ref[0]


@pytest.mark.parametrize("n_rays,S,jitter", [(37, 100, True), (130, 64, False), (9, 200, False), (65, 16, True), (1, 128, False)])
def test_train_bf16_mode_ragged_shapes(n_rays, S, jitter, checkpoints, poses):
    """BF16 mode on shapes that leave padded tiles (S = 100, 200), several rays per tile (S = 16), a partial last
    64-sample slab of the operand blocks and a single ray: rgb / loss against the bf16-forward emulation, per-tensor
    gradients within 3 % of its autograd; a second call accumulates."""
    import nerf_dbr_b200 as nb
    from nerf_dbr_b200.host import ops
    from nerf_dbr_b200.host import lib as L
    w = checkpoints["semi30"]["fine_model"]
    ro, rd = O.camera_rays(poses["generic"], 23, 11)
    gen = torch.Generator().manual_seed(1000 + n_rays)
    idx = torch.randperm(ro.reshape(-1, 3).shape[0], generator=gen)[:n_rays]
    ro, rd = ro.reshape(-1, 3)[idx].contiguous(), rd.reshape(-1, 3)[idx].contiguous()
    tgt = torch.rand(n_rays, 3, generator=gen)
    tr = torch.rand(n_rays, S, generator=gen) if jitter else None
    loss_ref, g_ref = _bf16_forward_autograd(w, ro, rd, tgt, S, tr)
    m = nb.NeRFModel().cuda()
    m.load_state_dict(w)
    with Watchdog() as wd:
        loss, rgb = ops.train_fwd_bwd(m, ro.cuda(), rd.cuda(), tgt.cuda(), S, None if tr is None else tr.cuda(), mode=L.BF16)
        torch.cuda.synchronize()
        assert int(wd.word.item()) == 0
    assert torch.isfinite(rgb).all()
    assert abs(float(loss) - loss_ref) <= 2e-3 * max(loss_ref, 1e-12)
    rows, first = [], {}
    for name, p in m.named_parameters():
        got, ref = p.grad.cpu().double(), g_ref[name].double()
        rows.append((float((got - ref).norm()) / max(float(ref.norm()), 1e-30), name))
        first[name] = p.grad.clone()
    rows.sort(reverse=True)
    print("bf16 ragged vs bf16-forward autograd:", [(f"{e:.1e}", k) for e, k in rows[:4]])
    assert rows[0][0] <= 3e-2, rows[0]
    ops.train_fwd_bwd(m, ro.cuda(), rd.cuda(), tgt.cuda(), S, None if tr is None else tr.cuda(), mode=L.BF16)
    for name, p in m.named_parameters():
        assert float((p.grad - 2 * first[name]).norm()) <= 1e-3 * max(float(first[name].norm()), 1e-20), name


def test_split_phases_equal_single_call(checkpoints, poses):
    """nerf_b200_train_fwd_bwd_ex: the activation phase and the weight-gradient phase as two calls (with an SM limit on
    each) accumulate the same gradients, loss and colours as the single call, up to the order of fp32 atomic
    additions; split phases are refused where they cannot work (FP32 mode)."""
    import nerf_dbr_b200 as nb
    from nerf_dbr_b200.host import lib as L, ops
    ck = O.seeded_checkpoint(11, 30.0)
    n = 1024
    ro, rd = O.camera_rays(poses["generic"], 64, 16)
    ro, rd = ro.reshape(-1, 3).contiguous().cuda(), rd.reshape(-1, 3).contiguous().cuda()
    g = torch.Generator().manual_seed(3)
    tgt, tr = torch.rand(n, 3, generator=g).cuda(), torch.rand(n, 64, generator=g).cuda()
    out = []
    with Watchdog() as wd:
        for split in (False, True):
            coarse, _ = models_from(ck)
            tp = ops.TrainPass(coarse, ro, rd, tgt, 64, tr, mode=L.BF16)
            if split:
                tp.run(L.TRAIN_ACTIVATIONS, sm_limit=100)
                torch.cuda.synchronize()                      # the phases may be arbitrarily far apart in time
                tp.run(L.TRAIN_WEIGHT_GRADS, sm_limit=48)
            else:
                tp.run()
            torch.cuda.synchronize()
            assert int(wd.word.item()) == 0
            out.append((float(tp.loss), tp.rgb.cpu(), [p.grad.cpu().double() for p in coarse.parameters()]))
    (l0, c0, g0), (l1, c1, g1) = out
    # the loss is one fp32 atomicAdd per ray (1024 of them into a sum of ~340): the order moves it by ~1e-6 relative;
    # the per-ray colours below are the bit-exact check
    assert abs(l0 - l1) <= 5e-6 * abs(l0)
    assert torch.equal(c0, c1)
    for a, b in zip(g0, g1):
        assert float((a - b).norm()) <= 1e-4 * max(float(a.norm()), 1e-20)
    coarse, _ = models_from(ck)
    with pytest.raises(nb.NerfB200Error) as e:
        ops.TrainPass(coarse, ro, rd, tgt, 64, tr, mode=L.FP32).run(L.TRAIN_ACTIVATIONS)
    assert e.value.code == -2


def test_flat_gradient_bucket_survives_zero_grad_to_none(checkpoints, poses):
    """B200TrainStep keeps all 44 gradients as views of one flat buffer (one memset, in-place all-reduce).  The
    reference's loop calls ``optimizer.zero_grad()`` (set_to_none=True by default), which drops the views: the next
    step must re-attach them and produce the same gradients as a fresh object."""
    from nerf_dbr_b200.host import lib as L
    from nerf_dbr_b200.host.trainer import B200TrainStep
    ck = O.seeded_checkpoint(5, 30.0)
    ro, rd = O.camera_rays(poses["generic"], 16, 8)
    ro, rd = ro.reshape(-1, 3).contiguous().cuda(), rd.reshape(-1, 3).contiguous().cuda()
    g = torch.Generator().manual_seed(2)
    tgt, tr = torch.rand(128, 3, generator=g).cuda(), torch.rand(128, 64, generator=g).cuda()
    coarse, fine = models_from(ck)
    step = B200TrainStep(coarse, fine, 64, 128, mode=L.FP32)
    opt = torch.optim.Adam(step.parameters(), lr=0.0)          # lr 0: the weights stay put, the protocol is exercised
    step(ro, rd, tgt, t_rand=tr)
    flat = step._flat
    assert all(p.grad.data_ptr() >= flat.data_ptr() and p.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4 for p in step.parameters())
    first = [p.grad.clone() for p in step.parameters()]
    opt.step()
    opt.zero_grad()                                            # grads -> None
    assert all(p.grad is None for p in step.parameters())
    step(ro, rd, tgt, t_rand=tr)
    assert step._flat is flat
    for a, p in zip(first, step.parameters()):
        assert p.grad.data_ptr() >= flat.data_ptr()
        assert float((p.grad - a).norm()) <= 1e-5 * max(float(a.norm()), 1e-20)


def test_checkpoint_round_trip_reference_format(tmp_path, checkpoints, poses):
    """A checkpoint written after CUDA training steps has the reference's keys (trainer.py:374-388), loads through the
    renderer's SharedNeRFModel path (base_renderer.py:42-48) and resumes training bit-identically."""
    import nerf_dbr_b200 as nb
    from nerf_dbr_b200.host.trainer import B200TrainStep, load_checkpoint, save_checkpoint
    ck = O.seeded_checkpoint(5, 30.0)
    coarse, fine = models_from(ck)
    step = B200TrainStep(coarse, fine, 32, 64)
    opt = torch.optim.Adam(step.parameters(), lr=5e-4)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.1 ** (1 / 250000))
    ro, rd = O.camera_rays(poses["generic"], 16, 8)
    ro, rd = ro.reshape(-1, 3).cuda(), rd.reshape(-1, 3).cuda()
    g = torch.Generator().manual_seed(3)
    tgt, tr = torch.rand(128, 3, generator=g).cuda(), torch.rand(128, 32, generator=g).cuda()
    losses = []
    for _ in range(2):
        loss, _, _ = step(ro, rd, tgt, t_rand=tr)
        opt.step(); sched.step(); losses.append(float(loss))
    path = str(tmp_path / "checkpoint_epoch_2.pth")
    save_checkpoint(path, step, opt, sched, {"n_rays": 128}, losses, [])
    raw = torch.load(path, weights_only=False)
    assert set(raw) == {"coarse_model", "fine_model", "optimizer", "scheduler", "config", "train_losses", "val_losses"}
    r = nb.B200Renderer("fp32")
    r.setup(path)                                            # the renderer-side loader
    rgb, _ = r.render_image(poses["generic"], (16, 8), 16)
    ref, _ = O.render_image({k: v.cpu() for k, v in fine.state_dict().items()}, poses["generic"], 16, 8, 16)
    assert (rgb.cpu() - ref).abs().max() <= 1e-4
    # resume: a fresh trainer loaded from the file takes the same next step
    c2, f2 = models_from(O.seeded_checkpoint(1))
    step2 = B200TrainStep(c2, f2, 32, 64)
    opt2 = torch.optim.Adam(step2.parameters(), lr=5e-4)
    tl, _ = load_checkpoint(path, step2, opt2)
    assert tl == losses
    l1, _, _ = step(ro, rd, tgt, t_rand=tr); opt.step()
    l2, _, _ = step2(ro, rd, tgt, t_rand=tr); opt2.step()
    # fp32 atomics (loss and gradient accumulation) make a step reproducible to rounding, not to the bit
    assert abs(float(l1) - float(l2)) <= 1e-6 * float(l1)
    num = sum(float((a - b).double().norm() ** 2) for a, b in zip(step.parameters(), step2.parameters()))
    den = sum(float(a.double().norm() ** 2) for a in step.parameters())
    assert (num / den) ** 0.5 <= 1e-5


@pytest.mark.parametrize("mode", [0, 1])
def test_multi_chunk_batch_equals_its_chunks(mode):
    """A batch larger than one workspace chunk (524,288 samples): 4600 rays x 128 samples = 588,800 samples run as a
    full 4096-ray chunk plus a ragged 504-ray chunk inside ONE call (chunk-offset ray pointers, a smaller last chunk,
    the pad-slab memsets, a fork/join of the auxiliary wgrad streams per chunk).  Gradients accumulate (+=) and are
    scaled by the global ray count, so the call must equal the two chunks run as separate calls -- each of those fits
    one chunk, the path every other test pins to the reference."""
    from nerf_dbr_b200.host import ops
    from nerf_dbr_b200.host.synthetic import seeded_models
    n, S = 4600, 128
    g = torch.Generator().manual_seed(3)
    ro = (torch.zeros(n, 3) + torch.tensor([0.0, 0.0, 4.0])).cuda()
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g) * 0.2 + torch.tensor([0.0, 0.0, -1.0]), dim=-1).cuda()
    tgt, tr = torch.rand(n, 3, generator=g).cuda(), torch.rand(n, S, generator=g).cuda()

    def run(parts):
        _, fine = seeded_models(5, 30.0, "cuda")
        total, first, rgbs = 0.0, 0, []
        for count in parts:
            sl = slice(first, first + count)
            loss, rgb = ops.train_fwd_bwd(fine, ro[sl], rd[sl], tgt[sl], S, tr[sl].contiguous(), n_rays_global=n, mode=mode)
            total += float(loss)
            rgbs.append(rgb)
            first += count
        return total, [p.grad.double().clone() for p in fine.parameters()], torch.cat(rgbs), [k for k, _ in fine.named_parameters()]

    with Watchdog() as wd:
        l_one, g_one, rgb_one, names = run([n])
        l_two, g_two, rgb_two, _ = run([4096, n - 4096])
        torch.cuda.synchronize()
        assert int(wd.word.item()) == 0, hex(int(wd.word.item()) & 0xffffffff)
    assert abs(l_one - l_two) <= 1e-5 * abs(l_two)
    assert torch.equal(rgb_one, rgb_two)
    rel = sorted(((float((a - b).norm()) / max(float(b.norm()), 1e-30), k) for a, b, k in zip(g_one, g_two, names)), reverse=True)
    print(f"mode {mode}: multi-chunk vs separate chunks, worst gradient tensors:", [(f"{e:.1e}", k) for e, k in rel[:3]])
    assert rel[0][0] <= 1e-4, rel[0]              # summation order only (fp32 atomics / split reduction)
