"""GPU: NeRFTrainer._render_rays as a torch.autograd.Function (SURVEY 8b training boundary): the reference's
unchanged loss / backward / optimizer code on top of the fused kernels."""
import pytest
import torch
import torch.nn.functional as F

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
REL = 2e-3          # per-tensor relative L2 against the reference's autograd (as tests/test_gpu_train.py)


def _oracle_rgb(w, ro, rd, S, tr):
    wt = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    pts, z = O.sample_along_rays(ro, rd, S, t_rand=tr)
    sg, col = O.mlp(wt, pts.reshape(-1, 3), rd[:, None, :].expand_as(pts).reshape(-1, 3))
    return wt, O.composite(sg.reshape(-1, S, 1), col.reshape(-1, S, 3), z, rd)[0]


@pytest.mark.parametrize("loss_kind", ["mse", "weighted_cubic"])
def test_autograd_function_matches_reference_autograd(loss_kind, poses):
    import nerf_dbr_b200 as nb
    ck = O.seeded_checkpoint(5, 30.0)
    ro, rd = O.camera_rays(poses["generic"], 16, 9)
    ro, rd = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
    n = ro.shape[0]
    g = torch.Generator().manual_seed(7)
    tgt, tr, wgt = torch.rand(n, 3, generator=g), torch.rand(n, 64, generator=g), torch.rand(n, 3, generator=g)

    def loss_fn(rgb_c, rgb_f, tgt, wgt):
        if loss_kind == "mse":                                # the reference's train_step (trainer.py:117-121)
            return F.mse_loss(rgb_c, tgt) + F.mse_loss(rgb_f, tgt)
        return (wgt * rgb_c).sum() / n + ((rgb_f - 0.3) ** 3).mean()      # any other differentiable loss

    wc, rgb_c_ref = _oracle_rgb(ck["coarse_model"], ro, rd, 64, tr)
    wf, rgb_f_ref = _oracle_rgb(ck["fine_model"], ro, rd, 128, None)
    loss_ref = loss_fn(rgb_c_ref, rgb_f_ref, tgt, wgt)
    loss_ref.backward()

    coarse, fine = nb.NeRFModel().cuda(), nb.NeRFModel().cuda()
    coarse.load_state_dict(ck["coarse_model"]); fine.load_state_dict(ck["fine_model"])
    rgb_c, rgb_f = nb.render_rays_autograd(coarse, fine, ro.cuda(), rd.cuda(), 64, 128, t_rand=tr.cuda())
    assert rgb_c.requires_grad and rgb_f.requires_grad
    assert (rgb_c.detach().cpu() - rgb_c_ref.detach()).abs().max() <= 1e-4
    assert (rgb_f.detach().cpu() - rgb_f_ref.detach()).abs().max() <= 1e-4
    loss = loss_fn(rgb_c, rgb_f, tgt.cuda(), wgt.cuda())
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 1e-5 * max(abs(float(loss_ref.detach())), 1e-6)
    worst = []
    for model, ref in ((coarse, wc), (fine, wf)):
        for name, p in model.named_parameters():
            got, want = p.grad.cpu().double(), ref[name].grad.double()
            worst.append((float((got - want).norm()) / max(float(want.norm()), 1e-30), name))
    worst.sort(reverse=True)
    print("autograd boundary vs reference autograd:", [(f"{e:.1e}", k) for e, k in worst[:4]])
    assert worst[0][0] <= REL, worst[0]
    # the unchanged reference optimizer code runs on these gradients
    params = list(coarse.parameters()) + list(fine.parameters())
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    torch.optim.Adam(params, lr=5e-4).step()
