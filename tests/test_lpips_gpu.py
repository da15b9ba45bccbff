"""GPU: LPIPS/VGG16 on the library's convolution kernels (reference models/lpips.py:86-125) against the value
and gradient the REAL reference produced for the same weights (tests/golden/lpips_2x32x48.npz, written by
oracle/make_golden_lpips.py) and against the CPU oracle on odd sizes."""
import numpy as np
import pytest
import torch

from oracle import papr_oracle as O

pytestmark = pytest.mark.gpu


def _net(seed):
    from papr_b200.lpips import LPNet
    P = O.init_lpips_params(seed=seed)
    with pytest.warns(UserWarning):
        net = LPNet(pretrained=None, lin_path="/nonexistent").cuda()
    net.load_trunk(P)
    with torch.no_grad():
        for k in range(5):
            net.lins[k].weight.copy_(P[f"lins.{k}.weight"])
    return net, P


def test_lpips_matches_reference_golden(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "lpips_2x32x48.npz"))
    net, P = _net(int(g["seed"]))
    lins = torch.from_numpy(g["lins"])
    o = 0
    with torch.no_grad():
        for k, c in enumerate((64, 128, 256, 512, 512)):      # the reference's shipped vgg.pth linear weights
            net.lins[k].weight.copy_(lins[o:o + c].reshape(1, c, 1, 1)); o += c
    in0 = torch.from_numpy(g["in0"]).cuda().requires_grad_(True)
    in1 = torch.from_numpy(g["in1"]).cuda()
    val = net(in0, in1)
    val.backward()
    want, gw = float(g["loss"]), torch.from_numpy(g["grad_in0"])
    rel = abs(float(val) - want) / want
    gl2 = float((in0.grad.cpu() - gw).norm() / gw.norm())
    cos = float(torch.nn.functional.cosine_similarity(in0.grad.cpu().flatten(), gw.flatten(), dim=0))
    print(f"lpips {float(val):.6f} vs reference {want:.6f} (rel {rel:.2e}); grad rel L2 {gl2:.3e} cosine {cos:.5f}")
    assert rel <= 2e-2 and gl2 <= 0.15 and cos >= 0.99          # bf16 trunk (13 layers) against the fp32 reference


@pytest.mark.parametrize("H,W,batch", [(37, 29, 1), (36, 28, 1), (64, 80, 2), (16, 16, 1), (160, 160, 1)])
def test_lpips_matches_oracle_on_other_sizes(H, W, batch):
    net, P = _net(5)
    g = torch.Generator().manual_seed(H)
    in0 = torch.rand(batch, H, W, 3, generator=g)
    in1 = torch.rand(batch, H, W, 3, generator=g)
    a = in0.cuda().requires_grad_(True)
    val = net(a, in1.cuda())
    val.backward()
    b = in0.clone().requires_grad_(True)
    want = O.lpips(P, b, in1)
    want.backward()
    rel = abs(float(val) - float(want)) / float(want)
    cos = float(torch.nn.functional.cosine_similarity(a.grad.cpu().flatten(), b.grad.flatten(), dim=0))
    print(f"lpips {H}x{W}x{batch}: {float(val):.6f} vs oracle {float(want):.6f} (rel {rel:.2e}); grad cosine {cos:.5f}")
    assert rel <= 2e-2 and cos >= 0.97          # 13 bf16 layers: a few percent of the ReLU masks flip against fp32
    assert float(net(in1.cuda(), in1.cuda())) == 0.0            # identical images: exactly zero


def test_training_step_with_lpips_loss():
    """train.py:171 with the shipped loss mix (mse 1.0 + lpips 0.01, default.yml:155-158) through get_loss."""
    from papr_b200.config import make_config
    from papr_b200.lpips import get_loss
    from papr_b200.model import PAPR
    from papr_b200.scene import learned_like_cloud, synthetic_scene
    cfg = make_config("chair", geoms=dict(points=dict(init_num=1500)))
    model = PAPR(cfg, device="cuda").cuda()
    cloud = learned_like_cloud(1500, cfg.dataset.coord_scale, seed=1)
    with torch.no_grad():
        model.points.copy_(cloud["points"]); model.pc_feats.copy_(cloud["pc_feats"]); model.points_influ_scores.copy_(cloud["points_influ_scores"])
    with pytest.warns(UserWarning):
        loss_fn = get_loss(cfg.training.losses, lin_path="/nonexistent").cuda()
    assert set(loss_fn.losses_and_weights.keys()) == {"mse/1e+00", "lpips/1e-02"}
    b = {k: v.cuda() for k, v in synthetic_scene(48, 48, cfg.dataset.coord_scale, n_views=1, seed=2).items()}
    model.clear_grad()
    out = model.last_act(model(b["rays_o"], b["rays_d"], b["c2w"]))
    loss = loss_fn(out, b["target"])
    loss.backward()
    mse = torch.mean((out - b["target"]) ** 2)
    assert float(loss) > float(mse) and torch.isfinite(loss)
    assert model.points.grad is not None and float(model.points.grad.abs().max()) > 0
    model.step(0)
