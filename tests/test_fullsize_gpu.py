"""GPU: the BASELINE-size workload (800x800 rays, 30,000 points, K=20), checked through size-independent properties and
against the CPU oracle on a random sample of rays (the oracle cannot run the whole frame in seconds)."""
import pytest
import torch

from oracle import papr_oracle as O
from papr_b200.config import make_config

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frame():
    from papr_b200.model import PAPR
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = make_config("chair", use_amp=False)
    P = 30000
    cfg.geoms.points["init_num"] = P
    params = O.init_params(cfg, P, seed=5, cloud="shell")
    rays_o, rays_d, c2w = O.synthetic_rays(800, 800, cfg.dataset.coord_scale, n_views=1, seed=2)
    model = PAPR(cfg, device="cuda", precision="bf16").cuda()
    model.load_my_state_dict({k: v.clone() for k, v in params.items()})
    with torch.no_grad():
        fused, attn = model.evaluate(rays_o.cuda(), rays_d.cuda(), c2w.cuda())
        idx = model._idx32.clone()
    torch.cuda.synchronize()
    return dict(cfg=cfg, params=params, rays_o=rays_o, rays_d=rays_d, model=model, fused=fused, attn=attn, idx=idx)


def test_fullsize_topk_exact_on_sampled_rays(frame):
    g = torch.Generator().manual_seed(0)
    pick = torch.randint(0, 800 * 800, (3000,), generator=g)
    rd = frame["rays_d"].reshape(-1, 3)[pick].reshape(1, 1, -1, 3)
    want, _ = O.select_topk(frame["rays_o"], rd, frame["params"]["points"], 20)
    got = frame["idx"].reshape(-1, 20)[pick.cuda()].cpu().long()
    assert torch.equal(got, want.reshape(-1, 20))


def test_fullsize_topk_structural_properties(frame):
    idx = frame["idx"].reshape(-1, 20)
    assert int(idx.min()) >= 0 and int(idx.max()) < 30000
    srt = torch.sort(idx, dim=-1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all()), "a ray selected the same point twice"
    # the selected points are ordered by distance: recompute the distances on the device in float64
    pts = frame["params"]["points"].cuda().double()
    o = frame["rays_o"].cuda().double()[0]
    d = frame["rays_d"].cuda().double().reshape(-1, 3)
    sub = torch.arange(0, d.shape[0], 97, device="cuda")
    v = pts[idx[sub].long()] - o
    dd = d[sub].unsqueeze(1)
    t = (v * dd).sum(-1, keepdim=True) / ((dd * dd).sum(-1, keepdim=True) + 1e-6)
    dist = (v - dd * t).norm(dim=-1)
    assert bool((dist[:, 1:] >= dist[:, :-1] - 1e-5).all())


def test_fullsize_attention_matches_oracle_on_sampled_rays(frame):
    g = torch.Generator().manual_seed(1)
    pick = torch.randint(0, 800 * 800, (1024,), generator=g)
    rd = frame["rays_d"].reshape(-1, 3)[pick].reshape(1, 32, 32, 3)
    idx = frame["idx"].reshape(-1, 20)[pick.cuda()].cpu().long().reshape(1, 32, 32, 20)
    with torch.no_grad():
        want = O.attention_features(frame["params"], frame["cfg"], frame["rays_o"], rd, idx=idx)
    got_fused = frame["fused"].reshape(-1, 32)[pick.cuda()].cpu()
    got_attn = frame["attn"].reshape(-1, 21)[pick.cuda()].cpu()
    assert float((got_attn - want["attn"].reshape(-1, 21)).abs().max()) <= 7e-3      # stated bf16 tolerance (test_model_gpu.py)
    scale = float(want["fused"].abs().max())
    assert float((got_fused - want["fused"].reshape(-1, 32)).abs().max()) <= 1.5e-2 * scale


def test_fullsize_attention_weights_are_a_distribution(frame):
    attn = frame["attn"].reshape(-1, 21)
    assert bool(torch.isfinite(attn).all()) and bool(torch.isfinite(frame["fused"]).all())
    assert float((attn.sum(-1) - 1).abs().max()) < 1e-5 and float(attn.min()) >= 0


def test_fullsize_tiles_equal_full_frame(frame):
    """Rays are independent: evaluating a 100x100 tile (test.py's tiling) gives the full-frame values for those rays."""
    m = frame["model"]
    ro, rd = frame["rays_o"].cuda(), frame["rays_d"].cuda()
    with torch.no_grad():
        f, a = m.evaluate(ro, rd[:, 300:400, 500:600].contiguous(), None)
    assert torch.equal(m._idx32, frame["idx"][:, 300:400, 500:600])
    assert float((f - frame["fused"][:, 300:400, 500:600]).abs().max()) == 0.0
    assert float((a - frame["attn"][:, 300:400, 500:600]).abs().max()) == 0.0


# ----------------------------------------------------------------------------- RGB and gradients at the full size
TILE = (352, 416, 384, 448)      # interior rows/cols of the checked tile (64 x 64), away from the frame border
HALO = 16                        # UNet influence radius is 13 px: tile + halo reproduces the full frame


def _tile_oracle(frame, with_grad):
    """The CPU oracle on tile + halo (96 x 96 rays of the 800x800 frame, all 30,000 points).  The loss is the MSE over the
    tile interior only, so its gradients equal those of the same masked loss on the whole frame."""
    r0, r1, c0, c1 = TILE
    rd = frame["rays_d"][:, r0 - HALO:r1 + HALO, c0 - HALO:c1 + HALO].contiguous()
    tgt = torch.rand(1, r1 - r0, c1 - c0, 3, generator=torch.Generator().manual_seed(11))
    params = {k: v.clone().requires_grad_(with_grad and v.dtype.is_floating_point and k != "bkg_feats")
              for k, v in frame["params"].items()}
    with torch.set_grad_enabled(with_grad):
        out = O.forward(params, frame["cfg"], frame["rays_o"], rd)
        rgb = out["rgb"][:, HALO:-HALO, HALO:-HALO]
        loss = ((rgb - tgt) ** 2).mean()
        if with_grad:
            loss.backward()
    grads = {k: params[k].grad for k in ("points", "points_influ_scores", "pc_feats")} if with_grad else None
    return rd, tgt, rgb.detach(), float(loss), grads, out["fused"].detach(), out["attn"].detach()


@pytest.fixture(scope="module")
def tile_truth(frame):
    return _tile_oracle(frame, with_grad=True)


def test_fullsize_rgb_and_gradients_bf16_whole_frame(frame, tile_truth):
    """Product path on the WHOLE 800x800 frame (P=30k): forward + backward of a loss masked to the tile; RGB on the tile
    and the point / feature / influence gradients against the CPU oracle (stated bf16 tolerances, tests/test_model_gpu.py)."""
    from tests.parity import rel_err
    r0, r1, c0, c1 = TILE
    _, tgt, want_rgb, want_loss, want_g, _, _ = tile_truth
    m = frame["model"]
    m.clear_grad()
    rgb = m(frame["rays_o"].cuda(), frame["rays_d"].cuda(), None)
    tile = rgb[:, r0:r1, c0:c1]
    loss = ((tile - tgt.cuda()) ** 2).mean()
    loss.backward()
    e_rgb = float((tile.detach().cpu() - want_rgb).abs().max())
    print(f"full frame bf16: tile rgb max-abs err {e_rgb:.2e}, loss {loss.item():.6f} vs {want_loss:.6f}")
    assert e_rgb <= 9e-3             # stated bf16 tolerance; measured 1.5e-3 here
    assert abs(loss.item() - want_loss) <= 5e-4
    for attr in ("points", "points_influ_scores", "pc_feats"):
        got, want = getattr(m, attr).grad.cpu(), want_g[attr]
        l2 = float((got - want).double().norm()) / float(want.double().norm())
        cos = float(torch.nn.functional.cosine_similarity(got.flatten().double(), want.flatten().double(), dim=0))
        print(f"   {attr}: relative L2 {l2:.3e} cosine {cos:.5f} max-abs/max {rel_err(got, want):.3e}")
        assert l2 <= 0.1 and cos >= 0.995, (attr, l2, cos)     # measured: L2 <= 4.6e-2, cosine >= 0.9989
        assert int((got != 0).sum()) > 0 and bool(((want != 0) | (got.abs() <= 1e-3 * float(want.abs().max()))).all()), \
            "a point outside the tile's receptive field received a gradient"
    m.clear_grad()


def test_fullsize_rgb_and_gradients_fp32_tile(frame, tile_truth):
    """Parity mode (fp32-accurate GEMMs on the library's own tensor-core kernels) on the same tile + halo rays: attention
    weights / features 1e-5, RGB 1e-4 (north star: 1e-3), gradients within the reference's own fp32 noise."""
    from papr_b200.model import PAPR
    from tests.parity import rel_err
    rd, tgt, want_rgb, want_loss, want_g, want_fused, want_attn = tile_truth
    cfg = frame["cfg"]
    m = PAPR(cfg, device="cuda", precision="fp32").cuda()
    m.load_my_state_dict({k: v.clone() for k, v in frame["params"].items()})
    m.clear_grad()
    rgb = m(frame["rays_o"].cuda(), rd.cuda(), None)
    tile = rgb[:, HALO:-HALO, HALO:-HALO]
    loss = ((tile - tgt.cuda()) ** 2).mean()
    loss.backward()
    with torch.no_grad():
        fused, attn = m.evaluate(frame["rays_o"].cuda(), rd.cuda(), None)
    e_rgb = float((tile.detach().cpu() - want_rgb).abs().max())
    e_attn = float((attn.squeeze(-1).cpu() - want_attn).abs().max())
    e_fused = rel_err(fused.squeeze(-2).cpu(), want_fused)
    print(f"tile fp32: rgb {e_rgb:.2e} attn {e_attn:.2e} fused {e_fused:.2e} loss {loss.item():.7f} vs {want_loss:.7f}")
    assert e_attn <= 1e-5 and e_fused <= 1e-5 and e_rgb <= 1e-4
    assert abs(loss.item() - want_loss) <= 1e-5
    for attr in ("points", "points_influ_scores", "pc_feats"):
        got, want = getattr(m, attr).grad.cpu(), want_g[attr]
        e = rel_err(got, want)
        print(f"   {attr}: max-abs/max vs reference fp32 {e:.3e}")
        assert e <= 5e-3, (attr, e)


# ----------------------------------------------------------------------------- configs[3]: Caterpillar shape
def test_caterpillar_shape_select_exact_and_chunked_training_equals_unchunked():
    """BASELINE configs[3] shape (1920x1080 frame, P=100,000 points, L=4, coord_scale 30; configs/t2/Caterpillar.yml:2-22):
    the grid selection over the whole frame is bit-exact against the C oracle on sampled rays, and training a crop of the
    frame in checkpointed ray chunks gives the unchunked result (outputs equal, gradients up to atomics reordering)."""
    from papr_b200.model import PAPR
    from papr_b200 import ops
    cfg = make_config("caterpillar", use_amp=False)
    P = 100000
    cfg.geoms.points["init_num"] = P
    params = O.init_params(cfg, P, seed=8, cloud="shell")
    rays_o, rays_d, c2w = O.synthetic_rays(1080, 1920, cfg.dataset.coord_scale, n_views=1, seed=3)
    idx = ops.select_topk(rays_o.cuda(), rays_d.cuda(), params["points"].cuda(), 20)
    pick = torch.randint(0, 1080 * 1920, (1500,), generator=torch.Generator().manual_seed(4))
    want, _ = O.select_topk(rays_o, rays_d.reshape(-1, 3)[pick].reshape(1, 1, -1, 3), params["points"], 20)
    assert torch.equal(idx.reshape(-1, 20)[pick.cuda()].cpu().long(), want.reshape(-1, 20))
    # The same with the candidate list cut to K + 1 entries (one neighbour of margin) and at its full 32: at this scale
    # eps |v|^2 = 0.016 is thirty times the gap between neighbouring keys, so a phase-1 key that is not the exact key
    # times den up to rounding (the eps (v.d)^2 / den term) loses true neighbours here -- and did, until it was added.
    import os
    for last in ("20", "31"):
        os.environ["PAPR_SELECT_LAST"] = last
        try:
            alt = ops.select_topk(rays_o.cuda(), rays_d.cuda(), params["points"].cuda(), 20)
        finally:
            os.environ.pop("PAPR_SELECT_LAST", None)
        assert torch.equal(alt, idx), last
    model = PAPR(cfg, device="cuda", precision="bf16").cuda()
    model.load_my_state_dict({k: v.clone() for k, v in params.items()})
    assert model.proximity_attn.L == 4 and model.proximity_attn.dk == 81 and model.proximity_attn.dv == 118
    crop = rays_d[:, 400:580, 800:1120].contiguous().cuda()          # 180 x 320 rays
    tgt = torch.rand(1, 180, 320, 3, generator=torch.Generator().manual_seed(6)).cuda()

    def run(chunk):
        model.ray_chunk = chunk
        model.clear_grad()
        out = model(rays_o.cuda(), crop, None)
        torch.mean((out - tgt) ** 2).backward()
        return out.detach().clone(), {k: getattr(model, k).grad.detach().clone() for k in ("points", "pc_feats", "points_influ_scores")}

    out_a, g_a = run(10 ** 9)
    out_b, g_b = run(12800)            # 5 chunks, ragged last one
    assert float((out_a - out_b).abs().max()) < 1e-5
    for k in g_a:
        assert float((g_a[k] - g_b[k]).abs().max()) <= 2e-2 * float(g_a[k].abs().max()), k
    # against the oracle on a small tile of the crop (features are what the L=4 row kernels produce)
    tile = rays_d[:, 480:496, 900:924].contiguous()
    with torch.no_grad():
        f, a = model.evaluate(rays_o.cuda(), tile.cuda(), None)
        w = O.attention_features(params, cfg, rays_o, tile, idx=model.select_k_ind.cpu())
    assert float((a.squeeze(-1).cpu() - w["attn"]).abs().max()) <= 7e-3
    assert float((f.squeeze(-2).cpu() - w["fused"]).abs().max()) <= 1.5e-2 * float(w["fused"].abs().max())
