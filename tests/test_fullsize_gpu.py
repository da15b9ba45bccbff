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
    assert float((got_attn - want["attn"].reshape(-1, 21)).abs().max()) <= 2e-2
    scale = float(want["fused"].abs().max())
    assert float((got_fused - want["fused"].reshape(-1, 32)).abs().max()) <= 4e-2 * scale


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
