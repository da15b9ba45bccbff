"""GPU: the between-step kernels -- exact kNN and stream-compaction prune at P=100k against
scipy / boolean masks, on-device ray generation against the reference's get_rays formula, the fused multi-tensor Adam
against torch.optim.Adam, and the batched weight-image pack against the per-layer one."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(P, seed=0):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(P, 3, generator=g)
    p = p / p.norm(dim=-1, keepdim=True) * 24.0 + 0.6 * torch.randn(P, 3, generator=g)
    return p.float()


@pytest.mark.parametrize("P,Q,k", [(100000, 100000, 10), (100000, 1000, 4), (777, 777, 32), (40, 40, 11)])
def test_knn_matches_scipy_kdtree(P, Q, k):
    from scipy.spatial import KDTree
    from papr_b200 import ops
    pts = _cloud(P, seed=P)
    qs = pts[torch.randperm(P, generator=torch.Generator().manual_seed(1))[:Q]] if Q < P else pts
    dist, idx = ops.knn(pts.cuda(), qs.cuda(), k)
    want_d, want_i = KDTree(pts.numpy().astype(np.float64)).query(qs.numpy().astype(np.float64), k=k)
    dist, idx = dist.cpu().numpy(), idx.cpu().numpy()
    assert np.allclose(dist, want_d, rtol=1e-13, atol=0)
    same = idx == want_i
    if not same.all():          # only exact distance ties may be resolved differently
        r, c = np.nonzero(~same)
        assert np.all(dist[r, c] == want_d[r, c])
        for q in np.unique(r):
            assert sorted(idx[q]) == sorted(want_i[q]) or dist[q, k - 1] == want_d[q, k - 1]
    assert np.all(np.diff(dist, axis=1) >= 0)


@pytest.mark.parametrize("P,F", [(100000, 64), (100000, 0), (1000, 128), (5, 64)])
@pytest.mark.parametrize("keep_less", [False, True])
def test_prune_compact_matches_boolean_mask(P, F, keep_less):
    from papr_b200 import ops
    g = torch.Generator().manual_seed(P + F)
    pts = torch.randn(P, 3, generator=g).cuda()
    influ = torch.randn(P, 1, generator=g).cuda()
    feats = torch.randn(P, F, generator=g).cuda() if F else None
    for thresh in (0.0, -10.0, 10.0):
        mask = influ[:, 0] < thresh if keep_less else influ[:, 0] > thresh
        p2, i2, f2, n = ops.prune_compact(pts, influ, feats, thresh, keep_less)
        assert n == int(mask.sum()) and p2.shape == (n, 3) and i2.shape == (n, 1)
        assert torch.equal(p2, pts[mask]) and torch.equal(i2, influ[mask])
        if F:
            assert torch.equal(f2, feats[mask])


def test_model_prune_and_add_at_100k_follow_reference_rules():
    """PAPR.prune_points / add_points on a 100,000-point cloud on the device == the reference rules evaluated with a host
    mask and scipy's KDTree (models/model.py:335-394, models/utils.py:9-109)."""
    from scipy.spatial import KDTree
    from papr_b200.config import make_config
    from papr_b200.model import PAPR
    P = 100000
    cfg = make_config("caterpillar", geoms=dict(points=dict(init_num=P)), max_num_pts=200000)
    model = PAPR(cfg, device="cuda").cuda()
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        model.points.copy_(_cloud(P, seed=5)); model.points_influ_scores.copy_(torch.randn(P, 1, generator=g))
        model.pc_feats.copy_(torch.randn(P, 64, generator=g))
    keep = (model.points_influ_scores[:, 0] > 0).cpu()
    before = {k: getattr(model, k).detach().cpu().clone() for k in ("points", "points_influ_scores", "pc_feats")}
    n = int(model.prune_points(0.0))
    assert n == P - int(keep.sum())
    for k, v in before.items():
        assert torch.equal(getattr(model, k).detach().cpu(), v[keep]), k
    pts = model.points.detach().cpu().numpy().astype(np.float64)
    influ, feats = model.points_influ_scores.detach().cpu(), model.pc_feats.detach().cpu()
    np.random.seed(7)
    added = model.add_points(500)
    assert added == 500
    # the reference rule with scipy: top-knn-std sampling over k=10 neighbours, then random convex combination of 3 neighbours
    tree = KDTree(pts)
    nd, _ = tree.query(pts, k=10)
    order = np.argsort(nd.std(axis=-1), kind="stable")[-500:]
    _, ni = tree.query(pts[order], k=4)
    ni = ni[:, 1:]
    np.random.seed(7)
    w = np.random.uniform(0, 1, (500, 3)).astype(np.float32)
    w = w / w.sum(axis=-1, keepdims=True)
    want = (torch.from_numpy(pts.astype(np.float32))[torch.from_numpy(ni)] * torch.from_numpy(w)[..., None]).sum(-2)
    got = model.points.detach().cpu()[-500:]
    # a different argsort tie order would permute rows: compare as sets of rows
    d = torch.cdist(got.double(), want.double())
    assert float(d.min(dim=1).values.max()) < 1e-4 and float(d.min(dim=0).values.max()) < 1e-4
    want_influ = (influ[torch.from_numpy(ni)] * torch.from_numpy(w)[..., None]).sum(-2)
    assert abs(float(model.points_influ_scores.detach().cpu()[-500:].sum()) - float(want_influ.sum())) < 1e-2
    assert model.pc_feats.shape == (model.points.shape[0], 64)


@pytest.mark.parametrize("H,W,window", [(800, 800, None), (1080, 1920, (300, 480, 1000, 1180)), (13, 7, None), (100, 160, (0, 100, 80, 160))])
def test_device_ray_generation_matches_get_rays(H, W, window):
    import math
    from papr_b200 import ops
    from papr_b200.scene import get_rays, look_at_poses
    c2w = look_at_poses(3, seed=H + W)
    focal = 0.5 * W / math.tan(0.5 * 0.6911)
    want_o, want_d = get_rays(H, W, focal, c2w)
    h0, h1, w0, w1 = window or (0, H, 0, W)
    ro, rd = ops.generate_rays(c2w.cuda(), H, W, focal, window=window, coord_scale=10.0)
    assert rd.shape == (3, h1 - h0, w1 - w0, 3)
    assert torch.equal(ro.cpu(), want_o * 10.0)
    err = float((rd.cpu() - want_d[:, h0:h1, w0:w1]).abs().max())
    assert err <= 2.5e-7, err            # the same fp32 formula; at most an ulp or two from linspace / sum order
    assert float((rd.norm(dim=-1) - 1).abs().max()) < 1e-6


def test_fused_adam_matches_torch_adam_and_its_state_dict():
    from papr_b200.optim import FlatAdam, FlatAdamBucket
    torch.manual_seed(0)
    shapes = [(1000, 3), (1000, 1), (257,), (256, 117), (3, 128, 1, 1), (1,)]
    ref_p = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    cfgs = [dict(lr=2e-3, weight_decay=0.0), dict(lr=3e-4, weight_decay=1e-2), dict(lr=1e-3, weight_decay=0.0)]
    split = [ref_p[:1], ref_p[1:4], ref_p[4:]]
    ref_opts = [torch.optim.Adam(ps, **c) for ps, c in zip(split, cfgs)]
    bucket = FlatAdamBucket("cuda")
    our_opts = [FlatAdam(ps, bucket, **c) for ps, c in zip([our_p[:1], our_p[1:4], our_p[4:]], cfgs)]
    bucket.finalize()
    for step in range(25):
        bucket.zero_grad()
        for o in ref_opts:
            o.zero_grad()
        for a, b in zip(ref_p, our_p):
            g = torch.randn_like(a) * (10.0 ** (step % 5 - 3))
            a.grad = g.clone()
            b.grad.add_(g)                       # accumulate into the flat view, as autograd does
        if step == 7:                            # a learning-rate change by a scheduler
            for o in ref_opts + our_opts:
                o.param_groups[0]["lr"] *= 0.5
        for o in ref_opts:
            o.step()
        bucket.step_all(our_opts)
    for a, b in zip(ref_p, our_p):
        assert float((a - b).abs().max()) <= 2e-6 * max(1.0, float(a.abs().max())), (a.shape, float((a - b).abs().max()))
    # state-dict layout interchanges with torch.optim.Adam in both directions
    sd = our_opts[1].state_dict()
    fresh = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in our_p[1:4]], lr=1.0)
    fresh.load_state_dict(sd)
    assert float(fresh.state[fresh.param_groups[0]["params"][0]]["step"]) == 25.0
    ref_sd = ref_opts[1].state_dict()
    our_opts[1].load_state_dict(ref_sd)
    assert our_opts[1]._step_count_adam == 25
    i0 = our_opts[1]._indices()[0]
    assert torch.allclose(bucket.view(bucket.flat_m, i0), ref_opts[1].state[ref_p[1]]["exp_avg"], rtol=1e-4, atol=1e-9)
    # only the stepped group moves when a single optimiser is stepped (the torch.optim interface)
    before = [p.detach().clone() for p in our_p]
    for b in our_p:
        b.grad.fill_(1.0)
    our_opts[0].step()
    assert not torch.equal(before[0], our_p[0]) and all(torch.equal(x, y) for x, y in zip(before[1:], our_p[1:]))


def test_model_training_with_fused_optimizer_matches_torch_adam():
    """Three training steps of the whole model: fused one-launch Adam over the flat bucket vs one torch.optim.Adam per group."""
    from papr_b200.config import make_config
    from papr_b200.model import PAPR
    from papr_b200.scene import learned_like_cloud, synthetic_scene
    cfg = make_config("chair", geoms=dict(points=dict(init_num=1200)),
                      training=dict(lr=dict(attn=dict(warmup=0), generator=dict(warmup=0), feats=dict(warmup=0), points_influ_scores=dict(warmup=0))))
    scene = {k: v.cuda() for k, v in synthetic_scene(32, 32, cfg.dataset.coord_scale, n_views=1, seed=2).items()}
    cloud = learned_like_cloud(1200, cfg.dataset.coord_scale, seed=1)
    models = []
    for fused in (True, False):
        torch.manual_seed(4)
        m = PAPR(cfg, device="cuda", fused_optimizer=fused).cuda()
        with torch.no_grad():
            m.points.copy_(cloud["points"]); m.pc_feats.copy_(cloud["pc_feats"]); m.points_influ_scores.copy_(cloud["points_influ_scores"])
        m.init_optimizers(0)
        assert (m._flat is not None) == fused
        models.append(m)
    models[1].load_state_dict(models[0].state_dict())
    fused, plain = models
    for step in range(3):
        fused.clear_grad()
        out = fused(scene["rays_o"], scene["rays_d"], scene["c2w"], step)
        torch.mean((out - scene["target"]) ** 2).backward()          # autograd accumulates into the flat bucket's views
        # the torch.optim.Adam model gets the very same gradients (atomics make a second backward differ in the last bits,
        # and Adam's first steps move every weight by lr * sign(g): a flipped sign of a ~0 gradient would dominate)
        plain.clear_grad()
        for (n, a), (_, b) in zip(fused.named_parameters(), plain.named_parameters()):
            if a.grad is not None:
                assert a.grad.data_ptr() == fused._flat.flat_g.data_ptr() + 4 * fused._flat.offsets[[id(q) for q in fused._flat.params].index(id(a))], n
                b.grad = a.grad.detach().clone()
        fused.step(step)
        plain.step(step)
    assert fused._flat.optimizers[0]._step_count_adam == 3
    for (n, a), (_, b) in zip(fused.named_parameters(), plain.named_parameters()):
        scale = max(float(b.abs().max()), 1.0)
        assert float((a - b).abs().max()) <= 2e-6 * scale, (n, float((a - b).abs().max()), scale)
    assert abs(models[0].attn_lr - models[1].attn_lr) < 1e-12 and abs(models[0].pts_lr - models[1].pts_lr) < 1e-12


def test_batched_weight_images_equal_per_layer_pack():
    from papr_b200 import ops
    from papr_b200.config import make_config
    from papr_b200.model import PAPR
    m = PAPR(make_config("chair", geoms=dict(points=dict(init_num=100))), device="cuda").cuda()
    wi = m.proximity_attn.weight_images
    wi.refresh()
    torch.cuda.synchronize()
    assert len(wi.images) == 2 * (5 + 5 + 8)
    for name, n in (("k", 5), ("q", 5), ("v", 8)):
        lins = getattr(m.proximity_attn.embed, f"embed_{name}").mlp.linears()
        in_pad = ops.pad_cols(lins[0].weight.shape[1])
        fwd, bwd = wi.stack(name, n)
        for i, lin in enumerate(lins):
            n_out, n_in = lin.weight.shape
            want_f = ops.pack_weight(lin.weight, (n_out + 31) // 32 * 32, (n_in + 15) // 16 * 16)
            want_t = ops.pack_weight(lin.weight, n_in if i > 0 else in_pad, (n_out + 15) // 16 * 16, transpose=True)
            for r in range(ops.WEIGHT_REPLICAS):
                assert torch.equal(fwd[i][r], want_f) and torch.equal(bwd[i][r], want_t), (name, i, r)
    with torch.no_grad():                      # the cache follows in-place weight updates
        lins[0].weight.mul_(2.0)
    wi.refresh()
    assert torch.equal(wi.stack("v", 8)[0][0][0], ops.pack_weight(lins[0].weight, 256, 144))
