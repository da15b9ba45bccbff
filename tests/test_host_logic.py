"""CPU tests of the host side: C-ABI surface, config presets, LR schedules, model container, bookkeeping."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

from papr_b200.config import Config, make_config, merge

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_exports_every_declared_symbol(built_library):
    from papr_b200 import _lib
    header = open(os.path.join(ROOT, "include", "papr_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char \*)\s*(papr_[a-z0-9_]+)\s*\(", header, re.M))
    assert declared, "no declarations found"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/papr_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.lib()
    assert lib.papr_abi_version() == 2
    assert lib.papr_status_string(-1) == b"invalid argument"


def test_missing_library_fails_loudly(monkeypatch):
    from papr_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpapr_b200.so")
    with pytest.raises(_lib.PaprError):
        _lib.lib()


def test_ops_refuse_cpu_tensors():
    from papr_b200 import ops
    with pytest.raises(AssertionError):
        ops.select_topk(torch.zeros(1, 3), torch.zeros(1, 2, 2, 3), torch.zeros(100, 3), 20)


def test_config_presets_and_merge_semantics():
    c = make_config("chair")
    assert c.geoms.points.init_num == 10000 and c.training.add_start == 10000 and c.geoms.points.select_k == 20
    t = make_config("caterpillar")
    assert t.models.attn.embed.k_L == [4, 4, 4] and t.geoms.background.constant == 4.0 and t.use_amp is False
    e = make_config("caterpillar_exposure")
    assert e.exposure_control.use and e.training.lr.lr_factor == 0.2
    c.geoms.points["init_num"] = 7
    assert c.geoms.points.init_num == 7 and "max_points" not in c and "max_num_pts" in c
    base = dict(test=dict(datasets=[dict(name="testset", path="a", factor=1)]))
    merge(base, dict(test=dict(datasets=[dict(name="testset", path="b"), dict(name="extra", path="c")])))
    assert base["test"]["datasets"][0] == dict(name="testset", path="b", factor=1)
    assert base["test"]["datasets"][1] == dict(name="extra", path="c", factor=1)
    assert isinstance(Config(base).test.datasets[0], Config)


@pytest.mark.parametrize("kind,warmup", [("cosine-hlfperiod", 10), ("cosine", 0), ("linear", 5), ("cosine", 7), ("linear", 0),
                                         ("cosine-hlfperiod", 0)])
def test_schedule_matches_torch_sequential_lr(kind, warmup):
    """The closed form equals the reference's SequentialLR([LinearLR, decay]) composition (models/utils.py:260-322),
    including an O(1) fast-forward to an arbitrary step (reference: .step() x total_steps, model.py:175-179)."""
    import torch.optim.lr_scheduler as S
    from papr_b200.schedule import create_learning_rate_fn
    max_steps, base = 60, 3e-4
    opt = Config(type=kind, base_lr=base, factor=1, warmup=warmup, weight_decay=0)

    def reference():
        p = torch.nn.Parameter(torch.zeros(1))
        o = torch.optim.Adam([p], lr=base)
        w = S.LinearLR(o, start_factor=1e-16 if warmup > 0 else 1.0, end_factor=1.0, total_iters=warmup)
        if kind == "linear":
            d = S.LinearLR(o, start_factor=1.0, end_factor=0.0, total_iters=max_steps - warmup)
        else:
            T = max(max_steps - warmup, 1) * (2 if kind == "cosine-hlfperiod" else 1)
            d = S.CosineAnnealingLR(o, T_max=T)
        return o, S.SequentialLR(o, schedulers=[w, d], milestones=[warmup])

    o_ref, s_ref = reference()
    p = torch.nn.Parameter(torch.zeros(1))
    o = torch.optim.Adam([p], lr=base)
    s = create_learning_rate_fn(o, max_steps, opt)
    lrs = []
    for t in range(max_steps):
        assert math.isclose(o.param_groups[0]["lr"], o_ref.param_groups[0]["lr"], rel_tol=1e-9, abs_tol=1e-18), (t, kind)
        lrs.append(o.param_groups[0]["lr"])
        o.step(); o_ref.step(); s.step(); s_ref.step()
    for jump in (1, warmup, 23, 59):
        p2 = torch.nn.Parameter(torch.zeros(1))
        o2 = torch.optim.Adam([p2], lr=base)
        s2 = create_learning_rate_fn(o2, max_steps, opt, start_step=jump)
        assert math.isclose(o2.param_groups[0]["lr"], lrs[jump], rel_tol=1e-9, abs_tol=1e-18)
        assert math.isclose(s2.get_last_lr()[0], lrs[jump], rel_tol=1e-9, abs_tol=1e-18)
    assert create_learning_rate_fn(o, max_steps, Config(type="none", base_lr=1, factor=1, warmup=0, weight_decay=0)) is None


def _small_model(**over):
    from papr_b200.model import PAPR
    cfg = make_config("chair", geoms=dict(points=dict(init_num=343)), **over)
    return PAPR(cfg, device="cpu"), cfg


def test_model_container_matches_reference_surface(tmp_path):
    model, cfg = _small_model()
    sd = model.state_dict()
    for key in ("points", "points_influ_scores", "bkg_feats", "pc_feats", "select_k",
                "proximity_attn.embed.embed_k.innorm.a_2", "proximity_attn.embed.embed_q.outnorm.b_2",
                "proximity_attn.embed.embed_k.mlp.model.9.weight", "proximity_attn.embed.embed_v.mlp.model.15.bias",
                "proximity_attn.attention_layer.w_k.weight", "proximity_attn.attention_layer.w_q.bias",
                "renderer.inc.double_conv.0.weight", "renderer.down2.maxpool_conv.1.double_conv.0.bias",
                "renderer.up1.up.weight", "renderer.up2.conv.double_conv.0.weight", "renderer.outc.conv.bias"):
        assert key in sd, key
    assert len(sd) == 69
    assert sd["proximity_attn.embed.embed_k.mlp.model.1.weight"].shape == (256, 117)
    assert sd["proximity_attn.embed.embed_v.mlp.model.1.weight"].shape == (256, 142)
    assert sd["proximity_attn.embed.embed_q.mlp.model.1.weight"].shape == (256, 39)
    assert sd["proximity_attn.embed.embed_v.mlp.model.15.weight"].shape == (32, 256)
    assert sum(p.numel() for p in model.proximity_attn.parameters()) == 1139288
    assert sum(p.numel() for p in model.renderer.parameters()) == 3643395
    assert model.points.shape == (343, 3) and float(model.points.abs().max()) == pytest.approx(12.0)
    assert set(model.optimizers) == {"points", "attn", "points_influ_scores", "pc_feats", "renderer"}
    assert model.select_k.dtype == torch.int32 and int(model.select_k) == 20 and model.bkg_score.shape == (1,)
    assert not model.bkg_feats.requires_grad
    # optimiser step + lr bookkeeping (model.py:439-460)
    for p in model.parameters():
        if p.requires_grad:
            p.grad = torch.ones_like(p)
    before = model.points.detach().clone()
    model.step(0)
    assert not torch.equal(before, model.points) and model.pts_lr > 0 and model.attn_lr >= 0
    model.clear_grad()
    # checkpoint round trip incl. a changed point count (model.py:562-641)
    model.save(5, str(tmp_path))
    other, _ = _small_model()
    other.prune_points(-1.0) if False else None
    other.points = torch.nn.Parameter(torch.zeros(10, 3))
    step = other.load(str(tmp_path), load_optimizer=False)
    assert step == 5 and other.points.shape == (343, 3)
    for k, v in model.state_dict().items():
        assert torch.equal(v, other.state_dict()[k]), k


def test_exposure_and_mlp_decode_variants_construct():
    model, _ = _small_model(exposure_control=dict(use=True))
    assert "mapping_mlp.model.model.15.weight" in model.state_dict() and "mapping_mlp" in model.optimizers
    assert model.mapping_mlp(torch.zeros(128)).shape == (64,)
    assert float(model.mapping_mlp(torch.zeros(128)).min()) >= 1.0      # relu+1 (mlp.py:62-78)
    m2, _ = _small_model(models=dict(use_renderer=False, attn=dict(embed=dict(value=dict(d_ff_out=3)))))
    assert not hasattr(m2, "renderer") and "renderer" not in m2.optimizers
    with pytest.raises(AssertionError):
        _small_model(models=dict(use_renderer=False))
    with pytest.raises((KeyError, AttributeError)):       # generator.type 'mlp' has no config block in the stock YAML, as in the reference
        _small_model(models=dict(renderer=dict(generator=dict(type="mlp"))))


def test_forward_without_gpu_raises_instead_of_falling_back():
    model, _ = _small_model()
    rays_o, rays_d = torch.zeros(1, 3), torch.nn.functional.normalize(torch.randn(1, 4, 4, 3), dim=-1)
    with pytest.raises((AssertionError, RuntimeError)):
        model(rays_o, rays_d, torch.eye(4)[None])


def test_prune_and_add_points_follow_reference_rules():
    model, cfg = _small_model()
    torch.manual_seed(0)
    with torch.no_grad():
        model.points_influ_scores.copy_(torch.randn(343, 1))
        model.points.copy_(torch.randn(343, 3) * 5)          # the lattice init has exact distance ties; avoid them here
    keep = int((model.points_influ_scores[:, 0] > 0).sum())
    n = model.prune_points(0.0)
    assert int(n) == 343 - keep and model.points.shape[0] == keep == model.pc_feats.shape[0] == model.points_influ_scores.shape[0]
    # growth: same rule as reference add_points_knn (models/utils.py:9-109), checked against scipy's KDTree here
    from scipy.spatial import KDTree
    pts = model.points.detach().clone()
    influ, feats = model.points_influ_scores.detach().clone(), model.pc_feats.detach().clone()
    np.random.seed(3)
    added = model.add_points(25)
    assert added == 25 and model.points.shape[0] == keep + 25
    np.random.seed(3)
    tree = KDTree(pts.numpy())
    nd, _ = tree.query(pts.numpy(), k=10)
    inds = np.argsort(nd.std(axis=-1))[-25:]
    _, ni = tree.query(pts.numpy()[inds], k=4)
    ni = ni[:, 1:]
    w = np.random.uniform(0, 1, (25, 3)).astype(np.float32)
    w /= w.sum(axis=-1, keepdims=True)
    want = (pts.numpy()[ni] * w[:, :, None]).sum(-2)
    got = model.points.detach()[keep:].numpy()
    assert np.allclose(np.sort(got, axis=0), np.sort(want, axis=0), atol=1e-5)
    want_f = (feats.numpy()[ni] * w[:, :, None]).sum(-2)
    assert np.allclose(np.sort(model.pc_feats.detach()[keep:].numpy(), axis=0), np.sort(want_f, axis=0), atol=1e-5)
    # optimisers are rebuilt around the new tensors at the right schedule position (train.py:207-250)
    model.clear_optimizer(); model.clear_scheduler(); model.init_optimizers(1234)
    assert model.optimizers["points"].param_groups[0]["params"][0] is model.points
    lr = model.optimizers["attn"].param_groups[0]["lr"]
    assert lr == pytest.approx(3e-4 * (1e-16 + (1 - 1e-16) * 1234 / 10000))


def test_unet_matches_oracle_restatement():
    from oracle import papr_oracle as O
    from papr_b200.renderer import SmallUNet
    torch.manual_seed(0)
    net = SmallUNet(32, 3, compute_dtype=torch.float32, affine_layer=0)
    params = {"renderer." + k: v for k, v in net.state_dict().items()}
    x = torch.randn(1, 32, 12, 20)
    g, b = torch.rand(32) + 0.5, torch.randn(32)
    assert torch.allclose(net(x, gamma=g, beta=b), O.unet(params, x, g, b, affine_layer=0), atol=1e-5)


def test_row_sharding_covers_frame_with_halo():
    from papr_b200.dist import shard_rows
    for H, world in ((800, 8), (800, 2), (1080, 8), (100, 3)):
        covered = []
        for r in range(world):
            r0, r1, h0, h1 = shard_rows(H, world, r)
            assert h0 <= r0 <= r1 <= h1 <= H and h0 % 4 == 0
            assert (r0 - h0 >= 16 or h0 == 0) and (h1 - r1 >= 16 or h1 == H)
            covered += list(range(r0, r1))
        assert covered == list(range(H))


def test_get_rays_matches_reference_when_available():
    """papr_b200.scene.get_rays against the reference's dataset/utils.py:81-96 (only where /root/reference is mounted;
    the formula itself is also pinned by the ray directions stored in tests/golden/*.npz)."""
    import math
    import os
    import torch
    from papr_b200.scene import get_rays, look_at_poses
    path = "/root/reference/dataset/utils.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    import sys
    sys.path.insert(0, "/root/reference")
    import importlib
    import types
    stubs, mod = [], None
    try:
        for _ in range(8):          # image-IO packages the loaders import (imageio, cv2, ...) are not needed by get_rays
            try:
                mod = importlib.import_module("dataset.utils")
                break
            except ModuleNotFoundError as e:
                sys.modules[e.name] = types.ModuleType(e.name)
                stubs.append(e.name)
    except Exception as e:
        pytest.skip(f"reference dataset utils not importable: {e}")
    finally:
        sys.path.remove("/root/reference")
        for name in stubs:
            sys.modules.pop(name, None)
    if mod is None:
        pytest.skip("reference dataset utils not importable")
    c2w = look_at_poses(3, seed=4)
    H, W = 37, 53
    focal = 0.5 * W / math.tan(0.5 * 0.6911)
    ro, rd = get_rays(H, W, focal, c2w)
    ro_ref, rd_ref = mod.get_rays(H, W, focal, focal, c2w)
    assert torch.equal(ro, ro_ref)
    assert float((rd - rd_ref).abs().max()) <= 1e-6
