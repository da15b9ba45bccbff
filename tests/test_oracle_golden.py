"""CPU: the oracle restatement against the golden vectors recorded from the real reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import papr_oracle as O
from tests.parity import GOLDEN_CASES, golden_params, load_golden, rel_err, topk_sets_match


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_reference_outputs(golden_dir, name):
    g = load_golden(golden_dir, name)
    cfg, params = golden_params(g)
    rays_o, rays_d = torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])
    code = torch.from_numpy(g["shading_code"]) if g["shading_code"].size else None
    K = int(cfg.geoms.points.select_k)
    idx, kth = O.select_topk(rays_o, rays_d, params["points"], K, cfg.eps)
    assert torch.equal(torch.sort(idx, -1).values, torch.from_numpy(g["idx_sorted"]).long())
    assert torch.equal(kth, torch.from_numpy(g["kth"]))
    with torch.no_grad():
        out = O.forward(params, cfg, rays_o, rays_d, shading_code=code, idx=idx)
    # the reference's candidate order is topk(sorted=False)'s; compare attention order-independently
    a = torch.sort(out["attn"][..., :K], -1).values
    b = torch.sort(torch.from_numpy(g["attn"])[..., :K], -1).values
    assert float((a - b).abs().max()) <= 1e-6
    assert rel_err(out["fused"], torch.from_numpy(g["fused"])) <= 2e-5
    assert float((out["rgb"] - torch.from_numpy(g["rgb"])).abs().max()) <= 1e-5


def test_oracle_gradients_match_reference(golden_dir):
    g = load_golden(golden_dir, "chair_2views_8x8_p500")
    cfg, params = golden_params(g)
    rays_o, rays_d = torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])
    pg = {k: v.clone().requires_grad_(v.dtype.is_floating_point and k != "bkg_feats") for k, v in params.items()}
    out = O.forward(pg, cfg, rays_o, rays_d)
    loss = ((out["rgb"] - torch.from_numpy(g["target"])) ** 2).mean()
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-6
    for key, name in (("grad_points", "points"), ("grad_influ", "points_influ_scores"), ("grad_pc_feats", "pc_feats")):
        assert rel_err(pg[name].grad, torch.from_numpy(g[key])) <= 2e-3, name
    for nm, norm in zip(g["wgrad_names"], g["wgrad_norms"]):
        gn = float(pg[str(nm)].grad.double().norm())
        assert abs(gn - norm) <= 1e-3 * max(norm, 1e-9), nm


def test_select_oracle_matches_reference_sets(golden_dir):
    g = load_golden(golden_dir, "select_24x24x2_p3000")
    from papr_b200.config import make_config
    cfg = make_config("chair")
    params = O.init_params(cfg, int(g["P"]), seed=2, cloud="shell")
    rays_o, rays_d = torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])
    idx, kth = O.select_topk(rays_o, rays_d, params["points"], 20)
    assert torch.equal(torch.sort(idx, -1).values, torch.from_numpy(g["idx_sorted"]).long())
    assert torch.equal(kth, torch.from_numpy(g["kth"]))
    # the C restatement and the literal tensor restatement agree bit for bit
    sub = (slice(0, 1), slice(0, 6), slice(0, 6))
    d_c = O.select_distances(rays_o[:1], rays_d[sub], params["points"])
    d_t = O.select_distances_torch(rays_o[:1], rays_d[sub], params["points"])
    assert torch.equal(d_c, d_t)


def test_select_oracle_edge_cases():
    pts = torch.randn(10, 3)
    rays_o, rays_d, _ = O.synthetic_rays(8, 8, 10.0)
    idx, kth = O.select_topk(rays_o, rays_d, pts, 20)        # K >= P bypass (model.py:326-327)
    assert idx.shape == (1, 8, 8, 10) and torch.equal(idx[0, 0, 0], torch.arange(10))
    dup = torch.cat([pts, pts])                               # exact duplicates: ties resolved by index
    idx, _ = O.select_topk(rays_o, rays_d, dup, 4)
    assert bool((idx[..., 1::2] - idx[..., 0::2] == 10).all())   # (distance, index) order: each point, then its copy
    empty_o, empty_d = torch.zeros(1, 3), torch.zeros(1, 0, 4, 3)
    idx, _ = O.select_topk(empty_o, empty_d, pts, 3)
    assert idx.shape == (1, 0, 4, 3)


def test_markstein_division_matches_ieee():
    """The select kernel's division sequence (see select.cu) vs hardware division on adversarial inputs."""
    import ctypes
    lib = O._select_lib()
    rng = np.random.default_rng(0)
    n = 4_000_000
    cases = [
        (rng.standard_normal(n) * 40, 1 + rng.standard_normal(n) * 1e-6 + 1e-6),
        (rng.standard_normal(n) * np.exp(rng.uniform(-20, 20, n)), np.exp(rng.uniform(-10, 10, n))),
        ((rng.integers(0, 1 << 23, n, dtype=np.uint32) | 0x3F800000).view(np.float32),
         (rng.integers(0, 1 << 23, n, dtype=np.uint32) | 0x3F800000).view(np.float32)),
        ((rng.integers(0, 1 << 23, n, dtype=np.uint32) | 0x3F800000).view(np.float32),
         np.float32(2) - rng.integers(1, 4096, n).astype(np.float32) * np.float32(2 ** -23)),
    ]
    fp = ctypes.POINTER(ctypes.c_float)
    for s, den in cases:
        s = np.ascontiguousarray(s, dtype=np.float32)
        den = np.ascontiguousarray(den, dtype=np.float32)
        assert lib.papr_oracle_markstein_mismatches(s.ctypes.data_as(fp), den.ctypes.data_as(fp), n) == 0
