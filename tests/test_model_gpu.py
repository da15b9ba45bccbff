"""GPU parity of the whole hot path (papr_b200.PAPR) against the reference's golden vectors and the CPU oracle.

Two precisions are checked:
  fp32 : every CUDA-core kernel of the library + torch fp32 GEMMs.  Tolerances are the north-star ones: attention
         weights / aggregated features 1e-5 relative, RGB 1e-3 max-abs (we hold RGB to 1e-4), gradients 1e-3 relative.
         Gradients w.r.t. points / point features are limited by the REFERENCE's own fp32 noise (its CPU fp32 gradient
         is 2.7e-3 away from a float64 evaluation on chair_12x12_p800: the PE derivative multiplies by 2^5 and
         cancels), so they are checked two ways: within 5e-3 of the reference's fp32 values, and no further from a
         float64 evaluation of the oracle than max(1e-3, 3x the reference's own fp32 distance from it).
  bf16 : the product path (tcgen05 bf16 GEMMs, fp32 accumulation).  Stated bf16 tolerance = 2x the worst error measured
         over all fixtures on B200 (tools/error_budget.py, profiles/r02_error_budget.md): candidate attention weights
         2e-3 absolute (measured 8.3e-4), background weight 7e-3 (3.2e-3), aggregated features 1.5e-2 relative to their
         scale (7.5e-3), RGB 9e-3 max-abs (4.3e-3).  The RGB error is owed to the bf16 key stack (a 3e-3 error of the
         background weight multiplies the colour); the bf16 UNet alone contributes <= 8.7e-4.  Gradients: relative L2
         error <= 0.2 and cosine similarity >= 0.98 against the reference (measured: L2 0.004-0.15, cosine >= 0.99); the
         worst single entry may be off by up to 0.35 of the gradient's max-abs (measured 0.03-0.27 on these tiny fixtures).
  The fp32 mode's GEMMs run on the library's own tensor-core kernels (papr_b200/split_gemm.py), not on torch.matmul.
"""
import pytest
import torch

from oracle import papr_oracle as O
from tests.parity import GOLDEN_CASES, golden_params, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _build(cfg, params, precision):
    from papr_b200.model import PAPR
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg.geoms.points["init_num"] = int(params["points"].shape[0])
    model = PAPR(cfg, device="cuda", precision=precision).cuda()
    model.load_my_state_dict({k: v.clone() for k, v in params.items()})
    return model


def _inputs(g):
    code = torch.from_numpy(g["shading_code"]).cuda() if g["shading_code"].size else None
    return (torch.from_numpy(g["rays_o"]).cuda(), torch.from_numpy(g["rays_d"]).cuda(), torch.from_numpy(g["c2w"]).cuda(),
            torch.from_numpy(g["target"]).cuda(), code)


def _fp64_oracle_grads(cfg, params, g):
    """Gradients of the same loss from a float64 evaluation of the oracle (the tie-breaker for fp32 noise)."""
    dt = torch.float64
    pg = {k: (v.to(dt) if v.dtype.is_floating_point else v).clone().requires_grad_(
        v.dtype.is_floating_point and k != "bkg_feats") for k, v in params.items()}
    rays_o, rays_d = torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])
    idx, _ = O.select_topk(rays_o, rays_d, params["points"], int(cfg.geoms.points.select_k), cfg.eps)
    code = torch.from_numpy(g["shading_code"]).to(dt) if g["shading_code"].size else None
    out = O.forward(pg, cfg, rays_o.to(dt), rays_d.to(dt), shading_code=code, idx=idx)
    ((out["rgb"] - torch.from_numpy(g["target"]).to(dt)) ** 2).mean().backward()
    return {k: pg[k].grad for k in ("points", "points_influ_scores", "pc_feats")}


def _supported(name, precision):
    return True


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_and_evaluate_match_reference(golden_dir, name, precision):
    if not _supported(name, precision):
        pytest.skip("value.skip_layers is handled by the fp32 path only")
    g = load_golden(golden_dir, name)
    cfg, params = golden_params(g)
    model = _build(cfg, params, precision)
    rays_o, rays_d, c2w, tgt, code = _inputs(g)
    with torch.no_grad():
        fused, attn = model.evaluate(rays_o, rays_d, c2w)
        idx = model.select_k_ind
        rgb = model(rays_o, rays_d, c2w, step=-1, shading_code=code)
    K = int(cfg.geoms.points.select_k)
    assert fused.shape == (*rays_d.shape[:3], 1, cfg.models.attn.embed.value.d_ff_out)
    assert attn.shape == (*rays_d.shape[:3], K + 1, 1) and idx.dtype == torch.int64
    assert model.selected_points.shape == (*rays_d.shape[:3], K, 3)
    # top-K sets: bit-exact against the reference (no ties in these fixtures)
    assert torch.equal(torch.sort(idx.cpu(), -1).values, torch.from_numpy(g["idx_sorted"]).long())
    # the reference's attn is ordered by ITS topk order; ours by (distance, index): compare order-independently
    want_attn, want_fused, want_rgb = (torch.from_numpy(g[k]) for k in ("attn", "fused", "rgb"))
    got_attn = attn.squeeze(-1).cpu()
    a_got = torch.sort(got_attn[..., :K], -1).values
    a_want = torch.sort(want_attn[..., :K], -1).values
    e_attn = float((a_got - a_want).abs().max())
    e_bkg = float((got_attn[..., K] - want_attn[..., K]).abs().max())
    e_fused = rel_err(fused.squeeze(-2).cpu(), want_fused)
    e_rgb = float((rgb.cpu() - want_rgb).abs().max())
    print(f"{name} [{precision}] attn {e_attn:.2e} bkg {e_bkg:.2e} fused {e_fused:.2e} rgb {e_rgb:.2e}")
    if precision == "fp32":
        assert e_attn <= 1e-5 and e_bkg <= 1e-5 and e_fused <= 1e-5 and e_rgb <= 1e-4
    else:
        assert e_attn <= 2e-3 and e_bkg <= 7e-3 and e_fused <= 1.5e-2 and e_rgb <= 9e-3


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gradients_match_reference(golden_dir, name, precision):
    if not _supported(name, precision):
        pytest.skip("value.skip_layers is handled by the fp32 path only")
    g = load_golden(golden_dir, name)
    cfg, params = golden_params(g)
    model = _build(cfg, params, precision)
    rays_o, rays_d, c2w, tgt, code = _inputs(g)
    model.clear_grad()
    rgb = model(rays_o, rays_d, c2w, step=-1, shading_code=code)
    loss = torch.mean((model.last_act(rgb) - tgt) ** 2)
    model.scaler.scale(loss).backward()
    tol = 5e-3 if precision == "fp32" else 0.35     # bf16: worst single entry; the L2 / cosine checks are the stable ones
    assert abs(loss.item() - float(g["loss"])) <= (1e-5 if precision == "fp32" else 2e-2)
    errs = {}
    truth = _fp64_oracle_grads(cfg, params, g) if precision == "fp32" else None
    for key, attr in (("grad_points", "points"), ("grad_influ", "points_influ_scores"), ("grad_pc_feats", "pc_feats")):
        got = getattr(model, attr).grad.cpu()
        want = torch.from_numpy(g[key])
        errs[attr] = rel_err(got, want)
        l2 = float((got - want).double().norm()) / max(float(want.double().norm()), 1e-30)
        print(f"   {attr}: relative L2 error {l2:.3e}, max-abs/max {errs[attr]:.3e}")
        if precision == "bf16":
            assert l2 <= 0.2, (name, attr, "relative L2", l2)
        if float(want.abs().max()) > 0:
            cos = float(torch.nn.functional.cosine_similarity(got.flatten().double(), want.flatten().double(), dim=0))
            assert cos >= (0.9999 if precision == "fp32" else 0.98), (name, attr, "cosine", cos)
        if truth is not None:
            e64 = rel_err(got.double(), truth[attr])
            ref64 = rel_err(want.double(), truth[attr])       # the reference's own fp32 noise
            print(f"   {attr}: vs reference fp32 {errs[attr]:.2e}; vs float64 oracle: ours {e64:.2e}, reference {ref64:.2e}")
            assert e64 <= max(1e-3, 3 * ref64), (name, attr, "vs float64 oracle", e64, ref64)
    named = dict(model.named_parameters())
    worst = ("", 0.0)
    for nm, norm, sample in zip(g["wgrad_names"], g["wgrad_norms"], g["wgrad_samples"]):
        p = named[str(nm)]
        assert p.grad is not None, nm
        gn = float(p.grad.double().norm())
        e = abs(gn - norm) / max(norm, 1e-12)
        n = min(sample.size, p.grad.numel())
        es = float((p.grad.reshape(-1)[:n].cpu() - torch.from_numpy(sample[:n])).abs().max()) / max(norm / p.grad.numel() ** 0.5, 1e-12)
        if norm > 1e-7 and max(e, es * 0.1) > worst[1]:
            worst = (str(nm), max(e, es * 0.1))
    print(f"{name} [{precision}] grads {errs} worst weight grad {worst}")
    for k, e in errs.items():
        assert e <= tol, (k, e)
    assert worst[1] <= tol * 3, worst


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_against_oracle_medium(precision):
    """A larger seeded problem (2 views, 48x40 rays, 6000 points, ragged sizes) straight against the CPU oracle."""
    from tests.parity import golden_config
    cfg = golden_config("chair")
    params = O.init_params(cfg, 6001, seed=9, cloud="shell")
    rays_o, rays_d, c2w = O.synthetic_rays(200, 200, cfg.dataset.coord_scale, n_views=2, seed=4, h0=60, h1=108, w0=70, w1=110)
    model = _build(cfg, params, precision)
    with torch.no_grad():
        rgb = model(rays_o.cuda(), rays_d.cuda(), c2w.cuda())
        idx = model.select_k_ind.cpu()
        fused, attn = model.evaluate(rays_o.cuda(), rays_d.cuda(), c2w.cuda())
    with torch.no_grad():
        want = O.forward(params, cfg, rays_o, rays_d, idx=idx)      # same candidate order as ours
    e_attn = float((attn.squeeze(-1).cpu() - want["attn"]).abs().max())
    e_fused = rel_err(fused.squeeze(-2).cpu(), want["fused"])
    e_rgb = float((rgb.cpu() - want["rgb"]).abs().max())
    o_idx, _ = O.select_topk(rays_o, rays_d, params["points"], 20)
    assert torch.equal(idx, o_idx)
    print(f"medium [{precision}] attn {e_attn:.2e} fused {e_fused:.2e} rgb {e_rgb:.2e}")
    if precision == "fp32":
        assert e_attn <= 1e-5 and e_fused <= 1e-5 and e_rgb <= 1e-4
    else:
        assert e_attn <= 7e-3 and e_fused <= 1.5e-2 and e_rgb <= 9e-3


@pytest.mark.parametrize("K,P", [(1, 50), (31, 200), (32, 300), (20, 21)])
def test_edge_candidate_counts(K, P):
    """select_k extremes the kernels accept (1 <= K <= 32; P barely above K) against the oracle."""
    from tests.parity import golden_config
    cfg = golden_config("chair")
    cfg.geoms.points["select_k"] = K
    params = O.init_params(cfg, P, seed=K, cloud="cube")
    rays_o, rays_d, c2w = O.synthetic_rays(64, 64, cfg.dataset.coord_scale, n_views=1, seed=K, h0=20, h1=28, w0=30, w1=42)
    model = _build(cfg, params, "fp32")
    with torch.no_grad():
        fused, attn = model.evaluate(rays_o.cuda(), rays_d.cuda(), c2w.cuda())
        idx = model.select_k_ind.cpu()
        want = O.forward(params, cfg, rays_o, rays_d, idx=idx)
    o_idx, _ = O.select_topk(rays_o, rays_d, params["points"], K)
    assert torch.equal(idx, o_idx)
    assert float((attn.squeeze(-1).cpu() - want["attn"]).abs().max()) <= 1e-5
    assert rel_err(fused.squeeze(-2).cpu(), want["fused"]) <= 1e-5


def test_all_points_bypass_when_k_exceeds_cloud():
    """model.py:326-327: select_k >= P uses every point for every ray."""
    from tests.parity import golden_config
    cfg = golden_config("chair")
    params = O.init_params(cfg, 12, seed=3, cloud="cube")
    rays_o, rays_d, c2w = O.synthetic_rays(32, 32, cfg.dataset.coord_scale, n_views=1, seed=1, h0=8, h1=16, w0=8, w1=16)
    model = _build(cfg, params, "fp32")
    with torch.no_grad():
        fused, attn = model.evaluate(rays_o.cuda(), rays_d.cuda(), c2w.cuda())
        want = O.forward(params, cfg, rays_o, rays_d)
    assert attn.shape[-2] == 13 and model.select_k_ind.shape[-1] == 12
    assert float((attn.squeeze(-1).cpu() - want["attn"]).abs().max()) <= 1e-5
    assert rel_err(fused.squeeze(-2).cpu(), want["fused"]) <= 1e-5


def test_k32_gradients_match_oracle_autograd():
    """K = 32 uses every lane of the blend kernels' warp (the background token is a warp-uniform scalar): forward and the
    influence / feature / point gradients against the CPU oracle's autograd, parity mode."""
    from tests.parity import golden_config
    cfg = golden_config("chair")
    cfg.geoms.points["select_k"] = 32
    params = O.init_params(cfg, 400, seed=11, cloud="shell")
    rays_o, rays_d, c2w = O.synthetic_rays(64, 64, cfg.dataset.coord_scale, n_views=1, seed=3, h0=24, h1=36, w0=20, w1=36)
    tgt = torch.rand(1, 12, 16, 3, generator=torch.Generator().manual_seed(2))
    model = _build(cfg, params, "fp32")
    model.clear_grad()
    rgb = model(rays_o.cuda(), rays_d.cuda(), c2w.cuda())
    torch.mean((rgb - tgt.cuda()) ** 2).backward()
    pg = {k: v.clone().requires_grad_(v.dtype.is_floating_point and k != "bkg_feats") for k, v in params.items()}
    want = O.forward(pg, cfg, rays_o, rays_d, idx=model.select_k_ind.cpu())
    torch.mean((want["rgb"] - tgt) ** 2).backward()
    assert float((rgb.detach().cpu() - want["rgb"].detach()).abs().max()) <= 1e-4
    for key in ("points", "points_influ_scores", "pc_feats"):
        e = rel_err(getattr(model, key).grad.cpu(), pg[key].grad)
        assert e <= 5e-3, (key, e)
