"""
TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REAL reference.

Run in the build container (where /root/reference exists):  python oracle/make_golden.py
It imports zvict/papr read-only from /root/reference with the two shims of SURVEY.md section 8(c)
(stub ``lpips`` module; drop the removed ``verbose=`` kwarg of torch LR schedulers), builds
``models.PAPR`` on CPU in fp32, loads the oracle's seeded parameters into it (so the reference and
the restatement see identical numbers), runs forward / evaluate / backward, and stores the results.
It also asserts, before writing anything, that oracle/papr_oracle.py reproduces every stored tensor
-- bit-exactly for the distance/top-K stage.  The fixtures travel to the GPU box; the reference does not.
"""
import copy
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("PAPR_REFERENCE", "/root/reference")

from oracle import papr_oracle as O            # noqa: E402
from papr_b200.config import make_config      # noqa: E402


def import_reference():
    sys.modules.setdefault("lpips", types.ModuleType("lpips"))
    import torch.optim.lr_scheduler as S
    for name in ("LinearLR", "CosineAnnealingLR", "ExponentialLR", "StepLR", "SequentialLR"):
        cls = getattr(S, name)
        if getattr(cls, "_papr_shim", False):
            continue
        orig = cls.__init__

        def init(self, *a, _orig=orig, **kw):
            kw.pop("verbose", None)
            _orig(self, *a, **kw)
        cls.__init__ = init
        cls._papr_shim = True
    sys.path.insert(0, REF)
    import models  # noqa: F401  (reference package)
    return models


class RefArgs(dict):
    """Stand-in for the reference's utils.DictAsMember (utils.py:14-19; utils.py itself needs matplotlib)."""
    def __getattr__(self, name):
        v = self[name]
        return RefArgs(v) if isinstance(v, dict) else v


def build_reference(models, cfg, params):
    args = RefArgs(copy.deepcopy(dict(cfg)))
    args["geoms"]["points"]["init_num"] = int(params["points"].shape[0])
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = models.get_model(args, device="cpu")
    sd = model.state_dict()
    for k, v in params.items():
        if k in ("points", "points_influ_scores", "pc_feats"):
            continue
        assert k in sd, k
        assert tuple(sd[k].shape) == tuple(v.shape), (k, sd[k].shape, v.shape)
    missing = [k for k in sd if k not in params]
    assert not missing, missing
    with contextlib.redirect_stdout(io.StringIO()):
        model.load_my_state_dict({k: v.clone() for k, v in params.items()})
    return model


from tests.parity import golden_config as variant  # noqa: E402


CASES = [
    # name, variant, H, W, P, n_views, cloud
    ("chair_12x12_p800", "chair", 12, 12, 800, 1, "cube"),
    ("chair_2views_8x8_p500", "chair", 8, 8, 500, 2, "shell"),
    ("caterpillar_exposure_12x16_p600", "caterpillar_exposure", 12, 16, 600, 1, "cube"),
    ("lego_like_8x8_p400", "lego_like", 8, 8, 400, 1, "cube"),
    ("no_renderer_8x8_p400", "no_renderer", 8, 8, 400, 1, "cube"),
    ("hotdog_like_8x8_p400", "hotdog_like", 8, 8, 400, 1, "shell"),
]

SAMPLE = 96   # leading entries of every weight gradient that are stored


def run_case(models, name, var, H, W, P, n_views, cloud):
    cfg = variant(var)
    params = O.init_params(cfg, P, seed=1, cloud=cloud)
    rays_o, rays_d, c2w = O.synthetic_rays(H * 8, W * 8, cfg.dataset.coord_scale, n_views=n_views, seed=3,
                                           h0=H * 3, h1=H * 4, w0=W * 3, w1=W * 4)
    model = build_reference(models, cfg, params)
    code = None
    if cfg.exposure_control.use:
        code = torch.randn(cfg.exposure_control.shading_code_dim, generator=torch.Generator().manual_seed(5))
    tgt = torch.rand(n_views, H, W, 3, generator=torch.Generator().manual_seed(7))

    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        rgb = model(rays_o, rays_d, c2w, step=-1, shading_code=code)
        loss = torch.mean((model.last_act(rgb) - tgt) ** 2)
        loss.backward()
        with torch.no_grad():
            fused, attn = model.evaluate(rays_o, rays_d, c2w, step=-1, shading_code=code)
            ref_idx = model.select_k_ind.clone()
            dist = model._calculate_global_distances.__func__  # noqa (kept for clarity)
            feat = O.select_distances_torch(rays_o, rays_d, model.points.detach(), cfg.eps)

    # ---- the restatement must reproduce the reference before anything is written
    o_dist = O.select_distances(rays_o, rays_d, params["points"], cfg.eps)
    assert torch.equal(o_dist, feat), f"{name}: C oracle distances differ from torch CPU"
    o_idx, kth = O.select_topk(rays_o, rays_d, params["points"], int(cfg.geoms.points.select_k), cfg.eps)
    ref_set = torch.sort(ref_idx, dim=-1).values
    o_set = torch.sort(o_idx, dim=-1).values
    diff = (ref_set != o_set).any(-1)
    if diff.any():   # only ties at the K-th distance may differ
        d_ref = torch.gather(feat, -1, ref_idx).max(-1).values
        assert torch.equal(d_ref[diff], kth[diff]), f"{name}: top-K sets differ beyond ties"
    pg = {k: v.clone().requires_grad_(v.dtype.is_floating_point and k != "bkg_feats") for k, v in params.items()}
    out = O.forward(pg, cfg, rays_o, rays_d, shading_code=code, idx=ref_idx)
    o_loss = torch.mean((out["rgb"] - tgt) ** 2)
    o_loss.backward()

    def close(a, b, what, rtol=2e-5, atol=2e-6):
        err = (a - b).abs().max().item()
        ref = b.abs().max().item()
        assert err <= atol + rtol * ref, f"{name}: {what} err {err:.3e} (scale {ref:.3e})"
        return err

    close(out["rgb"], rgb.detach(), "rgb")
    close(out["fused"], fused.squeeze(-2), "fused")
    close(out["attn"], attn.squeeze(-1), "attn")
    close(o_loss.detach(), loss.detach(), "loss")
    ref_named = dict(model.named_parameters())
    grads = {}
    for k, p in ref_named.items():
        if p.grad is None:
            continue
        close(pg[k].grad, p.grad, f"grad[{k}]", rtol=2e-4, atol=1e-9)
        grads[k] = p.grad

    rec = dict(
        H=H, W=W, P=P, n_views=n_views, variant=var, cloud=cloud,
        params_checksum=O.params_checksum(params),
        rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), c2w=c2w.numpy(), target=tgt.numpy(),
        shading_code=(code.numpy() if code is not None else np.zeros(0, np.float32)),
        idx_sorted=ref_set.numpy().astype(np.int32), kth=kth.numpy(),
        rgb=rgb.detach().numpy(), fused=fused.squeeze(-2).numpy(), attn=attn.squeeze(-1).numpy(),
        loss=np.float32(loss.item()),
        grad_points=grads["points"].numpy(), grad_influ=grads["points_influ_scores"].numpy(),
        grad_pc_feats=grads["pc_feats"].numpy(),
    )
    names, norms, samples = [], [], []
    for k in sorted(grads):
        if k in ("points", "points_influ_scores", "pc_feats"):
            continue
        g = grads[k].reshape(-1)
        names.append(k)
        norms.append(float(g.double().norm()))
        s = torch.zeros(SAMPLE)
        s[: min(SAMPLE, g.numel())] = g[:SAMPLE]
        samples.append(s.numpy())
    rec["wgrad_names"] = np.array(names)
    rec["wgrad_norms"] = np.array(norms, dtype=np.float64)
    rec["wgrad_samples"] = np.stack(samples).astype(np.float32)
    return rec


def select_case(models):
    """Bigger selection-only fixture: reference distances vs the C restatement, bit for bit."""
    cfg = variant("chair")
    P, H, W = 3000, 24, 24
    params = O.init_params(cfg, P, seed=2, cloud="shell")
    rays_o, rays_d, c2w = O.synthetic_rays(96, 96, cfg.dataset.coord_scale, n_views=2, seed=11, h0=40, h1=40 + H, w0=30, w1=30 + W)
    model = build_reference(models, cfg, params)
    with torch.no_grad():
        ref_idx = model._calculate_global_distances(rays_o, rays_d, model.points)
    feat = O.select_distances_torch(rays_o, rays_d, params["points"], cfg.eps)
    o_dist = O.select_distances(rays_o, rays_d, params["points"], cfg.eps)
    assert torch.equal(o_dist, feat), "select: C oracle distances differ from torch CPU"
    o_idx, kth = O.select_topk(rays_o, rays_d, params["points"], 20, cfg.eps)
    ref_set, o_set = torch.sort(ref_idx, -1).values, torch.sort(o_idx, -1).values
    diff = (ref_set != o_set).any(-1)
    if diff.any():
        d_ref = torch.gather(feat, -1, ref_idx).max(-1).values
        assert torch.equal(d_ref[diff], kth[diff])
    kth_ref = torch.gather(feat, -1, ref_idx).max(-1).values
    return dict(H=H, W=W, P=P, n_views=2, params_checksum=O.params_checksum(params),
                rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), idx_sorted=ref_set.numpy().astype(np.int32),
                kth=kth_ref.numpy(), n_tie_rays=int(diff.sum()))


def main():
    torch.set_num_threads(8)
    models = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    rec = select_case(models)
    np.savez_compressed(os.path.join(out_dir, "select_24x24x2_p3000.npz"), **rec)
    print("select_24x24x2_p3000 ok; rays with tie-resolved differences:", rec["n_tie_rays"])
    for case in CASES:
        rec = run_case(models, *case)
        np.savez_compressed(os.path.join(out_dir, case[0] + ".npz"), **rec)
        print(case[0], "ok  loss", float(rec["loss"]))


if __name__ == "__main__":
    main()
