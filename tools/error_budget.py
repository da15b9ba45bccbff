"""Where the bf16 product path's RGB error comes from (VERDICT r1, weak #1b).  Runs on a GPU box:

  python tools/error_budget.py > gpurun_out/error_budget.json

For every golden fixture and two seeded medium problems it reports, against the reference's values (golden) or the CPU
oracle (medium), max-abs errors of: attention weights, aggregated features (relative to their scale) and RGB for
  bf16          : the product path (bf16 tcgen05 stacks + bf16 UNet)
  bf16+unet32   : bf16 stacks, the UNet evaluated in fp32 on the bf16 path's features  -> error owed to the stacks
  fp32+unetbf16 : parity-mode features, bf16 UNet                                      -> error owed to the UNet
  fp32          : parity mode (split-bf16 GEMMs on the library's kernels, fp32 UNet)
  ref_autocast  : the oracle port of the REFERENCE's own path on the same GPU under torch.autocast(bfloat16) -- what the
                  reference's `use_amp: True, amp_dtype: bfloat16` mode (models/attn.py:248, models/model.py:24-25) gives
                  on these inputs through cuBLAS / cuDNN; the yardstick for a bf16 tolerance
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import papr_oracle as O  # noqa: E402
from tests.parity import GOLDEN_CASES, golden_config, golden_params, load_golden, rel_err  # noqa: E402


def build(cfg, params, precision):
    from papr_b200.model import PAPR
    cfg.geoms.points["init_num"] = int(params["points"].shape[0])
    m = PAPR(cfg, device="cuda", precision=precision).cuda()
    m.load_my_state_dict({k: v.clone() for k, v in params.items()})
    return m


def errors(cfg, params, rays_o, rays_d, c2w, code, want):
    K = int(cfg.geoms.points.select_k)
    out = {}
    models = {p: build(cfg, params, p) for p in ("bf16", "fp32")}
    feats = {}
    with torch.no_grad():
        for p, m in models.items():
            fused, attn = m.evaluate(rays_o, rays_d, c2w)
            rgb = m(rays_o, rays_d, c2w, step=-1, shading_code=code)
            feats[p] = (fused, attn)
            a = attn.squeeze(-1).cpu()
            out[p] = dict(attn=float((torch.sort(a[..., :K], -1).values - torch.sort(want["attn"][..., :K], -1).values).abs().max()),
                          bkg=float((a[..., K] - want["attn"][..., K]).abs().max()),
                          fused=rel_err(fused.squeeze(-2).cpu(), want["fused"]),
                          rgb=float((rgb.cpu() - want["rgb"]).abs().max()))
        if cfg.models.use_renderer:
            for name, feat_p, unet_p in (("bf16+unet32", "bf16", "fp32"), ("fp32+unetbf16", "fp32", "bf16")):
                fused, attn = feats[feat_p]
                m = models[unet_p]
                gamma = beta = None
                if code is not None and m.mapping_mlp is not None:
                    aff = m.mapping_mlp(code)
                    gamma, beta = aff[: aff.shape[-1] // 2], aff[aff.shape[-1] // 2:]
                fg = m.renderer(fused.squeeze(-2).permute(0, 3, 1, 2), gamma=gamma, beta=beta).permute(0, 2, 3, 1)
                bk = attn[..., K, :]
                rgb = fg * (1 - bk) + m.bkg_feats.reshape(1, 1, 1, -1) * bk
                out[name] = dict(rgb=float((rgb.cpu() - want["rgb"]).abs().max()))
    # the reference's own mixed-precision mode on the same inputs (same candidate sets), stock PyTorch kernels
    with torch.no_grad():
        pc = {k: v.cuda() for k, v in params.items()}
        idx = models["fp32"].select_k_ind.long()
        r = O.forward(pc, cfg, rays_o, rays_d, shading_code=code, idx=idx, autocast_dtype=torch.bfloat16)
        a = r["attn"].float().cpu()
        out["ref_autocast"] = dict(attn=float((torch.sort(a[..., :K], -1).values - torch.sort(want["attn"][..., :K], -1).values).abs().max()),
                                   bkg=float((a[..., K] - want["attn"][..., K]).abs().max()),
                                   fused=rel_err(r["fused"].float().cpu(), want["fused"]),
                                   rgb=float((r["rgb"].float().cpu() - want["rgb"]).abs().max()))
    return out


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {}
    gdir = os.path.join(ROOT, "tests", "golden")
    for name in GOLDEN_CASES:
        g = load_golden(gdir, name)
        cfg, params = golden_params(g)
        code = torch.from_numpy(g["shading_code"]).cuda() if g["shading_code"].size else None
        want = {k: torch.from_numpy(g[k]) for k in ("attn", "fused", "rgb")}
        res[name] = errors(cfg, params, torch.from_numpy(g["rays_o"]).cuda(), torch.from_numpy(g["rays_d"]).cuda(),
                           torch.from_numpy(g["c2w"]).cuda(), code, want)
    for tag, P, seed, win in (("medium_2x48x40_p6001", 6001, 9, (60, 108, 70, 110)), ("medium_1x64x64_p30000", 30000, 5, (300, 364, 420, 484))):
        cfg = golden_config("chair")
        params = O.init_params(cfg, P, seed=seed, cloud="shell")
        nv = 2 if P == 6001 else 1
        size = 200 if P == 6001 else 800
        rays_o, rays_d, c2w = O.synthetic_rays(size, size, cfg.dataset.coord_scale, n_views=nv, seed=4, h0=win[0], h1=win[1], w0=win[2], w1=win[3])
        m = build(cfg, params, "bf16")
        with torch.no_grad():
            m.evaluate(rays_o.cuda(), rays_d.cuda(), c2w.cuda())
            idx = m.select_k_ind.cpu()
            w = O.forward(params, cfg, rays_o, rays_d, idx=idx)
        want = dict(attn=w["attn"], fused=w["fused"], rgb=w["rgb"])
        res[tag] = errors(cfg, params, rays_o.cuda(), rays_d.cuda(), c2w.cuda(), None, want)
    worst = {}
    for case, d in res.items():
        for mode, e in d.items():
            for k, v in e.items():
                worst.setdefault(mode, {}).setdefault(k, 0.0)
                worst[mode][k] = max(worst[mode][k], v)
    print(json.dumps(dict(cases=res, worst=worst), indent=1))


if __name__ == "__main__":
    main()
