"""what the callers do right after the path -- the cIMLE shading-code resampling of
exposure_control_finetune.py (reference utils.py:406-494) and the depth / foreground / background-mask outputs of
test.py:86-126 -- restructured around the B200 path.

The reference evaluates the frame in 100x100 tiles (its distance temporaries need it), then runs the UNet once per
sampled code (20 sequential passes with an empty_cache() in between).  Here the code-independent part -- selection,
attention, blend -- runs ONCE on the whole frame, and the code-dependent part is ONE batched decode: the 20 FiLM-modulated
copies of the feature map go through the renderer as a batch of 20.
"""
import numpy as np
import torch


def affine_terms(model, codes):
    """mapping_mlp over a batch of shading codes (S, dim) -> gamma (S, C), beta (S, C)  (model.py:497-499)."""
    aff = torch.stack([model.mapping_mlp(c) for c in codes]) if codes.dim() == 2 else model.mapping_mlp(codes).unsqueeze(0)
    half = aff.shape[-1] // 2
    return aff[:, :half], aff[:, half:]


def composite(model, fg, attn):
    """test.py:91-100 / model.py:539-545: fg (B,H,W,3), attn (1|B,H,W,K+1,1) -> rgb (B,H,W,3)."""
    K = attn.shape[-2] - 1
    bkg_attn = attn[..., K, :]
    bkg = model.bkg_feats.reshape(1, 1, 1, -1)
    if model.args.models.normalize_topk_attn:
        return fg * (1 - bkg_attn) + bkg * bkg_attn
    return fg + bkg * bkg_attn


@torch.no_grad()
def render_outputs(model, rays_o, rays_d, c2w=None, shading_code=None):
    """One frame through the path with every by-product test.py writes: dict(rgb (N,H,W,3) after last_act, foreground
    (N,H,W,3), bkg_mask (N,H,W) = background attention weight, depth (N,H,W) = attention-weighted distance of the
    selected points from the camera plane (test.py:120-126), attn, fused)."""
    N, H, W, _ = rays_d.shape
    fused, attn = model.evaluate(rays_o, rays_d, c2w)
    gamma = beta = None
    if shading_code is not None and model.mapping_mlp is not None:
        g, b = affine_terms(model, shading_code)
        gamma, beta = g[0], b[0]
    if model.args.models.use_renderer:
        fg = model.renderer(fused.squeeze(-2).permute(0, 3, 1, 2), gamma=gamma, beta=beta).permute(0, 2, 3, 1)
    else:
        fg = fused.squeeze(-2)
    rgb = model.last_act(composite(model, fg, attn))
    # depth: |p . od - D| / |od| with od = -o, D = od . o, weighted by the (un-renormalised) attention; background at 0
    od = -rays_o.reshape(N, 1, 1, 1, 3)
    D = (od * rays_o.reshape(N, 1, 1, 1, 3)).sum(-1)
    sel = model.selected_points
    dists = ((sel * od).sum(-1) - D).abs() / od.norm(dim=-1)
    K = sel.shape[-2]
    depth = (attn.squeeze(-1)[..., :K] * dists).sum(-1)
    return dict(rgb=rgb, foreground=fg, bkg_mask=attn[..., K, 0], depth=depth, attn=attn, fused=fused)


@torch.no_grad()
def resample_shading_codes(shading_codes, model, img_id, rays_o, rays_d, img, loss_fn=None, c2w=None, batch=None):
    """utils.py:406-494: draw `shading_code_num_samples` codes, keep the one whose render is closest to `img`
    (by loss or PSNR, args.exposure_control.shading_code_resample_select_by) in shading_codes[img_id].
    Returns (best index, losses (S,), psnrs (S,), sampled codes)."""
    E = model.args.exposure_control
    S = E.shading_code_num_samples
    codes = torch.randn(S, E.shading_code_dim, device=rays_d.device) * E.shading_code_scale
    N, H, W, _ = rays_d.shape
    assert N == 1, "one image at a time, as the reference"
    fused, attn = model.evaluate(rays_o, rays_d, c2w)              # code-independent: once, whole frame
    gamma, beta = affine_terms(model, codes)
    x = fused.squeeze(-2).permute(0, 3, 1, 2)
    batch = batch or S
    losses, psnrs = [], []
    for s0 in range(0, S, batch):
        s1 = min(s0 + batch, S)
        fg = model.renderer(x.expand(s1 - s0, -1, -1, -1), gamma=gamma[s0:s1], beta=beta[s0:s1]).permute(0, 2, 3, 1)
        rgb = model.last_act(composite(model, fg, attn))
        mse = ((rgb - img) ** 2).reshape(s1 - s0, -1).mean(-1)
        if loss_fn is not None:
            losses.append(torch.stack([loss_fn(rgb[i:i + 1], img) for i in range(s1 - s0)]).reshape(-1))
        else:
            losses.append(mse)
        psnrs.append(-10.0 * torch.log(mse) / np.log(10.0))
    losses, psnrs = torch.cat(losses), torch.cat(psnrs)
    best = int(torch.argmin(losses)) if E.shading_code_resample_select_by == "loss" else int(torch.argmax(psnrs))
    shading_codes[img_id] = codes[best]
    return best, losses, psnrs, codes
