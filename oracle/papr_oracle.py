"""
TEST INFRASTRUCTURE ONLY -- CPU oracle of the PAPR per-ray proximity-attention rendering path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product (``papr_b200``) never does: it fails loudly when its CUDA
library is missing instead of falling back to anything in here.

What it is: a functional, fp32, CPU restatement of the reference algorithm (zvict/papr, files under
/root/reference/models) written against a flat ``params`` dict that uses the reference's state_dict
key names.  It uses ``torch`` CPU tensors as the array library because the reference *is* PyTorch
CPU code: matching its rounding sequence (separately rounded products, ``(a+b)+c`` sums, IEEE
division, FMA norm) is what makes the top-K stage bit-exact.

Parity pinning: ``oracle/make_golden.py`` imports the real reference from /root/reference (read
only), loads the same seeded parameters into it, and records its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors.  The reference has
no golden vectors or tests of its own, so that is the only pin there can be.

Stage map (reference file:line):
  select_topk            models/model.py:258-283, 312-333
  ray_point_geometry     models/model.py:285-310 + models/utils.py:255-257
  posenc                 models/utils.py:232-242
  layer_norm             models/attn.py:30-42
  mlp_chain              models/mlp.py:12-59 (+ attn.py:90-117 FeedForward)
  embed / scores         models/attn.py:165-197, 212-226, 45-54
  blend                  models/model.py:519-545 (forward) / 473-492 (evaluate)
  unet                   models/unet.py:208-258
  mapping_mlp            models/mlp.py:62-78, model.py:497-499
  forward / evaluate     models/model.py:494-560 / 462-492
  init_params            models/model.py:18-115, attn.py:204-208, mlp.py:43-45 (shapes + init laws)
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _select_lib():
    """Load (building on demand with gcc) the C select oracle."""
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "libselect_oracle.so")
        src = os.path.join(_HERE, "select_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        lib = ctypes.CDLL(so)
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
        lib.papr_oracle_distances.argtypes = [fp, fp, fp, ctypes.c_int64, ctypes.c_int64, ctypes.c_float, fp]
        lib.papr_oracle_topk.argtypes = [fp, fp, fp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                         ctypes.c_float, ip, fp]
        lib.papr_oracle_markstein_mismatches.argtypes = [fp, fp, ctypes.c_int64]
        lib.papr_oracle_markstein_mismatches.restype = ctypes.c_int64
        _LIB = lib
    return _LIB


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _per_ray_origins(rays_o, rays_d):
    N, H, W, _ = rays_d.shape
    o = rays_o.reshape(N, 1, 1, 3).expand(N, H, W, 3)
    return (np.ascontiguousarray(o.reshape(-1, 3).numpy(), dtype=np.float32),
            np.ascontiguousarray(rays_d.reshape(-1, 3).numpy(), dtype=np.float32))


# ----------------------------------------------------------------------------- select (a1)
def select_distances(rays_o, rays_d, points, eps=1e-6):
    """(N,H,W,P) fp32 distances, model.py:272-279, through the C restatement (small sizes)."""
    N, H, W, _ = rays_d.shape
    o, d = _per_ray_origins(rays_o, rays_d)
    p = np.ascontiguousarray(points.detach().numpy(), dtype=np.float32)
    out = np.empty((o.shape[0], p.shape[0]), dtype=np.float32)
    _select_lib().papr_oracle_distances(_fptr(o), _fptr(d), _fptr(p), o.shape[0], p.shape[0], eps, _fptr(out))
    return torch.from_numpy(out).reshape(N, H, W, -1)


def select_distances_torch(rays_o, rays_d, points, eps=1e-6):
    """Literal tensor restatement of model.py:272-279 (materialises (N,H,W,P,3); small sizes)."""
    N, H, W, _ = rays_d.shape
    P = points.shape[0]
    d = rays_d.unsqueeze(-2)
    o = rays_o.reshape(N, 1, 1, 1, 3)
    v = points.reshape(1, 1, 1, P, 3) - o
    proj = d * (torch.sum(v * d, dim=-1) / (torch.sum(d * d, dim=-1) + eps)).unsqueeze(-1)
    return torch.norm(v - proj, dim=-1)


def select_topk(rays_o, rays_d, points, K, eps=1e-6):
    """model.py:312-333.  Returns (idx int64 (N,H,W,K') ordered by (distance, index), kth fp32 (N,H,W)).

    ``kth`` is the distance of the last selected point; the reference's ``topk(sorted=False)``
    leaves order and tie resolution unspecified, so comparisons must be on sets with ties at
    ``kth`` allowed (see tests/parity.py)."""
    N, H, W, _ = rays_d.shape
    P = points.shape[0]
    if K >= P or K < 0:   # model.py:326-327
        idx = torch.arange(P).expand(N, H, W, -1)
        return idx, torch.full((N, H, W), float("inf"))
    o, d = _per_ray_origins(rays_o, rays_d)
    p = np.ascontiguousarray(points.detach().numpy(), dtype=np.float32)
    idx = np.empty((o.shape[0], K), dtype=np.int32)
    kth = np.empty((o.shape[0],), dtype=np.float32)
    lib = _select_lib()
    lib.papr_oracle_topk(_fptr(o), _fptr(d), _fptr(p), o.shape[0], P, K, eps,
                         idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _fptr(kth))
    return torch.from_numpy(idx.astype(np.int64)).reshape(N, H, W, K), torch.from_numpy(kth).reshape(N, H, W)


def select_topk_torch(rays_o, rays_d, points, K, eps=1e-6):
    """model.py:258-283 as the reference runs it on any device: the materialised distance tensor + torch.topk
    (largest=False, sorted=False).  Used for the GPU-reference timing of bench.py (tile-sized inputs only)."""
    if K >= points.shape[0] or K < 0:
        N, H, W, _ = rays_d.shape
        return torch.arange(points.shape[0], device=points.device).expand(N, H, W, -1)
    return torch.topk(select_distances_torch(rays_o, rays_d, points, eps), K, dim=-1, largest=False, sorted=False)[1]


# ----------------------------------------------------------------------------- geometry (a3)
def ray_point_geometry(rays_o, rays_d, sel_points, eps=1e-6):
    """model.py:302-305: returns (proj, D) = (vec_p2o, vec_p2r), both (N,H,W,K,3)."""
    N = rays_d.shape[0]
    rays = (rays_d / (torch.norm(rays_d, dim=-1, keepdim=True) + eps)).unsqueeze(-2)
    v = sel_points - rays_o.reshape(N, 1, 1, 1, 3)
    proj = rays * (torch.sum(v * rays, dim=-1) / (torch.sum(rays * rays, dim=-1) + eps)).unsqueeze(-1)
    return proj, v - proj


# ----------------------------------------------------------------------------- PE / LN / MLP (a5-a7)
def posenc(x, L, factor=2.0, mult=1.0, without_self=False):
    """models/utils.py:232-242: per coordinate [x, sin(2^0 x), cos(2^0 x), ...], grouped per coordinate."""
    parts = [] if without_self else [x]
    for i in range(L):
        for fn in (torch.sin, torch.cos):
            parts.append(fn(factor ** i * x * mult))
    return torch.flatten(torch.stack(parts, -1), start_dim=-2, end_dim=-1)


def layer_norm(x, a, b, eps):
    """attn.py:39-42: unbiased std, eps added to std."""
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return a * (x - mean) / (std + eps) + b


def _act(name, x):
    if name == "relu":
        return torch.relu(x)
    if name == "leakyrelu":
        return F.leaky_relu(x, 0.2)
    if name == "none":
        return x
    if name == "relu+1":
        return torch.relu(x) + 1.0
    raise NotImplementedError(name)


def mlp_chain(params, prefix, x, n_layer, act, last_act, skip_layers=()):
    """mlp.py:47-59.  Linear i lives at ``{prefix}.model.{2i+1}``; skip layers re-concatenate the input."""
    inp = x
    for i in range(n_layer):
        if i in skip_layers:
            x = torch.cat([x, inp], dim=-1)
        x = F.linear(x, params[f"{prefix}.model.{2 * i + 1}.weight"], params[f"{prefix}.model.{2 * i + 1}.bias"])
        x = _act(last_act if i == n_layer - 1 else act, x)
    return x


def feed_forward(params, prefix, x, opt, eps):
    """attn.py:113-117 with dropout 0 and residual_ff false."""
    if opt.norm == "layernorm":
        x = layer_norm(x, params[f"{prefix}.innorm.a_2"], params[f"{prefix}.innorm.b_2"], eps)
    x = mlp_chain(params, f"{prefix}.mlp", x, opt.n_ff_layer, opt.ff_act, opt.ff_last_act, tuple(opt.skip_layers))
    if opt.norm == "layernorm":
        x = layer_norm(x, params[f"{prefix}.outnorm.a_2"], params[f"{prefix}.outnorm.b_2"], eps)
    return x


# ----------------------------------------------------------------------------- attention (a4,a8)
def proximity_attention(params, cfg, sel_points, proj, D, rays_d, sel_feats):
    """model.py:396-437 + attn.py:165-197, 212-226.  Returns (embedv (R,K,C), scores (R,K))."""
    A = cfg.models.attn
    E = A.embed
    eps = cfg.eps
    pe = lambda f, L: posenc(f, L, E.pe_factor, E.pe_mult_factor, without_self=(E.embed_type == 2))
    k_feats = [pe(f, E.k_L[i]) for i, f in enumerate([sel_points.detach(), proj, D])]
    q_feats = [pe(rays_d.unsqueeze(-2), E.q_L[0])]
    v_feats = [pe(f, E.v_L[i]) for i, f in enumerate([proj, D])]
    if cfg.geoms.point_feats.use_ink:
        k_feats.append(sel_feats)
    if cfg.geoms.point_feats.use_inq:
        q_feats.append(sel_feats)
    if cfg.geoms.point_feats.use_inv:
        v_feats.append(sel_feats)
    k = torch.cat(k_feats, -1).flatten(0, 2)   # (R, K, dk)
    q = torch.cat(q_feats, -1).flatten(0, 2)   # (R, 1, dq)
    v = torch.cat(v_feats, -1).flatten(0, 2)   # (R, K, dv)
    k = feed_forward(params, "proximity_attn.embed.embed_k", k, E.key, eps)
    q = feed_forward(params, "proximity_attn.embed.embed_q", q, E.query, eps)
    v = feed_forward(params, "proximity_attn.embed.embed_v", v, E.value, eps)
    key = F.linear(k, params["proximity_attn.attention_layer.w_k.weight"], params["proximity_attn.attention_layer.w_k.bias"])
    query = F.linear(q, params["proximity_attn.attention_layer.w_q.weight"], params["proximity_attn.attention_layer.w_q.bias"])
    scores = torch.matmul(query, key.transpose(-2, -1)) / math.sqrt(query.shape[-1])   # (R,1,K)
    scores = _act(A.score_act, scores)
    return v, scores.squeeze(1)


# ----------------------------------------------------------------------------- blend (a9)
def blend(cfg, embedv, scores, influ, bkg_score):
    """model.py:519-534: returns (fused (R,C), attn (R,K+1) incl. background, un-renormalised)."""
    scores = scores * influ
    scores = torch.cat([scores, bkg_score.reshape(1, 1).expand(scores.shape[0], 1)], dim=-1)
    attn = torch.softmax(scores, dim=-1)
    topk = attn[:, :-1]
    if cfg.models.normalize_topk_attn:
        topk = topk / topk.sum(-1, keepdim=True)
    fused = (embedv * topk.unsqueeze(-1)).sum(1)
    return fused, attn


# ----------------------------------------------------------------------------- decode (a10, a11)
def unet(params, x, gamma=None, beta=None, affine_layer=-1, prefix="renderer"):
    """unet.py:208-258 for the shipped SmallUNet (single=True, bilinear=False, norm='none')."""
    def conv(name, t, pad=1):
        return F.conv2d(t, params[f"{prefix}.{name}.weight"], params[f"{prefix}.{name}.bias"], padding=pad)

    def film(t, stage):
        if affine_layer == stage:
            C = t.shape[1]
            t = t * gamma.reshape(1, C, 1, 1) + beta.reshape(1, C, 1, 1)
        return t

    def up(name, t, skip):
        t = F.conv_transpose2d(t, params[f"{prefix}.{name}.up.weight"], params[f"{prefix}.{name}.up.bias"], stride=2)
        dy, dx = skip.shape[2] - t.shape[2], skip.shape[3] - t.shape[3]
        t = F.pad(t, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return torch.relu(conv(f"{name}.conv.double_conv.0", torch.cat([skip, t], dim=1)))

    x = film(x, 0)
    x1 = film(torch.relu(conv("inc.double_conv.0", x)), 1)
    x2 = film(torch.relu(conv("down1.maxpool_conv.1.double_conv.0", F.max_pool2d(x1, 2))), 2)
    x3 = film(torch.relu(conv("down2.maxpool_conv.1.double_conv.0", F.max_pool2d(x2, 2))), 3)
    y = film(up("up1", x3, x2), 4)
    y = film(up("up2", y, x1), 5)
    return conv("outc.conv", y, pad=0)


def mapping_mlp(params, cfg, code):
    M = cfg.exposure_control.mapping_mlp
    return mlp_chain(params, "mapping_mlp.model", code, M.num_layers, M.act, M.last_act)


# ----------------------------------------------------------------------------- whole path
def attention_features(params, cfg, rays_o, rays_d, idx=None, autocast_dtype=None):
    """Everything up to the blend.  Returns dict(idx, sel_points, fused (N,H,W,C), attn (N,H,W,K+1)).
    Device-agnostic: with CUDA tensors the selection is the reference's own torch formulation (select_topk_torch) and
    ``autocast_dtype`` wraps the attention block as attn.py:248 does (bench.py's GPU-reference leg)."""
    N, H, W, _ = rays_d.shape
    points = params["points"]
    if idx is None:
        if points.is_cuda:
            idx = select_topk_torch(rays_o, rays_d, points.detach(), int(cfg.geoms.points.select_k), cfg.eps)
        else:
            idx, _ = select_topk(rays_o, rays_d, points.detach(), int(cfg.geoms.points.select_k), cfg.eps)
    sel_points = points[idx]
    proj, D = ray_point_geometry(rays_o, rays_d, sel_points, cfg.eps)
    sel_feats = params["pc_feats"][idx]
    with torch.autocast(device_type="cuda", dtype=autocast_dtype or torch.bfloat16, enabled=autocast_dtype is not None):
        embedv, scores = proximity_attention(params, cfg, sel_points, proj, D, rays_d, sel_feats)
    embedv, scores = embedv.float(), scores.float()
    influ = params["points_influ_scores"][idx].reshape(N * H * W, -1)
    bkg_score = torch.tensor(float(cfg.geoms.background.constant), dtype=torch.float32, device=points.device)
    fused, attn = blend(cfg, embedv, scores, influ, bkg_score)
    return dict(idx=idx, sel_points=sel_points, fused=fused.reshape(N, H, W, -1), attn=attn.reshape(N, H, W, -1),
                embedv=embedv, scores=scores)


def forward(params, cfg, rays_o, rays_d, shading_code=None, idx=None, autocast_dtype=None):
    """model.py:494-560 -> rgb (N,H,W,3).  autocast_dtype: run the attention block and the UNet under torch.autocast
    (the reference's use_amp path, attn.py:248 / unet.py), CUDA only."""
    out = attention_features(params, cfg, rays_o, rays_d, idx, autocast_dtype)
    fused, attn = out["fused"], out["attn"]
    gamma = beta = None
    affine_layer = cfg.models.renderer.generator.small_unet.affine_layer
    if shading_code is not None and cfg.exposure_control.use:
        affine = mapping_mlp(params, cfg, shading_code)
        gamma, beta = affine[: affine.shape[-1] // 2], affine[affine.shape[-1] // 2:]
    if cfg.models.use_renderer:
        with torch.autocast(device_type="cuda", dtype=autocast_dtype or torch.bfloat16, enabled=autocast_dtype is not None):
            fg = unet(params, fused.permute(0, 3, 1, 2), gamma, beta, affine_layer)
        fg = fg.float().permute(0, 2, 3, 1)
    else:
        fg = fused
    bkg_attn = attn[..., -1:]
    bkg = params["bkg_feats"].reshape(1, 1, 1, 3)
    if cfg.models.normalize_topk_attn:
        rgb = fg * (1 - bkg_attn) + bkg * bkg_attn
    else:
        rgb = fg + bkg * bkg_attn
    out["rgb"] = rgb
    return out


def evaluate(params, cfg, rays_o, rays_d):
    """model.py:462-492 -> (fused (N,H,W,1,C), attn (N,H,W,K+1,1))."""
    out = attention_features(params, cfg, rays_o, rays_d)
    return out["fused"].unsqueeze(-2), out["attn"].unsqueeze(-1)


# ----------------------------------------------------------------------------- parameters
def _xavier(gen, out_f, in_f, receptive=1):
    bound = math.sqrt(6.0 / ((in_f + out_f) * receptive))
    return (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound


def _uniform(gen, shape, bound):
    return (torch.rand(*shape, generator=gen) * 2 - 1) * bound


def embed_dims(cfg):
    """attn.py:136-146 -> (dk, dq, dv) input widths of the three stacks."""
    E = cfg.models.attn.embed
    pf = cfg.geoms.point_feats
    own = 1 if E.embed_type == 1 else 0
    dk = sum(3 * own + 3 * 2 * L for L in E.k_L) + (pf.dim if pf.use_ink else 0)
    dq = sum(3 * own + 3 * 2 * L for L in E.q_L) + (pf.dim if pf.use_inq else 0)
    dv = sum(3 * own + 3 * 2 * L for L in E.v_L) + (pf.dim if pf.use_inv else 0)
    return dk, dq, dv


def init_params(cfg, num_points, seed=1, influ="uniform", cloud="cube"):
    """Seeded parameters with the reference's shapes and init laws (xavier-uniform Linear weights
    mlp.py:43-45 / attn.py:207-208, U(+-1/sqrt(fan_in)) biases and conv weights as torch defaults,
    pc_feats ~ N(0,1) model.py:88).  Not the reference's RNG stream: make_golden.py loads *these*
    tensors into the reference model, so both sides see identical numbers.
    influ='uniform' draws U(0,1) influence scores (the 0.0 init of model.py:62-63 hides bugs)."""
    g = torch.Generator().manual_seed(seed)
    E = cfg.models.attn.embed
    cs = cfg.dataset.coord_scale
    P = {}
    scale = [s * cs for s in cfg.geoms.points.init_scale]
    pts = (torch.rand(num_points, 3, generator=g) * 2 - 1) * torch.tensor(scale)
    if cloud == "shell":   # points near a sphere of radius 0.8*scale: a "learned-like" surface cloud
        pts = pts / pts.norm(dim=-1, keepdim=True) * (0.8 * scale[0]) + 0.02 * scale[0] * torch.randn(num_points, 3, generator=g)
    P["points"] = pts.float()
    P["points_influ_scores"] = (torch.rand(num_points, 1, generator=g) if influ == "uniform"
                                else torch.full((num_points, 1), float(cfg.geoms.points.influ_init_val)))
    P["bkg_feats"] = torch.tensor(cfg.geoms.background.init_color, dtype=torch.float32)[None, :]
    P["pc_feats"] = torch.randn(num_points, cfg.geoms.point_feats.dim, generator=g)
    P["select_k"] = torch.tensor(cfg.geoms.points.select_k, dtype=torch.int32)
    dk, dq, dv = embed_dims(cfg)
    for name, d_in, opt in (("k", dk, E.key), ("q", dq, E.query), ("v", dv, E.value)):
        pre = f"proximity_attn.embed.embed_{name}"
        if opt.norm == "layernorm":
            # not the reference's ones/zeros: random affine terms so that a_2/b_2 handling is tested
            P[f"{pre}.innorm.a_2"] = 1 + 0.1 * torch.randn(d_in, generator=g)
            P[f"{pre}.innorm.b_2"] = 0.1 * torch.randn(d_in, generator=g)
            P[f"{pre}.outnorm.a_2"] = 1 + 0.1 * torch.randn(opt.d_ff_out, generator=g)
            P[f"{pre}.outnorm.b_2"] = 0.1 * torch.randn(opt.d_ff_out, generator=g)
        for i in range(opt.n_ff_layer):
            fi = d_in if i == 0 else opt.d_ff
            if i in opt.skip_layers:
                fi += d_in
            fo = opt.d_ff_out if i == opt.n_ff_layer - 1 else opt.d_ff
            P[f"{pre}.mlp.model.{2 * i + 1}.weight"] = _xavier(g, fo, fi)
            P[f"{pre}.mlp.model.{2 * i + 1}.bias"] = _uniform(g, (fo,), 1 / math.sqrt(fi))
    dm = cfg.models.attn.d_model
    for name, fi in (("w_k", E.key.d_ff_out), ("w_q", E.query.d_ff_out)):
        P[f"proximity_attn.attention_layer.{name}.weight"] = _xavier(g, dm, fi)
        P[f"proximity_attn.attention_layer.{name}.bias"] = _uniform(g, (dm,), 1 / math.sqrt(fi))
    if cfg.models.use_renderer:
        C = E.value.d_ff_out
        convs = [("inc.double_conv.0", 128, C, 3), ("down1.maxpool_conv.1.double_conv.0", 256, 128, 3),
                 ("down2.maxpool_conv.1.double_conv.0", 512, 256, 3), ("up1.conv.double_conv.0", 256, 512, 3),
                 ("up2.conv.double_conv.0", 128, 256, 3), ("outc.conv", 3, 128, 1)]
        for name, co, ci, ks in convs:
            b = 1 / math.sqrt(ci * ks * ks)
            P[f"renderer.{name}.weight"] = _uniform(g, (co, ci, ks, ks), b)
            P[f"renderer.{name}.bias"] = _uniform(g, (co,), b)
        for name, ci in (("up1.up", 512), ("up2.up", 256)):   # ConvTranspose2d weight is (in, out, 2, 2)
            b = 1 / math.sqrt((ci // 2) * 4)
            P[f"renderer.{name}.weight"] = _uniform(g, (ci, ci // 2, 2, 2), b)
            P[f"renderer.{name}.bias"] = _uniform(g, (ci // 2,), b)
    if cfg.exposure_control.use:
        M = cfg.exposure_control.mapping_mlp
        for i in range(M.num_layers):
            fi = cfg.exposure_control.shading_code_dim if i == 0 else M.dim
            fo = M.out_dim if i == M.num_layers - 1 else M.dim
            P[f"mapping_mlp.model.model.{2 * i + 1}.weight"] = _xavier(g, fo, fi)
            P[f"mapping_mlp.model.model.{2 * i + 1}.bias"] = _uniform(g, (fo,), 1 / math.sqrt(fi))
    return P


def params_checksum(params):
    """Order-independent fingerprint used by the golden fixtures to prove both sides used the same numbers."""
    tot = 0.0
    for k in sorted(params):
        t = params[k].double()
        tot += float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64).reshape(t.shape).remainder(97.0)).sum())
    return tot


# ----------------------------------------------------------------------------- synthetic scene
def synthetic_rays(H, W, coord_scale, n_views=1, seed=1, radius=4.0, camera_angle_x=0.6911,
                   h0=0, h1=None, w0=0, w1=None):
    """Cameras on a sphere looking at the origin; rays per dataset/utils.py:81-96 (pixel-centre dirs
    (x,-y,-1) rotated by c2w, unit norm) and origins scaled per dataset/dataset.py:19-25."""
    g = torch.Generator().manual_seed(seed)
    focal = 0.5 * W / math.tan(0.5 * camera_angle_x)
    c2ws = []
    for _ in range(n_views):
        th = float(torch.rand(1, generator=g)) * 2 * math.pi
        ph = (0.15 + 0.5 * float(torch.rand(1, generator=g))) * math.pi / 2
        pos = torch.tensor([math.cos(th) * math.cos(ph), math.sin(th) * math.cos(ph), math.sin(ph)]) * radius
        fwd = -pos / pos.norm()
        right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0]))
        right = right / right.norm()
        up = torch.linalg.cross(right, fwd)
        c2w = torch.eye(4)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up, -fwd, pos
        c2ws.append(c2w)
    c2w = torch.stack(c2ws).float()
    width = torch.linspace(0, W / focal, steps=W + 1, dtype=torch.float32)
    height = torch.linspace(0, H / focal, steps=H + 1, dtype=torch.float32)
    y, x = torch.meshgrid(height, width, indexing="ij")
    px, py = width[1] - width[0], height[1] - height[0]
    x = (x - W / focal / 2 + px / 2)[:-1, :-1]
    y = -(y - H / focal / 2 + py / 2)[:-1, :-1]
    dirs = torch.stack([x, y, -torch.ones_like(x)], -1)
    dirs4 = torch.cat([dirs, torch.zeros_like(dirs[..., :1])], -1)
    rays_d = torch.sum(dirs4.unsqueeze(0).unsqueeze(-2) * c2w.reshape(-1, 1, 1, 4, 4), -1)[..., :3]
    rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    rays_o = c2w[:, :3, 3] * coord_scale
    h1 = H if h1 is None else h1
    w1 = W if w1 is None else w1
    return rays_o.contiguous(), rays_d[:, h0:h1, w0:w1].contiguous(), c2w


# ----------------------------------------------------------------------------- LPIPS / VGG16
VGG16_CFG = [(3, 64), (64, 64), "P", (64, 128), (128, 128), "P", (128, 256), (256, 256), (256, 256), "P",
             (256, 512), (512, 512), (512, 512), "P", (512, 512), (512, 512), (512, 512)]
VGG16_TAPS = (1, 3, 6, 9, 12)          # conv index (0-based over the 13 convs) whose ReLU output LPIPS reads: relu1_2 ... relu5_3
VGG16_FEATURE_INDEX = (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)      # torchvision vgg16().features index of each conv


def init_lpips_params(seed=1, lin_path=None):
    """Seeded stand-in for the ImageNet VGG16 weights (not available offline) with torchvision's shapes, keyed like the
    reference's LPNet state dict (models/lpips.py:8-28: net.slice{1..5}.{features index}.{weight,bias}), plus the learnt
    LPIPS linear weights lins.{k}.weight (loaded from the reference's vgg.pth when lin_path is given, else seeded)."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    slices = {0: 1, 2: 1, 5: 2, 7: 2, 10: 3, 12: 3, 14: 3, 17: 4, 19: 4, 21: 4, 24: 5, 26: 5, 28: 5}
    ci = 0
    for item in VGG16_CFG:
        if item == "P":
            continue
        cin, cout = item
        idx = VGG16_FEATURE_INDEX[ci]
        std = math.sqrt(2.0 / (cin * 9))                       # kaiming-normal keeps activations O(1) through 13 layers
        P[f"net.slice{slices[idx]}.{idx}.weight"] = torch.randn(cout, cin, 3, 3, generator=g) * std
        P[f"net.slice{slices[idx]}.{idx}.bias"] = 0.05 * torch.randn(cout, generator=g)
        ci += 1
    if lin_path is not None and os.path.exists(lin_path):
        w = torch.load(lin_path, map_location="cpu")
        for k in range(5):
            P[f"lins.{k}.weight"] = w[f"lin{k}.model.1.weight"].float().clone()
    else:
        for k, c in enumerate((64, 128, 256, 512, 512)):
            P[f"lins.{k}.weight"] = torch.rand(1, c, 1, 1, generator=g) * 0.5
    return P


def vgg16_features(params, x):
    """models/lpips.py:8-48: the five ReLU taps of torchvision's vgg16().features."""
    slices = {0: 1, 2: 1, 5: 2, 7: 2, 10: 3, 12: 3, 14: 3, 17: 4, 19: 4, 21: 4, 24: 5, 26: 5, 28: 5}
    outs, ci = [], 0
    for item in VGG16_CFG:
        if item == "P":
            x = F.max_pool2d(x, 2)
            continue
        idx = VGG16_FEATURE_INDEX[ci]
        x = torch.relu(F.conv2d(x, params[f"net.slice{slices[idx]}.{idx}.weight"], params[f"net.slice{slices[idx]}.{idx}.bias"], padding=1))
        if ci in VGG16_TAPS:
            outs.append(x)
        ci += 1
    return outs


def lpips(params, in0, in1):
    """models/lpips.py:103-125 (LPNet.forward): in0, in1 (N,H,W,3) in [0,1] -> scalar."""
    shift = torch.tensor([-.030, -.088, -.188], dtype=in0.dtype, device=in0.device)[None, :, None, None]
    scale = torch.tensor([.458, .448, .450], dtype=in0.dtype, device=in0.device)[None, :, None, None]
    a = ((2 * in0.permute(0, 3, 1, 2) - 1) - shift) / scale
    b = ((2 * in1.permute(0, 3, 1, 2) - 1) - shift) / scale
    fa, fb = vgg16_features(params, a), vgg16_features(params, b)

    def norm(t, eps=1e-10):
        return t / (torch.sqrt(torch.sum(t ** 2, dim=1, keepdim=True) + eps) + eps)
    val = 0
    for k in range(5):
        d = (norm(fa[k]) - norm(fb[k])) ** 2
        val = val + torch.sum(params[f"lins.{k}.weight"] * d, 1, keepdim=True).mean([2, 3], keepdim=True)
    return val.squeeze().mean()
