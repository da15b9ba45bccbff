"""GPU: the prologue kernels agree -- the warp-per-ray kernel with exact fp32 taps, the same kernel on bf16 only, and the
row-per-lane bf16 kernels -- forward
(key / value stack inputs: models/model.py:285-310, 396-437; utils.py:232-257; attn.py:39-42) and backward."""
import os
import types

import pytest
import torch

from papr_b200 import ops
from papr_b200 import attention as A

pytestmark = pytest.mark.gpu


SHAPES = [(6, 64), (4, 64), (6, 128), (4, 128)]      # (PE order, feature width): nerfsyn / Tanks&Temples x default / materials.yml


def _case(R=301, K=20, P=700, views=1, seed=0, L=6, F=64):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    S = 1 + 2 * L
    dk, dv = 9 * S, 6 * S + F
    sh = types.SimpleNamespace(R=R, rays_per_view=R // views, K=K, M=R * K, L=L, F=F, dk=dk, dv=dv,
                               dk_pad=ops.pad_cols(dk), dv_pad=ops.pad_cols(dv), eps=1e-6)
    rays_o = torch.randn(views, 3, device=dev, generator=g) * 3
    rays_d = torch.nn.functional.normalize(torch.randn(R, 3, device=dev, generator=g), dim=-1)
    points = torch.randn(P, 3, device=dev, generator=g)
    feats = torch.randn(P, F, device=dev, generator=g)
    idx = torch.randint(0, P, (R, K), device=dev, generator=g, dtype=torch.int32)
    ln_a = 1 + 0.1 * torch.randn(dk, device=dev, generator=g)
    ln_b = 0.1 * torch.randn(dk, device=dev, generator=g)
    return sh, rays_o, rays_d, points, feats, idx, ln_a, ln_b


def _with_env(flag, fn):
    if flag:
        os.environ["PAPR_PROLOGUE_GENERIC"] = "1"
    try:
        return fn()
    finally:
        os.environ.pop("PAPR_PROLOGUE_GENERIC", None)


@pytest.mark.parametrize("L,F", SHAPES)
@pytest.mark.parametrize("R,views", [(301, 1), (96, 2), (5, 1)])
def test_prologue_forward_families_agree(R, views, L, F):
    sh, rays_o, rays_d, points, feats, idx, ln_a, ln_b = _case(R=R, views=views, L=L, F=F)
    _, _, k32, v32 = A._prologue_fwd(sh, rays_o, rays_d, points, feats, idx, ln_a, ln_b, taps=True)
    outs = {}
    for name, flag in (("generic", True), ("rows", False)):
        kin, vin, _, _ = _with_env(flag, lambda: A._prologue_fwd(sh, rays_o, rays_d, points, feats, idx, ln_a, ln_b))
        outs[name] = (kin.to_f32(sh.M, sh.dk), vin.to_f32(sh.M, sh.dv), kin.to_f32(kin.rows_pad, kin.cols_pad),
                      vin.to_f32(vin.rows_pad, vin.cols_pad))
    for name, (k, v, kfull, vfull) in outs.items():
        # bf16 rounding of the fp32 kernel's values: half an ulp = 2^-9 relative, plus 1e-5 absolute for the statistics
        assert float(((k - k32).abs() - 4e-3 * k32.abs()).max()) <= 2e-5, name
        assert float(((v - v32).abs() - 4e-3 * v32.abs()).max()) <= 2e-5, name
        # padding rows / columns of the tile-blocked tensors are zero
        assert float(kfull[sh.M:].abs().sum()) == 0.0 and float(vfull[sh.M:].abs().sum()) == 0.0, name
        assert float(kfull[:, sh.dk:].abs().sum()) == 0.0 and float(vfull[:, sh.dv:].abs().sum()) == 0.0, name
    # the two bf16 kernels differ only where a value sits on a rounding boundary
    dk = (outs["rows"][0] - outs["generic"][0]).abs()
    assert float((dk > 0).float().mean()) < 0.05 and float((dk - 8e-3 * outs["rows"][0].abs()).max()) <= 2e-5
    dpe = 6 * (1 + 2 * L)
    assert torch.equal(outs["rows"][1][:, dpe:], outs["generic"][1][:, dpe:])       # gathered point features: copies


@pytest.mark.parametrize("L,F", SHAPES)
@pytest.mark.parametrize("R,views", [(301, 1), (96, 2)])
def test_prologue_backward_families_agree(R, views, L, F):
    sh, rays_o, rays_d, points, feats, idx, ln_a, ln_b = _case(R=R, views=views, seed=1, L=L, F=F)
    g = torch.Generator(device="cuda").manual_seed(5)
    dk32 = torch.randn(sh.M, sh.dk, device="cuda", generator=g)
    dv32 = torch.randn(sh.M, sh.dv, device="cuda", generator=g)
    dkin = ops.Blocked.from_f32(dk32, cols_pad=sh.dk_pad)
    dvin = ops.Blocked.from_f32(dv32, cols_pad=sh.dv_pad)
    # reference: the exact fp32 kernel fed with the bf16-rounded gradients the other two read
    ref = A._prologue_bwd(sh, rays_o, rays_d, points, idx, ln_a, None, None,
                          dkin.to_f32(sh.M, sh.dk), dvin.to_f32(sh.M, sh.dv), points.shape[0])
    for name, flag in (("generic", True), ("rows", False)):
        got = _with_env(flag, lambda: A._prologue_bwd(sh, rays_o, rays_d, points, idx, ln_a, dkin, dvin, None, None, points.shape[0]))
        for what, a, b in zip(("g_points", "g_feats", "g_a2", "g_b2"), got, ref):
            scale = float(b.abs().max())
            tol = 2e-2 if (what == "g_a2" and name == "rows") else 2e-3      # rows kernel sums bf16-rounded d kin * z products
            assert float((a - b).abs().max()) <= tol * scale, (name, what, float((a - b).abs().max()), scale)


@pytest.mark.parametrize("R,K,relu", [(301, 20, True), (77, 16, False), (40, 31, True), (33, 32, True)])
def test_score_rows_kernel_matches_warp_per_ray(R, K, relu):
    """Raw scores / LayerNorm statistics by the row-per-lane kernel against the warp-per-ray kernel (same bf16 h5) and
    against a float64 evaluation of ua . LN(h5) + c' (attn.py:39-42, 212-226 after the key-head fold)."""
    g = torch.Generator(device="cuda").manual_seed(R)
    M, C, P = R * K, 32, 500
    sh = types.SimpleNamespace(R=R, K=K, M=M, C=C, score_relu=relu, normalize=True, bkg_score=5.0, eps=1e-6)
    h32 = torch.randn(M, 256, device="cuda", generator=g) * 2 + 0.5
    h5 = ops.Blocked.from_f32(h32)
    hq = h5.to_f32(M, 256).double()
    ua = torch.randn(R, 256, device="cuda", generator=g) * 0.2
    cprime = torch.randn(R, device="cuda", generator=g)
    influ = torch.rand(P, device="cuda", generator=g)
    idx = torch.randint(0, P, (R, K), device="cuda", generator=g, dtype=torch.int32)
    v = torch.randn(ops.pad_rows(M), C, device="cuda", generator=g)
    new = A._score_blend_fwd(sh, h5, None, ua, cprime, influ, idx, v)
    os.environ["PAPR_SCORE_WARP"] = "1"
    try:
        old = A._score_blend_fwd(sh, h5, None, ua, cprime, influ, idx, v)
    finally:
        os.environ.pop("PAPR_SCORE_WARP", None)
    mean = hq.mean(-1, keepdim=True)
    std = hq.std(-1, keepdim=True)
    z = (hq - mean) / (std + 1e-6)
    raw = (z * ua.double().repeat_interleave(K, 0)).sum(-1) + cprime.double().repeat_interleave(K)
    if relu:
        raw = raw.clamp_min(0)
    for name, out in (("rows", new), ("warp", old)):
        fused, attn, sc, stats = out
        assert float((sc.double() - raw).abs().max()) <= 2e-5 * max(1.0, float(raw.abs().max())), name
        assert float((stats[:, 0].double() - mean[:, 0]).abs().max()) <= 1e-5, name
        assert float((stats[:, 1].double() * (std[:, 0] + 1e-6) - 1).abs().max()) <= 1e-5, name
    assert float((new[0] - old[0]).abs().max()) <= 1e-4 * float(old[0].abs().max())
    assert float((new[1] - old[1]).abs().max()) <= 1e-5


@pytest.mark.parametrize("R,K", [(301, 20), (5000, 20), (77, 16), (64, 13), (90, 24), (40, 31), (33, 5)])
def test_key_score_bwd_kernel_families_agree(R, K):
    """d h5 / zsum / dssum of the score backward (the chain through ua . LN(h5), attn.py:39-42): the block-major kernel and
    the warp-per-row kernel against a float64 evaluation on the same bf16 h5.  (A third family that staged a ray's rows
    through shared memory with cp.async one ray ahead, 8 warps per SM, was correct but slower: 4.1 vs 3.6 ms at 800x800.)"""
    g = torch.Generator(device="cuda").manual_seed(R + K)
    M = R * K
    sh = types.SimpleNamespace(R=R, K=K, M=M, C=32, score_relu=True, normalize=True, bkg_score=5.0, eps=1e-6)
    h32 = torch.randn(M, 256, device="cuda", generator=g) * 2 + 0.5
    h5 = ops.Blocked.from_f32(h32)
    hq = h5.to_f32(M, 256).double()
    ua = torch.randn(R, 256, device="cuda", generator=g) * 0.2
    cprime = torch.randn(R, device="cuda", generator=g)
    influ = torch.rand(500, device="cuda", generator=g)
    idx = torch.randint(0, 500, (R, K), device="cuda", generator=g, dtype=torch.int32)
    v = torch.randn(ops.pad_rows(M), 32, device="cuda", generator=g)
    stats = A._score_blend_fwd(sh, h5, None, ua, cprime, influ, idx, v)[3]
    d_score = torch.randn(M, device="cuda", generator=g)
    mean, std = hq.mean(-1, keepdim=True), hq.std(-1, keepdim=True)
    z = (hq - mean) / (std + 1e-6)
    uar, ds = ua.double().repeat_interleave(K, 0), d_score.double()[:, None]
    dz = ds * uar                                            # score = ua . z + c'
    m1 = dz.mean(-1, keepdim=True)
    m2 = (dz * z).sum(-1, keepdim=True) / (255.0 * std)
    want = (dz - m1) / (std + 1e-6) - z * m2                 # through z = (h - mean) / (std + eps), unbiased std
    want_zsum = (ds * z).view(R, K, 256).sum(1)
    outs = {}
    for name, env in (("blocks", None), ("rows", "PAPR_KEY_SCORE_ROWWISE")):
        if env:
            os.environ[env] = "1"
        try:
            dh5, _, zsum, dssum, _ = A._key_score_bwd(sh, d_score, h5, None, stats, ua, want_bias=False)
        finally:
            if env:
                os.environ.pop(env, None)
        torch.cuda.synchronize()
        got = dh5.to_f32(M, 256).double()
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 6e-3 * scale, (name, float((got - want).abs().max()), scale)      # bf16 output
        assert float((zsum.double() - want_zsum).abs().max()) <= 2e-4 * max(1.0, float(want_zsum.abs().max())), name
        assert float((dssum.double() - d_score.double().view(R, K).sum(1)).abs().max()) <= 1e-5, name
        pad = dh5.to_f32(ops.pad_rows(M), 256)[M:]
        assert not bool(pad.any()), name                     # padding rows of the tile-blocked gradient are zero
        outs[name] = got
    assert float((outs["rows"] - outs["blocks"]).abs().max()) <= 8e-3 * float(want.abs().max())       # one bf16 ulp at the top


@pytest.mark.parametrize("L", [6, 4])
@pytest.mark.parametrize("R", [1, 255, 257, 5000])
def test_query_prologue_matches_torch_posenc_and_innorm(L, R):
    """papr_query_prologue_fwd / _bwd against the torch formulation it replaces (utils.py:232-242 + attn.py:30-42)."""
    from papr_b200.nn import LayerNorm
    g = torch.Generator(device="cuda").manual_seed(R + L)
    D = 3 * (1 + 2 * L)
    rays_d = torch.nn.functional.normalize(torch.randn(R, 3, device="cuda", generator=g), dim=-1)
    if R > 4:
        rays_d[3] = torch.tensor([0.0, 0.0, 1.0], device="cuda")          # an axis-aligned direction
    ln = LayerNorm(D, 1e-6).cuda()
    with torch.no_grad():
        ln.a_2.copy_(1 + 0.1 * torch.randn(D, device="cuda", generator=g))
        ln.b_2.copy_(0.1 * torch.randn(D, device="cuda", generator=g))
    want = ln(A.posenc(rays_d.double(), L).float())
    got = A._QueryPrologueFn.apply(rays_d, ln.a_2, ln.b_2, L, ln.eps)
    assert got.shape == want.shape
    # the double-angle recurrence doubles the rounding error per octave: 2^L * 2^-24 on O(1) values, times 1/std
    assert float((got - want).abs().max()) <= 2e-4
    w = torch.randn(R, D, device="cuda", generator=g)
    ga, gb = torch.autograd.grad((got * w).sum(), [ln.a_2, ln.b_2])
    wa, wb = torch.autograd.grad((want * w).sum(), [ln.a_2, ln.b_2])
    scale = max(1.0, float(wa.abs().max()), float(wb.abs().max()))
    assert float((ga - wa).abs().max()) <= 2e-4 * scale * max(1.0, R ** 0.5)
    assert float((gb - wb).abs().max()) <= 1e-5 * scale * max(1.0, R ** 0.5)


def test_query_prologue_rejects_unsupported_orders():
    from papr_b200._lib import PaprError
    rays_d = torch.randn(8, 3, device="cuda")
    a = torch.ones(3 * 11, device="cuda")
    with pytest.raises(PaprError):
        A._QueryPrologueFn.apply(rays_d, a, a, 5, 1e-6)
