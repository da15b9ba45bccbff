"""GPU parity of stage a1 (fused distance + top-K) against the CPU oracle and the reference's golden vectors."""
import numpy as np
import pytest
import torch

from oracle import papr_oracle as O
from papr_b200.config import make_config
from tests.parity import load_golden, topk_sets_match

pytestmark = pytest.mark.gpu


def _run(rays_o, rays_d, points, K, eps=1e-6, cull=None):
    from papr_b200 import ops
    idx = ops.select_topk(rays_o.cuda(), rays_d.cuda(), points.cuda(), K, eps, cull=cull)
    torch.cuda.synchronize()
    return idx.cpu()


def _compare(rays_o, rays_d, points, K, eps=1e-6, cull=None):
    got = _run(rays_o, rays_d, points, K, eps, cull)
    want, kth = O.select_topk(rays_o, rays_d, points, K, eps)
    assert got.dtype == torch.int32 and got.shape == want.shape
    if torch.equal(got.long(), want):      # same order too: (distance, index) ascending
        return 0
    dist = O.select_distances(rays_o, rays_d, points, eps)
    ok, n = topk_sets_match(got, want, lambda i: torch.gather(dist, -1, i), kth)
    assert ok, "top-K sets differ beyond ties at the K-th distance"
    return n


def test_select_golden_reference_sets(golden_dir):
    """Index sets equal the sets the real reference produced (fixture written by oracle/make_golden.py)."""
    g = load_golden(golden_dir, "select_24x24x2_p3000")
    cfg = make_config("chair")
    params = O.init_params(cfg, int(g["P"]), seed=2, cloud="shell")
    assert abs(O.params_checksum(params) - float(g["params_checksum"])) < 1e-6 * abs(float(g["params_checksum"]))
    rays_o, rays_d = torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"])
    got = _run(rays_o, rays_d, params["points"], 20)
    got_sorted = torch.sort(got, -1).values
    ref_sorted = torch.from_numpy(g["idx_sorted"])
    diff = (got_sorted != ref_sorted).any(-1)
    if diff.any():
        dist = O.select_distances(rays_o, rays_d, params["points"])
        ok, _ = topk_sets_match(got, ref_sorted, lambda i: torch.gather(dist, -1, i), torch.from_numpy(g["kth"]))
        assert ok


@pytest.mark.parametrize("H,W,P,K,views,cloud", [
    (16, 16, 3000, 20, 1, "cube"), (7, 5, 257, 20, 3, "shell"), (33, 31, 2049, 30, 2, "cube"),
    (4, 4, 33, 32, 1, "cube"), (1, 1, 21, 20, 1, "cube"), (40, 40, 30000, 20, 1, "shell"), (3, 3, 4100, 1, 2, "cube"),
])
@pytest.mark.parametrize("cull", [False, "grid"])
def test_select_matches_oracle(H, W, P, K, views, cloud, cull):
    cfg = make_config("chair")
    params = O.init_params(cfg, P, seed=H * 131 + P, cloud=cloud)
    rays_o, rays_d, _ = O.synthetic_rays(H * 4, W * 4, cfg.dataset.coord_scale, n_views=views, seed=P, h0=H, h1=2 * H, w0=W, w1=2 * W)
    _compare(rays_o, rays_d, params["points"], K, cull=cull)


def test_select_exact_ties_and_lattice():
    """The reference's cube-lattice init (model.py:246-256) gives many exactly tied distances."""
    cfg = make_config("chair")
    xs = np.linspace(-12, 12, 12)
    pts = torch.tensor(np.array([[i, j, k] for i in xs for j in xs for k in xs]), dtype=torch.float32)
    rays_o, rays_d, _ = O.synthetic_rays(64, 64, 10.0, n_views=1, seed=5, h0=20, h1=44, w0=20, w1=44)
    # axis-aligned rays through the lattice: plenty of exact ties
    rays_d[0, 0, 0] = torch.tensor([0.0, 0.0, -1.0])
    rays_d[0, 0, 1] = torch.tensor([1.0, 0.0, 0.0])
    for cull in (False, "grid", "morton"):
        _compare(rays_o, rays_d, pts, 20, cull=cull)


def test_select_unnormalised_directions_and_duplicates():
    """rays_d is used as given (model.py:277: den = d.d + eps); duplicated points must tie-break by index."""
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(500, 3, generator=g) * 5
    pts = torch.cat([pts, pts[:100]])            # exact duplicates
    rays_o = torch.randn(2, 3, generator=g) * 20
    rays_d = torch.randn(2, 9, 9, 3, generator=g) * torch.rand(2, 9, 9, 1, generator=g) * 3
    want, _ = O.select_topk(rays_o, rays_d, pts, 20)
    for cull in (False, "grid"):
        got = _run(rays_o, rays_d, pts, 20, cull=cull)
        assert torch.equal(got.long(), want), cull          # identical order: (key, index)


def test_select_rejects_bad_arguments():
    from papr_b200 import ops
    with pytest.raises(ValueError):
        ops.select_topk(torch.zeros(1, 3).cuda(), torch.zeros(1, 2, 2, 3).cuda(), torch.zeros(10, 3).cuda(), 20)
    with pytest.raises(ValueError):
        ops.select_topk(torch.zeros(1, 3).cuda(), torch.zeros(1, 2, 2, 3).cuda(), torch.zeros(100, 3).cuda(), 33)


@pytest.mark.parametrize("P,cloud", [(5000, "cube"), (30000, "shell"), (1031, "shell")])
def test_culled_and_plain_kernels_agree_exactly(P, cloud):
    """The spatially culled kernel (Morton groups + bounding spheres) returns the plain scan's result bit for bit."""
    from papr_b200 import ops
    cfg = make_config("chair")
    params = O.init_params(cfg, P, seed=P, cloud=cloud)
    rays_o, rays_d, _ = O.synthetic_rays(256, 256, cfg.dataset.coord_scale, n_views=2, seed=3, h0=100, h1=164, w0=90, w1=150)
    a = ops.select_topk(rays_o.cuda(), rays_d.cuda(), params["points"].cuda(), 20, cull=True)
    b = ops.select_topk(rays_o.cuda(), rays_d.cuda(), params["points"].cuda(), 20, cull=False)
    c = ops.select_topk(rays_o.cuda(), rays_d.cuda(), params["points"].cuda(), 20, cull="grid")
    assert torch.equal(a, b) and torch.equal(c, b)
    want, _ = O.select_topk(rays_o[:1], rays_d[:1, :8, :8], params["points"], 20)
    assert torch.equal(a[:1, :8, :8].cpu().long(), want)


def test_culled_kernel_with_ties_duplicates_and_wide_rays():
    """Lattice ties, exact duplicates and a warp whose rays point in very different directions (culling disabled or
    useless) still give the oracle's (distance, index) order."""
    from papr_b200 import ops
    xs = np.linspace(-12, 12, 11)
    pts = torch.tensor(np.array([[i, j, k] for i in xs for j in xs for k in xs]), dtype=torch.float32)
    pts = torch.cat([pts, pts[:200]])
    g = torch.Generator().manual_seed(4)
    rays_o = torch.randn(2, 3, generator=g) * 25
    rays_d = torch.randn(2, 6, 7, 3, generator=g)
    rays_d[0, 0, :4] = torch.tensor([[0.0, 0.0, -1.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, -2.0, 0.0]])
    want, _ = O.select_topk(rays_o, rays_d, pts, 20)
    for cull in (True, "grid"):
        got = ops.select_topk(rays_o.cuda(), rays_d.cuda(), pts.cuda(), 20, cull=cull).cpu().long()
        assert torch.equal(got, want), cull


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_grid_kernel_with_points_all_around_the_camera(seed):
    """Screen-space culling must stay exact when the cloud surrounds the camera: points behind it (the distance is to the
    LINE, so they count), points perpendicular to the viewing direction (unbounded gnomonic coordinates), a point at the
    ray origin itself, several views with different frames, and a view whose rays point everywhere (no culling possible)."""
    g = torch.Generator().manual_seed(seed)
    P = 5000
    pts = torch.randn(P, 3, generator=g) * 30
    rays_o, rays_d, _ = O.synthetic_rays(96, 96, 10.0, n_views=3, seed=seed, h0=40, h1=56, w0=30, w1=50)
    pts[7] = rays_o[0]
    pts[8] = rays_o[1] + 1e-3
    rays_d[2] = torch.randn(16, 20, 3, generator=g)           # the third view: directions all over the sphere
    want, _ = O.select_topk(rays_o, rays_d, pts, 20)
    got = _run(rays_o, rays_d, pts, 20, cull="grid")
    assert torch.equal(got.long(), want)


def test_grid_kernel_matches_plain_scan_on_a_dense_frame():
    """200x200 rays of the 800x800 frame against 30,000 and 100,000 points: the grid kernel and the plain scan agree on every
    ray (the oracle cannot run this many), also with a tiny eps and K = 32."""
    from papr_b200 import ops
    cfg = make_config("chair")
    for P, K, eps in ((30000, 20, 1e-6), (100000, 32, 1e-12)):
        params = O.init_params(cfg, P, seed=P, cloud="shell")
        rays_o, rays_d, _ = O.synthetic_rays(800, 800, cfg.dataset.coord_scale, n_views=1, seed=7, h0=250, h1=450, w0=300, w1=500)
        a = ops.select_topk(rays_o.cuda(), rays_d.cuda(), params["points"].cuda(), K, eps, cull="grid")
        b = ops.select_topk(rays_o.cuda(), rays_d.cuda(), params["points"].cuda(), K, eps, cull=False)
        assert torch.equal(a, b), (P, K)


@pytest.mark.parametrize("P,views,stripe", [(30000, 1, None), (5000, 3, None), (30000, 1, (700, 800)), (1031, 2, None)])
def test_grid_build_kernels_structure_and_selection(P, views, stripe):
    """papr_select_grid_build (csrc/select_grid_build.cu): the structure it leaves is a valid input of the grid kernel --
    every point of every view stored exactly once, inside the range of the cell its own gnomonic image falls into, v and
    eps|v|^2 as the plain scan stages them, each cell's depth no larger than that of any of its points -- and the selection
    on it equals the selection on the torch-built structure and the plain scan, also for a stripe of the frame that misses
    the object (all points beyond the rays' extent)."""
    from papr_b200 import ops
    cfg = make_config("chair")
    params = O.init_params(cfg, P, seed=P, cloud="shell")
    h0, h1 = stripe if stripe else (300, 364)
    rays_o, rays_d, _ = O.synthetic_rays(800, 800, cfg.dataset.coord_scale, n_views=views, seed=3, h0=h0, h1=h1, w0=200, w1=328)
    ro, rd, pts = rays_o.cuda(), rays_d.cuda().contiguous(), params["points"].cuda()
    N, R, eps = views, rd.shape[1] * rd.shape[2], 1e-6
    sv, perm, cells, vw, G = ops.view_grids(ro, rd.reshape(N, R, 3), pts, eps)
    torch.cuda.synchronize()
    perm_v = perm.view(N, P).long()
    assert torch.equal(perm_v.sort(dim=1).values, torch.arange(P, device="cuda").expand(N, P))
    v = pts[perm_v] - ro[:, None, :]
    sv = sv.view(N, P, 4)
    assert torch.equal(sv[..., :3], v)
    vn2 = torch.addcmul(torch.addcmul(v[..., 0] * v[..., 0], v[..., 1], v[..., 1]), v[..., 2], v[..., 2])
    torch.testing.assert_close(sv[..., 3], eps * vn2, rtol=1e-6, atol=0)
    cells = cells.view(N, G * G, 4)
    start, end = cells[..., 0].long(), cells[..., 1].long()
    assert torch.equal(start[:, 1:], end[:, :-1]) and (start[:, 0] == 0).all() and (end[:, -1] == P).all()
    # cell of every stored point, recomputed from the view parameters the kernel will read
    cell_of_pos = torch.searchsorted(end.contiguous(), torch.arange(P, device="cuda").expand(N, P).contiguous(), right=True)
    e1, e2, c = vw[:, 0:3], vw[:, 3:6], vw[:, 6:9]
    w3 = (v * c[:, None]).sum(-1)
    valid = w3.abs() > 1.01e-3 * vn2.sqrt()
    gx = (v * e1[:, None]).sum(-1) / w3
    gy = (v * e2[:, None]).sum(-1) / w3
    fx = ((gx - vw[:, None, 9]) * vw[:, None, 13])
    fy = ((gy - vw[:, None, 10]) * vw[:, None, 14])
    cx, cy = cell_of_pos % G, cell_of_pos // G
    tol = 1e-3                                              # cells: a point may sit on an edge up to rounding
    inside = ((fx >= cx - tol) | (cx == 0)) & ((fx <= cx + 1 + tol) | (cx == G - 1)) & ((fy >= cy - tol) | (cy == 0)) & ((fy <= cy + 1 + tol) | (cy == G - 1))
    assert bool((inside | ~valid).all())
    zmin = cells[..., 2].contiguous().view(torch.float32)
    depth_ok = zmin.gather(1, cell_of_pos) <= w3.abs() * (1 + 1e-5)
    assert bool(depth_ok.all())
    assert bool((vw[:, 15] <= zmin.amin(1)).all()) and bool((vw[:, 16] >= vn2.amax(1)).all())
    a = ops.select_topk(ro, rd, pts, 20, eps, cull="grid")
    b = ops.select_topk(ro, rd, pts, 20, eps, cull="grid_torch")
    c_ = ops.select_topk(ro, rd, pts, 20, eps, cull=False)
    assert torch.equal(a, c_) and torch.equal(b, c_)
