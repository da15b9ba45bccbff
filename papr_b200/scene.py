"""Procedural NeRF-synthetic-format scene (no datasets are available offline).

Cameras sit on a sphere of radius 4 (x coord_scale) looking at the origin with NeRF-synthetic's camera_angle_x; rays
follow the reference's ``get_rays`` convention (dataset/utils.py:81-96: pixel-centre directions (x, -y, -1) rotated by
c2w and unit-normalised; origin = coord_scale * c2w[:3,3], dataset/dataset.py:19-25).
"""
import math

import torch


def look_at_poses(n_views, seed=1, radius=4.0):
    g = torch.Generator().manual_seed(seed)
    poses = []
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
        poses.append(c2w)
    return torch.stack(poses).float()


def get_rays(H, W, focal, c2w):
    """rays_o (N,3) [unscaled], rays_d (N,H,W,3) unit norm."""
    xs = torch.linspace(0, W / focal, steps=W + 1, dtype=torch.float32)
    ys = torch.linspace(0, H / focal, steps=H + 1, dtype=torch.float32)
    y, x = torch.meshgrid(ys, xs, indexing="ij")
    x = (x - W / focal / 2 + (xs[1] - xs[0]) / 2)[:-1, :-1]
    y = -(y - H / focal / 2 + (ys[1] - ys[0]) / 2)[:-1, :-1]
    dirs = torch.stack([x, y, -torch.ones_like(x)], -1)
    rays_d = torch.einsum("hwj,nij->nhwi", dirs, c2w[:, :3, :3])
    return c2w[:, :3, 3].clone(), rays_d / torch.norm(rays_d, dim=-1, keepdim=True)


def synthetic_scene(H, W, coord_scale, n_views=1, seed=1, camera_angle_x=0.6911):
    """dict(rays_o (N,3), rays_d (N,H,W,3), c2w (N,4,4), target (N,H,W,3)) on the CPU."""
    c2w = look_at_poses(n_views, seed)
    focal = 0.5 * W / math.tan(0.5 * camera_angle_x)
    rays_o, rays_d = get_rays(H, W, focal, c2w)
    g = torch.Generator().manual_seed(seed + 17)
    target = torch.rand(n_views, H, W, 3, generator=g)
    return dict(rays_o=(rays_o * coord_scale).contiguous(), rays_d=rays_d.contiguous(), c2w=c2w, target=target)


def learned_like_cloud(num_points, coord_scale, seed=1, feat_dim=64):
    """A surface-like cloud (noisy shell of radius 0.8*coord_scale) with N(0,1) features and U(0,1) influence scores."""
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(num_points, 3, generator=g)
    p = p / p.norm(dim=-1, keepdim=True) * (0.8 * coord_scale) + 0.02 * coord_scale * torch.randn(num_points, 3, generator=g)
    return dict(points=p.float(), pc_feats=torch.randn(num_points, feat_dim, generator=g),
                points_influ_scores=torch.rand(num_points, 1, generator=g))
