"""Python wrappers over the C ABI: allocate outputs with torch, pass raw device pointers and the current stream."""
import torch

from ._lib import check, lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t):
    assert t.is_cuda, "papr_b200 ops need CUDA tensors (there is no CPU path)"
    return t.detach().contiguous().float()


def select_topk(rays_o, rays_d, points, K, eps=1e-6):
    """Stage a1 (reference models/model.py:258-283): int32 (N,H,W,K) nearest-point indices per ray,
    ordered by (distance, index).  rays_o (N,3), rays_d (N,H,W,3), points (P,3), all CUDA fp32."""
    N, H, W, _ = rays_d.shape
    P = points.shape[0]
    if not (1 <= K <= 32) or K >= P:
        raise ValueError(f"select_topk needs 1 <= K <= 32 and K < P (K={K}, P={P})")
    ro, rd, pts = _f32c(rays_o), _f32c(rays_d), _f32c(points)
    idx = torch.empty((N, H, W, K), dtype=torch.int32, device=rd.device)
    check(lib().papr_select_topk(ro.data_ptr(), rd.data_ptr(), pts.data_ptr(), N, H * W, P, K, float(eps),
                                 idx.data_ptr(), _stream()), "papr_select_topk")
    return idx
