"""Comparison helpers shared by the parity tests."""
import numpy as np
import torch


def load_golden(golden_dir, name):
    import os
    rec = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    return {k: rec[k] for k in rec.files}


def topk_sets_match(idx_a, idx_b, dist_of, kth):
    """Tie-aware equality of two top-K index tensors (..., K).

    The reference's topk(sorted=False) leaves order and tie resolution unspecified, so sets are compared and an
    index may differ only if its distance (``dist_of(idx)`` -> same shape, oracle fp32 distances) equals the K-th
    distance ``kth`` (...,) of that ray.  Returns (ok, n_rays_with_tie_differences)."""
    a = torch.sort(idx_a.long(), dim=-1).values
    b = torch.sort(idx_b.long(), dim=-1).values
    diff = (a != b).any(-1)
    if not diff.any():
        return True, 0
    da, db = dist_of(idx_a.long()), dist_of(idx_b.long())
    ok = True
    for d, idx_mine, idx_other in ((da, idx_a.long(), idx_b.long()), (db, idx_b.long(), idx_a.long())):
        extra = ~(idx_mine.unsqueeze(-1) == idx_other.unsqueeze(-2)).any(-1)   # entries not in the other set
        bad = extra & (d != kth.unsqueeze(-1)) & diff.unsqueeze(-1)
        ok = ok and not bool(bad.any())
    return ok, int(diff.sum())
