"""Point growing (stage a14): reference models/utils.py:9-109 ``add_points_knn`` on the GPU.

The reference copies the cloud to the host and queries a scipy KDTree; here the k-nearest-neighbour search is the
library's exact brute-force kernel (``papr_knn``: float64 distances, one warp per query, nothing of size P x P ever
exists), which returns the same neighbours as the KDTree up to exact distance ties.  The random convex weights come
from numpy's global RNG with the reference's call (np.random.uniform(0, 1, (Nq, k))) so that a seeded run draws the same
numbers.
"""
import numpy as np
import torch

from . import ops


def knn(points, queries, k, chunk=4096):
    """(dist float64 (Q,k), idx int64 (Q,k)) of the k nearest `points` for every query, ascending (self included).
    CUDA tensors go through papr_knn; host tensors (the CPU-constructed container of tests/test_host_logic.py) use torch."""
    if points.is_cuda and k <= 32:
        return ops.knn(points, queries, k)
    p64 = points.double()
    dists, inds = [], []
    for s in range(0, queries.shape[0], chunk):
        d = torch.cdist(queries[s:s + chunk].double(), p64, compute_mode="donot_use_mm_for_euclid_dist")
        dd, ii = torch.topk(d, k, dim=-1, largest=False, sorted=True)
        dists.append(dd)
        inds.append(ii)
    return torch.cat(dists), torch.cat(inds)


def add_points_knn(coords, influ_scores, add_num, k, comb_type="mean", sample_type="random", sample_k=10,
                   point_features=None):
    """Returns (new_coords, n_new, new_influ_scores, new_features) on coords.device."""
    N = coords.shape[0]
    dev = coords.device
    if N <= add_num and "random" in comb_type:
        inds = torch.from_numpy(np.random.choice(N, add_num, replace=True)).to(dev)
    elif N <= add_num:
        inds = torch.arange(N, device=dev)
    elif sample_type == "random":
        inds = torch.from_numpy(np.random.choice(N, add_num, replace=False)).to(dev)
    elif sample_type.startswith("top-knn-"):
        assert k >= 2
        nd, _ = knn(coords, coords, sample_k)
        stat = {"std": lambda t: t.std(dim=-1, unbiased=False), "mean": lambda t: t.mean(-1),
                "max": lambda t: t.max(-1).values, "min": lambda t: t.min(-1).values}[sample_type[8:]](nd)
        inds = torch.argsort(stat, stable=True)[-add_num:]
    elif sample_type == "influ-scores-max":
        inds = torch.argsort(influ_scores.squeeze(), stable=True)[-add_num:]
    elif sample_type == "influ-scores-min":
        inds = torch.argsort(influ_scores.squeeze(), stable=True)[:add_num]
    else:
        raise NotImplementedError(sample_type)
    query = coords[inds]

    feats = None
    if comb_type == "duplicate":
        noise = np.random.randn(3).astype(np.float32)
        noise = noise / np.linalg.norm(noise) * k
        new_coords = query + torch.from_numpy(noise).to(dev)
        new_influ = influ_scores[inds]
        if point_features is not None:
            feats = point_features[inds]
        return new_coords, len(new_coords), new_influ, feats

    nd, ni = knn(coords, query, k + 1)
    nd, ni = nd[:, 1:].float(), ni[:, 1:]
    Q = query.shape[0]
    if comb_type == "mean":
        w = torch.full((Q, k), 1.0 / k, device=dev)
    elif comb_type == "random":
        w = np.random.uniform(0, 1, (Q, k)).astype(np.float32)
        w = torch.from_numpy(w / w.sum(axis=-1, keepdims=True)).to(dev)
    elif comb_type == "random-softmax":
        w = torch.softmax(torch.from_numpy(np.random.randn(Q, k).astype(np.float32)), dim=-1).to(dev)
    elif comb_type == "weighted":
        w = 1 / (nd + 1e-6)
        w = w / w.sum(-1, keepdim=True)
    else:
        raise NotImplementedError(comb_type)
    w3 = w.reshape(Q, k, 1)
    new_coords = (coords[ni] * w3).sum(-2)
    new_influ = (influ_scores[ni] * w3).sum(-2)
    if point_features is not None:
        feats = (point_features[ni] * w3).sum(-2)
    return new_coords, len(new_coords), new_influ, feats
