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


def golden_config(name):
    """Config variants the golden fixtures were generated with (oracle/make_golden.py CASES)."""
    from papr_b200.config import make_config
    if name == "chair":
        return make_config("chair", use_amp=False)
    if name == "caterpillar_exposure":
        return make_config("caterpillar_exposure", use_amp=False)
    if name == "lego_like":     # configs/nerfsyn/lego.yml:10-15 style: leakyrelu + value skip layer
        emb = dict(key=dict(ff_act="leakyrelu"), query=dict(ff_act="leakyrelu"),
                   value=dict(ff_act="leakyrelu", skip_layers=[5]))
        return make_config("chair", use_amp=False, models=dict(attn=dict(embed=emb)))
    if name == "no_renderer":   # models.use_renderer false + value d_ff_out 3 (model.py:77-79)
        return make_config("chair", use_amp=False,
                           models=dict(use_renderer=False, attn=dict(embed=dict(value=dict(d_ff_out=3)))))
    if name == "hotdog_like":   # select_k 30 (hotdog.yml), feature dim 128 (materials.yml)
        return make_config("chair", use_amp=False, geoms=dict(points=dict(select_k=30), point_feats=dict(dim=128)))
    raise KeyError(name)


GOLDEN_CASES = ["chair_12x12_p800", "chair_2views_8x8_p500", "caterpillar_exposure_12x16_p600",
                "lego_like_8x8_p400", "no_renderer_8x8_p400", "hotdog_like_8x8_p400"]


def golden_params(g):
    """Regenerate the seeded parameters a fixture was produced with and check their fingerprint."""
    from oracle import papr_oracle as O
    cfg = golden_config(str(g["variant"]))
    params = O.init_params(cfg, int(g["P"]), seed=1, cloud=str(g["cloud"]))
    chk = float(g["params_checksum"])
    assert abs(O.params_checksum(params) - chk) <= 1e-9 * abs(chk), "seeded parameters differ from the fixture's"
    return cfg, params


def rel_err(a, b):
    """max |a-b| relative to the scale of b."""
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
