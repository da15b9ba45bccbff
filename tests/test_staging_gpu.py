"""GPU: a frame render replayed as a CUDA graph (papr_b200.staging.GraphedCall) equals the eagerly enqueued one bit for bit,
also on inputs it was not captured with."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_graphed_render_is_bit_identical():
    from papr_b200.config import make_config
    from papr_b200.model import PAPR
    from papr_b200.scene import learned_like_cloud, synthetic_scene
    from papr_b200.staging import GraphedCall
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    cfg = make_config("chair")
    P = 4000
    cfg.geoms.points["init_num"] = P
    model = PAPR(cfg, device=dev).to(dev)
    cloud = learned_like_cloud(P, cfg.dataset.coord_scale)
    with torch.no_grad():
        model.points.copy_(cloud["points"]); model.pc_feats.copy_(cloud["pc_feats"])
        model.points_influ_scores.copy_(cloud["points_influ_scores"])
    b = {k: v.to(dev) for k, v in synthetic_scene(96, 128, cfg.dataset.coord_scale).items()}
    ro, rd = b["rays_o"], b["rays_d"][:, 16:80].contiguous()           # a 64-row stripe: the sharded-render case
    fn = lambda o, d: model(o, d, None, step=-1)
    with torch.no_grad():
        ref = fn(ro, rd).clone()
    g = GraphedCall(fn, [ro, rd])
    assert torch.equal(g(ro, rd), ref)
    rd2 = rd.flip(2).contiguous()
    with torch.no_grad():
        ref2 = fn(ro, rd2).clone()
    assert torch.equal(g(ro, rd2), ref2)
    assert not torch.equal(ref, ref2)
