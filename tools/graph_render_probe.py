import os, sys, time, torch
sys.path.insert(0, "/root/repo")
from papr_b200 import ops
from papr_b200.config import make_config
from papr_b200.model import PAPR
from papr_b200.scene import learned_like_cloud, synthetic_scene
from papr_b200.staging import GraphedCall
dev = torch.device("cuda", 0); torch.manual_seed(1)
cfg = make_config("chair"); P = 30000; cfg.geoms.points["init_num"] = P
model = PAPR(cfg, device=dev).to(dev)
cloud = learned_like_cloud(P, cfg.dataset.coord_scale)
with torch.no_grad():
    model.points.copy_(cloud["points"]); model.pc_feats.copy_(cloud["pc_feats"]); model.points_influ_scores.copy_(cloud["points_influ_scores"])
b = {k: v.to(dev) for k, v in synthetic_scene(800, 800, cfg.dataset.coord_scale).items()}
for rows in (800, 148):
    rd = b["rays_d"][:, 300:300 + rows].contiguous() if rows < 800 else b["rays_d"]
    ro = b["rays_o"]
    fn = lambda o, d: model(o, d, None)
    with torch.no_grad():
        ref = fn(ro, rd).clone()
    g = GraphedCall(fn, [ro, rd])
    out = g(ro, rd)
    torch.cuda.synchronize()
    print("rows", rows, "graph == eager:", torch.equal(out, ref), float((out - ref).abs().max()))
    rd2 = rd.flip(2).contiguous()
    with torch.no_grad():
        ref2 = fn(ro, rd2).clone()
    out2 = g(ro, rd2); torch.cuda.synchronize()
    print("   other input:", torch.equal(out2, ref2))
    for name, call in (("eager", lambda: fn(ro, rd)), ("graph", lambda: g(ro, rd))):
        with torch.no_grad():
            for _ in range(3): call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): call()
            e1.record(); torch.cuda.synchronize()
        print(f"   {name}: {e0.elapsed_time(e1) / 10:.2f} ms/frame")
