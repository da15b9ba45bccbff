"""Render time of row stripes of the 800x800 frame on one GPU (what a rank of an N-GPU render does): fixed cost vs rows."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from papr_b200 import ops
from papr_b200.config import make_config
from papr_b200.model import PAPR
from papr_b200.scene import learned_like_cloud, synthetic_scene
dev = torch.device("cuda", 0); torch.manual_seed(1)
cfg = make_config("chair"); P = 30000; cfg.geoms.points["init_num"] = P
model = PAPR(cfg, device=dev).to(dev)
cloud = learned_like_cloud(P, cfg.dataset.coord_scale)
with torch.no_grad():
    model.points.copy_(cloud["points"]); model.pc_feats.copy_(cloud["pc_feats"]); model.points_influ_scores.copy_(cloud["points_influ_scores"])
b = {k: v.to(dev) for k, v in synthetic_scene(800, 800, cfg.dataset.coord_scale).items()}
for rows in (800, 432, 232, 132, 64):
    rd = b["rays_d"][:, 300:300 + rows].contiguous() if rows < 800 else b["rays_d"]
    def render():
        with torch.no_grad():
            return model(b["rays_o"], rd, None)
    for _ in range(3): render()
    torch.cuda.synchronize()
    ops.STATS.reset(); ops.STATS.timing = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(5): render()
    e1.record(); host = (time.perf_counter() - t0) / 5 * 1e3
    torch.cuda.synchronize(); ops.STATS.timing = False
    k = ops.STATS.summary()
    own = sum(v["ms"] for v in k.values()) / 5
    print(f"rows {rows}: {e0.elapsed_time(e1) / 5:.2f} ms/frame (host enqueue {host:.2f} ms, own kernels {own:.2f} ms, {ops.STATS.count // 5} own launches)")
    if rows in (132, 800): print("   ", {n: round(v["ms"] / 5, 3) for n, v in sorted(k.items(), key=lambda kv: -kv[1]["ms"])[:9]})
