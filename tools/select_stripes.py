"""Selection time for row stripes of the 800x800 frame (what the ranks of a row-sharded render / training step see)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from papr_b200 import ops
from papr_b200.scene import learned_like_cloud, synthetic_scene
dev = torch.device("cuda", 0)
cloud = learned_like_cloud(30000, 10.0)
pts = cloud["points"].to(dev)
b = {k: v.to(dev) for k, v in synthetic_scene(800, 800, 10.0).items()}
full = ops.select_topk(b["rays_o"], b["rays_d"], pts, 20)
for r0, rows in ((0, 800), (0, 132), (100, 148), (326, 148), (552, 148), (668, 132)):
    rd = b["rays_d"][:, r0:r0 + rows].contiguous()
    for _ in range(2): idx = ops.select_topk(b["rays_o"], rd, pts, 20)
    torch.cuda.synchronize()
    ops.STATS.reset(); ops.STATS.timing = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): idx = ops.select_topk(b["rays_o"], rd, pts, 20)
    e1.record(); torch.cuda.synchronize(); ops.STATS.timing = False
    k = ops.STATS.summary()["papr_select_topk_grid"]["ms"] / 3
    print(f"rows {r0:3d}..{r0 + rows:3d}: kernel {k:6.3f} ms, with the grid build {e0.elapsed_time(e1) / 3:6.3f} ms, identical to the full-frame result: {torch.equal(idx, full[:, r0:r0 + rows])}")
