"""One full-size training step (and one render) between cudaProfilerStart/Stop, for ncu (--profile-from-start off)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from papr_b200.config import make_config
from papr_b200.model import PAPR
from papr_b200.scene import learned_like_cloud, synthetic_scene
HW = int(os.environ.get("HW", 800)); P = int(os.environ.get("P", 30000)); WARM = int(os.environ.get("WARM", 1))
dev = torch.device("cuda", 0)
torch.manual_seed(1)
cfg = make_config("chair"); cfg.geoms.points["init_num"] = P
model = PAPR(cfg, device=dev).to(dev)
cloud = learned_like_cloud(P, cfg.dataset.coord_scale)
with torch.no_grad():
    model.points.copy_(cloud["points"]); model.pc_feats.copy_(cloud["pc_feats"]); model.points_influ_scores.copy_(cloud["points_influ_scores"])
model.init_optimizers(0)
b = {k: v.to(dev) for k, v in synthetic_scene(HW, HW, cfg.dataset.coord_scale).items()}
def step():
    model.clear_grad()
    out = model(b["rays_o"], b["rays_d"], b["c2w"])
    loss = torch.mean((out - b["target"]) ** 2)
    loss.backward(); model.step()
for _ in range(WARM): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
if os.environ.get("RENDER", "1") == "1":
    with torch.no_grad(): model(b["rays_o"], b["rays_d"], b["c2w"])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
