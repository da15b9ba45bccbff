"""torch.profiler breakdown of one full-size training step (which torch-side kernels surround the library's)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import profile, ProfilerActivity
from papr_b200.config import make_config
from papr_b200.model import PAPR
from papr_b200.scene import learned_like_cloud, synthetic_scene
HW = int(os.environ.get("HW", 800)); P = 30000
dev = torch.device("cuda", 0); torch.manual_seed(1)
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
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
    if e.device_type.name == "CUDA" and t > 0: rows.append((t / 1e3, e.count, e.key[:100]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device ms {tot:.1f}")
for t, n, k in rows[:45]: print(f"{t:8.3f} ms x{n:4d} {k}")

# which aten ops (with input shapes) own the torch-side device time
ops_rows = []
for e in prof.key_averages(group_by_input_shape=True):
    t = getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0)
    if e.device_type.name == "CPU" and t > 50 and e.key.startswith("aten::"):
        ops_rows.append((t / 1e3, e.count, e.key, str(e.input_shapes)[:120]))
ops_rows.sort(reverse=True)
print("aten ops by self device time:")
for t, n, k, sh in ops_rows[:30]: print(f"{t:8.3f} ms x{n:4d} {k:28s} {sh}")
