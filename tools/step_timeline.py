"""Host-enqueue time vs device time of one full-size training step, phase by phase (is the step host-bound?)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from papr_b200.config import make_config
from papr_b200.model import PAPR
from papr_b200.scene import learned_like_cloud, synthetic_scene
HW = int(os.environ.get("HW", 800)); P = int(os.environ.get("P", 30000))
dev = torch.device("cuda", 0); torch.manual_seed(1)
cfg = make_config("chair"); cfg.geoms.points["init_num"] = P
model = PAPR(cfg, device=dev).to(dev)
cloud = learned_like_cloud(P, cfg.dataset.coord_scale)
with torch.no_grad():
    model.points.copy_(cloud["points"]); model.pc_feats.copy_(cloud["pc_feats"]); model.points_influ_scores.copy_(cloud["points_influ_scores"])
model.init_optimizers(0)
b = {k: v.to(dev) for k, v in synthetic_scene(HW, HW, cfg.dataset.coord_scale).items()}
ev = lambda: torch.cuda.Event(enable_timing=True)
def step(sync_phases=False):
    marks, host = [ev() for _ in range(5)], []
    t0 = time.perf_counter(); marks[0].record()
    model.clear_grad()
    idx = model._get_points(b["rays_o"], b["rays_d"])
    marks[1].record(); host.append(time.perf_counter() - t0)
    if sync_phases: torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = model(b["rays_o"], b["rays_d"], b["c2w"])
    loss = torch.mean((out - b["target"]) ** 2)
    marks[2].record(); host.append(time.perf_counter() - t0)
    if sync_phases: torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss.backward()
    marks[3].record(); host.append(time.perf_counter() - t0)
    if sync_phases: torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.step()
    marks[4].record(); host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    dev_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(4)]
    return [h * 1e3 for h in host], dev_ms
for _ in range(3): step()
for mode in (False, True):
    hs, ds = [], []
    for _ in range(4):
        h, d = step(mode); hs.append(h); ds.append(d)
    h = [sum(x[i] for x in hs) / len(hs) for i in range(4)]; d = [sum(x[i] for x in ds) / len(ds) for i in range(4)]
    print(f"sync between phases={mode}: host enqueue ms [select(extra), forward, backward, step] = {[round(x, 2) for x in h]} (sum {sum(h):.1f}); "
          f"device ms between marks = {[round(x, 2) for x in d]} (sum {sum(d):.1f})")
print("mem: allocated %.1f GB reserved %.1f GB" % (torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
st = torch.cuda.memory_stats()
print("cudaMalloc retries:", st.get("num_alloc_retries"), "ooms:", st.get("num_ooms"), "segments:", st.get("segment.all.current"))

import cProfile, pstats, io
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
model.clear_grad()
out = model(b["rays_o"], b["rays_d"], b["c2w"])
loss = torch.mean((out - b["target"]) ** 2)
pr.disable()
torch.cuda.synchronize()
sio = io.StringIO()
pstats.Stats(pr, stream=sio).sort_stats("cumulative").print_stats(45)
print(sio.getvalue()[:9000])
