"""Times the select kernel alone at the C2 shape (800x800 rays, P=30k, K=20) with CUDA events."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from papr_b200 import ops
H = W = int(os.environ.get("HW", 800)); P = int(os.environ.get("P", 30000))
g = torch.Generator().manual_seed(0)
pts = ((torch.rand(P, 3, generator=g) * 2 - 1) * 12).cuda()
d = torch.randn(1, H, W, 3, generator=g); d = (d / d.norm(dim=-1, keepdim=True)).cuda()
o = torch.tensor([[30.0, 20.0, 15.0]]).cuda()
for _ in range(3):
    idx = ops.select_topk(o, d, pts, 20)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); idx = ops.select_topk(o, d, pts, 20); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
pairs = H * W * P
print(json.dumps({"select_ms": ms, "all": ts, "pairs_per_s": pairs / ms * 1e3, "flop17_tflops": pairs * 17 / ms / 1e9}))
