"""Times the select kernel alone at the C2 shape (800x800 rays, P=30k, K=20) with CUDA events."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from papr_b200 import ops
H = W = int(os.environ.get("HW", 800)); P = int(os.environ.get("P", 30000))
g = torch.Generator().manual_seed(0)
pts = ((torch.rand(P, 3, generator=g) * 2 - 1) * 12).cuda()
if os.environ.get('CLOUD') == 'shell':
    pts = torch.randn(P, 3, generator=g); pts = (pts / pts.norm(dim=-1, keepdim=True) * 8 + 0.2 * torch.randn(P, 3, generator=g)).cuda()
CULL = {'0': False, '1': True, 'grid': 'grid'}.get(os.environ.get('CULL', ''), None)
if os.environ.get('RAYS') == 'random':
    d = torch.randn(1, H, W, 3, generator=g); d = (d / d.norm(dim=-1, keepdim=True)).cuda()
    o = torch.tensor([[30.0, 20.0, 15.0]]).cuda()
else:
    from papr_b200.scene import synthetic_scene
    sc = synthetic_scene(H, int(os.environ.get("WID", W)), 10.0)
    W = int(os.environ.get("WID", W))
    d, o = sc['rays_d'].cuda(), sc['rays_o'].cuda()
for _ in range(3):
    idx = ops.select_topk(o, d, pts, 20, cull=CULL)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); idx = ops.select_topk(o, d, pts, 20, cull=CULL); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sorted(ts)[len(ts) // 2]
pairs = H * W * P
print(json.dumps({"select_ms": ms, "all": ts, "pairs_per_s": pairs / ms * 1e3, "flop17_tflops": pairs * 17 / ms / 1e9}))
ops.STATS.reset(); ops.STATS.timing = True
idx = ops.select_topk(o, d, pts, 20, cull=CULL); torch.cuda.synchronize()
print({k: round(v["ms"], 3) for k, v in ops.STATS.summary().items()})
import ctypes
from papr_b200._lib import lib
fb = getattr(ctypes.CDLL(lib()._name), "papr_debug_select_fallbacks"); fb.restype = ctypes.c_longlong
fb(); idx = ops.select_topk(o, d, pts, 20, cull=CULL); torch.cuda.synchronize()
print("fallback rays:", fb(), "of", H * W)
