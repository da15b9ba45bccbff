"""Diagnostic dump for the tcgen05 building blocks (prints error structure instead of asserting)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from papr_b200 import ops
torch.manual_seed(0)
def bf(t): return t.bfloat16().float()
for (rows, N, K) in [(128, 256, 64), (128, 256, 256), (256, 64, 128), (1000, 32, 256)]:
    x = torch.randn(rows, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5
    yb, yf, _ = ops.linear_bf16(ops.Blocked.from_f32(x), ops.pack_weight(w, N, K), N, K, out_f32=True)
    torch.cuda.synchronize()
    want = bf(x) @ bf(w).t()
    e = (yf[:rows] - want).abs()
    print(f"linear rows={rows} N={N} K={K}: max err {e.max().item():.4g}, frac bad {(e > 1e-2).float().mean().item():.3f}")
    if e.max() > 1e-2:
        print("  bad rows:", (e > 1e-2).any(1).nonzero().flatten()[:16].tolist(), " bad cols:", (e > 1e-2).any(0).nonzero().flatten()[:16].tolist())
        print("  got[0,:8]", yf[0, :8].tolist(), "\n  want[0,:8]", want[0, :8].tolist())
        for kk in range(0, K, 16):   # which K slices contribute?
            part = bf(x[:, kk:kk + 16]) @ bf(w[:, kk:kk + 16]).t()
            print(f"   k-slice {kk}: corr {torch.corrcoef(torch.stack([yf[:rows].flatten(), part.flatten()]))[0, 1].item():.3f}")
for (rows, A, B) in [(64, 128, 64), (128, 256, 256), (640, 256, 192)]:
    a = torch.randn(rows, A, device="cuda"); b = torch.randn(rows, B, device="cuda")
    out = torch.zeros(A, B, device="cuda")
    ops.wgrad_bf16(ops.Blocked.from_f32(a), ops.Blocked.from_f32(b), out, A, B)
    torch.cuda.synchronize()
    want = bf(a).t() @ bf(b)
    e = (out - want).abs()
    print(f"wgrad rows={rows} A={A} B={B}: max err {e.max().item():.4g}, frac bad {(e > 5e-2).float().mean().item():.3f}")
    if e.max() > 5e-2:
        print("  bad a:", (e > 5e-2).any(1).nonzero().flatten()[:16].tolist(), " bad b:", (e > 5e-2).any(0).nonzero().flatten()[:16].tolist())
        print("  got[0,:8]", out[0, :8].tolist(), "\n  want[0,:8]", want[0, :8].tolist())
        print("  got^T match?", (out - (bf(b).t() @ bf(a)).t()).abs().max().item() if A == B else "n/a")
