"""Per-job timeline of the fused stack kernel (cluster 0) from in-kernel clock stamps."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from papr_b200 import ops
from papr_b200._lib import lib
L = int(os.environ.get("LAYERS", 8)); rows = 128 * 4 * 74 * int(os.environ.get("QUADS", 6))
torch.manual_seed(0)
x = ops.Blocked.from_f32(torch.randn(rows, 256, device="cuda"))
layers = []
stash = os.environ.get("STASH", "0") == "1"
for i in range(L):
    w = torch.randn(256, 256, device="cuda") / 16
    layers.append(dict(w_image=ops.pack_weight(w, 256, 256, replicas=int(os.environ.get('REPS', 1))), N=256, bias=torch.zeros(256, device="cuda"), act=i < L - 1,
                       out_blocked=ops.Blocked(rows, 256, "cuda") if (stash or i == L - 1) else None))
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
h = lib(); fn = getattr(ctypes.CDLL(h._name), "papr_debug_stack_trace"); fn.argtypes = [ctypes.c_void_p]
for _ in range(2): ops.stack_bf16(x, 256, layers)
fn(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.stack_bf16(x, 256, layers); e1.record(); torch.cuda.synchronize()
fn(None)
ms = e0.elapsed_time(e1)
clk = buf.cpu()[4000:4004]
fin = [int(v) - int(clk[1]) for v in buf.cpu()[3900:3974] if int(v)]
if fin:
    fin.sort()
    print(f"CTA pairs finish (ms after the start of pair 0): first {fin[0] / 1e6:.3f}, median {fin[len(fin) // 2] / 1e6:.3f}, last {fin[-1] / 1e6:.3f} "
          f"-> the static tile assignment costs {100.0 * (fin[-1] - sum(fin) / len(fin)) / fin[-1]:.1f}% of the launch")
if int(clk[3]) > int(clk[1]):
    print(f"SM clock during the launch (CTA 0, clock64 / globaltimer): {(int(clk[2]) - int(clk[0])) / (int(clk[3]) - int(clk[1])):.3f} GHz over {(int(clk[3]) - int(clk[1])) / 1e6:.3f} ms")
print(f"kernel {ms:.3f} ms for {rows // 128 * L} tile-layers -> {ms * 1e-3 * 1.9e9 / (rows // 128 * L / 148):.0f} cycles/tile-layer/SM (at 1.9 GHz)")
t = buf.cpu()[:1024].reshape(4, 16, 2, 8); tw = buf.cpu()[2048:2048 + 256].reshape(4, 16, 2, 2); tc = buf.cpu()[3072:3072 + 128].reshape(4, 16, 2)
t0 = int(t[0, 0, 0, 0])
print("quad layer slot | mma_start mma_commit(issue) | L:epi_start epi_done arrived | P:epi_start epi_done arrived   (cycles since first MMA start)")
for q in range(2):
    for l in range(L):
        for s in range(2):
            r = [int(v) - t0 if int(v) else -1 for v in t[q, l, s]]
            print(f"{q} {l} {s} | {r[0]:7d} {r[1]:7d} | {r[2]:7d} {r[3]:7d} {r[4]:7d} | {r[5]:7d} {r[6]:7d} {r[7]:7d} | issuer waited: tile {int(tw[q, l, s, 0]):5d} weights {int(tw[q, l, s, 1]):5d}")
# steady state (second quad): job period from accumulator-complete stamps, issuer waits
done = [int(t[1, l, s, 2]) for l in range(L) for s in range(2)]
per = [(b - a) for a, b in zip(done[:-1], done[1:])]
ww = [int(tw[1, l, s, 1]) for l in range(L) for s in range(2)]
wt = [int(tw[1, l, s, 0]) for l in range(L) for s in range(2)]
epi = [int(t[1, l, s, 3]) - int(t[1, l, s, 2]) for l in range(L) for s in range(2)]
stw = [int(t[1, l, s, 7]) - int(t[1, l, s, 2]) for l in range(L) for s in range(2)]
body = [int(t[1, l, s, 5]) - int(t[1, l, s, 7]) for l in range(L) for s in range(2)]
fen = [int(t[1, l, s, 6]) - int(t[1, l, s, 5]) for l in range(L) for s in range(2)]
arr = [int(t[1, l, s, 4]) - int(t[1, l, s, 3]) for l in range(L) for s in range(2)]
print(f"   epilogue of warp 4: stash-read wait {sum(stw) / len(stw):5.0f}  drain+math+st.shared {sum(body) / len(body):5.0f}  fence.proxy.async {sum(fen) / len(fen):5.0f}  syncwarp+arrive {sum(arr) / len(arr):5.0f}")
print(f"SUMMARY stash={int(stash)} nsplit={os.environ.get('PAPR_STACK_NSPLIT', '1')} share_w={os.environ.get('PAPR_STACK_SHARE_W', '0')} epi={os.environ.get('PAPR_DBG_STACK_EPI', '0')}: "
      f"cycles/job {sum(per) / len(per):6.0f}  epilogue {sum(epi) / len(epi):6.0f}  issuer waits: tile {sum(wt) / len(wt):5.0f} weights {sum(ww) / len(ww):5.0f}  kernel {ms:.3f} ms")
