"""Sustained run of the fused stack kernel at full size with SM clock / power sampled alongside (pynvml): separates
"cycles per job" from "clock under the power cap"."""
import os, sys, threading, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
from papr_b200 import ops

L = int(os.environ.get("LAYERS", 8)); rows = 128 * 4 * 74 * int(os.environ.get("QUADS", 160))
stash = os.environ.get("STASH", "0") == "1"
secs = float(os.environ.get("SECS", 3))
torch.manual_seed(0)
x = ops.Blocked.from_f32(torch.randn(rows, 256, device="cuda"))
layers = []
for i in range(L):
    w = torch.randn(256, 256, device="cuda") / 16
    layers.append(dict(w_image=ops.pack_weight(w, 256, 256, replicas=ops.WEIGHT_REPLICAS), N=256, bias=torch.zeros(256, device="cuda"), act=i < L - 1,
                       out_blocked=ops.Blocked(rows, 256, "cuda") if (stash or i == L - 1) else None,
                       sign_bits_out=torch.empty((rows, 4), dtype=torch.int64, device="cuda") if (stash and i < L - 1) else None))
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False
def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.05)
for _ in range(3): ops.stack_bf16(x, 256, layers)
torch.cuda.synchronize()
th = threading.Thread(target=sampler); th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0; t0 = time.time(); e0.record()
while time.time() - t0 < secs:
    for _ in range(5): ops.stack_bf16(x, 256, layers)
    n += 5; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize(); stop = True; th.join()
ms = e0.elapsed_time(e1) / n
jobs_per_cluster = rows / 256 * L / 74
clk = sorted(s[0] for s in samples[len(samples) // 3:]); pw = sorted(s[1] for s in samples[len(samples) // 3:])
mhz = clk[len(clk) // 2]
print(f"stash={int(stash)} rows={rows} L={L}: {ms:.3f} ms/launch, {ms * 1e3 / jobs_per_cluster:.3f} us/job, SM clock median {mhz} MHz "
      f"(min {clk[0]}, max {clk[-1]}), power median {pw[len(pw) // 2]:.0f} W -> {ms * 1e-3 * mhz * 1e6 / jobs_per_cluster:.0f} cycles/job; "
      f"{2 * rows * 256 * 256 * L / ms / 1e9:.0f} TFLOP/s")
