"""Feasibility probe for passing the backward's dZ tiles between kernels through L2 instead of HBM.

Runs the training-mode stack kernel (6 x 256->256, stash + sign bits) and the weight-gradient kernel on 3.2 M rows under
the debug switches PAPR_DBG_STACK_RING (stash tiles written to a small ring: tile % R), PAPR_DBG_STACK_GRID and
PAPR_DBG_WGRAD_GRID (cap the number of CTAs), plus a write-only / read-only / copy HBM rate for reference.  The ring
results are garbage by construction: this only measures time.  One configuration per process (the switches are read once).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from papr_b200 import ops  # noqa: E402


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device("cuda:0")
    what = sys.argv[1]
    rows = 400 * 400 * 20
    L = 6
    if what == "hbm":
        n = 8 << 30
        a = torch.empty(n, dtype=torch.uint8, device=dev)
        b = torch.empty(n, dtype=torch.uint8, device=dev)
        print(f"memset (write only)  {n / timed(lambda: a.zero_()) / 1e6:8.1f} GB/s")
        print(f"copy (read + write)  {2 * n / timed(lambda: b.copy_(a)) / 1e6:8.1f} GB/s")
        af = a.view(torch.float32)
        print(f"sum (read only)      {n / timed(lambda: af.sum()) / 1e6:8.1f} GB/s")
        return
    torch.manual_seed(0)
    ws = [torch.randn(256, 256, device=dev) / 16 for _ in range(L)]
    bs = [torch.zeros(256, device=dev) for _ in range(L)]
    x = ops.Blocked.from_f32(torch.randn(rows, 256, device=dev))
    imgs = [ops.pack_weight(w, 256, 256, replicas=ops.WEIGHT_REPLICAS) for w in ws]
    ring = int(os.environ.get("PAPR_DBG_STACK_RING", "0"))
    out_rows = ring * 128 if ring else rows
    outs = [ops.Blocked(out_rows, 256, dev) for _ in range(L)]
    bits = [torch.empty((rows, 4), dtype=torch.int64, device=dev) for _ in range(L)]
    if what == "stack":
        layers = [dict(w_image=imgs[i], N=256, bias=bs[i], act=True, out_blocked=outs[i], sign_bits_out=bits[i]) for i in range(L)]
        ms = timed(lambda: ops.stack_bf16(x, 256, layers))
        fl = 2.0 * rows * 256 * 256 * L
        print(f"stack fwd+stash ring={ring} grid={os.environ.get('PAPR_DBG_STACK_GRID', 'all')}: {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  "
              f"stash {rows * 512.0 * L / ms / 1e6:7.1f} GB/s")
    elif what in ("dgrad", "dgrad_nocs"):
        layers = [dict(w_image=imgs[i], N=256, out_blocked=outs[i], sign_bits_in=bits[i], colsum=bs[i] if what == "dgrad" else None) for i in range(L)]
        for b_ in bits:
            b_.random_()
        ms = timed(lambda: ops.stack_bf16(x, 256, layers))
        fl = 2.0 * rows * 256 * 256 * L
        print(f"stack {what} ring={ring} grid={os.environ.get('PAPR_DBG_STACK_GRID', 'all')}: {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s")
    elif what == "wgrad":
        full = [ops.Blocked(rows, 256, dev) for _ in range(2)]
        gw = torch.zeros(256, 256, device=dev)
        ms = timed(lambda: ops.wgrad_bf16(full[0], full[1], gw, 256, 256))
        fl = 2.0 * rows * 256 * 256
        print(f"wgrad grid={os.environ.get('PAPR_DBG_WGRAD_GRID', 'all')}: {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  {rows * 1024.0 / ms / 1e6:7.1f} GB/s")


if __name__ == "__main__":
    main()
