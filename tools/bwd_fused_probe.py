"""Times the backward of one MLP stack both ways -- dgrad stack launch + per-layer weight-gradient launches against the
single fused launch (papr_stack_bwd_fused) -- on the key-stack (6 x 256) and value-stack (142 -> 7 x 256 -> 32) shapes of the
chair configuration, with a sweep over the number of SMs given to the dgrad side."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from papr_b200 import attention as A, ops  # noqa: E402


def timed(fn, n=3):
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
    rows = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 400 * 400 * 20
    torch.manual_seed(0)
    for name, dims in (("key", (117, 256, 256, 256, 256, 256, 256)), ("value", (142, 256, 256, 256, 256, 256, 256, 256, 32))):
        ws = [torch.randn(n, k, device="cuda") * (2.0 / k) ** 0.5 for k, n in zip(dims[:-1], dims[1:])]
        bs = [torch.zeros(n, device="cuda") for n in dims[1:]]
        x = ops.Blocked.from_f32(torch.randn(rows, dims[0], device="cuda"))
        inputs, bits, out = A._stack_forward_fused(x, ws, bs, 0.0, dims[0], True, False)
        dz = ops.Blocked.from_f32(torch.randn(rows, dims[-1], device="cuda") * 0.1)
        fl = sum(4.0 * rows * k * n for k, n in zip(dims[:-1], dims[1:])) - 2.0 * rows * dims[0] * dims[1]
        A.BWD_FUSED = False
        for wb in (False, True, False, True):
            A.WGRAD_BIAS = wb
            ms = timed(lambda: A._stack_backward(dz, inputs, bits, ws, 0.0, None, dims[0], ops.pad_cols(dims[0])))
            print(f"{name}: two-kernel backward, bias gradients in {'wgrad' if wb else 'dgrad'}: {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
        A.BWD_FUSED = True
        for prod in ((0, 72, 80, 88) if "--fused" in sys.argv else ()):
            A.BWD_PROD_CTAS = prod
            ms = timed(lambda: A._stack_backward(dz, inputs, bits, ws, 0.0, None, dims[0], ops.pad_cols(dims[0])))
            print(f"{name}: fused, {prod:3d} dgrad SMs   {ms:7.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
        del inputs, bits, out, x, dz


if __name__ == "__main__":
    main()
