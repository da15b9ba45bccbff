"""fp32-accurate Linear layers on the library's own bf16 tensor-core kernels (the GEMMs of ``precision="fp32"``).

A fp32 number is the exact sum of three bf16 numbers (x = x1 + x2 + x3, 8 significant bits each), and a product of two
bf16 numbers is exact in fp32, so

    x . w = sum_{i+j<=4} x_i . w_j  + O(2^-24 |x||w|)        (six of the nine partial products)

is computed by six launches of ``papr_linear_bf16`` (tcgen05, fp32 accumulation in TMEM) chained through the kernel's
fp32 ``addend`` input; bias and activation ride on the last launch.  The weight gradient is six accumulating launches
of ``papr_wgrad_bf16`` and the data gradient is the same forward with the transposed weight.  This is the parity mode
of the product kernels: the 1e-5 attention / feature assertions and the 1e-3 RGB assertion of the tests exercise
``linear_kernel`` and ``wgrad_kernel`` themselves (reference models/mlp.py:53-58, attn.py:217-218 and their autograd),
not a library matmul.  It is ~6x the tensor work plus fp32 round trips, so it is for tests, not for speed.
"""
import torch
import torch.nn.functional as F

from . import ops

# (i, j): x_i . w_j, smallest partial products first so that they are not absorbed by the large one
_PAIRS = ((2, 0), (1, 1), (0, 2), (1, 0), (0, 1), (0, 0))
_KMAX = 256


def split3(t):
    """t (fp32) -> three fp32 tensors holding bf16-representable values with t == t1 + t2 + t3 exactly."""
    t1 = t.to(torch.bfloat16).float()
    r = t - t1
    t2 = r.to(torch.bfloat16).float()
    t3 = (r - t2).to(torch.bfloat16).float()
    return t1, t2, t3


def _chunks(n):
    return [(s, min(s + _KMAX, n)) for s in range(0, n, _KMAX)]


def split_linear(x, w, bias=None, slope=None):
    """act(x @ w.T + bias) for x (M, n_in) fp32, w (n_out, n_in) fp32, n_out <= 256; fp32-accurate (see module doc).
    slope: None = no activation, 0.0 = relu, > 0 = leaky relu."""
    M, n_in = x.shape
    n_out = w.shape[0]
    if n_out > _KMAX:       # e.g. the data gradient of a skip layer (398 inputs): one pass per 256 output features
        return torch.cat([split_linear(x, w[o0:o1], None if bias is None else bias[o0:o1], slope)
                          for o0, o1 in _chunks(n_out)], dim=1)
    N = (n_out + 31) // 32 * 32
    act = slope is not None
    b = None
    if bias is not None or act:
        b = torch.zeros(N, dtype=torch.float32, device=x.device)
        if bias is not None:
            b[:n_out] = bias.detach().float()
    acc = None
    chunks = _chunks(n_in)
    for ci, (k0, k1) in enumerate(chunks):
        K = (k1 - k0 + 15) // 16 * 16
        xs = [ops.Blocked.from_f32(t) for t in split3(x[:, k0:k1].contiguous())]
        ws = [ops.pack_weight(t, N, K) for t in split3(w[:, k0:k1].float().contiguous())]
        for pi, (i, j) in enumerate(_PAIRS):
            last = ci == len(chunks) - 1 and pi == len(_PAIRS) - 1
            _, acc, _ = ops.linear_bf16(xs[i], ws[j], N, K, bias=b if last else None, act=act and last,
                                        slope=(slope or 0.0) if last else 0.0, out_blocked=False, out_f32=True, addend=acc)
    return acc[:M, :n_out]


def split_wgrad(g, x):
    """g.T @ x for g (M, n_out), x (M, n_in) fp32 -> (n_out, n_in) fp32 through papr_wgrad_bf16."""
    n_out, n_in = g.shape[1], x.shape[1]
    out = torch.zeros((n_out, n_in), dtype=torch.float32, device=g.device)
    for a0, a1 in _chunks(n_out):
        ga = g[:, a0:a1].contiguous()
        gs = [ops.Blocked.from_f32(t, cols_pad=max(ops.pad_cols(a1 - a0), 128 * ((a1 - a0 + 127) // 128))) for t in split3(ga)]
        for b0, b1 in _chunks(n_in):
            xs = [ops.Blocked.from_f32(t) for t in split3(x[:, b0:b1].contiguous())]
            for i, j in _PAIRS:
                ops.wgrad_bf16(gs[i], xs[j], out[a0:a1, b0:b1], a1 - a0, b1 - b0)
    return out


class SplitLinearFn(torch.autograd.Function):
    """y = act(x W^T + b), forward and backward on the bf16 tensor-core kernels at fp32 accuracy."""

    @staticmethod
    def forward(ctx, x, weight, bias, slope):
        x = x.contiguous().float()
        y = split_linear(x, weight.detach(), bias, slope)
        ctx.slope = slope
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight.detach(), y if slope is not None else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        g = gy
        if ctx.slope is not None:       # relu'(0) = 0 as torch's threshold_backward
            g = torch.where(y > 0, gy, gy * ctx.slope)
        g = g.contiguous().float()
        gx = split_linear(g, weight.t().contiguous()) if ctx.needs_input_grad[0] else None
        gw = split_wgrad(g, x) if ctx.needs_input_grad[1] else None
        gb = g.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None


def linear(x, lin, slope=None):
    """nn.Linear (+ relu / leaky relu) on 2-D fp32 input through SplitLinearFn."""
    return SplitLinearFn.apply(x, lin.weight, lin.bias, slope)


def matmul_t(x, w):
    """x @ w.T without bias (w given as (n_out, n_in))."""
    return SplitLinearFn.apply(x, w, None, None)


def mlp_forward(mlp, x):
    """papr_b200.nn.MLP.forward (reference models/mlp.py:47-59) with every Linear on the split tensor-core path."""
    from .nn import activation_slope
    inp = x
    for i, lin in enumerate(mlp.linears()):
        if i in mlp.skip_layers:
            x = torch.cat([x, inp], dim=-1)
        name = mlp.last_act_type if i == mlp.num_layers - 1 else mlp.act_type
        try:
            slope = activation_slope(name)
            x = linear(x, lin, slope)
        except NotImplementedError:      # an activation the kernel epilogue does not have: apply it in torch
            x = mlp.model[2 * i + 2](linear(x, lin, None))
    return x
