"""The SmallUNet decode of the fp32 PARITY mode on the library's own tensor-core kernels (reference models/unet.py:196-258).

Same idea as papr_b200/split_gemm.py: a fp32 number is the exact sum of three bf16 numbers, so a convolution at fp32
accuracy is six launches of ``papr_conv_bf16`` (bf16 operands, fp32 accumulation in TMEM) chained through the kernel's fp32
addend, over three split copies of the activation planes and of the weight images; the weight gradient is six accumulating
launches of ``papr_conv_wgrad_bf16``.  The 3x3 convolutions go through ``SplitConv3x3Fn`` below; the 1x1 head and the GEMM
half of the transposed convolutions are per-pixel Linear layers and reuse ``split_gemm.SplitLinearFn``; pooling, ReLU,
concatenation and the pixel shuffle are fp32 torch elementwise ops.  ~6x the tensor work of the bf16 path: for tests.
"""
import torch
import torch.nn.functional as F

from . import ops, split_gemm
from . import unet as U

_PAIRS = split_gemm._PAIRS


def _split_planes(x_hwc, copies):
    """(H, W, C) fp32 -> three Planes holding the bf16 split of x."""
    H, W, C = x_hwc.shape
    out = []
    for t in split_gemm.split3(x_hwc.contiguous()):
        t = t.contiguous()
        p = U.Planes(H, W, max(64, (C + 63) // 64 * 64), copies, x_hwc.device)
        ops.call("papr_unet_pack_input", t.data_ptr(), t.stride(1), C, None, None, p.ptr(), U._ref(p.raster()), copies, p.cbs,
                 nbytes=H * W * C * 10.0)
        out.append(p)
    return out


def _split_images(mat):
    """fp32 (rows, K) -> three lists of weight-image tiles + what must stay alive until the pack kernel has run."""
    descs, keep, tiles3 = [], [], []
    for t in split_gemm.split3(mat.contiguous()):
        tiles, d, keepalive = U._pack_matrix(t)
        tiles3.append(tiles)
        descs.extend(d)
        keep.append(keepalive)
    keep.append(U._launch_pack(descs, mat.device))
    return tiles3, keep


def _conv_rows(planes3, tiles3, cbs, sign, n_out, bias, relu):
    """sum over the six split pairs of conv(planes_i, weights_j) -> fp32 (H, W, n_out)."""
    ref = planes3[0]
    Npad = sum(N for _, N, _ in tiles3[0])
    acc = None
    for pi, (i, j) in enumerate(_PAIRS):
        last = pi == len(_PAIRS) - 1
        out = torch.empty((ref.L, Npad), dtype=torch.float32, device=ref.buf.device)
        U._conv(planes3[i], planes3[i].ptr(), cbs, 9, sign, tiles3[j], bias if last else None, relu and last, None, out_f32=out,
                addend=acc)
        acc = out
    return acc[: (ref.H + 2) * ref.Wp].view(ref.H + 2, ref.Wp, Npad)[1:ref.H + 1, 1:ref.W + 1, :n_out]


class SplitConv3x3Fn(torch.autograd.Function):
    """relu?(conv2d(x, w, b, padding=1)) for x (B, C, H, W) fp32 at fp32 accuracy on papr_conv_bf16 / papr_conv_wgrad_bf16."""

    @staticmethod
    def forward(ctx, x, w, b, relu):
        Bn, C, H, W = x.shape
        co, ci = w.shape[:2]
        cip = max(64, (ci + 63) // 64 * 64)
        wf = U._pad_k(w.detach().float().permute(0, 2, 3, 1), cip).reshape(co, 9 * cip)
        tiles3, keep = _split_images(wf)
        xh = x.detach().float().permute(0, 2, 3, 1).contiguous()
        ys, saved = [], []
        for n in range(Bn):
            planes3 = _split_planes(xh[n], 3)
            ys.append(_conv_rows(planes3, tiles3, cip // 64, 1, co, b.detach().float() if b is not None else None, relu))
            saved.append(planes3)
        y = torch.stack(ys).permute(0, 3, 1, 2)
        ctx.relu, ctx.has_bias, ctx.planes, ctx.keep = relu, b is not None, saved, keep
        ctx.save_for_backward(w.detach(), y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        w, y = ctx.saved_tensors
        co, ci = w.shape[:2]
        cop = max(64, (co + 63) // 64 * 64)
        g = gy.float()
        if ctx.relu:
            g = torch.where(y > 0, g, torch.zeros_like(g))
        gh = g.permute(0, 2, 3, 1).contiguous()
        Bn, H, W, _ = gh.shape
        gx = gw = None
        if ctx.needs_input_grad[0]:
            wd = U._pad_k(w.float().permute(1, 2, 3, 0), cop).reshape(ci, 9 * cop)
            tiles3, keep = _split_images(wd)
        if ctx.needs_input_grad[1]:
            gw9 = torch.zeros((9, co, ci), dtype=torch.float32, device=g.device)
        gxs = []
        for n in range(Bn):
            dz3 = _split_planes(gh[n], 3)
            if ctx.needs_input_grad[0]:
                gxs.append(_conv_rows(dz3, tiles3, cop // 64, -1, ci, None, False))
            if ctx.needs_input_grad[1]:
                for i, j in _PAIRS:
                    for a0 in range(0, co, 256):
                        for b0 in range(0, ci, 256):
                            U._wgrad(dz3[i], a0 // 64, min(256, co - a0), ctx.planes[n][j], b0 // 64, min(256, ci - b0), 9,
                                     gw9[:, a0:, b0:], gw9.stride(0))
        if ctx.needs_input_grad[0]:
            gx = torch.stack(gxs).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            gw = gw9.permute(1, 2, 0).reshape(co, ci, 3, 3)
        gb = g.sum((0, 2, 3)) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        ctx.planes = None
        return gx, gw, gb, None


def _conv3(x, conv, relu=True):
    return SplitConv3x3Fn.apply(x, conv.weight, conv.bias, relu)


def _pixelwise_linear(x, weight, bias):
    """A 1x1 convolution = a Linear layer over pixels: (B, C, H, W) x (Cout, C) -> (B, Cout, H, W), split tensor-core path."""
    Bn, C, H, W = x.shape
    rows = x.permute(0, 2, 3, 1).reshape(-1, C)
    y = split_gemm.SplitLinearFn.apply(rows, weight, bias, None)
    return y.reshape(Bn, H, W, -1).permute(0, 3, 1, 2)


def _conv_transpose2x2(x, up):
    """ConvTranspose2d(kernel 2, stride 2) (unet.py:60-76): per-pixel Linear to (a, b, co) channels + pixel shuffle."""
    ci, co = up.weight.shape[:2]
    wg = up.weight.permute(2, 3, 1, 0).reshape(4 * co, ci)                  # rows (a, b, co)
    Bn, _, H, W = x.shape
    y = _pixelwise_linear(x, wg, None)                                     # (B, 4*co, H, W)
    y = y.reshape(Bn, 2, 2, co, H, W).permute(0, 3, 4, 1, 5, 2).reshape(Bn, co, 2 * H, 2 * W)
    return y + up.bias.reshape(1, co, 1, 1)


def unet_forward_fp32(m, x, gamma=None, beta=None):
    """papr_b200.renderer.SmallUNet.forward in the parity mode (FiLM at any stage, unet.py:208-258)."""
    def film(t, stage):
        return m._film(t, gamma, beta) if m.affine_layer == stage else t

    def up(block, t, skip):
        t = _conv_transpose2x2(t, block.up)
        dy, dx = skip.shape[2] - t.shape[2], skip.shape[3] - t.shape[3]
        if dy or dx:
            t = F.pad(t, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return _conv3(torch.cat([skip, t], dim=1), block.conv.double_conv[0])

    x = film(x.float(), 0)
    x1 = film(_conv3(x, m.inc.double_conv[0]), 1)
    x2 = film(_conv3(F.max_pool2d(x1, 2), m.down1.maxpool_conv[1].double_conv[0]), 2)
    x3 = film(_conv3(F.max_pool2d(x2, 2), m.down2.maxpool_conv[1].double_conv[0]), 3)
    y = film(up(m.up1, x3, x2), 4)
    y = film(up(m.up2, y, x1), 5)
    out = _pixelwise_linear(y, m.outc.conv.weight.reshape(m.outc.conv.weight.shape[0], -1), m.outc.conv.bias)
    return m.last_act(out)
