"""Proximity attention (reference models/attn.py ProximityAttention + the K/Q/V assembly and blend of
models/model.py:396-437, 519-534) on the library's CUDA kernels.

Data flow for one batch of rays (R rays, K candidates each, M = R*K rows):

  select_topk (model.PAPR)            idx (R,K)
  per ray, PyTorch + tcgen05 GEMMs    query stack -> q' ; ua = (q' W_k / sqrt d) * a2_k ; c' = ...      (5% of the work)
  papr_attn_prologue_fwd              kin (M,128) , vin (M,192)    tile-blocked bf16
  papr_linear_bf16 x5 / x8            key stack -> h5 ; value stack -> v (fp32)                          (tcgen05)
  papr_score_blend_fwd                fused (R,C), attn (R,K+1)

The backward of the per-row part is one torch.autograd.Function (RowAttentionFn) that calls the matching backward
kernels (papr_blend_bwd, papr_key_score_bwd, papr_linear_bf16 as dgrad, papr_wgrad_bf16, papr_attn_prologue_bwd).
The key stack's output LayerNorm, w_k and the dot with the query are folded algebraically:
  score = q'.(W_k LN(h5) + b_k)/sqrt d = ua . z(h5) + c'   with z the normalised h5,
which removes one 256x256 GEMM per row; its parameters therefore get their gradients through ua and c' (autograd).

precision="fp32" is the parity mode: the same CUDA-core kernels (fp32 taps instead of bf16 tiles) and every GEMM on the
library's own tensor-core kernels at fp32 accuracy through a three-way bf16 split (papr_b200/split_gemm.py), so the
1e-5 assertions of the tests exercise linear_kernel / wgrad_kernel themselves.  The bf16 path is the product.
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, split_gemm
from .nn import AttentionLayer, Embeddings, activation_slope

#: GEMM engine of precision="fp32": "split" = the library's tensor-core kernels (three-way bf16 split); "torch" = torch
#: fp32 matmuls, kept only to A/B the split path in tests/test_tc_gpu.py
FP32_GEMM = "split"


def posenc(x, L, factor=2.0, mult=1.0):
    """models/utils.py:232-242 (embed_type 1), used for the per-ray query features."""
    parts = [x]
    for i in range(L):
        a = (factor ** i) * x * mult
        parts += [torch.sin(a), torch.cos(a)]
    return torch.flatten(torch.stack(parts, -1), start_dim=-2)


def _ptr(t):
    return t.data_ptr() if t is not None else None


class _LinearBf16Fn(torch.autograd.Function):
    """y = act(x W^T + b) for per-ray (fp32 in / fp32 out) tensors on the tcgen05 kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, slope):
        n_out, n_in = weight.shape
        K = (n_in + 15) // 16 * 16
        N = (n_out + 31) // 32 * 32
        xb = ops.Blocked.from_f32(x)
        img = ops.pack_weight(weight, N, K)
        b = bias if n_out == N else F.pad(bias, (0, N - n_out))
        _, y, _ = ops.linear_bf16(xb, img, N, K, bias=b.detach().contiguous(), act=slope is not None,
                                  slope=slope or 0.0, out_blocked=False, out_f32=True)
        y = y[: x.shape[0], :n_out]
        ctx.save_for_backward(weight, y)
        ctx.xb, ctx.slope, ctx.dims = xb, slope, (N, K, n_out, n_in)
        return y

    @staticmethod
    def backward(ctx, gy):
        weight, y = ctx.saved_tensors
        N, K, n_out, n_in = ctx.dims
        g = gy
        if ctx.slope is not None:
            g = torch.where(y > 0, gy, gy * ctx.slope)
        g = g.contiguous()
        gb = ops.Blocked.from_f32(g, cols_pad=max(ops.pad_cols(n_out), 128))
        Kd = (n_out + 15) // 16 * 16
        Nd = ops.pad_cols(n_in)
        img_t = ops.pack_weight(weight, Nd, Kd, transpose=True)
        gb_lin = gb if gb.cols_pad == ops.pad_cols(Kd) else ops.Blocked.from_f32(g)
        _, gx, _ = ops.linear_bf16(gb_lin, img_t, Nd, Kd, out_blocked=False, out_f32=True)
        gw = torch.zeros(weight.shape, dtype=torch.float32, device=weight.device)
        ops.wgrad_bf16(gb, ctx.xb, gw, n_out, n_in)
        return gx[: g.shape[0], :n_in], gw, g.sum(0), None


def _linear(x, lin, slope, precision):
    if precision == "fp32":
        if FP32_GEMM == "split":
            return split_gemm.linear(x, lin, slope)
        y = F.linear(x, lin.weight, lin.bias)
        if slope is not None:
            y = F.leaky_relu(y, slope) if slope else F.relu(y)
        return y
    return _LinearBf16Fn.apply(x, lin.weight, lin.bias, slope)


def _mlp_fp32(mlp, x):
    """An MLP stack in the fp32 parity mode."""
    return split_gemm.mlp_forward(mlp, x) if FP32_GEMM == "split" else mlp(x)


class _Shape:
    """Static description of one call (sizes + flags) shared by forward and backward."""

    def __init__(self, attn, R, rays_per_view, K):
        self.R, self.rays_per_view, self.K = R, rays_per_view, K
        self.M = R * K
        self.L, self.F, self.C = attn.L, attn.F, attn.C
        self.dk, self.dv = attn.dk, attn.dv
        self.dk_pad, self.dv_pad = ops.pad_cols(attn.dk), ops.pad_cols(attn.dv)
        self.eps = attn.eps
        self.k_slope, self.v_slope = attn.k_slope, attn.v_slope
        self.score_relu, self.normalize, self.bkg_score = attn.score_relu, attn.normalize, attn.bkg_score
        self.k_skip = tuple(attn.embed.embed_k.mlp.skip_layers)
        self.v_skip = tuple(attn.embed.embed_v.mlp.skip_layers)
        # grad mode of the CALLER: inside autograd.Function.forward it always reads False, and ctx.needs_input_grad
        # mirrors requires_grad regardless of torch.no_grad(), so this is what decides whether a backward stash is kept
        self.grad = torch.is_grad_enabled()
        self.k_images = attn.weight_images.stack("k", len(attn.embed.embed_k.mlp.linears()))
        self.v_images = attn.weight_images.stack("v", len(attn.embed.embed_v.mlp.linears()))


def _prologue_fwd(sh, rays_o, rays_d, points, feats, idx, ln_a, ln_b, taps=False):
    dev = rays_d.device
    kin = ops.Blocked(sh.M, sh.dk, dev)
    vin = ops.Blocked(sh.M, sh.dv, dev)
    kin32 = torch.empty((sh.M, sh.dk), device=dev) if taps else None
    vin32 = torch.empty((sh.M, sh.dv), device=dev) if taps else None
    ops.call("papr_attn_prologue_fwd",
        rays_o.data_ptr(), rays_d.data_ptr(), points.data_ptr(), _ptr(feats), idx.data_ptr(), ln_a.data_ptr(),
        ln_b.data_ptr(), sh.R, sh.rays_per_view, sh.K, sh.L, sh.F, sh.eps, kin.data_ptr(), sh.dk_pad, vin.data_ptr(),
        sh.dv_pad, _ptr(kin32), _ptr(vin32),
        nbytes=sh.M * (2.0 * (sh.dk_pad + sh.dv_pad) + 12 + 4.0 * sh.F))
    return kin, vin, kin32, vin32


def _prologue_bwd(sh, rays_o, rays_d, points, idx, ln_a, dkin, dvin, dkin32, dvin32, n_points):
    dev = rays_d.device
    g_points = torch.zeros((n_points, 3), device=dev)
    g_feats = torch.zeros((n_points, sh.F), device=dev) if sh.F else None
    g_a = torch.zeros(sh.dk, device=dev)
    g_b = torch.zeros(sh.dk, device=dev)
    ops.call("papr_attn_prologue_bwd",
        rays_o.data_ptr(), rays_d.data_ptr(), points.data_ptr(), idx.data_ptr(), ln_a.data_ptr(), sh.R,
        sh.rays_per_view, sh.K, sh.L, sh.F, sh.eps, _ptr(dkin), sh.dk_pad, _ptr(dvin), sh.dv_pad, _ptr(dkin32),
        _ptr(dvin32), g_points.data_ptr(), _ptr(g_feats), g_a.data_ptr(), g_b.data_ptr(),
        nbytes=sh.M * (2.0 * (sh.dk_pad + sh.dv_pad) + 24 + 8.0 * sh.F))
    return g_points, g_feats, g_a, g_b


def _score_blend_fwd(sh, h5, h5_32, ua, cprime, influ, idx, v):
    dev = ua.device
    fused = torch.empty((sh.R, sh.C), device=dev)
    attn = torch.empty((sh.R, sh.K + 1), device=dev)
    sc = torch.empty((sh.M,), device=dev)
    stats = torch.empty((sh.M, 4), device=dev)      # per row: LayerNorm mean, 1/(std+eps), ua . z, unused
    ops.call("papr_score_blend_fwd",
        _ptr(h5), _ptr(h5_32), ua.data_ptr(), cprime.data_ptr(), influ.data_ptr(), idx.data_ptr(), v.data_ptr(),
        v.stride(0), sh.R, sh.K, sh.C, int(sh.score_relu), int(sh.normalize), sh.bkg_score, sh.eps, fused.data_ptr(),
        attn.data_ptr(), sc.data_ptr(), stats.data_ptr(),
        nbytes=sh.M * (512.0 + 4 * sh.C + 16) + sh.R * (1024.0 + 4 * sh.C),
        kernels=2 if (h5 is not None and h5_32 is None and sh.K >= 16) else 1)     # row scores, then per-ray softmax + blend
    return fused, attn, sc, stats


def _blend_bwd(sh, d_fused, d_attn, attn, sc, influ, idx, v, n_points):
    dev = d_fused.device
    dv = ops.Blocked(sh.M, sh.C, dev)
    d_score = torch.empty((sh.M,), device=dev)
    g_influ = torch.zeros((n_points,), device=dev)
    g_bv = torch.zeros((sh.C,), device=dev)
    ops.call("papr_blend_bwd",
        d_fused.data_ptr(), _ptr(d_attn), attn.data_ptr(), sc.data_ptr(), influ.data_ptr(), idx.data_ptr(),
        v.data_ptr(), v.stride(0), sh.R, sh.K, sh.C, int(sh.score_relu), int(sh.normalize), dv.data_ptr(),
        d_score.data_ptr(), g_influ.data_ptr(), g_bv.data_ptr(),
        nbytes=sh.M * (4.0 * sh.C + 128 + 16) + sh.R * 8.0 * sh.C)
    return dv, d_score, g_influ, g_bv


def _key_score_bwd(sh, d_score, h5, h5_32, stats, ua, tap=False, want_bias=True):
    """want_bias=False: g_b5 is left to the weight-gradient kernel of the last key layer (column sums of the dh5 tiles it
    streams anyway), which lets the bf16 call use the block-major kernel."""
    dev = ua.device
    dh5 = ops.Blocked(sh.M, 256, dev)
    dh5_32 = torch.empty((sh.M, 256), device=dev) if tap else None
    zsum = torch.empty((sh.R, 256), device=dev)
    dssum = torch.empty((sh.R,), device=dev)
    g_b5 = torch.zeros((256,), device=dev) if want_bias else None
    ops.call("papr_key_score_bwd",
        d_score.data_ptr(), _ptr(h5), _ptr(h5_32), stats.data_ptr(), ua.data_ptr(), sh.R, sh.K, sh.eps,
        dh5.data_ptr(), _ptr(dh5_32), zsum.data_ptr(), dssum.data_ptr(), _ptr(g_b5),
        nbytes=sh.M * (1024.0 + 12) + sh.R * 2048.0)
    return dh5, dh5_32, zsum, dssum, g_b5


def _stash_pack(seq):
    """Flatten a list of Blocked / tensor / None into (tensors for ctx.save_for_backward, python spec).  The big buffers
    have to go through save_for_backward: that is what torch.utils.checkpoint's saved-tensor hooks intercept, so a
    checkpointed ray chunk really drops its stash after the forward and rebuilds it during backward."""
    tensors, spec = [], []
    for it in seq:
        if it is None:
            spec.append(None)
        elif isinstance(it, ops.Blocked):
            spec.append((it.rows, it.cols, it.cols_pad))
            tensors.append(it.buf)
        else:
            spec.append("t")
            tensors.append(it)
    return tensors, spec


def _stash_unpack(tensors, spec):
    out, it = [], iter(tensors)
    for sp in spec:
        if sp is None:
            out.append(None)
        elif sp == "t":
            out.append(next(it))
        else:
            out.append(ops.Blocked.wrap(next(it), *sp))
    return out


class WeightImages:
    """bf16 weight images of the key / query / value stacks -- every layer, forward and transposed (dgrad), with the
    replicas the fused stack kernel wants -- kept in persistent buffers and rebuilt by ONE papr_pack_weight_batch launch
    per forward call instead of one papr_pack_weight launch per layer and direction (38 per training step)."""

    def __init__(self, attn):
        self.attn = attn
        self.images = {}          # (stack, layer, "f"|"t") -> uint8 (replicas, bytes)
        self._ptrs = None
        self._table = None
        self._n = 0

    def _specs(self):
        out = []
        for name in ("k", "q", "v"):
            mlp = getattr(self.attn.embed, f"embed_{name}").mlp
            lins = mlp.linears()
            if mlp.skip_layers or len(lins) > 8 or any(l.weight.shape[0] != 256 for l in lins[:-1]):
                continue          # those stacks run layer by layer and pack on the fly
            in_pad = ops.pad_cols(lins[0].weight.shape[1])
            for i, lin in enumerate(lins):
                n_out, n_in = lin.weight.shape
                out.append(((name, i, "f"), lin.weight, 0, (n_out + 31) // 32 * 32, (n_in + 15) // 16 * 16))
                out.append(((name, i, "t"), lin.weight, 1, n_in if i > 0 else in_pad, (n_out + 15) // 16 * 16))
        return out

    def refresh(self):
        import ctypes
        from ._lib import PackDesc
        specs = self._specs()
        if not specs:
            self.images = {}
            return
        ptrs = [w.data_ptr() for _, w, _, _, _ in specs]
        if ptrs != self._ptrs:
            dev = specs[0][1].device
            arr = (PackDesc * len(specs))()
            self.images = {}
            for j, (key, w, tr, N, K) in enumerate(specs):
                nbytes = (K + 63) // 64 * N * 128
                img = torch.empty((ops.WEIGHT_REPLICAS, nbytes), dtype=torch.uint8, device=dev)
                self.images[key] = img
                arr[j].w, arr[j].ld, arr[j].rows, arr[j].cols = w.data_ptr(), w.stride(0), w.shape[0], w.shape[1]
                arr[j].transpose, arr[j].N, arr[j].K, arr[j].replicas = tr, N, K, ops.WEIGHT_REPLICAS
                arr[j].scale, arr[j].rep_stride, arr[j].image = 1.0, img.stride(0), img.data_ptr()
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self._table = raw.to(dev)
            self._ptrs, self._n = ptrs, len(specs)
        ops.call("papr_pack_weight_batch", self._table.data_ptr(), self._n,
                 nbytes=sum(6.0 * w.numel() for _, w, _, _, _ in specs))

    def stack(self, name, n_layers):
        """(forward images, transposed images) of a stack, or (None, None) if it is not cached."""
        if (name, 0, "f") not in self.images:
            return None, None
        return ([self.images[(name, i, "f")] for i in range(n_layers)], [self.images[(name, i, "t")] for i in range(n_layers)])


def _pad_bias(b, N):
    b = b.detach().float().contiguous()
    return b if b.numel() == N else F.pad(b, (0, N - b.numel()))


FUSE_STACKS = True      # one cta_group::2 launch per MLP stack (papr_stack_bf16) instead of one launch per layer
# backward schedule of a stack (see _stack_backward); the environment overrides are tuning knobs for tools/ only
BWD_SLICE_ROWS = int(os.environ.get("PAPR_BWD_SLICE_ROWS", 8 << 20))    # rows per slice (bounds the dZ stash)
# dgrad of slice s+1 || weight gradients of slice s on two streams.  OFF: measured slower on B200 (113.8 vs 110.5 ms per step;
# the step is power-capped, so running the HBM-bound and the tensor-bound kernel together only lowers the clocks)
BWD_OVERLAP = os.environ.get("PAPR_BWD_OVERLAP", "0") != "0"
# the whole backward of a stack in one launch, dZ handed over through L2 (papr_stack_bwd_fused); ReLU / linear stacks only
BWD_FUSED = os.environ.get("PAPR_BWD_FUSED", "0") != "0"
# bias gradients of the hidden layers from the weight-gradient kernel (idle warps sum the dZ tiles it streams anyway)
# instead of the dgrad epilogue, where the column sums cost 12% of the kernel
WGRAD_BIAS = os.environ.get("PAPR_WGRAD_BIAS", "1") != "0"
BWD_PROD_CTAS = int(os.environ.get("PAPR_BWD_PROD_CTAS", 0))     # SMs on the dgrad side of the fused launch (0 = library default)
BWD_DGRAD_CTAS = int(os.environ.get("PAPR_BWD_DGRAD_CTAS", 88))  # SMs given to the dgrad stack kernel while both run
BWD_WGRAD_CTAS = int(os.environ.get("PAPR_BWD_WGRAD_CTAS", 60))  # ... and to the weight-gradient kernel
_SIDE_STREAMS = {}


def _side_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def _stack_forward_fused(x, weights, biases, slope, n_in0, save, last_f32, images=None):
    n_layers = len(weights)
    dev = weights[0].device
    K0 = (n_in0 + 15) // 16 * 16
    layers, inputs, bits_list = [], [x], []
    out = None
    for i, (w, b) in enumerate(zip(weights, biases)):
        n_out, n_in = w.shape
        K = (n_in + 15) // 16 * 16
        N = (n_out + 31) // 32 * 32
        last = i == n_layers - 1
        act = (not last) and slope is not None
        img = images[i] if images is not None else ops.pack_weight(w, N, K, replicas=ops.WEIGHT_REPLICAS)
        spec = dict(w_image=img, N=N, bias=_pad_bias(b, N), act=act)
        bits = None
        if save and act:
            bits = torch.empty((x.rows_pad, ops.pad_cols(N) // 64), dtype=torch.int64, device=dev)
            spec["sign_bits_out"] = bits
        bits_list.append(bits)
        if last and last_f32:
            out = torch.empty((x.rows_pad, N), dtype=torch.float32, device=dev)
            spec["out_f32"] = out
        elif last or save:                       # hidden outputs only leave the SM when backward needs them
            ob = ops.Blocked(x.rows, N, dev)
            spec["out_blocked"] = ob
            if last:
                out = ob
            else:
                inputs.append(ob)
        layers.append(spec)
    ops.stack_bf16(x, K0, layers, slope=slope or 0.0)
    return inputs, bits_list, out


def _stack_forward(x, weights, biases, slope, n_in0, save, last_f32=False, skip_layers=(), images=None):
    """Run one MLP stack on the tensor cores.  Returns (layer inputs, sign bits, last output: Blocked or fp32).
    A skip layer (mlp.py:54-55: input = cat[h, stack input]) is two GEMMs into one accumulator: the stack-input half
    is computed first in fp32 and handed to the main launch as its `addend`."""
    if FUSE_STACKS and not skip_layers and len(weights) <= 8 and all(w.shape[0] == 256 for w in weights[:-1]):
        return _stack_forward_fused(x, weights, biases, slope, n_in0, save, last_f32, images)
    inputs, bits_list = [], []
    h = x
    n_layers = len(weights)
    out = None
    K0 = (n_in0 + 15) // 16 * 16
    for i, (w, b) in enumerate(zip(weights, biases)):
        n_out = w.shape[0]
        n_in = w.shape[1] - (n_in0 if i in skip_layers else 0)
        K = (n_in + 15) // 16 * 16
        N = (n_out + 31) // 32 * 32
        last = i == n_layers - 1
        addend = None
        if i in skip_layers:
            _, addend, _ = ops.linear_bf16(x, ops.pack_weight(w[:, n_in:], N, K0), N, K0, out_blocked=False, out_f32=True)
        img = ops.pack_weight(w[:, :n_in], N, K)
        inputs.append(h)
        f32 = last and last_f32
        yb, yf, bits = ops.linear_bf16(h, img, N, K, bias=_pad_bias(b, N), act=(not last) and slope is not None,
                                       slope=slope or 0.0, out_blocked=not f32, out_f32=f32,
                                       sign_bits_out=save and not last and slope is not None, addend=addend)
        bits_list.append(bits)
        h = yb
        out = yf if f32 else yb
    return inputs, bits_list, out


def _stack_backward(dz, inputs, bits_list, weights, slope, g_bias_last, in_valid, in_pad, skip_layers=(), images_t=None,
                    last_bias_from_wgrad=False):
    """Backward of _stack_forward.  dz: Blocked gradient of the last layer's output.  Returns (d_input Blocked,
    [gW], [gb])."""
    n_layers = len(weights)
    gWs, gbs = [None] * n_layers, [None] * n_layers
    gbs[-1] = g_bias_last
    last_bias_here = last_bias_from_wgrad and g_bias_last is None     # the caller left the last layer's bias gradient to the
    if last_bias_here:                                                # weight-gradient kernel (column sums of dz)
        assert weights[-1].shape[0] >= 128, "a narrow last layer runs the weight gradient with swapped operands"
        gbs[-1] = torch.zeros((weights[-1].shape[0],), device=weights[0].device)
    if FUSE_STACKS and not skip_layers and n_layers <= 8 and all(w.shape[0] == 256 for w in weights[:-1]):
        # dgrad of the whole stack in one launch per slice of rows; the per-layer dZ tiles it stashes feed the weight-gradient
        # launches of that slice.  The two are software-pipelined over the slices on two streams: the dgrad of slice s+1
        # (bound by the HBM WRITE rate of its dZ stash, ~3.9 TB/s) runs on BWD_DGRAD_CTAS SMs while the weight gradients of
        # slice s (bound by HBM reads) run on the other BWD_WGRAD_CTAS SMs, so the tensor pipe and both HBM directions are
        # busy at once.  Two sets of dZ buffers alternate, so at most 2 x BWD_SLICE_ROWS rows of dZ stash (512 B per row
        # and layer) are alive next to the forward stash.
        dev = weights[0].device
        rows_pad = dz.rows_pad
        # one launch for everything, dZ handed over through L2: no dZ stash, so no slicing either
        fused_bwd = BWD_FUSED and n_layers >= 2 and not slope and all(w.shape[1] <= 256 for w in weights) and not last_bias_here
        n_slices = 1 if fused_bwd else max(1, -(-rows_pad // BWD_SLICE_ROWS))
        per = -(-(rows_pad // 128) // n_slices) * 128
        n_slices = -(-rows_pad // per)
        overlap = BWD_OVERLAP and n_slices > 1 and not fused_bwd
        Nd0 = in_pad
        d_in = ops.Blocked(dz.rows, Nd0, dev)
        for i in range(1, n_layers):
            gbs[i - 1] = torch.zeros((weights[i - 1].shape[0],), device=dev)
        for i in range(n_layers):
            gWs[i] = torch.zeros(weights[i].shape, dtype=torch.float32, device=dev)
        imgs = [images_t[i] if images_t is not None else
                ops.pack_weight(weights[i], weights[i].shape[1] if i > 0 else in_pad, (weights[i].shape[0] + 15) // 16 * 16,
                                transpose=True, replicas=ops.WEIGHT_REPLICAS) for i in range(n_layers)]
        K0 = (weights[-1].shape[0] + 15) // 16 * 16
        main = torch.cuda.current_stream(dev)
        side = _side_stream(dev) if overlap else main
        sets = [[None if fused_bwd else ops.Blocked(per, weights[i].shape[1], dev) for i in range(1, n_layers)]
                for _ in range(2 if overlap else 1)]
        consumed = [None] * len(sets)        # event: the weight gradients have read this set of dZ buffers
        for si, r0 in enumerate(range(0, rows_pad, per)):
            r1 = min(r0 + per, rows_pad)
            bufs = sets[si % len(sets)]
            if consumed[si % len(sets)] is not None:
                main.wait_event(consumed[si % len(sets)])
            layers, dzs = [], [dz.rows_view(r0, r1)]
            for i in range(n_layers - 1, -1, -1):
                Nd = weights[i].shape[1] if i > 0 else in_pad
                ob = (None if fused_bwd else bufs[i - 1].rows_view(0, r1 - r0)) if i > 0 else d_in.rows_view(r0, r1)
                spec = dict(w_image=imgs[i], N=Nd, out_blocked=ob)
                if i > 0:
                    if fused_bwd or not WGRAD_BIAS:
                        spec["colsum"] = gbs[i - 1]       # bias gradient in the dgrad epilogue
                    if slope is not None:
                        spec["sign_bits_in"] = bits_list[i - 1][r0:r1]
                layers.append(spec)
                dzs.append(ob)
            if fused_bwd:
                # one launch: dgrad CTAs hand every dZ tile to weight-gradient CTAs through L2 (csrc/stack_bwd.cu)
                ops.stack_bwd_fused(dzs[0], K0, layers, [dict(x=inputs[i].rows_view(r0, r1), gw=gWs[i], n_out=weights[i].shape[0],
                                                               n_in=weights[i].shape[1]) for i in range(n_layers)],
                                    producer_ctas=BWD_PROD_CTAS)
                continue
            ops.stack_bf16(dzs[0], K0, layers, slope=slope or 0.0, max_ctas=BWD_DGRAD_CTAS if overlap else 0)
            if overlap:
                produced = torch.cuda.Event()
                produced.record(main)
                side.wait_event(produced)
            with torch.cuda.stream(side):
                for i in range(n_layers - 1, -1, -1):
                    n_out, n_in = weights[i].shape
                    dzi, xi = dzs[n_layers - 1 - i], inputs[i].rows_view(r0, r1)
                    cap = BWD_WGRAD_CTAS if overlap else 0
                    if n_out < 128:
                        ops.wgrad_bf16(xi, dzi, gWs[i], n_in, n_out, transpose_out=True, max_ctas=cap)
                    else:
                        # ... and the bias gradient of a hidden layer: the column sums of the dZ tiles streamed here
                        ops.wgrad_bf16(dzi, xi, gWs[i], n_out, n_in, max_ctas=cap,
                                       a_colsum=gbs[i] if ((WGRAD_BIAS and i < n_layers - 1) or (last_bias_here and i == n_layers - 1)) else None)
                if overlap:
                    consumed[si % len(sets)] = torch.cuda.Event()
                    consumed[si % len(sets)].record(side)
            del layers, dzs
        if overlap:
            main.wait_stream(side)
        return d_in, gWs, gbs
    d_in_extra = None          # fp32 gradient reaching the stack input through skip connections
    for i in range(n_layers - 1, -1, -1):
        w = weights[i]
        n_out = w.shape[0]
        n_in = w.shape[1] - (in_valid if i in skip_layers else 0)
        gW = torch.zeros(w.shape, dtype=torch.float32, device=w.device)
        x = inputs[i]
        if n_out < 128:     # narrow output (value head): swap operands so that M = n_in
            ops.wgrad_bf16(x, dz, gW, n_in, n_out, transpose_out=True)
        else:
            ops.wgrad_bf16(dz, x, gW, n_out, n_in, a_colsum=gbs[i] if (last_bias_here and i == n_layers - 1) else None)
        Kd = (n_out + 15) // 16 * 16
        if i in skip_layers:
            ops.wgrad_bf16(dz, inputs[0], gW[:, n_in:], n_out, in_valid)
            _, extra, _ = ops.linear_bf16(dz, ops.pack_weight(w[:, n_in:], in_pad, Kd, transpose=True), in_pad, Kd,
                                          out_blocked=False, out_f32=True)
            d_in_extra = extra if d_in_extra is None else d_in_extra + extra
        gWs[i] = gW
        if i > 0:
            img_t = ops.pack_weight(w[:, :n_in], n_in, Kd, transpose=True)
            gb_prev = torch.zeros((weights[i - 1].shape[0],), device=w.device)
            dz, _, _ = ops.linear_bf16(dz, img_t, n_in, Kd, sign_bits_in=bits_list[i - 1] if slope is not None else None,
                                       slope=slope or 0.0, colsum=gb_prev)
            gbs[i - 1] = gb_prev
        else:
            img_t = ops.pack_weight(w, in_pad, Kd, transpose=True)
            dz, _, _ = ops.linear_bf16(dz, img_t, in_pad, Kd, addend=d_in_extra)
    return dz, gWs, gbs


class _StackBf16Fn(torch.autograd.Function):
    """A whole MLP stack (Linear+act ... Linear) for per-ray fp32 tensors: activations stay tile-blocked bf16 between
    layers, sign bits drive the dgrad masks and the dgrad epilogues produce the bias gradients."""

    @staticmethod
    def forward(ctx, x, slope, n, grad, images, *wb):
        ws, bs = wb[:n], wb[n:]
        xb = ops.Blocked.from_f32(x)
        save = grad and any(ctx.needs_input_grad)
        img_f, ctx.img_t = images if images is not None else (None, None)
        inputs, bits, y = _stack_forward(xb, [w.detach() for w in ws], bs, slope, x.shape[1], save, last_f32=True, images=img_f)
        n_out = ws[-1].shape[0]
        if save:
            tensors, ctx.spec = _stash_pack(list(inputs) + list(bits))
            ctx.slope, ctx.n, ctx.rows, ctx.n_in = slope, n, x.shape[0], x.shape[1]
            ctx.save_for_backward(*ws, *tensors)
        return y[: x.shape[0], :n_out]

    @staticmethod
    def backward(ctx, gy):
        saved = ctx.saved_tensors
        ws = saved[:ctx.n]
        stash = _stash_unpack(saved[ctx.n:], ctx.spec)
        inputs, bits = stash[:ctx.n], stash[ctx.n:]
        n_out = ws[-1].shape[0]
        g = gy.contiguous()
        dz = ops.Blocked.from_f32(g, cols_pad=max(ops.pad_cols(n_out), 128))
        # the last layer's bias gradient = column sums of dz: from the weight-gradient kernel's idle warps when it can
        # (papr_wgrad_bias_bf16), not from a separate reduction over the (R, n_out) tensor
        from_wgrad = n_out >= 128
        gb_last = None if from_wgrad else g.sum(0)
        in_pad = ops.pad_cols(ctx.n_in)
        dx, gWs, gbs = _stack_backward(dz, inputs, bits, ws, ctx.slope, gb_last, ctx.n_in, in_pad, images_t=ctx.img_t,
                                       last_bias_from_wgrad=from_wgrad)
        return (dx.to_f32(ctx.rows, ctx.n_in), None, None, None, None, *gWs, *gbs)


class _QueryPrologueFn(torch.autograd.Function):
    """rays_d (R,3) -> the query stack's input a_2 * normalise(pe(rays_d)) + b_2 (R, 3(1+2L)) in one launch
    (utils.py:232-242, attn.py:30-42); gradients for the affine pair only -- the ray direction is data."""

    @staticmethod
    def forward(ctx, rays_d, a2, b2, L, eps):
        R = rays_d.shape[0]
        rd = rays_d.detach().contiguous()
        D = 3 * (1 + 2 * L)
        q = torch.empty((R, D), device=rd.device)
        ops.call("papr_query_prologue_fwd", rd.data_ptr(), a2.detach().contiguous().data_ptr(), b2.detach().contiguous().data_ptr(),
                 R, L, float(eps), q.data_ptr(), nbytes=R * (12.0 + 4 * D))
        ctx.save_for_backward(rd)
        ctx.L, ctx.eps = L, eps
        return q

    @staticmethod
    def backward(ctx, dq):
        (rd,) = ctx.saved_tensors
        R = rd.shape[0]
        D = 3 * (1 + 2 * ctx.L)
        dq = dq.contiguous()
        g = torch.zeros((2, D), device=rd.device)
        ops.call("papr_query_prologue_bwd", rd.data_ptr(), dq.data_ptr(), R, ctx.L, float(ctx.eps), g[0].data_ptr(),
                 g[1].data_ptr(), nbytes=R * (12.0 + 4 * D))
        return None, g[0], g[1], None, None


class _QueryTailFn(torch.autograd.Function):
    """(q5, A, c0, w_c) -> (ua, w_c . z):  z = normalise(q5);  ua = z A^T + c0."""

    @staticmethod
    def forward(ctx, q5, A, c0, w_c, eps):
        R = q5.shape[0]
        dev = q5.device
        q5c = q5.detach().contiguous()
        wc = w_c.detach().contiguous()
        zb = ops.Blocked(R, 256, dev)
        stats = torch.empty((R, 2), device=dev)
        cprime = torch.empty((R,), device=dev)
        ops.call("papr_query_tail_fwd", q5c.data_ptr(), wc.data_ptr(), 0.0, float(eps), R, zb.data_ptr(),
                 stats.data_ptr(), cprime.data_ptr(), nbytes=R * (1024.0 + 512 + 12))
        _, ua, _ = ops.linear_bf16(zb, ops.pack_weight(A, 256, 256), 256, 256, bias=c0.detach().contiguous(),
                                   out_blocked=False, out_f32=True)
        ctx.save_for_backward(q5c, stats, wc, A.detach())
        ctx.zb, ctx.eps = zb, eps
        return ua[:R], cprime

    @staticmethod
    def backward(ctx, d_ua, d_c):
        q5, stats, wc, A = ctx.saved_tensors
        R = q5.shape[0]
        dev = q5.device
        d_ua = d_ua.contiguous()
        d_c = d_c.contiguous()
        dub = ops.Blocked.from_f32(d_ua)
        _, dz, _ = ops.linear_bf16(dub, ops.pack_weight(A, 256, 256, transpose=True), 256, 256, out_blocked=False, out_f32=True)
        gA = torch.zeros((256, 256), device=dev)
        g_c0 = torch.zeros(256, device=dev)
        ops.wgrad_bf16(dub, ctx.zb, gA, 256, 256, a_colsum=g_c0)       # d c0 = column sums of d ua, from the same launch
        dq5 = torch.empty((R, 256), device=dev)
        g_wc = torch.zeros(256, device=dev)
        g_cc = torch.zeros(1, device=dev)
        ops.call("papr_query_tail_bwd", q5.data_ptr(), stats.data_ptr(), wc.data_ptr(), dz.data_ptr(), dz.stride(0),
                 d_c.data_ptr(), float(ctx.eps), R, dq5.data_ptr(), g_wc.data_ptr(), g_cc.data_ptr(),
                 nbytes=R * (3 * 1024.0 + 12))
        ctx.zb = None
        return dq5, gA, g_c0, g_wc, None


class RowAttentionFn(torch.autograd.Function):
    """(points, pc_feats, influ, ua, c', key-in LayerNorm, key/value stack weights) -> (fused, attn)."""

    @staticmethod
    def forward(ctx, sh, rays_o, rays_d, idx, points, feats, influ, ua, cprime, ln_a, ln_b, nk, *wb):
        kw, kb = wb[:nk], wb[nk:2 * nk]
        nv = (len(wb) - 2 * nk) // 2
        vw, vb = wb[2 * nk:2 * nk + nv], wb[2 * nk + nv:]
        save = sh.grad and any(ctx.needs_input_grad)
        pts = points.detach().contiguous()
        fts = feats.detach().contiguous() if feats is not None else None
        kin, vin, _, _ = _prologue_fwd(sh, rays_o, rays_d, pts, fts, idx, ln_a.detach(), ln_b.detach())
        k_in, k_bits, h5 = _stack_forward(kin, [w.detach() for w in kw], kb, sh.k_slope, sh.dk, save, skip_layers=sh.k_skip,
                                          images=sh.k_images[0])
        v_in, v_bits, v = _stack_forward(vin, [w.detach() for w in vw], vb, sh.v_slope, sh.dv, save, last_f32=True,
                                         skip_layers=sh.v_skip, images=sh.v_images[0])
        infl = influ.detach().reshape(-1).contiguous()
        fused, attn, sc, stats = _score_blend_fwd(sh, h5, None, ua.detach().contiguous(), cprime.detach().contiguous(),
                                                  infl, idx, v)
        if save:
            ctx.sh, ctx.nk, ctx.nv = sh, nk, nv
            tensors, ctx.spec = _stash_pack(list(k_in) + list(k_bits) + [h5] + list(v_in) + list(v_bits))
            ctx.save_for_backward(rays_o, rays_d, idx, pts, infl, ua.detach(), ln_a.detach(), v, attn, sc, stats, *wb,
                                  *tensors)
        return fused, attn

    @staticmethod
    def backward(ctx, d_fused, d_attn):
        sh, nk, nv = ctx.sh, ctx.nk, ctx.nv
        saved = ctx.saved_tensors           # one access only (non-reentrant checkpointing unpacks on access)
        rays_o, rays_d, idx, pts, infl, ua, ln_a, v, attn, sc, stats = saved[:11]
        n_wb = 2 * nk + 2 * nv
        wb = saved[11:11 + n_wb]
        kw, vw = wb[:nk], wb[2 * nk:2 * nk + nv]
        st = _stash_unpack(saved[11 + n_wb:], ctx.spec)
        k_in, k_bits, h5 = st[:nk], st[nk:2 * nk], st[2 * nk]
        v_in, v_bits = st[2 * nk + 1:2 * nk + 1 + nv], st[2 * nk + 1 + nv:]
        P = pts.shape[0]
        d_attn_c = d_attn.contiguous() if d_attn is not None else None
        dv, d_score, g_influ, g_bv = _blend_bwd(sh, d_fused.contiguous(), d_attn_c, attn, sc, infl, idx, v, P)
        d_vin, gvW, gvb = _stack_backward(dv, v_in, v_bits, vw, sh.v_slope, g_bv, sh.dv, sh.dv_pad, sh.v_skip, images_t=sh.v_images[1])
        dh5, _, zsum, dssum, g_b5 = _key_score_bwd(sh, d_score, h5, None, stats, ua, want_bias=not WGRAD_BIAS)
        d_kin, gkW, gkb = _stack_backward(dh5, k_in, k_bits, kw, sh.k_slope, g_b5, sh.dk, sh.dk_pad, sh.k_skip, images_t=sh.k_images[1],
                                          last_bias_from_wgrad=g_b5 is None)
        g_points, g_feats, g_a, g_b = _prologue_bwd(sh, rays_o, rays_d, pts, idx, ln_a, d_kin, d_vin, None, None, P)
        return (None, None, None, None, g_points, g_feats, g_influ.reshape(-1, 1), zsum, dssum, g_a, g_b, None,
                *gkW, *gkb, *gvW, *gvb)


class _PrologueFp32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sh, rays_o, rays_d, idx, points, feats, ln_a, ln_b):
        pts = points.detach().contiguous()
        fts = feats.detach().contiguous() if feats is not None else None
        _, _, kin32, vin32 = _prologue_fwd(sh, rays_o, rays_d, pts, fts, idx, ln_a.detach(), ln_b.detach(), taps=True)
        ctx.sh = sh
        ctx.save_for_backward(rays_o, rays_d, idx, pts, ln_a.detach())
        return kin32, vin32

    @staticmethod
    def backward(ctx, dkin, dvin):
        rays_o, rays_d, idx, pts, ln_a = ctx.saved_tensors
        g_points, g_feats, g_a, g_b = _prologue_bwd(ctx.sh, rays_o, rays_d, pts, idx, ln_a, None, None,
                                                    dkin.contiguous(), dvin.contiguous(), pts.shape[0])
        return None, None, None, None, g_points, g_feats, g_a, g_b


class _ScoreBlendFp32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sh, idx, h5, v, ua, cprime, influ):
        infl = influ.detach().reshape(-1).contiguous()
        h5c, vc, uac = h5.detach().contiguous(), v.detach().contiguous(), ua.detach().contiguous()
        fused, attn, sc, stats = _score_blend_fwd(sh, None, h5c, uac, cprime.detach().contiguous(), infl, idx, vc)
        ctx.sh, ctx.P = sh, infl.shape[0]
        ctx.save_for_backward(idx, h5c, vc, uac, infl, attn, sc, stats)
        return fused, attn

    @staticmethod
    def backward(ctx, d_fused, d_attn):
        sh = ctx.sh
        idx, h5, v, ua, infl, attn, sc, stats = ctx.saved_tensors
        _, d_score, g_influ, _ = _blend_bwd(sh, d_fused.contiguous(), d_attn.contiguous(), attn, sc, infl, idx, v, ctx.P)
        _, dh5, zsum, dssum, _ = _key_score_bwd(sh, d_score, None, h5, stats, ua, tap=True)
        w = attn[:, :sh.K]
        if sh.normalize:
            w = w / w.sum(-1, keepdim=True)
        d_v = (w.unsqueeze(-1) * d_fused.unsqueeze(1)).reshape(sh.M, sh.C)
        return None, None, dh5, d_v, zsum, dssum, g_influ.reshape(-1, 1)


class _ChunkFn(torch.autograd.Function):
    """One checkpointed chunk of rays.  Forward runs the chunk WITHOUT autograd -- the inference form of every kernel: no
    layer inputs, sign bits or row statistics are written, the MLP stacks run at their no-stash speed -- and keeps only the
    chunk's inputs.  Backward runs the chunk again with the backward stash and back-propagates through it at once, so at
    most one chunk's stash is alive.  (torch.utils.checkpoint would run the first pass in training form and throw the
    stash away: a third of the first pass's HBM traffic for nothing.)  Gradients of the module's own parameters are
    accumulated into their .grad by the inner backward, exactly as with a re-entrant torch checkpoint; the gradients of
    the point cloud tensors are returned to the outer graph."""

    @staticmethod
    def forward(ctx, module, meta, rays_o, rd, idx, points, feats, influ):
        ctx.module, ctx.meta = module, meta
        ctx.save_for_backward(rays_o, rd, idx, points, feats, influ)
        with torch.no_grad():
            return module._rows(rays_o, rd, idx, points, feats, influ, *meta)

    @staticmethod
    def backward(ctx, d_fused, d_attn):
        rays_o, rd, idx, points, feats, influ = ctx.saved_tensors
        leaves = []
        for t, need in zip((points, feats, influ), ctx.needs_input_grad[5:8]):
            leaves.append(t.detach().requires_grad_(True) if (t is not None and need) else t)
        with torch.enable_grad():
            fused, attn = ctx.module._rows(rays_o, rd, idx, *leaves, *ctx.meta)
            outs, grads = [], []
            for o, g in ((fused, d_fused), (attn, d_attn)):
                if g is not None and o.requires_grad:
                    outs.append(o)
                    grads.append(g)
            if outs:
                torch.autograd.backward(outs, grads)
        return (None, None, None, None, None) + tuple(t.grad if (t is not None and t.requires_grad) else None for t in leaves)


class ProximityAttention(nn.Module):
    """Owns the reference's attention parameters (same state_dict keys) and runs the B200 path."""

    def __init__(self, args, point_feat_dim, eps=1e-6, bkg_score=5.0, normalize=True, precision="bf16"):
        super().__init__()
        A, E = args, args.embed
        if A.k_type != 1 or A.q_type != 1 or A.v_type != 1:
            raise ValueError("Invalid key/query/value type")
        if E.embed_type != 1:
            raise NotImplementedError("embed_type 2 (PE without the input itself) is not used by any shipped config")
        Ls = set(E.k_L) | set(E.q_L) | set(E.v_L)
        if len(Ls) != 1:
            raise NotImplementedError("the B200 path needs one positional-encoding order for all features")
        self.L = int(next(iter(Ls)))
        self.F = int(point_feat_dim)
        S = 1 + 2 * self.L
        self.dk, self.dq, self.dv = 9 * S, 3 * S, 6 * S + self.F
        self.C = E.value.d_ff_out
        self.eps = eps
        self.bkg_score, self.normalize = float(bkg_score), bool(normalize)
        self.pe_factor, self.pe_mult = E.pe_factor, E.pe_mult_factor
        if E.pe_factor != 2.0 or E.pe_mult_factor != 1.0:
            raise NotImplementedError("pe_factor/pe_mult_factor other than 2/1 are not used by any shipped config")
        if E.key.d_ff_out != 256 or A.d_model != 256 or E.query.d_ff_out != 256:
            raise NotImplementedError("the B200 score kernel is specialised to d_model = key/query width = 256")
        if E.key.norm != "layernorm" or E.query.norm != "layernorm" or E.value.norm != "none":
            raise NotImplementedError("only key/query layernorm + value none (the shipped setting) is supported")
        for part in (E.key, E.query, E.value):
            if part.ff_last_act != "none":
                raise NotImplementedError("ff_last_act other than 'none' is not used by any shipped config")
        self.k_slope = activation_slope(E.key.ff_act)
        self.v_slope = activation_slope(E.value.ff_act)
        self.q_slope = activation_slope(E.query.ff_act)
        if A.score_act not in ("relu", "none"):
            raise NotImplementedError("score_act must be relu or none")
        self.score_relu = A.score_act == "relu"
        self.d_model = A.d_model
        self.precision = precision
        self.embed = Embeddings(self.dk, self.dq, self.dv, E, eps)
        self.attention_layer = AttentionLayer(E, A.d_model, A.score_act)
        self.weight_images = WeightImages(self)

    # ------------------------------------------------------------------ per-ray query side (5% of the work)
    def query_terms(self, rays_d_flat, precision):
        """(R,3) -> ua (R,256), c' (R): the query stack, w_q and the fold of w_k / key outnorm (see module doc)."""
        fq = self.embed.embed_q
        if (precision != "fp32" and self.L in (4, 6) and hasattr(fq.innorm, "a_2") and rays_d_flat.is_cuda
                and not os.environ.get("PAPR_QUERY_TORCH")):
            q = _QueryPrologueFn.apply(rays_d_flat, fq.innorm.a_2, fq.innorm.b_2, self.L, fq.innorm.eps)
        else:
            q = fq.innorm(posenc(rays_d_flat, self.L))
        lins = fq.mlp.linears()
        if precision == "fp32":
            q = _mlp_fp32(fq.mlp, q)
        elif fq.mlp.skip_layers:
            raise NotImplementedError("skip_layers in the query stack are not used by any shipped config")
        else:
            imgs = self.weight_images.stack("q", len(lins))
            q = _StackBf16Fn.apply(q, self.q_slope, len(lins), torch.is_grad_enabled(), imgs if imgs[0] is not None else None,
                                   *[l.weight for l in lins], *[l.bias for l in lins])
        al = self.attention_layer
        scale = 1.0 / math.sqrt(self.d_model)
        on = self.embed.embed_k.outnorm
        if precision == "fp32":
            q = fq.outnorm(q)
            qp = _linear(q, al.w_q, None, precision)                           # q' (R,256)
            wk_t = al.w_k.weight.t()
            u = (split_gemm.matmul_t(qp, wk_t) if FP32_GEMM == "split" else qp @ al.w_k.weight) * scale
            ua = u * on.a_2
            cprime = (u * on.b_2).sum(-1) + (qp * al.w_k.bias).sum(-1) * scale
            return ua, cprime
        # Everything after the normalisation z = (q - mean)/(std + eps) is linear in z (attn.py:39-42, 217-218, 53-54):
        #   q' = W_q (a_q z + b_q) + c_q ;  u = W_k^T q' / sqrt(d) ;  ua = u a_k ;  c' = u . b_k2 + q' . c_k / sqrt(d)
        # so it is folded into ONE 256x256 matrix (tiny fp32 torch ops, differentiable) applied on the tensor cores.
        qn = fq.outnorm
        M1 = scale * (al.w_k.weight.t() @ al.w_q.weight)                        # u = M1 (a_q z) + u0
        q0 = al.w_q.weight @ qn.b_2 + al.w_q.bias
        u0 = scale * (al.w_k.weight.t() @ q0)
        A = on.a_2[:, None] * M1 * qn.a_2[None, :]
        c0 = on.a_2 * u0
        w_c = (M1.t() @ on.b_2 + scale * (al.w_q.weight.t() @ al.w_k.bias)) * qn.a_2
        c_const = on.b_2 @ u0 + scale * (al.w_k.bias @ q0)
        ua, cdot = _QueryTailFn.apply(q, A, c0, w_c, self.eps)
        cprime = cdot + c_const
        return ua, cprime

    def _rows(self, rays_o, rd, idx, points, feats, influ, precision, n_views, rays_per_view):
        """One batch of rays: rd (R,3), idx (R,K) -> fused (R,C), attn (R,K+1)."""
        K = idx.shape[-1]
        sh = _Shape(self, n_views * rays_per_view, rays_per_view, K)
        ua, cprime = self.query_terms(rd, precision)
        fk, fv = self.embed.embed_k, self.embed.embed_v
        if precision == "fp32":
            kin, vin = _PrologueFp32Fn.apply(sh, rays_o, rd, idx, points, feats, fk.innorm.a_2, fk.innorm.b_2)
            h5 = _mlp_fp32(fk.mlp, kin)
            v = _mlp_fp32(fv.mlp, vin)
            return _ScoreBlendFp32Fn.apply(sh, idx, h5, v, ua, cprime, influ)
        klin, vlin = fk.mlp.linears(), fv.mlp.linears()
        wb = [l.weight for l in klin] + [l.bias for l in klin] + [l.weight for l in vlin] + [l.bias for l in vlin]
        return RowAttentionFn.apply(sh, rays_o, rd, idx, points, feats, influ, ua, cprime, fk.innorm.a_2, fk.innorm.b_2,
                                    len(klin), *wb)

    #: bytes per (ray, candidate) row alive at the peak of a call in the bf16 path: with the backward stash (layer
    #: inputs, sign bits, row statistics) and without it (inference: kin, vin, h5, v only)
    STASH_BYTES_PER_ROW = 8192
    INFER_BYTES_PER_ROW = 1536

    def forward(self, rays_o, rays_d, idx, points, feats, influ, precision=None, ray_chunk=None):
        """rays_o (N,3), rays_d (N,H,W,3), idx int32 (N,H,W,K) -> fused (R,C), attn (R,K+1); differentiable.

        ray_chunk: process at most that many rays of one view at a time.  Under autograd each chunk is checkpointed
        (its forward is recomputed during backward), so the backward stash is bounded by one chunk instead of growing
        with the frame -- the replacement for the reference's fixed 160x160 training patches when a whole frame (or a
        1920x1080 one) is trained on at once.  None = automatic: chunk only if the working set would not fit in free
        memory (with or without grad)."""
        precision = precision or self.precision
        N, H, W, _ = rays_d.shape
        K = idx.shape[-1]
        if K > 32:
            raise NotImplementedError("the blend kernels hold a ray's K candidates in the lanes of one warp: K <= 32")
        rays_o = rays_o.detach().float().contiguous()
        rd = rays_d.detach().float().contiguous().reshape(N, H * W, 3)
        idx = idx.reshape(N, H * W, K).contiguous()
        grad = torch.is_grad_enabled()
        if ray_chunk is None and rd.is_cuda:
            free, _ = torch.cuda.mem_get_info(rd.device)
            free += torch.cuda.memory_reserved(rd.device) - torch.cuda.memory_allocated(rd.device)
            per_row = self.STASH_BYTES_PER_ROW if grad else self.INFER_BYTES_PER_ROW
            if precision == "fp32":
                per_row *= 6
            need = N * H * W * K * per_row
            if need > 0.8 * free:
                # Equal chunks sized from the device's TOTAL memory, so that every step of a run allocates the same block
                # sizes (a chunk size that follows the momentary free memory makes the caching allocator free and
                # re-allocate segments every step: 24 cudaMalloc + 57 synchronising cudaFree per step, 2.3x slower);
                # only when even that does not fit is the chunk cut to what is free right now.
                total = torch.cuda.get_device_properties(rd.device).total_memory
                n_chunks = max(2, -(-need // int(0.3 * total)))
                ray_chunk = -(-(-(-(N * H * W) // n_chunks)) // 128) * 128
                if ray_chunk * K * per_row > 0.5 * free:
                    ray_chunk = max(4096, int(0.4 * free / (K * per_row)) // 128 * 128)
        with torch.cuda.device(rd.device):      # the C ABI launches on the current device
            if precision != "fp32":
                self.weight_images.refresh()    # one launch: every weight image of the three stacks, both directions
            if not ray_chunk or (N == 1 and H * W <= ray_chunk) or (N * H * W <= ray_chunk):
                return self._rows(rays_o, rd.reshape(-1, 3), idx.reshape(-1, K), points, feats, influ, precision, N, H * W)
            fused, attn = [], []
            for v in range(N):
                for r0 in range(0, H * W, ray_chunk):
                    r1 = min(r0 + ray_chunk, H * W)
                    args = (rays_o[v:v + 1], rd[v, r0:r1], idx[v, r0:r1], points, feats, influ)
                    if grad:
                        f, a = _ChunkFn.apply(self, (precision, 1, r1 - r0), *args)
                    else:
                        f, a = self._rows(*args, precision, 1, r1 - r0)
                    fused.append(f)
                    attn.append(a)
            return torch.cat(fused), torch.cat(attn)
