"""Stage a10: the SmallUNet decode (reference models/unet.py:196-258) on the library's own kernels.

Every convolution is an implicit GEMM on tcgen05 (``papr_conv_bf16``: 3x3 / 1x1 forward and data gradient;
``papr_conv_wgrad_bf16``: weight gradient), the 2x2 transposed convolutions are 1x1 GEMMs followed by a pixel shuffle,
and the glue (max-pool, ReLU masks, skip concatenation, FiLM, bias gradients) is a handful of raster kernels
(``papr_unet_*``).  Feature maps live in "pixel planes" (see csrc/conv.cu): zero-padded rasters split into planes of 64
channels whose 128-row tiles are tensor-core operand tiles; the skip concatenations of unet.py:75 are free because the
two halves are simply different planes of one buffer.  No cuDNN / cuBLAS kernel runs in this path.

Data flow (C = channels, @ = resolution):
  fused (32 @ H) -[FiLM]-> P0 -conv 32>128-> U2[0:128] -pool-> P1 -conv 128>256-> U1[0:256] -pool-> P2 -conv 256>512-> X3
  X3 -convT 512>256-> U1[256:512] ; U1 -conv 512>256-> Y1 -convT 256>128-> U2[128:256] ; U2 -conv 256>128-> Y2 -1x1-> rgb
"""
import ctypes

import torch

from . import ops
from ._lib import PackDesc, Raster, SpreadArgs


class Planes:
    """A feature map of one image in the pixel-plane layout: uint8 (copies, planes of 64 channels, rows, 128)."""

    def __init__(self, H, W, C, copies, device, zero=True):
        """zero: True = memset the whole buffer; "border" = zero only what no raster kernel writes; False = leave as is."""
        self.H, self.W, self.C, self.copies = H, W, C, copies
        self.cbs = (C + 63) // 64
        self.Wp = (W + 2 + 7) // 8 * 8
        self.L = ((H + 2) * self.Wp + 127) // 128 * 128
        self.G0 = (self.Wp + 8 + 127) // 128 * 128
        self.S = self.G0 + self.L + self.G0
        alloc = torch.zeros if zero is True else torch.empty
        self.buf = alloc((copies, self.cbs, self.S, 128), dtype=torch.uint8, device=device)
        self.plane_bytes = self.S * 128
        self.copy_bytes = self.cbs * self.plane_bytes
        self.n_tiles = self.L // 128
        if zero == "border":
            # the producer of this map writes EVERY interior pixel of every channel block: only the guard rows and the
            # padding frame need zeros (papr_unet_zero_border), 3% of the bytes a full memset would write
            ops.call("papr_unet_zero_border", self.ptr(), _ref(self.raster()), copies, self.cbs, nbytes=0.03 * self.buf.numel())

    def ptr(self, copy=0, cb=0, row=0):
        return self.buf.data_ptr() + copy * self.copy_bytes + cb * self.plane_bytes + row * 128

    def mid(self, cb=0):
        """The unshifted copy (dx = 0): what the raster kernels and 1x1 convolutions read."""
        return self.ptr(copy=1 if self.copies == 3 else 0, cb=cb)

    def raster(self):
        r = Raster()
        r.H, r.W, r.Wp, r.row0, r.plane_bytes, r.copy_bytes = self.H, self.W, self.Wp, self.G0, self.plane_bytes, self.copy_bytes
        return r


def _ref(struct):
    return ctypes.cast(ctypes.pointer(struct), ctypes.c_void_p)


# ----------------------------------------------------------------------------------------------------- weight images
def _pack_matrix(mat, n_tile=256):
    """fp32 (rows, K) matrix, K a multiple of 64 -> list of (image uint8, N) per tile of <= 256 rows (N padded to 32)."""
    rows, K = mat.shape
    mat = mat.contiguous()
    tiles, descs, keep = [], [], []
    for r0 in range(0, rows, n_tile):
        n = min(n_tile, rows - r0)
        N = (n + 31) // 32 * 32
        img = torch.empty((K // 64) * N * 128, dtype=torch.uint8, device=mat.device)
        d = PackDesc()
        d.w, d.ld, d.rows, d.cols, d.transpose, d.N, d.K = mat.data_ptr() + r0 * K * 4, K, n, K, 0, N, K
        d.replicas, d.scale, d.rep_stride, d.image = 1, 1.0, 0, img.data_ptr()
        descs.append(d)
        tiles.append((img, N, n))
    return tiles, descs, mat


def _launch_pack(all_descs, device):
    arr = (PackDesc * len(all_descs))(*all_descs)
    table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device, non_blocking=True)
    ops.call("papr_pack_weight_batch", table.data_ptr(), len(all_descs), nbytes=0.0)
    return table


def _pad_k(t, k):
    return t if t.shape[-1] == k else torch.nn.functional.pad(t, (0, k - t.shape[-1]))


class _Weights:
    """bf16 weight images of every UNet layer (forward and data-gradient orientation) in PERSISTENT buffers: fp32 staging
    matrices with fixed addresses, the images, padded biases and the device-side descriptor table are built once per set of
    parameter tensors; `refresh` then costs one small copy per matrix plus ONE papr_pack_weight_batch launch -- and no
    host-to-device copy (a pageable one would synchronise the stream and stop the host from running ahead)."""

    def __init__(self, params):
        self.key = tuple(t.data_ptr() for t in params)
        self.stage, self.views = [], []          # staging matrix, function producing the source view from the live weights
        self.fwd, self.bwd, self.bias = {}, {}, {}
        descs = []

        def add(table, name, shape, view_fn):
            mat = torch.zeros(shape, dtype=torch.float32, device=params[0].device)
            tiles, d, _ = _pack_matrix(mat)
            table[name] = tiles
            descs.extend(d)
            self.stage.append(mat)
            self.views.append(view_fn)

        for i, (name, _) in enumerate(_NAMES):
            w = params[2 * i]
            if name in ("inc", "down1", "down2", "up1c", "up2c"):                         # (Cout, Cin, 3, 3)
                co, ci = w.shape[:2]
                cip, cop = (ci + 63) // 64 * 64, (co + 63) // 64 * 64
                add(self.fwd, name, (co, 9 * cip), lambda t, ci=ci, cip=cip: (t.permute(0, 2, 3, 1), (t.shape[0], 9, cip), ci))
                add(self.bwd, name, (ci, 9 * cop), lambda t, co=co, cop=cop: (t.permute(1, 2, 3, 0), (t.shape[1], 9, cop), co))
            elif name in ("up1t", "up2t"):                                                 # ConvTranspose2d: (Cin, Cout, 2, 2)
                ci, co = w.shape[:2]
                add(self.fwd, name, (4 * co, ci), lambda t: (t.permute(2, 3, 1, 0), (4 * t.shape[1], 1, t.shape[0]), t.shape[0]))
                add(self.bwd, name, (ci, 4 * co), lambda t: (t.permute(0, 2, 3, 1), (t.shape[0], 1, 4 * t.shape[1]), 4 * t.shape[1]))
            else:                                                                          # outc (3, 128, 1, 1)
                co, ci = w.shape[:2]
                add(self.fwd, name, (co, ci), lambda t: (t.reshape(t.shape[0], t.shape[1]), (t.shape[0], 1, t.shape[1]), t.shape[1]))
                add(self.bwd, name, (ci, 64), lambda t: (t.reshape(t.shape[0], t.shape[1]).t(), (t.shape[1], 1, 64), t.shape[0]))
            self.bias[name] = []
            n0 = 0
            for _, N, n in (self.fwd[name] if name not in ("up1t", "up2t") else []):      # the pixel shuffle adds the convT biases
                self.bias[name].append((torch.zeros(N, dtype=torch.float32, device=w.device), n0, n))
                n0 += n
        self.n_descs = len(descs)
        arr = (PackDesc * len(descs))(*descs)
        self.table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(params[0].device)      # once per parameter set

    def refresh(self, params):
        """Re-read the live parameters into the staging matrices and rebuild every image (one launch)."""
        k = 0
        for i, (name, _) in enumerate(_NAMES):
            w, b = params[2 * i].detach(), params[2 * i + 1].detach()
            for _ in range(2):
                src, shape3, valid = self.views[k](w)
                dst = self.stage[k].view(shape3)
                if src.dim() == 4:          # (rows, 3, 3, C) -> (rows, 9, C_pad)
                    dst[:, :, :valid].copy_(src.reshape(shape3[0], shape3[1], valid))
                else:
                    dst[:, 0, :valid].copy_(src)
                k += 1
            for buf, n0, n in self.bias[name]:
                buf[:n].copy_(b[n0:n0 + n])
        ops.call("papr_pack_weight_batch", self.table.data_ptr(), self.n_descs, nbytes=sum(6.0 * m.numel() for m in self.stage))


# ----------------------------------------------------------------------------------------------------- kernel wrappers
def _conv(src, src_ptr, cbs, ntaps, sign, tiles, bias, act, out, out_cb0=0, out_f32=None, addend=None):
    """One convolution over all output-channel tiles.  src_ptr: plane 0 of copy dx=-1 (3x3) or of the unshifted copy (1x1).
    addend: fp32 (rows, ld) partial sums added before bias / activation (same column tiling as out_f32)."""
    n0 = 0
    for ti, (img, N, n) in enumerate(tiles):
        b = None
        if isinstance(bias, list):           # padded per-tile biases kept by _Weights
            b = bias[ti][0]
        elif bias is not None:
            b = torch.zeros(N, dtype=torch.float32, device=img.device)
            b[:n] = bias[n0:n0 + n]
        ops.call("papr_conv_bf16", src_ptr, src.copy_bytes, src.plane_bytes, src.G0, cbs, ntaps, src.Wp, sign, img.data_ptr(),
                 b.data_ptr() if b is not None else None, N, int(act), 0.0,
                 out.ptr(cb=out_cb0 + n0 // 64) if out is not None else None, out.plane_bytes if out is not None else 0,
                 out.G0 if out is not None else 0,
                 out_f32.data_ptr() + n0 * 4 if out_f32 is not None else None, out_f32.stride(0) if out_f32 is not None else 0,
                 addend.data_ptr() + n0 * 4 if addend is not None else None, addend.stride(0) if addend is not None else 0,
                 src.n_tiles, flops=2.0 * src.H * src.W * n * cbs * 64 * ntaps,
                 nbytes=src.L * 128.0 * (cbs * (3 if ntaps == 9 else 1) + N / 64))
        n0 += n


def _spread(src, src_cb0, cbs, dst=None, dst_cb0=0, add=None, add_cb0=0, mask=None, mask_cb0=0, pool_grad=None, pool_ref=None,
            pool_ref_cb0=0, gamma=None, beta=None, colsum=None):
    a = SpreadArgs()
    a.src, a.src_cb0 = src.mid(), src_cb0
    a.geom = src.raster()
    a.cbs = cbs
    if add is not None:
        a.add, a.add_cb0 = add.mid(), add_cb0
    if mask is not None:
        a.mask, a.mask_cb0 = mask.mid(), mask_cb0
    if pool_grad is not None:
        a.pool_grad, a.pool_geom = pool_grad.mid(), pool_grad.raster()
        a.pool_ref, a.pool_ref_cb0 = pool_ref.mid(), pool_ref_cb0
    if gamma is not None:
        a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
    if dst is not None:
        a.dst, a.dst_geom, a.dst_cb0, a.ncopies = dst.ptr(), dst.raster(), dst_cb0, dst.copies
    if colsum is not None:
        a.colsum = colsum.data_ptr()
    ops.call("papr_unet_spread", _ref(a), nbytes=src.H * src.W * cbs * 128.0 * (2 + (dst.copies if dst is not None else 0)))


def _pool(src, src_cb0, cbs, dst):
    ops.call("papr_unet_pool", src.mid(), _ref(src.raster()), src_cb0, dst.ptr(), _ref(dst.raster()), dst.copies, cbs,
             nbytes=src.H * src.W * cbs * 128.0 * 2)


def _wgrad(a, a_cb0, a_valid, b, b_cb0, b_valid, ntaps, out, tap_stride):
    """out[tap] (a_valid, ldc) += sum_rows A[row, a] * B_tap[row, b] over the raster rows of `a` (all taps in one launch)."""
    b_ptr = b.ptr(copy=0, cb=b_cb0, row=b.G0) if ntaps == 9 else b.mid(b_cb0) + b.G0 * 128
    ops.call("papr_conv_wgrad_bf16", a.mid(a_cb0) + a.G0 * 128, a.plane_bytes, a_valid, b_ptr, b.plane_bytes, b.copy_bytes, b_valid,
             ntaps, b.Wp, out.data_ptr(), out.stride(-2), tap_stride, a.L,
             flops=2.0 * a.H * a.W * a_valid * b_valid * ntaps,
             nbytes=a.L * 128.0 * ntaps * ((a_valid + 127) // 128 * 2 + (b_valid + 63) // 64))


def _conv_wgrad(dz, x, co, ci):
    """Weight gradient of a 3x3 convolution: dz (Cout channels, unshifted copy clean), x (3-copy input) -> (Cout, Cin, 3, 3)."""
    g = torch.zeros((9, co, ci), dtype=torch.float32, device=dz.buf.device)
    for a0 in range(0, co, 256):
        for b0 in range(0, ci, 256):
            _wgrad(dz, a0 // 64, min(256, co - a0), x, b0 // 64, min(256, ci - b0), 9, g[:, a0:, b0:], g.stride(0))
    return g.permute(1, 2, 0).reshape(co, ci, 3, 3)


def _gemm_wgrad(a, a_ch, b, b_ch):
    """(a_ch, b_ch) = A^T B for two maps of one raster (1x1 / transposed convolutions)."""
    g = torch.zeros((a_ch, b_ch), dtype=torch.float32, device=a.buf.device)
    for a0 in range(0, a_ch, 256):
        for b0 in range(0, b_ch, 256):
            _wgrad(a, a0 // 64, min(256, a_ch - a0), b, b0 // 64, min(256, b_ch - b0), 1, g[a0:, b0:], 0)
    return g


# ----------------------------------------------------------------------------------------------------- forward / backward
_NAMES = (("inc", "inc.double_conv.0"), ("down1", "down1.maxpool_conv.1.double_conv.0"), ("down2", "down2.maxpool_conv.1.double_conv.0"),
          ("up1t", "up1.up"), ("up1c", "up1.conv.double_conv.0"), ("up2t", "up2.up"), ("up2c", "up2.conv.double_conv.0"), ("outc", "outc.conv"))


def unet_forward_image(x_hwc, p, W, gamma, beta, keep):
    """One image: x_hwc (H, W, C_in) fp32 -> (rgb (H, W, 3) fp32, saved maps or None)."""
    H, Wd, Cin = x_hwc.shape
    dev = x_hwc.device
    H2, W2, H3, W3 = H // 2, Wd // 2, H // 4, Wd // 4
    P0 = Planes(H, Wd, 64, 3, dev, zero="border")
    ops.call("papr_unet_pack_input", x_hwc.data_ptr(), x_hwc.stride(1), Cin, gamma.data_ptr() if gamma is not None else None,
             beta.data_ptr() if beta is not None else None, P0.ptr(), _ref(P0.raster()), 3, 1, nbytes=H * Wd * (4.0 * Cin + 384))
    U2 = Planes(H, Wd, 256, 3, dev, zero="border" if (H == 2 * H2 and Wd == 2 * W2) else True)   # odd sizes: the up-sampled half has an unwritten rim
    T = Planes(H, Wd, 128, 1, dev, zero=False)
    _conv(P0, P0.ptr(), 1, 9, 1, W.fwd["inc"], W.bias["inc"], True, T)
    _spread(T, 0, 2, dst=U2, dst_cb0=0)
    P1 = Planes(H2, W2, 128, 3, dev, zero="border")
    _pool(U2, 0, 2, P1)
    U1 = Planes(H2, W2, 512, 3, dev, zero="border" if (H2 == 2 * H3 and W2 == 2 * W3) else True)
    T = Planes(H2, W2, 256, 1, dev, zero=False)
    _conv(P1, P1.ptr(), 2, 9, 1, W.fwd["down1"], W.bias["down1"], True, T)
    _spread(T, 0, 4, dst=U1, dst_cb0=0)
    P2 = Planes(H3, W3, 256, 3, dev, zero="border")
    _pool(U1, 0, 4, P2)
    X3 = Planes(H3, W3, 512, 1, dev, zero=False)
    _conv(P2, P2.ptr(), 4, 9, 1, W.fwd["down2"], W.bias["down2"], True, X3)
    # up1: transposed convolution = 1x1 GEMM to (a, b, co) channels + pixel shuffle into the upper planes of U1 (the concat)
    T = Planes(H3, W3, 1024, 1, dev, zero=False)
    _conv(X3, X3.ptr(), 8, 1, 1, W.fwd["up1t"], None, False, T)
    ops.call("papr_unet_convt_scatter", T.ptr(), _ref(T.raster()), 256, p["up1t.b"].data_ptr(), U1.ptr(), _ref(U1.raster()), 4, 3,
             (H2 - 2 * H3) // 2, (W2 - 2 * W3) // 2, nbytes=H2 * W2 * 256 * 8.0)
    Y1 = Planes(H2, W2, 256, 1, dev, zero=False)
    _conv(U1, U1.ptr(), 8, 9, 1, W.fwd["up1c"], W.bias["up1c"], True, Y1)
    T = Planes(H2, W2, 512, 1, dev, zero=False)
    _conv(Y1, Y1.ptr(), 4, 1, 1, W.fwd["up2t"], None, False, T)
    ops.call("papr_unet_convt_scatter", T.ptr(), _ref(T.raster()), 128, p["up2t.b"].data_ptr(), U2.ptr(), _ref(U2.raster()), 2, 3,
             (H - 2 * H2) // 2, (Wd - 2 * W2) // 2, nbytes=H * Wd * 128 * 8.0)
    Y2 = Planes(H, Wd, 128, 1, dev, zero=False)
    _conv(U2, U2.ptr(), 4, 9, 1, W.fwd["up2c"], W.bias["up2c"], True, Y2)
    out = torch.empty((Y2.L, 32), dtype=torch.float32, device=dev)
    _conv(Y2, Y2.ptr(), 2, 1, 1, W.fwd["outc"], W.bias["outc"], False, None, out_f32=out)
    rgb = out[: (H + 2) * Y2.Wp].view(H + 2, Y2.Wp, 32)[1:H + 1, 1:Wd + 1, :3]
    saved = (P0, U2, P1, U1, P2, X3, Y1, Y2) if keep else None
    return rgb, saved


def unet_backward_image(d_rgb, p, W, saved, Cin, need_input_grad):
    """One image: d_rgb (H, W, 3) fp32 -> (d_x (H, W, Cin) fp32 or None, dict of parameter gradients)."""
    P0, U2, P1, U1, P2, X3, Y1, Y2 = saved
    H, Wd = P0.H, P0.W
    H2, W2, H3, W3 = P1.H, P1.W, P2.H, P2.W
    dev = d_rgb.device
    g = {}
    d_rgb = d_rgb.contiguous()
    dO = Planes(H, Wd, 128, 1, dev, zero="border")
    ops.call("papr_unet_pack_input", d_rgb.data_ptr(), d_rgb.stride(1), 3, None, None, dO.ptr(), _ref(dO.raster()), 1, 2,
             nbytes=H * Wd * (12.0 + 256))
    # outc (1x1)
    g["outc.w"] = _gemm_wgrad(dO, 3, Y2, 128).reshape(3, 128, 1, 1)
    g["outc.b"] = d_rgb.sum((0, 1))
    T = Planes(H, Wd, 256, 1, dev, zero=False)
    _conv(dO, dO.ptr(), 1, 1, 1, W.bwd["outc"], None, False, T)
    dZ = Planes(H, Wd, 128, 3, dev, zero="border")
    g["up2c.b"] = torch.zeros(128, device=dev)
    _spread(T, 0, 2, dst=dZ, mask=Y2, colsum=g["up2c.b"])
    # up2.conv (256 -> 128)
    g["up2c.w"] = _conv_wgrad(dZ, U2, 128, 256)
    _conv(dZ, dZ.ptr(), 2, 9, -1, W.bwd["up2c"], None, False, T)                  # T = d U2: [d x1 (skip) | d up2]
    dSkip1 = T
    # up2.up (ConvTranspose 256 -> 128)
    Q = Planes(H2, W2, 512, 1, dev)
    g["up2t.b"] = torch.zeros(128, device=dev)
    ops.call("papr_unet_convt_gather", T.ptr(), _ref(T.raster()), 2, 128, Q.ptr(), _ref(Q.raster()), (H - 2 * H2) // 2, (Wd - 2 * W2) // 2,
             g["up2t.b"].data_ptr(), nbytes=H * Wd * 128 * 4.0)
    g["up2t.w"] = _gemm_wgrad(Q, 512, Y1, 256).reshape(2, 2, 128, 256).permute(3, 2, 0, 1)
    T2 = Planes(H2, W2, 512, 1, dev, zero=False)
    _conv(Q, Q.ptr(), 8, 1, 1, W.bwd["up2t"], None, False, T2)
    dZ = Planes(H2, W2, 256, 3, dev, zero="border")
    g["up1c.b"] = torch.zeros(256, device=dev)
    _spread(T2, 0, 4, dst=dZ, mask=Y1, colsum=g["up1c.b"])
    # up1.conv (512 -> 256)
    g["up1c.w"] = _conv_wgrad(dZ, U1, 256, 512)
    _conv(dZ, dZ.ptr(), 4, 9, -1, W.bwd["up1c"], None, False, T2)                 # T2 = d U1: [d x2 (skip) | d up1]
    # up1.up (ConvTranspose 512 -> 256)
    Q = Planes(H3, W3, 1024, 1, dev)
    g["up1t.b"] = torch.zeros(256, device=dev)
    ops.call("papr_unet_convt_gather", T2.ptr(), _ref(T2.raster()), 4, 256, Q.ptr(), _ref(Q.raster()), (H2 - 2 * H3) // 2, (W2 - 2 * W3) // 2,
             g["up1t.b"].data_ptr(), nbytes=H2 * W2 * 256 * 4.0)
    g["up1t.w"] = _gemm_wgrad(Q, 1024, X3, 512).reshape(2, 2, 256, 512).permute(3, 2, 0, 1)
    T3 = Planes(H3, W3, 512, 1, dev, zero=False)
    _conv(Q, Q.ptr(), 16, 1, 1, W.bwd["up1t"], None, False, T3)
    dZ = Planes(H3, W3, 512, 3, dev, zero="border")
    g["down2.b"] = torch.zeros(512, device=dev)
    _spread(T3, 0, 8, dst=dZ, mask=X3, colsum=g["down2.b"])
    # down2.conv (256 -> 512) and the pool in front of it
    g["down2.w"] = _conv_wgrad(dZ, P2, 512, 256)
    Tp = Planes(H3, W3, 256, 1, dev, zero=False)
    _conv(dZ, dZ.ptr(), 8, 9, -1, W.bwd["down2"], None, False, Tp)
    dZ = Planes(H2, W2, 256, 3, dev, zero="border")
    g["down1.b"] = torch.zeros(256, device=dev)
    _spread(T2, 0, 4, dst=dZ, pool_grad=Tp, pool_ref=U1, pool_ref_cb0=0, mask=U1, mask_cb0=0, colsum=g["down1.b"])
    # down1.conv (128 -> 256) and its pool
    g["down1.w"] = _conv_wgrad(dZ, P1, 256, 128)
    Tp = Planes(H2, W2, 128, 1, dev, zero=False)
    _conv(dZ, dZ.ptr(), 4, 9, -1, W.bwd["down1"], None, False, Tp)
    dZ = Planes(H, Wd, 128, 3, dev, zero="border")
    g["inc.b"] = torch.zeros(128, device=dev)
    _spread(dSkip1, 0, 2, dst=dZ, pool_grad=Tp, pool_ref=U2, pool_ref_cb0=0, mask=U2, mask_cb0=0, colsum=g["inc.b"])
    # inc (Cin -> 128)
    g["inc.w"] = _conv_wgrad(dZ, P0, 128, Cin)
    d_x = None
    if need_input_grad:
        out = torch.empty((dZ.L, 32), dtype=torch.float32, device=dev)
        _conv(dZ, dZ.ptr(), 2, 9, -1, W.bwd["inc"], None, False, None, out_f32=out)
        d_x = out[: (H + 2) * dZ.Wp].view(H + 2, dZ.Wp, 32)[1:H + 1, 1:Wd + 1, :Cin]
    return d_x, g


class UNetFn(torch.autograd.Function):
    """x (B, C_in, H, W) fp32 [, gamma, beta per image] -> (B, 3, H, W) fp32 through the library's UNet kernels."""

    @staticmethod
    def forward(ctx, x, gamma, beta, grad_enabled, *params):
        p = {}
        for i, (short, _) in enumerate(_NAMES):
            p[short + ".w"], p[short + ".b"] = params[2 * i].detach().float(), params[2 * i + 1].detach().float()
        Bn, Cin, H, Wd = x.shape
        if Cin > 32 or H < 4 or Wd < 4:
            raise NotImplementedError("the UNet kernels take up to 32 input channels and images of at least 4 x 4 pixels")
        keep = grad_enabled and any(ctx.needs_input_grad)
        W = _weights_for(params)
        xh = x.detach().permute(0, 2, 3, 1).contiguous().float()
        outs, saved = [], []
        for b in range(Bn):
            gm = gamma[b if gamma.shape[0] > 1 else 0].detach().float().contiguous() if gamma is not None else None
            bt = beta[b if beta.shape[0] > 1 else 0].detach().float().contiguous() if beta is not None else None
            rgb, sv = unet_forward_image(xh[b], p, W, gm, bt, keep)
            outs.append(rgb)
            saved.append(sv)
        if keep:
            ctx.p, ctx.W, ctx.saved, ctx.Cin = p, W, saved, Cin
            ctx.film = gamma is not None
            ctx.save_for_backward(xh, gamma, beta)
        return torch.stack(outs).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, d_out):
        xh, gamma, beta = ctx.saved_tensors
        d = d_out.permute(0, 2, 3, 1).float()
        Bn = d.shape[0]
        grads, dxs = None, []
        for b in range(Bn):
            dx, g = unet_backward_image(d[b], ctx.p, ctx.W, ctx.saved[b], ctx.Cin, ctx.needs_input_grad[0] or ctx.film)
            dxs.append(dx)
            grads = g if grads is None else {k: grads[k] + g[k] for k in g}
        d_x = d_gamma = d_beta = None
        if dxs[0] is not None:
            dxp = torch.stack(dxs)                                   # gradient w.r.t. the (FiLM-modulated) input
            if ctx.film:                                             # y = x * gamma + beta (unet.py:213-217)
                gam = gamma if gamma.shape[0] > 1 else gamma.expand(Bn, -1)
                d_gamma = (dxp * xh).sum((1, 2))
                d_beta = dxp.sum((1, 2))
                if gamma.shape[0] == 1:
                    d_gamma, d_beta = d_gamma.sum(0, keepdim=True), d_beta.sum(0, keepdim=True)
                dxp = dxp * gam.reshape(Bn, 1, 1, -1)
            d_x = dxp.permute(0, 3, 1, 2)
        flat = []
        for short, _ in _NAMES:
            flat += [grads[short + ".w"], grads[short + ".b"]]
        ctx.saved = None
        return (d_x, d_gamma, d_beta, None, *flat)


_WEIGHT_CACHE = {}


def _weights_for(params):
    """The persistent weight-image set of these parameter tensors, refreshed from their current values."""
    key = tuple(t.data_ptr() for t in params)
    W = _WEIGHT_CACHE.get(key)
    if W is None:
        if len(_WEIGHT_CACHE) > 8:
            _WEIGHT_CACHE.clear()
        W = _WEIGHT_CACHE[key] = _Weights(params)
    W.refresh(params)
    return W


def parameter_list(module):
    """(weight, bias) of every layer of a papr_b200.renderer.SmallUNet in _NAMES order."""
    out = []
    for _, path in _NAMES:
        m = module
        for part in path.split("."):
            m = m[int(part)] if part.isdigit() else getattr(m, part)
        out += [m.weight, m.bias]
    return out
