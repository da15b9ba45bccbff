"""Decode stage a10: the reference's SmallUNet (models/unet.py:182-258, built by models/renderer.py:21-34).

Same module tree as the reference so checkpoints interchange (renderer.inc.double_conv.0.*, renderer.down{1,2}.
maxpool_conv.1.double_conv.0.*, renderer.up{1,2}.{up,conv.double_conv.0}.*, renderer.outc.conv.*); the modules below only
OWN the parameters.  On the product path (CUDA, bf16) the forward and backward run on the library's implicit-GEMM tcgen05
convolution kernels (papr_b200/unet.py, csrc/conv.cu, csrc/unet_raster.cu) -- no cuDNN; the fp32 parity mode runs the
same kernels at fp32 accuracy through a three-way bf16 split (papr_b200/unet_fp32.py).  The torch modules are executed
only for FiLM at an inner stage (affine_layer 1..5, which no shipped config uses) in bf16, on the CPU, or with
`own_kernels = False` (A/B measurements).  Only the shipped variant (single conv blocks, transposed-conv upsampling, no
normalisation) exists.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .nn import MLP, make_activation


class ConvBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.double_conv = nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.double_conv(x)


class Down(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), ConvBlock(cin, cout))

    def forward(self, x):
        return self.maxpool_conv(x)


class Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.up = nn.ConvTranspose2d(cin, cin // 2, kernel_size=2, stride=2)
        self.conv = ConvBlock(cin, cout)

    def forward(self, x, skip):
        x = self.up(x)
        dy, dx = skip.shape[2] - x.shape[2], skip.shape[3] - x.shape[3]
        if dy or dx:
            x = F.pad(x, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
        return self.conv(torch.cat([skip, x], dim=1))


class OutConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=1)

    def forward(self, x):
        return self.conv(x)


class SmallUNet(nn.Module):
    def __init__(self, n_channels, n_classes, bilinear=False, single=True, norm="none", last_act="none",
                 affine_layer=-1, compute_dtype=torch.bfloat16):
        super().__init__()
        if bilinear or not single or norm != "none":
            raise NotImplementedError("only the shipped SmallUNet variant (single=True, bilinear=False, norm='none')")
        self.affine_layer = affine_layer
        self.compute_dtype = compute_dtype
        self.own_kernels = True         # False: cuDNN through torch (kept for A/B measurements in tools/, never the default)
        self.inc = ConvBlock(n_channels, 128)
        self.down1 = Down(128, 256)
        self.down2 = Down(256, 512)
        self.up1 = Up(512, 256)
        self.up2 = Up(256, 128)
        self.outc = OutConv(128, n_classes)
        self.last_act = make_activation(last_act)

    @staticmethod
    def _film(x, gamma, beta):
        """x * gamma + beta per channel (unet.py:213-247); gamma / beta (C,) or one pair per image (B, C)."""
        C = x.shape[1]
        assert gamma.shape[-1] == C and beta.shape[-1] == C and gamma.dim() in (1, 2)
        return x * gamma.reshape(-1, C, 1, 1).to(x.dtype) + beta.reshape(-1, C, 1, 1).to(x.dtype)

    def forward(self, x, log=False, gamma=None, beta=None):
        if self.affine_layer >= 0:
            assert gamma is not None and beta is not None
        if x.is_cuda and self.compute_dtype != torch.float32 and self.affine_layer in (-1, 0) and self.own_kernels:
            from . import unet as U
            g = b = None
            if self.affine_layer == 0:
                C = x.shape[1]
                g, b = gamma.reshape(-1, C).float(), beta.reshape(-1, C).float()
            out = U.UNetFn.apply(x, g, b, torch.is_grad_enabled(), *U.parameter_list(self))
            return self.last_act(out)
        if x.is_cuda and self.compute_dtype == torch.float32 and self.own_kernels:
            from .unet_fp32 import unet_forward_fp32      # parity mode: split-bf16 convolutions on the same tensor-core kernels
            return unet_forward_fp32(self, x, gamma, beta)
        amp = x.is_cuda and self.compute_dtype != torch.float32
        with torch.autocast(device_type="cuda", dtype=self.compute_dtype, enabled=amp):
            if amp:
                x = x.contiguous(memory_format=torch.channels_last)
            stages = []
            if self.affine_layer == 0:
                x = self._film(x, gamma, beta)
            x1 = self.inc(x)
            if self.affine_layer == 1:
                x1 = self._film(x1, gamma, beta)
            x2 = self.down1(x1)
            if self.affine_layer == 2:
                x2 = self._film(x2, gamma, beta)
            x3 = self.down2(x2)
            if self.affine_layer == 3:
                x3 = self._film(x3, gamma, beta)
            y = self.up1(x3, x2)
            if self.affine_layer == 4:
                y = self._film(y, gamma, beta)
            y = self.up2(y, x1)
            if self.affine_layer == 5:
                y = self._film(y, gamma, beta)
            out = self.last_act(self.outc(y))
        return out.float()


class MLPGenerator(nn.Module):
    """models/renderer.py:6-17: a per-pixel MLP decode (needs a user-written renderer.generator.mlp block)."""

    def __init__(self, inp_dim, opt, out_dim):
        super().__init__()
        self.mlp = MLP(inp_dim, opt.num_layers, opt.num_channels, out_dim, opt.act_type, opt.last_act_type,
                       opt.skip_layers, opt.use_wn, opt.half_layers, opt.residual_layers)

    def forward(self, x, gamma=None, beta=None):
        return self.mlp(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


def get_generator(args, in_c, out_c, compute_dtype=torch.bfloat16):
    if args.type == "small-unet":
        opt = args.small_unet
        return SmallUNet(in_c, out_c, bilinear=opt.bilinear, single=opt.single, norm=opt.norm, last_act=opt.last_act,
                         affine_layer=opt.affine_layer, compute_dtype=compute_dtype)
    if args.type == "mlp":
        return MLPGenerator(in_c, args.mlp, out_c)     # KeyError('mlp') with the stock configs, as in the reference
    raise NotImplementedError("generator type [{}] is not supported".format(args.type))
