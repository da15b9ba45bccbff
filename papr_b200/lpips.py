"""the LPIPS/VGG16 perceptual loss that follows the path in training (reference models/lpips.py:86-125,
models/__init__.py:27-52) on the library's convolution kernels.

The VGG16 trunk (13 3x3 convolutions + ReLU, four 2x2 max-pools; frozen weights) runs on ``papr_conv_bf16`` in the pixel-plane
layout of the UNet (csrc/conv.cu), forward for both images and the data-gradient chain for the prediction; the LPIPS head
(channel normalisation, squared difference, learnt 1x1 weights, spatial mean; ~0.1% of the work) stays in fp32 torch
elementwise ops so autograd provides its derivative.  Same module tree / state-dict keys as the reference's ``LPNet``
(``net.slice{1..5}.{i}.{weight,bias}``, ``lins.{k}.weight``, ``scaling_layer.{shift,scale}``).

The ImageNet weights cannot be downloaded offline: ``LPNet(pretrained=...)`` takes a state dict (or a torchvision
``vgg16().features`` state dict); without one the trunk is seeded random (a warning says so) -- fine for timing, not for
training quality.  The learnt linear weights are read from the reference's ``vgg.pth`` when it is found.
"""
import math
import os
import warnings

import torch
import torch.nn as nn

from . import ops
from . import unet as U

CFG = [(3, 64), (64, 64), "P", (64, 128), (128, 128), "P", (128, 256), (256, 256), (256, 256), "P",
       (256, 512), (512, 512), (512, 512), "P", (512, 512), (512, 512), (512, 512)]
TAPS = (1, 3, 6, 9, 12)                                  # conv index whose ReLU output is an LPIPS feature (relu1_2 ... relu5_3)
FEATURE_INDEX = (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)
SLICE_OF = {0: 1, 2: 1, 5: 2, 7: 2, 10: 3, 12: 3, 14: 3, 17: 4, 19: 4, 21: 4, 24: 5, 26: 5, 28: 5}


def _trunk_forward(x_hwc, W, biases, keep):
    """One image (H, W, 3) fp32, already scaled -> (five fp32 feature maps (h, w, C), saved planes or None)."""
    H, Wd, _ = x_hwc.shape
    dev = x_hwc.device
    cur = U.Planes(H, Wd, 64, 3, dev)
    ops.call("papr_unet_pack_input", x_hwc.data_ptr(), x_hwc.stride(1), 3, None, None, cur.ptr(), U._ref(cur.raster()), 3, 1,
             nbytes=H * Wd * 400.0)
    feats, saved, ci = [], [], 0
    for item in CFG:
        if item == "P":
            nxt = U.Planes(cur.H // 2, cur.W // 2, cur.C, 3, dev)
            U._pool(cur, 0, cur.cbs, nxt)
            cur = nxt
            continue
        cin, cout = item
        T = U.Planes(cur.H, cur.W, cout, 1, dev, zero=False)
        U._conv(cur, cur.ptr(), cur.cbs, 9, 1, W.fwd[ci], biases[ci], True, T)
        out = U.Planes(cur.H, cur.W, cout, 3, dev)
        U._spread(T, 0, out.cbs, dst=out)
        if ci in TAPS:
            f = torch.empty((cur.H, cur.W, cout), dtype=torch.float32, device=dev)
            ops.call("papr_unet_unpack", out.mid(), U._ref(out.raster()), out.cbs, f.data_ptr(), f.stride(1), cout,
                     nbytes=cur.H * cur.W * cout * 6.0)
            feats.append(f)
        if keep:
            saved.append(out)
        cur = out
        ci += 1
    return feats, (saved if keep else None)


def _trunk_backward(d_feats, W, saved):
    """Gradients w.r.t. the five feature maps (fp32 (h, w, C)) -> gradient w.r.t. the scaled input image (H, W, 3)."""
    dev = d_feats[0].device
    order = [i for i in CFG]
    conv_pos = [i for i, it in enumerate(order) if it != "P"]
    g = None                     # gradient planes (3 copies) w.r.t. the pre-activation of conv `ci`, produced top-down
    pending_pool = None          # gradient w.r.t. a pooled map (unshifted), to be routed into the conv output below it
    for ci in range(12, -1, -1):
        out = saved[ci]
        cout = out.C
        src = None
        if ci in TAPS:
            d = d_feats[TAPS.index(ci)].contiguous()
            src = U.Planes(out.H, out.W, cout, 1, dev)
            ops.call("papr_unet_pack_input", d.data_ptr(), d.stride(1), cout, None, None, src.ptr(), U._ref(src.raster()), 1, src.cbs,
                     nbytes=out.H * out.W * cout * 6.0)
        dz = U.Planes(out.H, out.W, cout, 3, dev)
        if g is not None and pending_pool is None:            # plain chain: gradient from the conv above, same resolution
            if src is None:
                U._spread(g, 0, out.cbs, dst=dz, mask=out)
            else:
                U._spread(src, 0, out.cbs, dst=dz, add=g, mask=out)
        elif pending_pool is not None:                        # a pool sits above this conv
            if src is None:
                src = U.Planes(out.H, out.W, cout, 1, dev)    # zeros
            U._spread(src, 0, out.cbs, dst=dz, pool_grad=pending_pool, pool_ref=out, mask=out)
            pending_pool = None
        else:                                                 # the top of the trunk
            U._spread(src, 0, out.cbs, dst=dz, mask=out)
        cin = CFG[conv_pos[ci]][0]
        if ci == 0:
            res = torch.empty((dz.L, 32), dtype=torch.float32, device=dev)
            U._conv(dz, dz.ptr(), dz.cbs, 9, -1, W.bwd[0], None, False, None, out_f32=res)
            return res[: (dz.H + 2) * dz.Wp].view(dz.H + 2, dz.Wp, 32)[1:dz.H + 1, 1:dz.W + 1, :3]
        gin = U.Planes(dz.H, dz.W, cin, 1, dev, zero=False)
        U._conv(dz, dz.ptr(), dz.cbs, 9, -1, W.bwd[ci], None, False, gin)
        if conv_pos[ci] > 0 and order[conv_pos[ci] - 1] == "P":
            pending_pool, g = gin, None                       # gin is the gradient w.r.t. the pooled map
        else:
            g = gin


class _TrunkWeights:
    """bf16 weight images of the 13 frozen convolutions (forward, and transposed for the data gradient): packed once."""

    def __init__(self, weights):
        descs, self.keep, self.fwd, self.bwd = [], [], [], []
        for w in weights:
            co, ci = w.shape[:2]
            cip, cop = (ci + 63) // 64 * 64, (co + 63) // 64 * 64
            for table, mat in ((self.fwd, U._pad_k(w.permute(0, 2, 3, 1), cip).reshape(co, 9 * cip)),
                               (self.bwd, U._pad_k(w.permute(1, 2, 3, 0), cop).reshape(ci, 9 * cop))):
                tiles, d, keepalive = U._pack_matrix(mat.float())
                table.append(tiles)
                descs.extend(d)
                self.keep.append(keepalive)
        self.table = U._launch_pack(descs, weights[0].device)


class _TrunkFn(torch.autograd.Function):
    """scaled image batch (B, H, W, 3) -> five feature maps (B, h, w, C) each; differentiable w.r.t. the images."""

    @staticmethod
    def forward(ctx, x, module, grad_enabled):
        W, biases = module._images()
        keep = grad_enabled and ctx.needs_input_grad[0]
        xs = x.detach().float().contiguous()
        per_image, saved = [], []
        for b in range(xs.shape[0]):
            f, s = _trunk_forward(xs[b], W, biases, keep)
            per_image.append(f)
            saved.append(s)
        if keep:
            ctx.W, ctx.saved = W, saved
        return tuple(torch.stack([per_image[b][k] for b in range(xs.shape[0])]) for k in range(5))

    @staticmethod
    def backward(ctx, *d_feats):
        outs = []
        for b in range(len(ctx.saved)):
            outs.append(_trunk_backward([d[b].float() for d in d_feats], ctx.W, ctx.saved[b]))
        ctx.saved = None
        return torch.stack(outs), None, None


class _Slice(nn.Module):
    pass


class _VGG16(nn.Module):
    """Parameter container with torchvision's numbering inside the reference's five slices (models/lpips.py:8-28)."""

    def __init__(self):
        super().__init__()
        for s in range(1, 6):
            setattr(self, f"slice{s}", nn.Module())
        ci = 0
        for item in CFG:
            if item == "P":
                continue
            idx = FEATURE_INDEX[ci]
            conv = nn.Conv2d(item[0], item[1], 3, padding=1)
            getattr(self, f"slice{SLICE_OF[idx]}").add_module(str(idx), conv)
            ci += 1
        for p in self.parameters():
            p.requires_grad = False

    def convs(self):
        return [getattr(getattr(self, f"slice{SLICE_OF[i]}"), str(i)) for i in FEATURE_INDEX]


class NetLinLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(1, c, 1, 1))


class LPNet(nn.Module):
    """reference models/lpips.py:86-125.  forward(in0, in1): (N, H, W, 3) images in [0, 1] -> scalar LPIPS distance."""

    def __init__(self, pretrained=None, lin_path=None, seed=0):
        super().__init__()
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188])[None, None, None, :])
        self.register_buffer("scale", torch.tensor([.458, .448, .450])[None, None, None, :])
        self.net = _VGG16()
        self.L = 5
        self.lins = nn.ModuleList([NetLinLayer(c) for c in (64, 128, 256, 512, 512)])
        self.random_trunk = pretrained is None
        if pretrained is None:
            warnings.warn("LPNet: no VGG16 weights given (ImageNet weights are not available offline): seeded random trunk")
            g = torch.Generator().manual_seed(seed)
            with torch.no_grad():
                for conv in self.net.convs():
                    conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * math.sqrt(2.0 / (conv.weight.shape[1] * 9)))
                    conv.bias.copy_(0.05 * torch.randn(conv.bias.shape, generator=g))
        else:
            self.load_trunk(pretrained)
        lin_path = lin_path or os.path.abspath(os.path.join(".", "vgg.pth"))          # where the reference looks for it
        if os.path.exists(lin_path):
            w = torch.load(lin_path, map_location="cpu")
            with torch.no_grad():
                for k in range(5):
                    self.lins[k].weight.copy_(w[f"lin{k}.model.1.weight"])
        for p in self.parameters():
            p.requires_grad = False
        self._cache = None

    def load_trunk(self, sd):
        """`sd`: the reference's `net.slice*.N.*` keys, or torchvision's `features.N.*` / `N.*` keys."""
        with torch.no_grad():
            for idx, conv in zip(FEATURE_INDEX, self.net.convs()):
                for cand in (f"net.slice{SLICE_OF[idx]}.{idx}", f"slice{SLICE_OF[idx]}.{idx}", f"features.{idx}", str(idx)):
                    if cand + ".weight" in sd:
                        conv.weight.copy_(sd[cand + ".weight"])
                        conv.bias.copy_(sd[cand + ".bias"])
                        break
                else:
                    raise KeyError(f"no weights for VGG16 features.{idx}")
        self._cache = None

    def _images(self):
        convs = self.net.convs()
        key = tuple((c.weight.data_ptr(), c.weight._version) for c in convs)
        if self._cache is None or self._cache[0] != key:
            self._cache = (key, _TrunkWeights([c.weight.detach() for c in convs]), [c.bias.detach().float() for c in convs])
        return self._cache[1], self._cache[2]

    def forward(self, in0, in1):
        if not in0.is_cuda:
            raise RuntimeError("papr_b200.lpips needs CUDA tensors: there is no CPU path")
        with torch.cuda.device(in0.device):
            a = ((2 * in0 - 1) - self.shift) / self.scale
            b = ((2 * in1 - 1) - self.shift) / self.scale
            fa = _TrunkFn.apply(a, self, torch.is_grad_enabled())
            with torch.no_grad():
                fb = _TrunkFn.apply(b.detach(), self, False)
            val = 0
            for k in range(5):
                na = fa[k] / (torch.sqrt(torch.sum(fa[k] ** 2, dim=-1, keepdim=True) + 1e-10) + 1e-10)
                nb = fb[k] / (torch.sqrt(torch.sum(fb[k] ** 2, dim=-1, keepdim=True) + 1e-10) + 1e-10)
                d = (na - nb) ** 2
                val = val + torch.sum(self.lins[k].weight.reshape(1, 1, 1, -1) * d, -1).mean((1, 2))
            return val.mean()


class BasicLoss(nn.Module):
    """reference models/__init__.py:8-20: weighted sum; keys are "name/weight" as in the reference."""

    def __init__(self, losses_and_weights):
        super().__init__()
        self.losses_and_weights = losses_and_weights

    def forward(self, pred, target):
        loss = 0
        for name_and_weight, loss_func in self.losses_and_weights.items():
            _, weight = name_and_weight.split("/")
            loss = loss + float(weight) * loss_func(pred, target)
        return loss


def get_loss(args, vgg_weights=None, lin_path=None):
    """reference models/__init__.py:27-52 for the losses the shipped configs use (mse, l1, lpips)."""
    losses = nn.ModuleDict()
    for name, weight in args.items():
        if weight <= 0:
            continue
        key = name + "/" + str(format(weight, ".0e"))
        if name == "mse":
            losses[key] = nn.MSELoss()
        elif name == "l1":
            losses[key] = nn.L1Loss()
        elif name == "lpips":
            losses[key] = LPNet(vgg_weights, lin_path).eval()
        else:
            raise NotImplementedError(f"loss [{name}] is not supported")
    return BasicLoss(losses)
