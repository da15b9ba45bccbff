"""TEST INFRASTRUCTURE.  Pins oracle.papr_oracle.lpips against the REAL reference LPNet.forward (models/lpips.py:86-125)
and writes tests/golden/lpips_2x32x48.npz.  Run in a container that has /root/reference:  python oracle/make_golden_lpips.py

The reference's LPNet.__init__ downloads ImageNet VGG16 weights (impossible offline), so the object is assembled by hand
from the reference's own classes (vgg16(pretrained=False), ScalingLayer, NetLinLayer) and filled with the oracle's seeded
VGG weights + the reference's shipped vgg.pth linear weights; its unmodified forward() is what gets recorded."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import papr_oracle as O  # noqa: E402

REF = "/root/reference"


def main():
    sys.modules.setdefault("lpips", types.ModuleType("lpips"))
    spec = importlib.util.spec_from_file_location("ref_lpips", os.path.join(REF, "models", "lpips.py"))
    R = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(R)
    P = O.init_lpips_params(seed=3, lin_path=os.path.join(REF, "vgg.pth"))
    obj = R.LPNet.__new__(R.LPNet)
    nn.Module.__init__(obj)
    obj.scaling_layer, obj.net, obj.L = R.ScalingLayer(), R.vgg16(pretrained=False, requires_grad=False), 5
    obj.lins = nn.ModuleList([R.NetLinLayer() for _ in range(5)])
    for k in range(5):
        obj.lins[k].weight = nn.Parameter(P[f"lins.{k}.weight"].clone())
    sd = obj.net.state_dict()
    for k in sd:
        sd[k].copy_(P["net." + k])
    g = torch.Generator().manual_seed(0)
    in0 = torch.rand(2, 32, 48, 3, generator=g, requires_grad=True)
    in1 = torch.rand(2, 32, 48, 3, generator=g)
    ref = obj.forward(in0, in1)
    ref.backward()
    gref = in0.grad.clone()
    in0.grad = None
    mine = O.lpips(P, in0, in1)
    mine.backward()
    assert abs(float(ref.detach()) - float(mine.detach())) <= 1e-7 and float((in0.grad - gref).abs().max()) <= 1e-9, "oracle.lpips differs from the reference"
    out = os.path.join(ROOT, "tests", "golden", "lpips_2x32x48.npz")
    np.savez_compressed(out, in0=in0.detach().numpy(), in1=in1.numpy(), loss=np.float64(float(ref.detach())), grad_in0=gref.numpy(),
                        lins=np.concatenate([P[f"lins.{k}.weight"].reshape(-1).numpy() for k in range(5)]), seed=np.int64(3))
    print("wrote", out, "loss", float(ref.detach()))


if __name__ == "__main__":
    main()
