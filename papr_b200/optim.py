"""the optimiser step of ``PAPR.step`` (reference models/model.py:117-179, 439-446) on the device
in one launch.

The reference keeps one ``torch.optim.Adam`` per parameter group (points, attention, influence scores, features,
renderer, mapping MLP, background) and steps them one after the other, each through a handful of foreach kernels.
Here every group still has its own optimiser object with the Adam state-dict layout (``model.optimizers[name]`` is
part of the surface train.py touches; checkpoints interchange with ``torch.optim.Adam``), but they share ONE flat fp32
gradient bucket and ONE pair of flat moment buffers, and ``FlatAdamBucket.step_all`` updates all of them with a single
``papr_adam_step`` launch.  Gradients are views into the bucket, so

* ``clear_grad`` is one memset,
* the multi-GPU all-reduce runs on the bucket itself (no pack / unpack copies), and its 1/world
  averaging is folded into the Adam launch as ``grad_scale``.
"""
import ctypes

import torch

from . import ops
from ._lib import AdamGroup


class FlatAdamBucket:
    """Flat gradient / moment storage shared by the FlatAdam optimisers of one model."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.optimizers = []          # FlatAdam objects, index = group id
        self.params = []              # every parameter, bucket order
        self.group_of = []
        self.offsets = [0]
        self.flat_g = self.flat_m = self.flat_v = None
        self._tables = None
        self._ptrs = None

    # -- construction ---------------------------------------------------------------------------------------------
    def add(self, optimizer):
        gid = len(self.optimizers)
        self.optimizers.append(optimizer)
        for group in optimizer.param_groups:
            for p in group["params"]:
                if p.dtype != torch.float32:
                    raise TypeError("FlatAdam handles fp32 parameters")
                self.params.append(p)
                self.group_of.append(gid)
                self.offsets.append(self.offsets[-1] + (p.numel() + 3) // 4 * 4)       # 16-byte aligned segments
        return gid

    def finalize(self):
        total = self.offsets[-1]
        dev = self.params[0].device if self.params else self.device
        self.device = dev
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self._tables = None
        self.bind_grads(keep_values=True)

    def view(self, flat, i):
        p = self.params[i]
        return flat[self.offsets[i]:self.offsets[i] + p.numel()].view(p.shape)

    # -- gradients ------------------------------------------------------------------------------------------------
    def _follow_device(self):
        if self.params and self.params[0].device != self.flat_g.device:       # model.to(device) after construction
            self.flat_g, self.flat_m, self.flat_v = (t.to(self.params[0].device) for t in (self.flat_g, self.flat_m, self.flat_v))
            self.device = self.flat_g.device
            self._tables = None
            for opt in self.optimizers:
                opt._rebind_state()

    def bind_grads(self, keep_values=True):
        """Make every parameter's .grad the matching view of the flat bucket (autograd then accumulates in place).
        keep_values: a gradient autograd allocated elsewhere is copied into its view first; a missing one reads as zero."""
        if self.flat_g is None:
            return
        self._follow_device()
        base = self.flat_g.data_ptr()
        for i, p in enumerate(self.params):
            if p.grad is not None and p.grad.data_ptr() == base + self.offsets[i] * 4:
                continue
            v = self.view(self.flat_g, i)
            if keep_values:
                if p.grad is None:
                    v.zero_()
                else:
                    v.copy_(p.grad)
            p.grad = v

    def zero_grad(self):
        self.bind_grads(keep_values=False)
        self.flat_g.zero_()

    # -- step -----------------------------------------------------------------------------------------------------
    def _device_tables(self):
        ptrs = [p.data_ptr() for p in self.params]
        if self._tables is None or ptrs != self._ptrs:
            dev = self.flat_g.device
            self._ptrs = ptrs
            self._tables = (torch.tensor(self.offsets, dtype=torch.int64, device=dev),
                            torch.tensor(ptrs, dtype=torch.int64, device=dev),
                            torch.tensor(self.group_of, dtype=torch.int32, device=dev))
        return self._tables

    def step_groups(self, enabled, grad_scale=1.0):
        """One launch: advance the groups whose index is in `enabled`."""
        if not self.params:
            return
        if not self.flat_g.is_cuda:
            raise RuntimeError("FlatAdam needs CUDA parameters: there is no CPU path")
        self.bind_grads()           # gradients autograd allocated elsewhere (e.g. after zero_grad(set_to_none)) move in
        arr = (AdamGroup * len(self.optimizers))()
        for gid, opt in enumerate(self.optimizers):
            g = opt.param_groups[0]
            on = gid in enabled
            if on:
                opt._step_count_adam += 1
                opt._touch_state()
            arr[gid].lr, arr[gid].beta1, arr[gid].beta2 = float(g["lr"]), float(g["betas"][0]), float(g["betas"][1])
            arr[gid].eps, arr[gid].weight_decay = float(g["eps"]), float(g["weight_decay"])
            arr[gid].step, arr[gid].enabled = int(opt._step_count_adam), int(on)
        offs, ptrs, grp = self._device_tables()
        with torch.cuda.device(self.flat_g.device):
            ops.call("papr_adam_step", offs.data_ptr(), ptrs.data_ptr(), grp.data_ptr(), len(self.params), self.offsets[-1],
                     self.flat_g.data_ptr(), self.flat_m.data_ptr(), self.flat_v.data_ptr(),
                     ctypes.cast(arr, ctypes.c_void_p), len(self.optimizers), float(grad_scale),
                     nbytes=28.0 * self.offsets[-1])
        for p in self.params:                       # the kernel wrote through raw pointers: tell autograd's version counter
            torch.autograd.graph.increment_version(p)

    def step_all(self, optimizers, grad_scale=1.0):
        self.step_groups({gid for gid, opt in enumerate(self.optimizers) if any(opt is o for o in optimizers)}, grad_scale)


class FlatAdam(torch.optim.Optimizer):
    """One parameter group of the model with torch.optim.Adam's interface and state-dict layout; the arithmetic runs in
    FlatAdamBucket's single launch."""

    def __init__(self, params, bucket, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdam takes one parameter group")
        self.bucket = bucket
        self._step_count_adam = 0
        self.gid = bucket.add(self)

    def _indices(self):
        if getattr(self, "_idx_cache", None) is None or self._idx_cache[0] != len(self.bucket.group_of):
            self._idx_cache = (len(self.bucket.group_of), [i for i, g in enumerate(self.bucket.group_of) if g == self.gid])
        return self._idx_cache[1]

    def _rebind_state(self):
        if not self.state:
            return
        for i in self._indices():
            p = self.bucket.params[i]
            st = self.state[p]
            st["exp_avg"], st["exp_avg_sq"] = self.bucket.view(self.bucket.flat_m, i), self.bucket.view(self.bucket.flat_v, i)

    def _touch_state(self):
        """Expose the Adam state in torch's layout (created at the first step, as torch does); the step counter is one
        tensor shared by the group's parameters."""
        if getattr(self, "_step_tensor", None) is None:
            self._step_tensor = torch.tensor(0.0, dtype=torch.float32)
        if len(self.state) < len(self._indices()):
            for i in self._indices():
                st = self.state[self.bucket.params[i]]
                if not st:
                    st["step"] = self._step_tensor
                    st["exp_avg"], st["exp_avg_sq"] = self.bucket.view(self.bucket.flat_m, i), self.bucket.view(self.bucket.flat_v, i)
        self._step_tensor.fill_(float(self._step_count_adam))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self.bucket.step_groups({self.gid})
        return loss

    def zero_grad(self, set_to_none=True):
        self.bucket.bind_grads(keep_values=False)
        for i in self._indices():
            self.bucket.view(self.bucket.flat_g, i).zero_()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = []
        for i in self._indices():
            p = self.bucket.params[i]
            st = self.state.get(p)
            if not st:
                continue
            self.bucket.view(self.bucket.flat_m, i).copy_(st["exp_avg"])
            self.bucket.view(self.bucket.flat_v, i).copy_(st["exp_avg_sq"])
            steps.append(int(float(st["step"])))
        self._step_count_adam = max(steps) if steps else 0
        self._step_tensor = torch.tensor(float(self._step_count_adam), dtype=torch.float32)
        for i in self._indices():
            st = self.state.get(self.bucket.params[i])
            if st:
                st["step"] = self._step_tensor
        self._rebind_state()
