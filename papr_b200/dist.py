"""Multi-GPU plumbing: one process per GPU (torchrun), rays sharded, parameters replicated.

Training: each rank runs the hot path on its own views/patches; one collective per step -- an all-reduce (mean) of a
single flat fp32 gradient bucket (point positions, features, influence scores, attention and renderer weights) over
NCCL/NVLink, after which every rank applies the identical Adam update.  Rendering: image tiles are sharded with a halo
for the UNet and need no communication.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, init_method="env://", rank=rank, world_size=world)
    return rank, world, local


def trainable_parameters(model):
    return [p for p in model.parameters() if p.requires_grad]


class GradBucket:
    """Flat fp32 gradient bucket: pack -> all_reduce(sum) -> scale by 1/world -> unpack (one collective per step)."""

    def __init__(self, params):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        self.total = sum(self.sizes)
        dev = self.params[0].device
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=dev)

    def matches(self, params):
        params = list(params)
        return len(params) == len(self.params) and all(a is b for a, b in zip(params, self.params))

    def allreduce_mean(self, group=None):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        views = self.flat.split(self.sizes)
        for p, v in zip(self.params, views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad.reshape(-1))
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.mul_(1.0 / world)
        for p, v in zip(self.params, views):
            if p.grad is None:
                p.grad = v.reshape(p.shape).clone()
            else:
                p.grad.copy_(v.reshape(p.shape))
        return self.total * 4


def allreduce_gradients(model, bucket=None):
    """Average gradients across ranks with one flat-bucket all-reduce; returns the (re)usable bucket.

    With the fused optimiser (papr_b200/optim.py) the gradients already ARE views of one flat fp32 bucket: the
    all-reduce (sum) runs on it in place -- no pack / unpack copies -- and the 1/world averaging is folded into the Adam
    launch of the following model.step()."""
    flat = getattr(model, "_flat", None)
    if flat is not None and flat.flat_g is not None:
        flat.bind_grads()
        world = dist.get_world_size() if dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(flat.flat_g, op=dist.ReduceOp.SUM)
            model._grad_scale = 1.0 / world
        return flat
    params = trainable_parameters(model)
    if bucket is None or not bucket.matches(params):
        bucket = GradBucket(params)      # rebuilt after prune/add (parameter tensors are replaced)
    bucket.allreduce_mean()
    return bucket


def shard_rows(H, world, rank, halo=16, align=4):
    """Row stripe [r0, r1) of an H-row image for `rank`, plus the haloed range [h0, h1) to render so that the UNet
    output of the interior equals the full-frame result (measured influence radius 13 px; origins multiples of 4)."""
    per = (H + world - 1) // world
    per = (per + align - 1) // align * align
    r0, r1 = min(rank * per, H), min((rank + 1) * per, H)
    h0 = max(0, (r0 - halo) // align * align)
    h1 = min(H, (r1 + halo + align - 1) // align * align)
    return r0, r1, h0, h1


def broadcast_point_cloud(model, src=0):
    """After prune/add (run on every rank with the same seed, or on rank 0 only), make the clouds identical."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    n = torch.tensor([model.points.shape[0]], device=model.points.device)
    dist.broadcast(n, src)
    names = ["points", "points_influ_scores"] + (["pc_feats"] if model.use_pc_feats else [])
    for name in names:
        p = getattr(model, name)
        if p.shape[0] != int(n):
            p = torch.nn.Parameter(torch.empty((int(n),) + tuple(p.shape[1:]), device=p.device), requires_grad=p.requires_grad)
            setattr(model, name, p)
        dist.broadcast(p.data, src)
