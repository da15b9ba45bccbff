"""GPU, 2 ranks over NCCL: two views sharded over two ranks give the single-process two-view gradients (SURVEY section 4:
"N-GPU sharded grads == 1-GPU grads"), through the flat gradient bucket that the fused optimiser steps from, and the
parameters stay identical on both ranks after the step.  Needs 2 GPUs (skipped otherwise)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(rank):
    sys.path.insert(0, ROOT)
    from papr_b200.config import make_config
    from papr_b200.model import PAPR
    from papr_b200.scene import learned_like_cloud, synthetic_scene
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    cfg = make_config("chair", geoms=dict(points=dict(init_num=2000)),
                      training=dict(lr=dict(attn=dict(warmup=0), generator=dict(warmup=0), feats=dict(warmup=0), points_influ_scores=dict(warmup=0))))
    torch.manual_seed(7)
    model = PAPR(cfg, device=dev).to(dev)
    cloud = learned_like_cloud(2000, cfg.dataset.coord_scale, seed=1)
    with torch.no_grad():
        model.points.copy_(cloud["points"]); model.pc_feats.copy_(cloud["pc_feats"]); model.points_influ_scores.copy_(cloud["points_influ_scores"])
    model.init_optimizers(0)
    scene = {k: v.to(dev) for k, v in synthetic_scene(48, 64, cfg.dataset.coord_scale, n_views=2, seed=3).items()}
    return model, scene


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    model, scene = _setup(rank)
    from papr_b200.dist import allreduce_gradients, init_from_env
    init_from_env("nccl")
    for p in model.parameters():
        dist.broadcast(p.data, 0)
    model.clear_grad()
    rgb = model(scene["rays_o"][rank:rank + 1], scene["rays_d"][rank:rank + 1], scene["c2w"][rank:rank + 1])
    torch.mean((rgb - scene["target"][rank:rank + 1]) ** 2).backward()
    allreduce_gradients(model)
    grads = {n: (p.grad * model._grad_scale).detach().cpu() for n, p in model.named_parameters() if p.grad is not None}
    model.step(0)
    params = {n: p.detach().cpu() for n, p in model.named_parameters()}
    out[rank] = (grads, params)
    dist.destroy_process_group()


def test_two_rank_nccl_gradients_equal_single_process_two_views():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    (g0, p0), (g1, p1) = out[0], out[1]
    for n in p0:
        assert torch.equal(p0[n], p1[n]), f"replicas diverged: {n}"
    for n in g0:
        assert torch.equal(g0[n], g1[n]), n
    model, scene = _setup(0)
    model.clear_grad()
    rgb = model(scene["rays_o"], scene["rays_d"], scene["c2w"])
    torch.mean((rgb - scene["target"]) ** 2).backward()
    worst = 0.0
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        want, got = p.grad.detach().cpu(), g0[n]
        scale = max(float(want.abs().max()), 1e-12)
        e = float((got - want).abs().max()) / scale
        worst = max(worst, e)
        assert e <= 3e-2, (n, e)          # fp32 atomics + bf16 column sums in a different order
    print(f"2-rank NCCL gradients vs single process: worst max-abs/max {worst:.2e}")
