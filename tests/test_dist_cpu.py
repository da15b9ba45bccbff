"""world_size-2 gloo tests of the N>1 host path: flat gradient bucket all-reduce and point-cloud broadcast."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from papr_b200.dist import allreduce_gradients, broadcast_point_cloud, init_from_env
    from papr_b200.config import make_config
    from papr_b200.model import PAPR
    r, w, _ = init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = PAPR(make_config("chair", geoms=dict(points=dict(init_num=64))), device="cpu")
    params = [p for p in model.parameters() if p.requires_grad]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1)) * (i % 3 + 1) if i % 5 else None
    bucket = allreduce_gradients(model)
    ok = True
    for i, p in enumerate(params):
        want = 1.5 * (i % 3 + 1) if i % 5 else 0.0
        ok = ok and bool(torch.allclose(p.grad, torch.full_like(p, want)))
    assert bucket.total == sum(p.numel() for p in params)
    # same bucket object is reused while the parameter tensors are unchanged, rebuilt after a prune
    assert allreduce_gradients(model, bucket) is bucket
    if rank == 0:
        with torch.no_grad():
            model.points_influ_scores.copy_(torch.linspace(-1, 1, 64)[:, None])
        model.prune_points(0.0)
    broadcast_point_cloud(model, src=0)
    ok = ok and model.points.shape[0] == 32 and model.pc_feats.shape[0] == 32
    gathered = [torch.zeros_like(model.points.data) for _ in range(world)]
    dist.all_gather(gathered, model.points.data)
    ok = ok and bool(torch.equal(gathered[0], gathered[1]))
    assert allreduce_gradients(model, bucket) is not bucket
    out[rank] = ok
    dist.destroy_process_group()


def test_gradient_bucket_and_cloud_broadcast_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out[0] and out[1]
