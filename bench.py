#!/usr/bin/env python
"""Headline benchmark of the PAPR hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours      : one training step = clear_grad -> PAPR.forward (select, proximity attention, UNet, composite) -> MSE ->
            backward -> (N>1: flat-bucket gradient all-reduce) -> Adam, on BASELINE.json configs[1]: chair.yml
            hyper-parameters, one full 800x800 frame of rays per GPU per step, 30,000 points, K=20, bf16 tcgen05 GEMMs.
            Weak scaling: every rank trains on its own view; `value` = total rays / max-over-ranks step time.
reference : the CPU oracle (oracle/papr_oracle.py, a pinned restatement of the reference's own PyTorch CPU path) on a
            bounded sample of the same workload with all host threads (rank 0 only).
Prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train rays/sec (fwd+bwd), whole job"
UNIT = "rays/s"
FLOP_PER_RAY_TRAIN = 102.5e6     # 34.18 Mflop forward x 3
FLOP_PER_RAY_FWD = 34.18e6


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def oracle_train_sample(hw, P, threads, steps, warmup):
    """Times the CPU oracle (fwd + bwd of the same loss) on an hw x hw tile of the workload.  Returns rays/s."""
    from oracle import papr_oracle as O
    from papr_b200.config import make_config
    torch.set_num_threads(threads)
    cfg = make_config("chair", use_amp=False)
    params = O.init_params(cfg, P, seed=1, cloud="shell")
    rays_o, rays_d, _ = O.synthetic_rays(800, 800, cfg.dataset.coord_scale, n_views=1, seed=1, h0=380, h1=380 + hw, w0=380, w1=380 + hw)
    tgt = torch.rand(1, hw, hw, 3, generator=torch.Generator().manual_seed(3))
    times = []
    for it in range(warmup + steps):
        pg = {k: v.clone().requires_grad_(v.dtype.is_floating_point and k != "bkg_feats") for k, v in params.items()}
        t0 = time.perf_counter()
        out = O.forward(pg, cfg, rays_o, rays_d)
        loss = ((out["rgb"] - tgt) ** 2).mean()
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return hw * hw / statistics.median(times), statistics.median(times)


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    hw = args.cpu_tile
    rays_s, sec = oracle_train_sample(hw, args.points, threads, max(1, args.steps), min(args.warmup, 1))
    sample = f"{hw}x{hw}-ray tile of the 800x800 frame, P={args.points}, fwd+bwd (MSE), fp32, oracle port"
    line = {
        "impl": "reference", "metric": METRIC, "value": rays_s, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(hw=args.hw, P=args.points), "sample": sample},
        "cpu_baseline": {"value": rays_s, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rays_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


WORKLOAD = "chair.yml train step, one {hw}x{hw} frame of rays per GPU, P={P} points, K=20, F=64, L=6, bf16 tcgen05 GEMMs"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hw", type=int, default=800)
    ap.add_argument("--points", type=int, default=30000)
    ap.add_argument("--cpu-tile", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return

    import torch.distributed as dist
    from papr_b200 import ops
    from papr_b200.config import make_config
    from papr_b200.dist import allreduce_gradients, init_from_env, shard_rows
    from papr_b200.model import PAPR
    from papr_b200.scene import learned_like_cloud, synthetic_scene

    rank, world, local = init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    steps, warmup = args.steps, max(args.warmup, 3)
    H = W = args.hw

    torch.manual_seed(1)
    cfg = make_config("chair")
    cfg.geoms.points["init_num"] = args.points
    model = PAPR(cfg, device=dev, precision="bf16").to(dev)
    cloud = learned_like_cloud(args.points, cfg.dataset.coord_scale, seed=1, feat_dim=cfg.geoms.point_feats.dim)
    with torch.no_grad():
        model.points.copy_(cloud["points"]); model.pc_feats.copy_(cloud["pc_feats"])
        model.points_influ_scores.copy_(cloud["points_influ_scores"])
    if world > 1:   # replicas must start identical
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    model.init_optimizers(0)

    scene = synthetic_scene(H, W, cfg.dataset.coord_scale, n_views=1, seed=1 + rank)
    host = {k: scene[k].pin_memory() for k in ("rays_o", "rays_d", "c2w", "target")}
    resident = {k: v.to(dev) for k, v in host.items()}
    bucket = [None]

    def train_step(b):
        model.clear_grad()
        out = model(b["rays_o"], b["rays_d"], b["c2w"], step=-1)
        loss = torch.mean((model.last_act(out) - b["target"]) ** 2)
        model.scaler.scale(loss).backward()
        if world > 1:
            bucket[0] = allreduce_gradients(model, bucket[0])
        model.step()
        return loss

    def timed(fn, n):
        """n calls between barrier + synchronize on both sides, CUDA events; returns max-over-ranks ms per call."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(warmup):
        train_step(resident)
    torch.cuda.synchronize()

    # ---- timed region: device-resident inputs, per-kernel CUDA events on the launching stream
    sampler = ClockSampler(local)
    sampler.start()
    ops.STATS.reset()
    ops.STATS.timing = True
    ms_step = timed(lambda: train_step(resident), steps)
    ops.STATS.timing = False
    clocks = sampler.stop()
    launches = ops.STATS.count
    kern = ops.STATS.summary()
    rays_per_step = H * W * world
    value = rays_per_step / ms_step * 1e3

    # ---- end to end: host (pinned) inputs copied in, loss read back, every step
    def e2e_step():
        b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        return train_step(b).item()
    e2e_step()
    ms_e2e = timed(e2e_step, steps)
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    # ---- render: forward only, rows sharded across ranks with a UNet halo, no communication
    r0, r1, h0, h1 = shard_rows(H, world, rank)
    def render():
        with torch.no_grad():
            rgb = model(resident["rays_o"], resident["rays_d"][:, h0:h1].contiguous(), resident["c2w"], step=-1)
            return rgb[:, r0 - h0:r1 - h0]
    for _ in range(2):
        render()
    ops.STATS.reset(); ops.STATS.timing = True
    ms_render = timed(render, steps)
    ops.STATS.timing = False
    kern_render = ops.STATS.summary()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    zero = dict(launches=0, ms=0.0, flops=0.0, bytes=0.0)
    tc_names = ("papr_stack_bf16", "papr_linear_bf16", "papr_wgrad_bf16")
    tc = [kern.get(n, zero) for n in tc_names]
    sel = kern.get("papr_select_topk", zero)
    tc_ms = max(sum(k["ms"] for k in tc), 1e-9)
    tc_flops = sum(k["flops"] for k in tc)
    tc_bytes = sum(k["bytes"] for k in tc)
    tc_launches = sum(k["launches"] for k in tc)
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12
    stack = kern.get("papr_stack_bf16", zero)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram bytes per launch from the committed ncu capture
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("tcgen05_kernels_bytes_per_step")
    roofline = {
        "bound": "tensor",
        "kernel": "tcgen05 GEMM kernels of the key/value/query stacks: stack_kernel (fused fwd + dgrad, cta_group::2), wgrad_kernel, linear_kernel",
        "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"],
        "peak_source": peaks["source"] + " bf16 sustained (kernels timed inside a long step)",
        "traffic": traffic, "traffic_unit": "bytes per step over these launches (ncu dram__bytes_read+write)",
        "algorithmic_bytes_per_step": tc_bytes / steps,
        "launches_per_step": tc_launches / steps, "ms_per_step": tc_ms / steps, "share_of_step": tc_ms / steps / ms_step,
        "hbm_gbs": tc_bytes / (tc_ms * 1e-3) / 1e9, "hbm_frac": tc_bytes / (tc_ms * 1e-3) / 1e9 / peaks["hbm"],
        "by_kernel": {n: {"ms_per_step": k["ms"] / steps, "tflops": k["flops"] / max(k["ms"], 1e-9) / 1e9,
                          "tensor_frac": k["flops"] / max(k["ms"], 1e-9) / 1e9 / peaks["tf_sustained"],
                          "hbm_gbs": k["bytes"] / max(k["ms"], 1e-9) / 1e6, "hbm_frac": k["bytes"] / max(k["ms"], 1e-9) / 1e6 / peaks["hbm"]}
                      for n, k in zip(tc_names, tc) if k["launches"]},
    }
    kernels = {k: dict(launches_per_step=v["launches"] / steps, ms_per_step=v["ms"] / steps,
                       tflops=v["flops"] / max(v["ms"], 1e-9) / 1e9, gbs=v["bytes"] / max(v["ms"], 1e-9) / 1e6)
               for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD.format(hw=H, P=args.points), "rays_per_step_per_gpu": H * W,
                   "l2": "inputs and activations (GBs per step) exceed the 126 MB L2; no explicit flush",
                   "optimizer_in_step": True},
        "clocks": clocks,
        "e2e": {"value": rays_per_step / ms_e2e * 1e3, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": roofline,
        "step_roofline": {"algorithmic_tflop_per_step": rays_per_step / world * FLOP_PER_RAY_TRAIN / 1e12,
                          "achieved_tflops_per_gpu": rays_per_step / world * FLOP_PER_RAY_TRAIN / (ms_step * 1e-3) / 1e12,
                          "frac_of_bf16_sustained": rays_per_step / world * FLOP_PER_RAY_TRAIN / (ms_step * 1e-3) / 1e12 / peaks["tf_sustained"]},
        "select": {"ms_per_step": sel["ms"] / steps, "pairs_per_s": sel["flops"] / 17.0 / max(sel["ms"], 1e-9) * 1e3,
                   "fp32_tflops_17flop": sel["flops"] / max(sel["ms"], 1e-9) / 1e9,
                   "algorithmic_hbm_gbs": sel["bytes"] / max(sel["ms"], 1e-9) / 1e6},
        "render": {"ms_per_frame": ms_render, "frame": f"{H}x{W}", "rows_per_gpu": h1 - h0,
                   "frac_of_gemm_floor": (H * W * FLOP_PER_RAY_FWD / world / (peaks["tf_sustained"] * 1e12) * 1e3) / ms_render,
                   "kernels_ms": {k: v["ms"] / steps for k, v in kern_render.items()}},
        "kernels": kernels,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rays_s, sec = oracle_train_sample(args.cpu_tile, args.points, threads, 2, 1)
        line["cpu_baseline"] = {"value": rays_s, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{args.cpu_tile}x{args.cpu_tile}-ray tile of the frame, P={args.points}, fwd+bwd, "
                                          f"fp32 oracle port, {sec:.2f} s/step"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
