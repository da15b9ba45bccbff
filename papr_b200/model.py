"""B200-native drop-in for the reference's ``models.PAPR`` (models/model.py:17-641).

Same constructor, attributes, methods and state_dict keys as the reference model, so train.py / test.py /
exposure_control_finetune.py keep working against it; the hot path underneath is the library's
CUDA kernels (papr_b200/csrc) reached through the C ABI of include/papr_b200.h:

  _get_points            -> papr_select_topk                      (model.py:258-283, 312-333)
  _get_kqv + attention   -> papr_b200.attention.ProximityAttention (model.py:285-310, 396-437; attn.py)
  blend                  -> papr_score_blend_fwd / papr_blend_bwd   (model.py:519-534)
  renderer               -> papr_b200.renderer.SmallUNet (cuDNN)    (unet.py:208-258)

There is no CPU path: constructing the model on a non-CUDA device works (parameters, optimisers, checkpoints), but
forward/evaluate raise unless the tensors are on a GPU and libpapr_b200.so is present.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from .attention import ProximityAttention
from .bookkeeping import add_points_knn
from .nn import MappingMLP, make_activation
from .optim import FlatAdam, FlatAdamBucket
from .renderer import get_generator
from .schedule import create_learning_rate_fn, reposition


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def _sphere_lattice(center, num_pts, scale):
    """Fibonacci sphere (model.py:194-207)."""
    i = np.arange(num_pts, dtype=np.float64)
    y = 1 - (i / float(num_pts - 1)) * 2
    radius = np.sqrt(1 - y * y)
    theta = math.pi * (3.0 - math.sqrt(5.0)) * i
    pts = np.stack([np.cos(theta) * radius * scale[0] + center[0], y * scale[1] + center[1],
                    np.sin(theta) * radius * scale[2] + center[2]], axis=-1)
    return torch.from_numpy(pts).float()


def _cube_lattice(center, num_pts, scale):
    """Regular lattice + uniform remainder (model.py:239-256); uses numpy's global RNG like the reference."""
    n = int(num_pts ** (1.0 / 3.0))
    axes = [np.linspace(-scale[a], scale[a], n) + center[a] for a in range(3)]
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    rest = num_pts - grid.shape[0]
    if rest > 0:
        extra = np.stack([np.random.uniform(-scale[a], scale[a], rest) + center[a] for a in range(3)], axis=-1)
        grid = np.concatenate([grid, extra], axis=0)
    return torch.from_numpy(grid).float()


class PAPR(nn.Module):
    def __init__(self, args, device="cuda", precision="bf16", verbose=False, ray_chunk=None, fused_optimizer=True):
        super().__init__()
        self.args = args
        self.fused_optimizer = fused_optimizer      # False: one torch.optim.Adam per group as the reference (A/B in tests)
        self.eps = args.eps
        self.device = device
        self.verbose = verbose
        self.precision = precision          # "bf16": tcgen05 GEMMs (product) | "fp32": parity mode
        self.ray_chunk = ray_chunk          # rays per checkpointed chunk in training (None = automatic, see attention.py)
        self.use_amp = args.use_amp
        self.amp_dtype = torch.float16 if args.amp_dtype == "float16" else torch.bfloat16
        # bf16 tensor-core path: no loss scaling needed; the object is kept because callers use scale/step/update
        self.scaler = torch.amp.GradScaler("cuda", enabled=False)

        point_opt = args.geoms.points
        pc_feat_opt = args.geoms.point_feats
        bkg_feat_opt = args.geoms.background
        self.exposure_opt = args.exposure_control
        self.register_buffer("select_k", torch.tensor(point_opt.select_k, device=device, dtype=torch.int32))
        self._select_k = int(point_opt.select_k)
        self.coord_scale = args.dataset.coord_scale

        if point_opt.load_path:
            if not point_opt.load_path.endswith((".pth", ".pt")):
                raise ValueError("point cloud file must be .pth/.pt")
            points = np.asarray(torch.load(point_opt.load_path, map_location="cpu")).astype(np.float32)
            np.random.shuffle(points)
            points = torch.from_numpy(points[: args.max_num_pts, :]).float()
        else:
            center = [c * self.coord_scale for c in point_opt.init_center]
            scale = [s * self.coord_scale for s in point_opt.init_scale]
            if point_opt.init_type == "sphere":
                points = _sphere_lattice(center, point_opt.init_num, scale)
            elif point_opt.init_type == "cube":
                points = _cube_lattice(center, point_opt.init_num, scale)
            else:
                raise NotImplementedError("Point init type [{:s}] is not found".format(point_opt.init_type))
        self.points = nn.Parameter(points, requires_grad=True)
        self.points_influ_scores = nn.Parameter(
            torch.ones(points.shape[0], 1, device=device) * point_opt.influ_init_val, requires_grad=True)

        self.mapping_mlp = None
        if self.exposure_opt.use:
            self.mapping_mlp = MappingMLP(self.exposure_opt.mapping_mlp, self.exposure_opt.shading_code_dim,
                                          self.exposure_opt.mapping_mlp.out_dim)

        compute_dtype = torch.float32 if precision == "fp32" else torch.bfloat16
        if args.models.use_renderer:
            feat_dim = args.models.attn.embed.value.d_ff_out
            self.renderer = get_generator(args.models.renderer.generator, in_c=feat_dim, out_c=3,
                                          compute_dtype=compute_dtype)
        else:
            assert args.models.attn.embed.value.d_ff_out == 3, \
                "Value embedding MLP should have output dim 3 if not using renderer"

        self.bkg_feats = nn.Parameter(torch.FloatTensor(bkg_feat_opt.init_color)[None, :],
                                      requires_grad=bkg_feat_opt.learnable)
        self.bkg_score = torch.tensor(bkg_feat_opt.constant, device=device, dtype=torch.float32).reshape(1)

        if pc_feat_opt.use_ink or pc_feat_opt.use_inq:
            raise NotImplementedError("point features as key/query inputs are not used by any shipped config")
        self.use_pc_feats = bool(pc_feat_opt.use_inv)
        if self.use_pc_feats:
            self.pc_feats = nn.Parameter(torch.randn(points.shape[0], pc_feat_opt.dim), requires_grad=True)

        self.last_act = make_activation(args.models.last_act)
        self.proximity_attn = ProximityAttention(
            args.models.attn, pc_feat_opt.dim if self.use_pc_feats else 0, eps=self.eps,
            bkg_score=bkg_feat_opt.constant, normalize=args.models.normalize_topk_attn, precision=precision)
        self._idx32 = None
        self.init_optimizers(total_steps=0)

    # ------------------------------------------------------------------ optimisers (model.py:117-192, 439-460)
    def init_optimizers(self, total_steps):
        lr_opt = self.args.training.lr
        f = lr_opt.lr_factor
        steps = self.args.training.steps
        groups = [("points", [self.points], lr_opt.points),
                  ("attn", list(self.proximity_attn.parameters()), lr_opt.attn),
                  ("points_influ_scores", [self.points_influ_scores], lr_opt.points_influ_scores)]
        if self.use_pc_feats:
            groups.append(("pc_feats", [self.pc_feats], lr_opt.feats))
        if self.mapping_mlp is not None:
            groups.append(("mapping_mlp", list(self.mapping_mlp.parameters()), lr_opt.mapping_mlp))
        if self.args.models.use_renderer:
            groups.append(("renderer", list(self.renderer.parameters()), lr_opt.generator))
        if self.bkg_feats is not None and self.args.geoms.background.learnable:
            groups.append(("bkg_feats", [self.bkg_feats], lr_opt.bkg_feats))
        self.optimizers, self.schedulers = {}, {}
        # On the GPU all groups share one flat gradient / moment bucket and step in ONE launch (papr_b200/optim.py,
        # SURVEY 8(f1)); a model still on the host keeps torch.optim.Adam (container behaviour only, as the reference).
        self._flat = FlatAdamBucket(self.points.device) if (self.points.is_cuda and self.fused_optimizer) else None
        self._grad_scale = 1.0
        self._opt_total_steps, self._opt_stepped = total_steps, False
        for name, params, opt in groups:
            if name in self.args.training.fix_keys:
                continue
            wd = 0 if name == "points" else opt.weight_decay
            if self._flat is not None:
                optim = FlatAdam(params, self._flat, lr=opt.base_lr * f, weight_decay=wd)
            else:
                optim = torch.optim.Adam(params, lr=opt.base_lr * f, weight_decay=wd)
            self.optimizers[name] = optim
            self.schedulers[name] = create_learning_rate_fn(optim, steps, opt, start_step=total_steps)
        if self._flat is not None:
            self._flat.finalize()

    def _apply(self, fn, *args, **kwargs):
        """model.to(device) / .cuda(): the optimisers built in __init__ (before the move) are rebuilt for the new device,
        as long as they have not been stepped -- this is what switches the usual `PAPR(args, device).to(device)` flow of
        train.py:304-307 to the fused optimiser."""
        out = super()._apply(fn, *args, **kwargs)
        if getattr(self, "optimizers", None) is not None and not getattr(self, "_opt_stepped", True):
            was_cuda = self._flat is not None
            if self.points.is_cuda != was_cuda:
                self.init_optimizers(self._opt_total_steps)
        return out

    def clear_optimizer(self):
        self.optimizers.clear()
        del self.optimizers
        self._flat = None

    def clear_scheduler(self):
        self.schedulers.clear()
        del self.schedulers

    def clear_grad(self):
        if self._flat is not None:
            self._flat.zero_grad()          # one memset; .grad tensors stay views of the flat bucket
            covered = {id(q) for q in self._flat.params}
            for q in self.parameters():     # parameters replaced behind the optimisers' back (load / prune / add without re-init)
                if id(q) not in covered and q.grad is not None:
                    q.grad = None
            return
        for optimizer in self.optimizers.values():
            if optimizer is not None:
                optimizer.zero_grad()

    def step(self, step=-1):
        self._opt_stepped = True
        if self._flat is not None and not self.scaler.is_enabled():
            # every group in one papr_adam_step launch; _grad_scale = 1/world after a summed all-reduce (papr_b200/dist.py)
            self._flat.step_all([o for o in self.optimizers.values() if o is not None], self._grad_scale)
            self._grad_scale = 1.0
            for optimizer in self.optimizers.values():       # keep LambdaLR's "optimizer.step() before lr_scheduler.step()" bookkeeping
                if optimizer is not None:
                    optimizer._opt_called = True
        else:
            for optimizer in self.optimizers.values():
                if optimizer is not None:
                    self.scaler.step(optimizer)
        for scheduler in self.schedulers.values():
            if scheduler is not None:
                scheduler.step()
        for attr, name in (("attn_lr", "attn"), ("pts_lr", "points")):
            lr = 0
            if name in self.optimizers:
                sched = self.schedulers[name]
                lr = sched.get_last_lr()[0] if sched is not None else self.optimizers[name].param_groups[0]["lr"]
            setattr(self, attr, lr)

    # ------------------------------------------------------------------ selection (model.py:258-333)
    def _get_points(self, rays_o, rays_d, c2w=None, step=-1):
        """int32 (N,H,W,K') indices of the K nearest points per ray (all points when K >= P, model.py:326-327)."""
        N, H, W, _ = rays_d.shape
        P = self.points.shape[0]
        if self._select_k >= P or self._select_k < 0:
            idx = torch.arange(P, device=self.points.device, dtype=torch.int32).expand(N, H, W, -1).contiguous()
        else:
            idx = ops.select_topk(rays_o, rays_d, self.points.detach(), self._select_k, self.eps)
        self._idx32 = idx
        return idx

    @property
    def select_k_ind(self):
        """int64 indices as the reference exposes them (model.py:464); converted once per selection, on first access."""
        if self._idx32 is None:
            return None
        cached = getattr(self, "_idx64", None)
        if cached is None or cached[0] is not self._idx32:
            cached = (self._idx32, self._idx32.long())
            self._idx64 = cached
        return cached[1]

    @property
    def selected_points(self):
        """(N,H,W,K,3) positions of the selected points (model.py:330-331), gathered on demand."""
        return None if self._idx32 is None else self.points.detach()[self._idx32.long(), :]

    # ------------------------------------------------------------------ bookkeeping (model.py:335-394)
    def prune_points(self, thresh):
        if self.points_influ_scores is None:
            return 0
        if self.args.training.prune_type not in ("<", ">"):
            raise ValueError("Invalid prune type")
        keep_less = self.args.training.prune_type == ">"
        if self.points.is_cuda:       # order-preserving stream compaction on the device (papr_prune_compact)
            P = self.points.shape[0]
            feats = self.pc_feats.detach() if self.use_pc_feats else None
            pts, influ, feats, kept = ops.prune_compact(self.points.detach(), self.points_influ_scores.detach(), feats,
                                                        float(thresh), keep_less)
            self.points = nn.Parameter(pts.clone(), requires_grad=self.points.requires_grad)
            self.points_influ_scores = nn.Parameter(influ.clone(), requires_grad=self.points_influ_scores.requires_grad)
            if self.use_pc_feats:
                self.pc_feats = nn.Parameter(feats.clone(), requires_grad=self.pc_feats.requires_grad)
            self._idx32 = None
            return torch.tensor(P - kept, device=pts.device)
        mask = self.points_influ_scores[:, 0] < thresh if keep_less else self.points_influ_scores[:, 0] > thresh
        n_pruned = torch.sum(mask == 0)
        self.points = nn.Parameter(self.points[mask, :], requires_grad=self.points.requires_grad)
        self.points_influ_scores = nn.Parameter(self.points_influ_scores[mask, :],
                                                requires_grad=self.points_influ_scores.requires_grad)
        if self.use_pc_feats:
            self.pc_feats = nn.Parameter(self.pc_feats[mask, :], requires_grad=self.pc_feats.requires_grad)
        self._idx32 = None
        return n_pruned

    def add_points(self, add_num):
        cur = self.points.shape[0]
        if "max_points" in self.args and self.args.max_points > 0 and (cur + add_num) >= self.args.max_points:
            add_num = self.args.max_points - cur
            if add_num <= 0:
                return 0
        po = self.args.geoms.points
        feats = self.pc_feats.detach() if self.use_pc_feats else None
        new_pts, n_new, new_influ, new_feats = add_points_knn(
            self.points.detach(), self.points_influ_scores.detach(), add_num=add_num, k=po.add_k,
            comb_type=po.add_type, sample_k=po.add_sample_k, sample_type=po.add_sample_type, point_features=feats)
        if n_new > 0:
            dev = self.points.device
            self.points = nn.Parameter(torch.cat([self.points.detach(), new_pts.to(dev)], 0),
                                       requires_grad=self.points.requires_grad)
            self.points_influ_scores = nn.Parameter(
                torch.cat([self.points_influ_scores.detach(), new_influ.to(dev)], 0),
                requires_grad=self.points_influ_scores.requires_grad)
            if self.use_pc_feats:
                self.pc_feats = nn.Parameter(torch.cat([self.pc_feats.detach(), new_feats.to(dev)], 0),
                                             requires_grad=self.pc_feats.requires_grad)
            self._idx32 = None
        return n_new

    # ------------------------------------------------------------------ the hot path (model.py:462-560)
    def _attend(self, rays_o, rays_d, c2w, step):
        if not rays_d.is_cuda:
            raise RuntimeError("papr_b200 needs CUDA tensors: there is no CPU path (and no fallback)")
        _lib.check_device(rays_d.device.index if rays_d.device.index is not None else torch.cuda.current_device())
        with torch.cuda.device(rays_d.device):      # the C ABI launches on the current device
            idx = self._get_points(rays_o, rays_d, c2w, step)
        feats = self.pc_feats if self.use_pc_feats else None
        return self.proximity_attn(rays_o, rays_d, idx, self.points, feats, self.points_influ_scores,
                                   precision=self.precision, ray_chunk=self.ray_chunk)

    def evaluate(self, rays_o, rays_d, c2w, step=-1, shading_code=None):
        N, H, W, _ = rays_d.shape
        fused, attn = self._attend(rays_o, rays_d, c2w, step)
        return fused.reshape(N, H, W, 1, -1), attn.reshape(N, H, W, -1, 1)

    def forward(self, rays_o, rays_d, c2w, step=-1, shading_code=None):
        gamma = beta = None
        if shading_code is not None and self.mapping_mlp is not None:
            affine = self.mapping_mlp(shading_code)
            half = affine.shape[-1] // 2
            gamma, beta = affine[:half], affine[half:]
        N, H, W, _ = rays_d.shape
        fused, attn = self._attend(rays_o, rays_d, c2w, step)
        fused = fused.reshape(N, H, W, -1)
        if self.args.models.use_renderer:
            fg = self.renderer(fused.permute(0, 3, 1, 2), gamma=gamma, beta=beta).permute(0, 2, 3, 1)
        else:
            fg = fused
        bkg_attn = attn[:, -1].reshape(N, H, W, 1)
        bkg = self.bkg_feats.reshape(1, 1, 1, -1)
        if self.args.models.normalize_topk_attn:
            rgb = fg * (1 - bkg_attn) + bkg * bkg_attn
        else:
            rgb = fg + bkg * bkg_attn
        if self.verbose and step >= 0 and step % 1000 == 0:
            print(" predict rgb:", step, rgb.shape, rgb.min().item(), rgb.max().item(), rgb.mean().item())
        return rgb

    # ------------------------------------------------------------------ checkpoints (model.py:562-641)
    def save(self, step, save_dir):
        torch.save({str(step): self.state_dict()}, os.path.join(save_dir, "model.pth"))
        torch.save({n: (o.state_dict() if o is not None else None) for n, o in self.optimizers.items()},
                   os.path.join(save_dir, "optimizers.pth"))
        torch.save({n: (s.state_dict() if s is not None else None) for n, s in self.schedulers.items()},
                   os.path.join(save_dir, "schedulers.pth"))
        torch.save(self.scaler.state_dict(), os.path.join(save_dir, "scaler.pth"))

    def load(self, load_dir, load_optimizer=False):
        if load_optimizer:
            osd = torch.load(os.path.join(load_dir, "optimizers.pth"))
            for name, optimizer in self.optimizers.items():
                if optimizer is not None:
                    optimizer.load_state_dict(osd[name])
            ssd = torch.load(os.path.join(load_dir, "schedulers.pth"))
            for name, scheduler in self.schedulers.items():
                if scheduler is None:
                    continue
                sd = ssd[name]
                if "lr_lambdas" in sd:
                    scheduler.load_state_dict(sd)
                else:
                    # a reference checkpoint (SequentialLR state): our schedule is a closed form of the step count,
                    # so repositioning it at the saved step reproduces the same learning rates
                    reposition(scheduler, int(sd["last_epoch"]))
        spath = os.path.join(load_dir, "scaler.pth")
        if os.path.exists(spath):
            sd = torch.load(spath)
            if sd:
                self.scaler.load_state_dict(sd)
        for step, state_dict in torch.load(os.path.join(load_dir, "model.pth")).items():
            self.load_my_state_dict(state_dict)
            return int(step)

    def load_my_state_dict(self, state_dict, exclude_keys=[]):
        own = self.state_dict()
        resized = ("points", "points_influ_scores", "pc_feats")
        for name, param in state_dict.items():
            if any(ex in name for ex in exclude_keys) or name in resized:
                continue
            if isinstance(param, nn.Parameter):
                param = param.data
            if name in own and own[name].shape == param.shape:
                own[name].copy_(param)
            elif self.verbose:
                print("Can't load", name)
        dev = self.points.device
        self.points = nn.Parameter(state_dict["points"].data.to(dev), requires_grad=self.points.requires_grad)
        if self.points_influ_scores is not None:
            self.points_influ_scores = nn.Parameter(state_dict["points_influ_scores"].data.to(dev),
                                                    requires_grad=self.points_influ_scores.requires_grad)
        if self.use_pc_feats:
            self.pc_feats = nn.Parameter(state_dict["pc_feats"].data.to(dev), requires_grad=self.pc_feats.requires_grad)
        self._select_k = int(self.select_k)      # the buffer may just have been overwritten (e.g. a K=30 checkpoint)
        self._idx32 = None
        # the point tables are NEW Parameter objects now: optimisers that have not been stepped yet are rebuilt around them
        # (the reference leaves its optimisers pointing at the old tensors until train.py re-creates them)
        if getattr(self, "optimizers", None) is not None and not getattr(self, "_opt_stepped", True):
            self.init_optimizers(self._opt_total_steps)


def get_model(args, device="cuda", **kw):
    """reference models/__init__.py:23-24"""
    return PAPR(args, device=device, **kw)
