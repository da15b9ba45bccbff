import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import papr_oracle as O
from papr_b200.config import make_config
from papr_b200.model import PAPR
cfgname = os.environ.get("CFG", "caterpillar")
cfg = make_config(cfgname, use_amp=False)
P = int(os.environ.get("P", 100000))
cfg.geoms.points["init_num"] = P
params = O.init_params(cfg, P, seed=8, cloud="shell")
rays_o, rays_d, c2w = O.synthetic_rays(1080, 1920, cfg.dataset.coord_scale, n_views=1, seed=3)
model = PAPR(cfg, device="cuda", precision="bf16").cuda()
model.load_my_state_dict({k: v.clone() for k, v in params.items()})
crop = rays_d[:, 400:580, 800:1120].contiguous().cuda()
tgt = torch.rand(1, 180, 320, 3, generator=torch.Generator().manual_seed(6)).cuda()
def run(chunk):
    model.ray_chunk = chunk
    model.clear_grad()
    out = model(rays_o.cuda(), crop, None)
    torch.mean((out - tgt) ** 2).backward()
    return out.detach().clone(), {k: getattr(model, k).grad.detach().clone() for k in ("points", "pc_feats", "points_influ_scores")}
res = {}
for name, chunk in (("a", 10**9), ("a2", 10**9), ("b", 12800), ("b2", 12800), ("c", 28800)):
    res[name] = run(chunk)
for x, y in (("a", "a2"), ("b", "b2"), ("a", "b"), ("a", "c")):
    for k in ("points", "pc_feats", "points_influ_scores"):
        ga, gb = res[x][1][k].double(), res[y][1][k].double()
        d = (ga - gb).abs()
        i = int(d.reshape(-1).argmax())
        print(x, y, k, "max|a|", float(ga.abs().max()), "max diff", float(d.max()), "rel L2", float((ga - gb).norm() / ga.norm()),
              "at", i, float(ga.reshape(-1)[i]), float(gb.reshape(-1)[i]), "nnz", int((ga != 0).sum()), int((gb != 0).sum()))
