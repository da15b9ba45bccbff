"""Debug: grid selection with a tight threshold lane (PAPR_SELECT_LAST) against the plain scan at the Caterpillar shape."""
import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import papr_oracle as O
from papr_b200 import ops
from papr_b200._lib import lib
from papr_b200.config import make_config
cfg = make_config("caterpillar", use_amp=False)
P = 100000
params = O.init_params(cfg, P, seed=8, cloud="shell")
rays_o, rays_d, c2w = O.synthetic_rays(1080, 1920, cfg.dataset.coord_scale, n_views=1, seed=3)
ro, rd, pts = rays_o.cuda(), rays_d.cuda().contiguous(), params["points"].cuda()
fb = getattr(ctypes.CDLL(lib()._name), "papr_debug_select_fallbacks"); fb.restype = ctypes.c_longlong
want = ops.select_topk(ro, rd, pts, 20, cull=False).reshape(-1, 20)
torch.cuda.synchronize()
for build in ("grid", "grid_torch"):
    for last in (31, 26, 23, 21):
        os.environ["PAPR_SELECT_LAST"] = str(last)
        fb()
        got = ops.select_topk(ro, rd, pts, 20, cull=build).reshape(-1, 20)
        torch.cuda.synchronize()
        bad = (got != want).any(-1).nonzero().flatten()
        print(build, "last", last, "fallbacks", fb(), "bad rays", bad.numel(), flush=True)
        if bad.numel() and last == 21 and build == "grid":
            sv, perm, cells, vw, G = ops.view_grids(ro, rd.reshape(1, -1, 3), pts, 1e-6)
            e1, e2, c = vw[0, 0:3].double(), vw[0, 3:6].double(), vw[0, 6:9].double()
            print("G", G, "gmin", vw[0, 9:11].tolist(), "cell", vw[0, 11:13].tolist(), "zmin_all", float(vw[0, 15]), "wmax", float(vw[0, 16]))
            for r in bad[:6].tolist():
                d = rd.reshape(-1, 3)[r].double(); o = ro[0].double()
                v = pts.double() - o
                t = (v @ d) / (d @ d + 1e-6)
                dist2 = ((v - t[:, None] * d) ** 2).sum(-1)
                order = dist2.argsort()
                g, w = got[r].tolist(), want[r].tolist()
                miss = [i for i in w if i not in g]
                extra = [i for i in g if i not in w]
                h = torch.stack([d @ e1, d @ e2]) / (d @ c)
                print("ray", r, "h", h.tolist(), "ray cell", ((h - vw[0, 9:11].double()) * vw[0, 13:15].double()).tolist())
                print("  want", w); print("  got ", g)
                for i in miss + extra:
                    vi = v[i]; w3 = float(vi @ c); gi = torch.stack([vi @ e1, vi @ e2]) / w3
                    rank = int((order == i).nonzero())
                    print("  point", i, "missing" if i in miss else "extra", "rank", rank, "dist2", float(dist2[i]), "w3", w3, "|v|", float(vi.norm()),
                          "g", gi.tolist(), "cell", ((gi - vw[0, 9:11].double()) * vw[0, 13:15].double()).tolist())
                print("  dist2 of ranks 18..24", dist2[order[18:25]].tolist())
os.environ.pop("PAPR_SELECT_LAST", None)
