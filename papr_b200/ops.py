"""Python wrappers over the C ABI: allocate outputs with torch, pass raw device pointers and the current stream."""
import torch

from ._lib import check, lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


class LaunchStats:
    """Counts the library's kernel launches; with timing on, brackets each launch with CUDA events on the launching
    stream and records its algorithmic work (flops, bytes) so bench.py can report per-kernel roofline numbers."""

    def __init__(self):
        self.count = 0
        self.timing = False
        self.records = []      # (name, start_event, end_event, flops, bytes)

    def reset(self):
        self.count = 0
        self.records = []

    def summary(self):
        """name -> dict(launches, ms, flops, bytes); call after torch.cuda.synchronize()."""
        out = {}
        for name, e0, e1, fl, by in self.records:
            d = out.setdefault(name, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += fl
            d["bytes"] += by
        return out


STATS = LaunchStats()


# entry points that are variants of one kernel are accounted under one name
_STAT_NAME = {"papr_wgrad_bf16_ex": "papr_wgrad_bf16", "papr_wgrad_bias_bf16": "papr_wgrad_bf16", "papr_stack_bf16_ex": "papr_stack_bf16"}


def call(name, *args, flops=0.0, nbytes=0.0, kernels=1):
    """Invoke one C-ABI entry point on the current stream (the stream pointer is appended).  `kernels`: how many kernels
    the entry point launches (for the launch count bench.py reports)."""
    fn = getattr(lib(), name)
    STATS.count += kernels
    if STATS.timing:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        status = fn(*args, _stream())
        e1.record()
        STATS.records.append((_STAT_NAME.get(name, name), e0, e1, flops, nbytes))
    else:
        status = fn(*args, _stream())
    check(status, name)


def _f32c(t):
    assert t.is_cuda, "papr_b200 ops need CUDA tensors (there is no CPU path)"
    return t.detach().contiguous().float()




def _bounding_spheres(grp, valid):
    """grp (G,n,3), valid (G,n,1) bool -> (G,4): centroid + radius covering the valid points (rounded up)."""
    cnt = valid.sum(1).clamp_min(1)
    centre = (grp * valid).sum(1) / cnt
    rad = (((grp - centre.unsqueeze(1)) * valid).norm(dim=-1)).max(1).values * (1.0 + 1e-5) + 1e-6 * centre.norm(dim=-1) + 1e-30
    return torch.cat([centre, rad.unsqueeze(-1)], dim=-1)


def morton_groups(points):
    """Sort points along a Morton curve, pad to a multiple of 256 and bound every group of 32 (and every super-group of
    256) consecutive points by a sphere.  Super-groups are then visited in bit-reversed curve order: successive
    super-groups are far apart in space, so a ray's thresholds tighten as fast as with randomly ordered points (no
    insertion storms while the curve creeps towards the ray) while every group stays compact for the culling test.
    Returns (sorted (P_pad,3) f32, perm (P_pad,) i32, spheres (P_pad/32,4), spheres8 (P_pad/256,4), pmax tensor ())."""
    P = points.shape[0]
    dev = points.device
    lo = points.min(0).values
    span = (points.max(0).values - lo).clamp_min(1e-20)
    q = ((points - lo) / span * 1023.0).clamp(0, 1023).long()

    def spread(v):      # 10 bits -> every third bit
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    order = torch.argsort(code)
    P_pad = (P + 255) // 256 * 256
    spts = torch.full((P_pad, 3), 1.0e18, dtype=torch.float32, device=dev)
    spts[:P] = points[order]
    perm = torch.full((P_pad,), -1, dtype=torch.int32, device=dev)
    perm[:P] = order.to(torch.int32)
    S = P_pad // 256
    bits = max(1, (S - 1).bit_length())
    si = torch.arange(S, device=dev)
    rev = torch.zeros_like(si)
    for b in range(bits):
        rev |= ((si >> b) & 1) << (bits - 1 - b)
    sorder = torch.argsort(rev)
    spts = spts.reshape(S, 256, 3)[sorder].reshape(-1, 3).contiguous()
    perm = perm.reshape(S, 256)[sorder].reshape(-1).contiguous()
    valid = (perm >= 0)
    spheres = _bounding_spheres(spts.reshape(-1, 32, 3), valid.reshape(-1, 32, 1)).contiguous()
    spheres8 = _bounding_spheres(spts.reshape(-1, 256, 3), valid.reshape(-1, 256, 1)).contiguous()
    return spts, perm, spheres, spheres8, points.norm(dim=-1).max()


GRID_MIN_POINTS = 1024     # below this the plain scan is as fast as building the grid


def grid_size(P):
    """Cells per axis of the selection grid: about eight points per cell."""
    return int(min(128, max(8, round((P / 8.0) ** 0.5))))


def view_grids(rays_o, rays_d, points, eps, G=None):
    """The acceleration structure of the screen-space grid selection (select_grid_kernel in csrc/select.cu), built by
    papr_select_grid_build (csrc/select_grid_build.cu: six small launches, no host synchronisation): per view, a camera
    frame around the mean ray direction, the gnomonic extent of the view's rays and of the points in front of the camera,
    the points binned on a G x G grid over that extent and stored cell by cell.
    rays_o (N,3), rays_d (N,R,3), points (P,3) -> (sorted_v (N*P,4), perm (N*P,) i32, cells (N*G*G,4) i32, views (N,20), G)."""
    N, R, _ = rays_d.shape
    P = points.shape[0]
    dev = points.device
    if G is None:
        G = grid_size(P)
    sv = torch.empty((N * P, 4), device=dev)
    perm = torch.empty((N * P,), dtype=torch.int32, device=dev)
    cells = torch.empty((N * G * G, 4), dtype=torch.int32, device=dev)
    views = torch.empty((N, 20), device=dev)
    nbytes = int(lib().papr_select_grid_workspace_bytes(N, P, G))
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=dev)
    call("papr_select_grid_build", rays_o.data_ptr(), rays_d.data_ptr(), points.data_ptr(), N, R, P, G, float(eps), sv.data_ptr(),
         perm.data_ptr(), cells.data_ptr(), views.data_ptr(), ws.data_ptr(), nbytes, kernels=6,
         nbytes=12.0 * N * R * 2 + N * P * (12.0 * 3 + 4 * 2 + 20))
    return sv, perm, cells, views, G


def view_grids_torch(rays_o, rays_d, points, eps, G=None):
    """The same structure from ~60 torch ops (stable sort by cell, searchsorted, scatter_reduce), batched over views: the
    first implementation, kept as the A/B partner of papr_select_grid_build in tests/test_select_gpu.py (the two differ in
    the last bits of the frame and in the order inside a cell; the selection they lead to is identical)."""
    N, R, _ = rays_d.shape
    P = points.shape[0]
    dev = points.device
    if G is None:
        G = grid_size(P)
    # (no torch.tensor(python data, device=...) in here: a pageable host-to-device copy synchronises the stream and would
    # stop the host from running ahead of the GPU)
    dn = rays_d / rays_d.norm(dim=-1, keepdim=True).clamp_min(1e-30)
    c = dn.mean(1)
    ax = torch.zeros((3, N, 3), device=dev)
    ax[0, :, 0] = 1.0
    ax[1, :, 1] = 1.0
    ax[2, :, 2] = 1.0
    c = torch.where(c.norm(dim=-1, keepdim=True) > 1e-6, c, ax[2])
    c = c / c.norm(dim=-1, keepdim=True)
    helper = torch.where(c[:, :1].abs() < 0.9, ax[0], ax[1])
    e1 = torch.linalg.cross(helper, c)
    e1 = e1 / e1.norm(dim=-1, keepdim=True)
    e2 = torch.linalg.cross(c, e1)
    B = torch.stack([e1, e2, c], dim=1)                                   # (N,3,3), rows = camera axes
    # gnomonic extent of the rays that point forward (the others cannot be bounded and scan everything)
    w = torch.einsum("nrj,nij->nri", dn, B)
    ok = w[..., 2] > 0.25
    h = w[..., :2] / torch.where(ok, w[..., 2], torch.ones_like(w[..., 2])).unsqueeze(-1)
    inf = torch.full_like(h, float("inf"))
    hmin = torch.where(ok.unsqueeze(-1), h, inf).amin(1)
    hmax = torch.where(ok.unsqueeze(-1), h, -inf).amax(1)
    none = ~ok.any(1, keepdim=True)
    hmin = torch.where(none, torch.full_like(hmin, -1.0), hmin)
    hmax = torch.where(none, torch.full_like(hmax, 1.0), hmax)
    # points in every view's frame
    v = points.unsqueeze(0) - rays_o.unsqueeze(1)                         # (N,P,3): the kernel's v = p - o, rounded once
    vn2 = torch.addcmul(torch.addcmul(v[..., 0] * v[..., 0], v[..., 1], v[..., 1]), v[..., 2], v[..., 2])
    pw = torch.einsum("npj,nij->npi", v, B)
    valid = pw[..., 2].abs() > 1e-3 * vn2.sqrt()
    g = pw[..., :2] / torch.where(valid, pw[..., 2], torch.ones_like(pw[..., 2])).unsqueeze(-1)
    # The grid covers the rays AND the points in front of the camera (at most two ray-spans beyond the rays on each side):
    # when the rays are a stripe of the frame that misses the object (one rank of a row-sharded render), a grid over the
    # rays alone would put every point into its semi-infinite border cells and the kernel would degenerate to a full scan.
    span0 = (hmax - hmin).clamp_min(1e-3)
    reach = 2.0 * span0.amax(-1, keepdim=True)
    front = (valid & (pw[..., 2] > 0.25 * vn2.sqrt())).unsqueeze(-1)
    pmin = torch.where(front, g, inf[:, :1].expand_as(g)).amin(1)
    pmax = torch.where(front, g, -inf[:, :1].expand_as(g)).amax(1)
    hmin = torch.maximum(torch.minimum(hmin, pmin), hmin - reach)
    hmax = torch.minimum(torch.maximum(hmax, pmax), hmax + reach)
    span = (hmax - hmin).clamp_min(1e-3)
    gmin = hmin - 0.02 * span
    cell = (span * 1.04) / G
    icell = 1.0 / cell
    cxy = ((g - gmin.unsqueeze(1)) * icell.unsqueeze(1)).floor().clamp(0, G - 1)
    cxy = torch.where(valid.unsqueeze(-1), cxy, torch.zeros_like(cxy)).long()
    cid = (torch.arange(N, device=dev).unsqueeze(1) * G + cxy[..., 1]) * G + cxy[..., 0]          # (N,P)
    az = torch.where(valid, pw[..., 2].abs(), torch.zeros_like(vn2))      # 0: the cell can never be skipped
    flat = cid.reshape(-1)
    order = torch.argsort(flat, stable=True)
    sorted_cid = flat[order]
    sv = torch.cat([v, (eps * vn2).unsqueeze(-1)], dim=-1).reshape(-1, 4)[order].contiguous()
    perm = (order % P).to(torch.int32).contiguous()
    bounds = torch.searchsorted(sorted_cid, torch.arange(N * G * G + 1, device=dev))
    view_start = (torch.arange(N, device=dev) * P).repeat_interleave(G * G)
    zmin = torch.full((N * G * G,), float("inf"), device=dev).scatter_reduce(0, flat, az.reshape(-1), "amin")
    cells = torch.stack([(bounds[:-1] - view_start).to(torch.int32), (bounds[1:] - view_start).to(torch.int32),
                         zmin.view(torch.int32), torch.zeros(N * G * G, dtype=torch.int32, device=dev)], dim=-1).contiguous()
    zmin_all = zmin.reshape(N, -1).amin(1) * (1.0 - 1e-6)
    views = torch.zeros((N, 20), device=dev)
    views[:, 0:3], views[:, 3:6], views[:, 6:9] = e1, e2, c
    views[:, 9:11], views[:, 11:13], views[:, 13:15] = gmin, cell, icell
    views[:, 15], views[:, 16] = zmin_all, vn2.amax(1)
    return sv, perm, cells, views.contiguous(), G


def select_topk(rays_o, rays_d, points, K, eps=1e-6, cull=None):
    """Stage a1 (reference models/model.py:258-283): int32 (N,H,W,K) nearest-point indices per ray,
    ordered by (distance, index).  rays_o (N,3), rays_d (N,H,W,3), points (P,3), all CUDA fp32.
    cull: None = automatic (the screen-space grid kernel for P >= GRID_MIN_POINTS, else the plain scan); "grid";
    "grid_torch" = the grid kernel on a structure built by torch ops (A/B partner); True / "morton" = the Morton-group
    kernel; False = plain scan.  All variants return the identical result."""
    N, H, W, _ = rays_d.shape
    P = points.shape[0]
    if not (1 <= K <= 32) or K >= P:
        raise ValueError(f"select_topk needs 1 <= K <= 32 and K < P (K={K}, P={P})")
    ro, rd, pts = _f32c(rays_o), _f32c(rays_d), _f32c(points)
    idx = torch.empty((N, H, W, K), dtype=torch.int32, device=rd.device)
    R = N * H * W
    work = dict(flops=17.0 * R * P, nbytes=12.0 * R + 12.0 * P + 4.0 * K * R)
    if cull is None:
        cull = "grid" if P >= GRID_MIN_POINTS else False
    if cull in ("grid", "grid_torch"):
        build = view_grids if cull == "grid" else view_grids_torch
        sv, perm, cells, views, G = build(ro, rd.reshape(N, H * W, 3), pts, float(eps))
        call("papr_select_topk_grid", ro.data_ptr(), rd.data_ptr(), sv.data_ptr(), perm.data_ptr(), cells.data_ptr(), views.data_ptr(),
             N, H * W, P, G, K, float(eps), idx.data_ptr(), **work)
    elif cull:
        spts, perm, spheres, spheres8, pmax = morton_groups(pts)
        pmax = pmax.reshape(1).float().contiguous()
        call("papr_select_topk_sorted", ro.data_ptr(), rd.data_ptr(), spts.data_ptr(), perm.data_ptr(), spheres.data_ptr(), spheres8.data_ptr(),
             N, H * W, spts.shape[0], P, K, float(eps), pmax.data_ptr(), idx.data_ptr(), **work)
    else:
        call("papr_select_topk", ro.data_ptr(), rd.data_ptr(), pts.data_ptr(), N, H * W, P, K, float(eps), idx.data_ptr(), **work)
    return idx


# --------------------------------------------------------------------------- tensor-core building blocks
def pad_rows(n):
    return (n + 127) // 128 * 128


def pad_cols(n):
    return (n + 63) // 64 * 64


class Blocked:
    """A tile-blocked bf16 activation (see include/papr_b200.h): raw uint8 storage + logical shape."""

    def __init__(self, rows, cols, device, zero=False):
        self.rows, self.cols = rows, cols
        self.rows_pad, self.cols_pad = pad_rows(rows), pad_cols(cols)
        alloc = torch.zeros if zero else torch.empty
        self.buf = alloc(self.rows_pad * self.cols_pad * 2, dtype=torch.uint8, device=device)

    def data_ptr(self):
        return self.buf.data_ptr()

    @staticmethod
    def wrap(buf, rows, cols, cols_pad):
        """A Blocked view of existing storage (the inverse of handing .buf to ctx.save_for_backward)."""
        out = Blocked.__new__(Blocked)
        out.rows, out.cols, out.cols_pad = rows, cols, cols_pad
        out.rows_pad = pad_rows(rows)
        out.buf = buf
        return out

    def rows_view(self, r0, r1):
        """Rows [r0, r1) (multiples of 128) as a Blocked over the same storage: a tile is contiguous in this layout."""
        assert r0 % 128 == 0 and (r1 % 128 == 0 or r1 == self.rows_pad) and 0 <= r0 < r1 <= self.rows_pad
        if r0 == 0 and r1 == self.rows_pad:
            return self
        per_row = self.cols_pad * 2
        return Blocked.wrap(self.buf[r0 * per_row:r1 * per_row], min(self.rows, r1) - r0 if self.rows > r0 else r1 - r0, self.cols, self.cols_pad)

    @staticmethod
    def from_f32(t, cols_pad=None):
        t = _f32c(t)
        out = Blocked(t.shape[0], t.shape[1], t.device)
        if cols_pad:
            out.cols_pad = cols_pad
            out.buf = torch.empty(out.rows_pad * cols_pad * 2, dtype=torch.uint8, device=t.device)
        call("papr_blocked_from_f32", t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), out.data_ptr(), out.rows_pad,
             out.cols_pad, nbytes=4.0 * t.numel() + 2.0 * out.rows_pad * out.cols_pad)
        return out

    def to_f32(self, rows=None, cols=None):
        rows = self.rows if rows is None else rows
        cols = self.cols if cols is None else cols
        out = torch.empty((rows, cols), dtype=torch.float32, device=self.buf.device)
        call("papr_blocked_to_f32", self.data_ptr(), self.cols_pad, out.data_ptr(), rows, cols, out.stride(0),
             nbytes=6.0 * rows * cols)
        return out


WEIGHT_REPLICAS = 8      # copies of every weight image handed to the fused stack kernel (see papr_stack_layer)


def pack_weight(w, N, K, transpose=False, scale=1.0, replicas=1):
    """bf16 weight image (uint8 tensor) for linear_bf16 from a torch Linear weight (out,in); `replicas` > 1 returns
    that many identical copies back to back (shape (replicas, bytes))."""
    if replicas > 1:
        one = pack_weight(w, N, K, transpose, scale)
        return one.unsqueeze(0).repeat(replicas, 1)
    w = _f32c(w)
    img = torch.empty(((K + 63) // 64) * N * 128, dtype=torch.uint8, device=w.device)
    call("papr_pack_weight", w.data_ptr(), w.stride(0), w.shape[0], w.shape[1], int(transpose), N, K, float(scale),
         img.data_ptr(), nbytes=6.0 * N * K)
    return img


def sign_bits_rowmajor(bits):
    """Sign-bit words as a (rows_pad, words) tensor indexed [row, 64-column group].  The kernels keep them as
    [128-row tile][group][row] so that a warp reads / writes 256 contiguous bytes; this is for tests and tools."""
    rp, ng = bits.shape
    return bits.view(rp // 128, ng, 128).permute(0, 2, 1).reshape(rp, ng)


def sign_bits_from_rowmajor(rm):
    """Inverse of sign_bits_rowmajor: (rows_pad, words) [row, group] -> the kernels' [tile][group][row] storage."""
    rp, ng = rm.shape
    return rm.view(rp // 128, 128, ng).permute(0, 2, 1).contiguous().view(rp, ng)


def linear_bf16(x, w_image, N, K, bias=None, act=False, slope=0.0, out_blocked=True, out_f32=False,
                sign_bits_out=False, sign_bits_in=None, colsum=None, addend=None):
    """Y = act(X W^T + b) on tcgen05 (see papr_linear_bf16).  Returns (Blocked|None, f32|None, bits|None)."""
    dev = x.buf.device
    rows_pad = x.rows_pad
    assert x.cols_pad == pad_cols(K), (x.cols_pad, K)
    yb = Blocked(x.rows, N, dev) if out_blocked else None
    yf = torch.empty((rows_pad, N), dtype=torch.float32, device=dev) if out_f32 else None
    bits = torch.empty((rows_pad, pad_cols(N) // 64), dtype=torch.int64, device=dev) if sign_bits_out else None
    nbytes = rows_pad * (2.0 * pad_cols(K) + (2.0 * pad_cols(N) if out_blocked else 0) + (4.0 * N if out_f32 else 0)
                         + (8.0 * pad_cols(N) / 64 if sign_bits_out else 0) + (8.0 * pad_cols(N) / 64 if sign_bits_in is not None else 0)
                         + (4.0 * N if addend is not None else 0))
    call("papr_linear_bf16",
         x.data_ptr(), w_image.data_ptr(), bias.data_ptr() if bias is not None else None,
         yb.data_ptr() if yb is not None else None, yf.data_ptr() if yf is not None else None, N,
         bits.data_ptr() if bits is not None else None, sign_bits_in.data_ptr() if sign_bits_in is not None else None,
         colsum.data_ptr() if colsum is not None else None,
         addend.data_ptr() if addend is not None else None, addend.stride(0) if addend is not None else 0,
         rows_pad, N, K, int(act), float(slope),
         flops=2.0 * x.rows * N * K, nbytes=nbytes)
    return yb, yf, bits


def wgrad_bf16(a, b, out, a_valid, b_valid, transpose_out=False, max_ctas=0, a_colsum=None):
    """out[a,b] += sum_rows A[row,a] B[row,b]  (out fp32, atomically accumulated), on at most max_ctas SMs (0 = all);
    a_colsum (fp32 [a_valid], optional) += sum_rows A[row, :] -- the bias gradient when A is dZ."""
    assert a.rows_pad == b.rows_pad
    if a_colsum is not None:
        assert a_colsum.dtype == torch.float32 and a_colsum.is_contiguous() and a_colsum.numel() >= a_valid
    call("papr_wgrad_bias_bf16", a.data_ptr(), a.cols_pad, b.data_ptr(), b.cols_pad, out.data_ptr(), out.stride(0),
         a_valid, b_valid, int(transpose_out), a.rows_pad, int(max_ctas), a_colsum.data_ptr() if a_colsum is not None else None,
         flops=2.0 * a.rows * a_valid * b_valid,
         nbytes=2.0 * a.rows_pad * (128 * ((a_valid + 127) // 128) + pad_cols(b_valid)))
    return out


def stack_bf16(x, K0, layers, slope=0.0, max_ctas=0):
    """Fused MLP stack (papr_stack_bf16).  `layers`: list of dicts with keys w_image, N and optionally bias, act,
    out_blocked (Blocked), out_f32 (tensor), sign_bits_out (tensor), sign_bits_in (tensor), colsum (tensor)."""
    import ctypes
    from ._lib import StackLayer
    arr = (StackLayer * len(layers))()
    flops, nbytes = 0.0, 2.0 * x.rows_pad * x.cols_pad
    K = K0
    for i, l in enumerate(layers):
        def ptr(key):
            v = l.get(key)
            return v.data_ptr() if v is not None else None
        arr[i].w_image, arr[i].bias = ptr("w_image"), ptr("bias")
        arr[i].out_blocked, arr[i].out_f32 = ptr("out_blocked"), ptr("out_f32")
        arr[i].ld_f32 = l["out_f32"].stride(0) if l.get("out_f32") is not None else 0
        arr[i].sign_bits_out, arr[i].sign_bits_in, arr[i].colsum = ptr("sign_bits_out"), ptr("sign_bits_in"), ptr("colsum")
        arr[i].N, arr[i].act = int(l["N"]), int(bool(l.get("act")))
        img = l["w_image"]
        arr[i].w_replicas = img.shape[0] if img.dim() == 2 else 1
        arr[i].w_replica_stride = img.stride(0) if img.dim() == 2 else 0
        flops += 2.0 * x.rows * l["N"] * K
        nbytes += x.rows_pad * ((2.0 * pad_cols(l["N"]) if l.get("out_blocked") is not None else 0)
                                + (4.0 * l["N"] if l.get("out_f32") is not None else 0)
                                + (8.0 * pad_cols(l["N"]) / 64 if (l.get("sign_bits_out") is not None or l.get("sign_bits_in") is not None) else 0))
        K = l["N"]
    call("papr_stack_bf16_ex", x.data_ptr(), K0, ctypes.cast(arr, ctypes.c_void_p), len(layers), x.rows_pad, float(slope),
         int(max_ctas), flops=flops, nbytes=nbytes)


_BWD_WORKSPACE = {}


def _bwd_workspace(dev):
    """The L2-resident hand-over ring of papr_stack_bwd_fused, one per device (launches on a stream reuse it in order)."""
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _BWD_WORKSPACE:
        n = int(lib().papr_stack_bwd_workspace_bytes())
        _BWD_WORKSPACE[key] = torch.empty(n + 1024, dtype=torch.uint8, device=dev)
    buf = _BWD_WORKSPACE[key]
    off = (-buf.data_ptr()) % 1024
    return buf.data_ptr() + off, buf.numel() - off


def stack_bwd_fused(dz, K0, layers, wlayers, producer_ctas=0):
    """Backward of a whole MLP stack in one launch (papr_stack_bwd_fused).  `layers`: the dgrad layer list of stack_bf16
    (dgrad order; out_blocked needed for the last one only); `wlayers` in forward order: dicts x (Blocked), gw (fp32
    (n_out, n_in) tensor), n_out, n_in."""
    import ctypes
    from ._lib import StackLayer, WgradLayer
    n = len(layers)
    arr = (StackLayer * n)()
    warr = (WgradLayer * n)()
    flops, nbytes = 0.0, 2.0 * dz.rows_pad * dz.cols_pad
    K = K0
    for i, l in enumerate(layers):
        def ptr(key):
            v = l.get(key)
            return v.data_ptr() if v is not None else None
        arr[i].w_image = ptr("w_image")
        arr[i].out_blocked = ptr("out_blocked") if i == n - 1 else None
        arr[i].sign_bits_in, arr[i].colsum = ptr("sign_bits_in"), ptr("colsum")
        arr[i].N, arr[i].act = int(l["N"]), 0
        img = l["w_image"]
        arr[i].w_replicas = img.shape[0] if img.dim() == 2 else 1
        arr[i].w_replica_stride = img.stride(0) if img.dim() == 2 else 0
        flops += 2.0 * dz.rows * l["N"] * K
        nbytes += dz.rows_pad * (8.0 * pad_cols(l["N"]) / 64 if l.get("sign_bits_in") is not None else 0)
        K = l["N"]
    nbytes += 2.0 * dz.rows_pad * pad_cols(layers[-1]["N"])
    for i, w in enumerate(wlayers):
        assert w["x"].rows_pad == dz.rows_pad and w["gw"].dtype == torch.float32
        warr[i].x_blocked, warr[i].x_cols = w["x"].data_ptr(), w["x"].cols_pad
        warr[i].gw, warr[i].ldw = w["gw"].data_ptr(), w["gw"].stride(0)
        warr[i].n_out, warr[i].n_in = int(w["n_out"]), int(w["n_in"])
        flops += 2.0 * dz.rows * w["n_out"] * w["n_in"]
        nbytes += 2.0 * dz.rows_pad * w["x"].cols_pad
    ws, ws_bytes = _bwd_workspace(dz.buf.device)
    call("papr_stack_bwd_fused", dz.data_ptr(), K0, ctypes.cast(arr, ctypes.c_void_p), ctypes.cast(warr, ctypes.c_void_p), n,
         dz.rows_pad, int(producer_ctas), ws, ws_bytes, flops=flops, nbytes=nbytes, kernels=1)


# --------------------------------------------------------------------------- bookkeeping / ray generation
def knn(points, queries, k):
    """Stage a14 (reference models/utils.py:21,73: KDTree.query): (dist float64 (Q,k), idx int64 (Q,k)) of the k nearest
    `points` for every query, ascending, ties by smaller index.  CUDA fp32 inputs, exact float64 distances."""
    pts, qs = _f32c(points), _f32c(queries)
    P, Q = pts.shape[0], qs.shape[0]
    if not (1 <= k <= 32) or k > P:
        raise ValueError(f"knn needs 1 <= k <= min(32, P) (k={k}, P={P})")
    dist = torch.empty((Q, k), dtype=torch.float64, device=pts.device)
    idx = torch.empty((Q, k), dtype=torch.int32, device=pts.device)
    with torch.cuda.device(pts.device):
        call("papr_knn", pts.data_ptr(), P, qs.data_ptr(), Q, k, dist.data_ptr(), idx.data_ptr(), flops=8.0 * P * Q,
             nbytes=12.0 * (P + Q) + 12.0 * Q * k)
    return dist, idx.long()


def prune_compact(points, influ, feats, thresh, keep_less=False):
    """Stage a14 (reference models/model.py:335-358): rows with influence > thresh (or < thresh) kept in order.
    Returns (points, influ (n,1), feats or None, n_kept int) -- one device->host read of the count."""
    pts, inf = _f32c(points), _f32c(influ).reshape(-1)
    fts = _f32c(feats) if feats is not None else None
    P = pts.shape[0]
    F = fts.shape[1] if fts is not None else 0
    dev = pts.device
    out_p, out_i = torch.empty_like(pts), torch.empty_like(inf)
    out_f = torch.empty_like(fts) if fts is not None else None
    scratch = torch.empty(((P + 255) // 256 + 1,), dtype=torch.int32, device=dev)
    n = torch.zeros((1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        call("papr_prune_compact", pts.data_ptr(), inf.data_ptr(), fts.data_ptr() if fts is not None else None, P, F,
             float(thresh), int(keep_less), out_p.data_ptr(), out_i.data_ptr(), out_f.data_ptr() if out_f is not None else None,
             scratch.data_ptr(), n.data_ptr(), nbytes=8.0 * P * (4 + F), kernels=3)
    kept = int(n.item())
    return out_p[:kept], out_i[:kept].reshape(-1, 1), (out_f[:kept] if out_f is not None else None), kept


def generate_rays(c2w, H, W, focal_x, focal_y=None, window=None, coord_scale=1.0):
    """SURVEY 8(f3) (reference dataset/utils.py:81-96 get_rays): rays of the pixel window (h0, h1, w0, w1) of H x W views
    on the device.  c2w (N,4,4) CUDA fp32 -> rays_o (N,3) = coord_scale * c2w[:, :3, 3], rays_d (N,h,w,3) unit norm."""
    c = _f32c(c2w).reshape(-1, 4, 4)
    h0, h1, w0, w1 = window if window is not None else (0, H, 0, W)
    N = c.shape[0]
    rays_o = torch.empty((N, 3), dtype=torch.float32, device=c.device)
    rays_d = torch.empty((N, h1 - h0, w1 - w0, 3), dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        call("papr_generate_rays", c.data_ptr(), N, H, W, float(focal_x), float(focal_y if focal_y is not None else focal_x),
             h0, w0, h1 - h0, w1 - w0, float(coord_scale), rays_o.data_ptr(), rays_d.data_ptr(), nbytes=12.0 * rays_d.numel() / 3)
    return rays_o, rays_d
