"""GPU: the fused MLP-stack kernel (cta_group::2) must reproduce the per-layer tcgen05 kernel bit for bit
(both round activations to bf16 between layers and accumulate in fp32)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _weights(dims, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ws, bs = [], []
    for k, n in zip(dims[:-1], dims[1:]):
        ws.append(torch.randn(n, k, device="cuda", generator=g) * (2.0 / k) ** 0.5)
        bs.append(torch.randn(n, device="cuda", generator=g) * 0.1)
    return ws, bs


@pytest.mark.parametrize("rows", [128, 512, 128 * 7, 128 * 601 + 5])
@pytest.mark.parametrize("dims", [(117, 256, 256, 256, 256, 256), (142, 256, 256, 256, 256, 256, 256, 256, 32), (39, 256, 256)])
def test_stack_forward_matches_per_layer(rows, dims):
    from papr_b200 import ops
    ws, bs = _weights(dims, rows + len(dims))
    g = torch.Generator(device="cuda").manual_seed(1)
    x = ops.Blocked.from_f32(torch.randn(rows, dims[0], device="cuda", generator=g))
    n = len(ws)
    K0 = (dims[0] + 15) // 16 * 16
    # reference: one launch per layer
    h, ref_out, ref_bits = x, [], []
    for i, (w, b) in enumerate(zip(ws, bs)):
        last = i == n - 1
        N, K = (w.shape[0] + 31) // 32 * 32, (w.shape[1] + 15) // 16 * 16
        yb, yf, bits = ops.linear_bf16(h, ops.pack_weight(w, N, K), N, K, bias=b, act=not last, slope=0.2,
                                       out_blocked=True, out_f32=last, sign_bits_out=not last)
        ref_out.append(yb); ref_bits.append(bits); h = yb
    ref_f32 = yf
    # fused
    layers, outs, bits_l = [], [], []
    for i, (w, b) in enumerate(zip(ws, bs)):
        last = i == n - 1
        N, K = (w.shape[0] + 31) // 32 * 32, (w.shape[1] + 15) // 16 * 16
        ob = ops.Blocked(rows, N, "cuda")
        bt = None if last else torch.zeros((x.rows_pad, ops.pad_cols(N) // 64), dtype=torch.int64, device="cuda")
        outs.append(ob); bits_l.append(bt)
        layers.append(dict(w_image=ops.pack_weight(w, N, K), N=N, bias=b, act=not last, out_blocked=ob, sign_bits_out=bt,
                           out_f32=torch.zeros((x.rows_pad, N), device="cuda") if last else None))
    ops.stack_bf16(x, K0, layers, slope=0.2)
    torch.cuda.synchronize()
    for i in range(n):
        a, b = outs[i].to_f32(), ref_out[i].to_f32()
        assert torch.equal(a, b), f"layer {i}: max diff {(a - b).abs().max().item()}"
        if bits_l[i] is not None:
            assert torch.equal(ops.sign_bits_rowmajor(bits_l[i])[:rows], ops.sign_bits_rowmajor(ref_bits[i])[:rows]), f"sign bits of layer {i}"
    assert torch.equal(layers[-1]["out_f32"][:rows], ref_f32[:rows])


@pytest.mark.parametrize("rows", [128 * 3, 128 * 300])
def test_stack_dgrad_matches_per_layer(rows):
    """Backward direction of an 8-layer value-like stack: masks from sign bits, bias-gradient column sums, dZ stash."""
    from papr_b200 import ops
    dims = (142, 256, 256, 256, 256, 256, 256, 256, 32)
    ws, _ = _weights(dims, 7)
    g = torch.Generator(device="cuda").manual_seed(2)
    dz0 = ops.Blocked.from_f32(torch.randn(rows, 32, device="cuda", generator=g))
    rows_pad = dz0.rows_pad
    bits = [torch.randint(-2 ** 62, 2 ** 62, (rows_pad, 4), device="cuda", generator=g, dtype=torch.int64) for _ in range(7)]
    # per-layer reference (layers 7..0)
    dz, ref, ref_cs = dz0, [], []
    for i in range(7, -1, -1):
        w = ws[i]
        Kd = (w.shape[0] + 15) // 16 * 16
        Nd = 256 if i > 0 else 192
        cs = torch.zeros(Nd, device="cuda") if i > 0 else None
        dz, _, _ = ops.linear_bf16(dz, ops.pack_weight(w, Nd, Kd, transpose=True), Nd, Kd,
                                   sign_bits_in=bits[i - 1] if i > 0 else None, slope=0.2, colsum=cs)
        ref.append(dz); ref_cs.append(cs)
    layers, outs, css = [], [], []
    for i in range(7, -1, -1):
        w = ws[i]
        Kd = (w.shape[0] + 15) // 16 * 16
        Nd = 256 if i > 0 else 192
        ob = ops.Blocked(rows, Nd, "cuda")
        cs = torch.zeros(Nd, device="cuda") if i > 0 else None
        outs.append(ob); css.append(cs)
        layers.append(dict(w_image=ops.pack_weight(w, Nd, Kd, transpose=True), N=Nd, out_blocked=ob,
                           sign_bits_in=bits[i - 1] if i > 0 else None, colsum=cs))
    ops.stack_bf16(dz0, 32, layers, slope=0.2)
    torch.cuda.synchronize()
    for j in range(8):
        assert torch.equal(outs[j].to_f32(), ref[j].to_f32()), f"dgrad step {j}"
        if css[j] is not None:
            assert (css[j] - ref_cs[j]).abs().max().item() <= 1e-3 * max(1.0, ref_cs[j].abs().max().item())


@pytest.mark.parametrize("rows", [128 * 5, 128 * 4 * 74 * 9 + 77])
@pytest.mark.parametrize("dims,last_f32", [((142, 256, 256, 256, 256, 256, 256, 256, 32), True), ((117, 256, 256, 256, 256, 256), False),
                                           ((39, 256, 256, 256), True)])
def test_stack_inference_mode_at_scale(rows, dims, last_f32):
    """Inference form of the fused stack: nothing leaves the SM except the last layer's output (fp32 row-major for the
    value / query stacks, a tile for the key stack), over many quads per cluster.  Round 1 only ever ran this form with a
    tile output on the last layer; with a pure fp32 output the stash-writer thread had nothing to wait for, ran ahead and
    broke the slot-release barrier (a hang at full frame size).  Must equal the per-layer kernels bit for bit."""
    from papr_b200 import ops
    ws, bs = _weights(dims, rows % 1000 + len(dims))
    g = torch.Generator(device="cuda").manual_seed(2)
    x = ops.Blocked.from_f32(torch.randn(rows, dims[0], device="cuda", generator=g))
    n = len(ws)
    K0 = (dims[0] + 15) // 16 * 16
    h = x
    for i, (w, b) in enumerate(zip(ws, bs)):
        last = i == n - 1
        N, K = (w.shape[0] + 31) // 32 * 32, (w.shape[1] + 15) // 16 * 16
        yb, yf, _ = ops.linear_bf16(h, ops.pack_weight(w, N, K), N, K, bias=b, act=not last, slope=0.0,
                                    out_blocked=not (last and last_f32), out_f32=last and last_f32)
        h = yb
    layers = []
    for i, (w, b) in enumerate(zip(ws, bs)):
        last = i == n - 1
        N, K = (w.shape[0] + 31) // 32 * 32, (w.shape[1] + 15) // 16 * 16
        spec = dict(w_image=ops.pack_weight(w, N, K, replicas=ops.WEIGHT_REPLICAS), N=N, bias=b, act=not last)
        if last and last_f32:
            spec["out_f32"] = torch.zeros((x.rows_pad, N), device="cuda")
        elif last:
            spec["out_blocked"] = ops.Blocked(rows, N, "cuda")
        layers.append(spec)
    for _ in range(3):
        ops.stack_bf16(x, K0, layers, slope=0.0)
    torch.cuda.synchronize()
    if last_f32:
        assert torch.equal(layers[-1]["out_f32"][:rows], yf[:rows])
    else:
        assert torch.equal(layers[-1]["out_blocked"].to_f32(), yb.to_f32())


@pytest.mark.parametrize("rows", [128 * 5, 128 * 300 + 9, 128 * 4 * 74 * 3 + 77])
@pytest.mark.parametrize("dims,relu", [((117, 256, 256, 256, 256, 256, 256), True), ((142, 256, 256, 256, 256, 256, 256, 256, 32), True),
                                       ((39, 256, 256), True), ((117, 256, 256, 256), False)])
def test_fused_backward_matches_two_kernel_backward(rows, dims, relu, monkeypatch):
    """papr_stack_bwd_fused (dgrad CTAs hand dZ tiles to weight-gradient CTAs through L2, one launch) against the dgrad stack
    launch followed by per-layer papr_wgrad_bf16: the input gradient bit for bit, weight / bias gradients up to the fp32
    summation order."""
    from papr_b200 import attention as A, ops
    ws, bs = _weights(dims, rows % 1000 + len(dims))
    g = torch.Generator(device="cuda").manual_seed(3)
    x = ops.Blocked.from_f32(torch.randn(rows, dims[0], device="cuda", generator=g))
    slope = 0.0 if relu else None
    in_pad = ops.pad_cols(dims[0])
    inputs, bits, out = A._stack_forward_fused(x, ws, bs, slope, dims[0], True, False)
    dz = ops.Blocked.from_f32(torch.randn(rows, dims[-1], device="cuda", generator=g) * 0.1)
    res = []
    for fused in (False, True):
        monkeypatch.setattr(A, "BWD_FUSED", fused)
        d_in, gWs, gbs = A._stack_backward(dz, inputs, bits, ws, slope, None, dims[0], in_pad)
        torch.cuda.synchronize()
        res.append((d_in.to_f32(), gWs, gbs))
    (d0, gW0, gb0), (d1, gW1, gb1) = res
    assert torch.equal(d0, d1), f"input gradient: max diff {(d0 - d1).abs().max().item()}"
    for i, (a, b) in enumerate(zip(gW0, gW1)):
        scale = max(a.abs().max().item(), 1e-6)
        assert (a - b).abs().max().item() <= 1e-4 * scale, f"gW[{i}]: {(a - b).abs().max().item()} of {scale}"
        assert a.abs().max().item() > 0
    for i, (a, b) in enumerate(zip(gb0, gb1)):
        if a is None:
            assert b is None
            continue
        assert (a - b).abs().max().item() <= 1e-3 * max(1.0, a.abs().max().item()), f"gb[{i}]"
