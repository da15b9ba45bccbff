"""GPU numerics of the tcgen05 building blocks (linear layer, weight gradient, layout helpers) against a plain
PyTorch fp32 evaluation of the same bf16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.bfloat16().float()


def test_blocked_roundtrip():
    from papr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(300, 142, device="cuda", generator=g)
    b = ops.Blocked.from_f32(x)
    assert b.rows_pad == 384 and b.cols_pad == 192
    y = b.to_f32()
    assert torch.equal(y, _bf(x))
    full = b.to_f32(384, 192)
    assert float(full[300:].abs().max()) == 0 and float(full[:, 142:].abs().max()) == 0


@pytest.mark.parametrize("N,K", [(256, 256), (256, 128), (256, 144), (32, 256), (256, 64), (64, 256), (128, 256),
                                 (192, 256), (256, 32), (32, 16)])
@pytest.mark.parametrize("rows", [128, 1000, 128 * 149 + 5])
def test_linear_forward(N, K, rows):
    from papr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N * 7 + K)
    x = torch.randn(rows, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    xb = ops.Blocked.from_f32(x)
    img = ops.pack_weight(w, N, K)
    yb, yf, bits = ops.linear_bf16(xb, img, N, K, bias=bias, act=True, slope=0.2, out_f32=True, sign_bits_out=True)
    torch.cuda.synchronize()
    pre = _bf(x) @ _bf(w).t() + bias
    want = torch.where(pre > 0, pre, 0.2 * pre)
    got32 = yf[:rows]
    err32 = (got32 - want).abs().max().item()
    assert err32 < 2e-3, f"fp32 output err {err32}"
    got = yb.to_f32()
    err = ((got - want).abs() / (want.abs() + 1)).max().item()
    assert err < 1e-2, f"bf16 output err {err}"
    # sign bits
    j = torch.arange(N, device="cuda")
    got_bits = (ops.sign_bits_rowmajor(bits)[:rows][:, j // 64] >> (j % 64)) & 1
    safe = pre.abs() > 1e-3
    assert torch.equal(got_bits[safe], (pre > 0).long()[safe])


def test_linear_dgrad_mask_and_colsum():
    from papr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    rows, N, K = 1000, 256, 192     # forward layer: K inputs -> N outputs; dgrad maps N -> K
    dz = torch.randn(rows, N, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / N ** 0.5
    mask = (torch.rand(rows, K, device="cuda", generator=g) > 0.5)
    bits = torch.zeros(ops.pad_rows(rows), ops.pad_cols(K) // 64, dtype=torch.int64, device="cuda")
    j = torch.arange(K, device="cuda")
    for wd in range(K // 64):
        sel = mask[:, wd * 64:(wd + 1) * 64].long()
        sh = torch.arange(64, device="cuda")
        # int64 holds bit 63 as the sign bit
        val = (sel[:, :63] << sh[:63]).sum(1) + torch.where(sel[:, 63] > 0, torch.tensor(-2 ** 63, device="cuda"), torch.tensor(0, device="cuda"))
        bits[:rows, wd] = val
    bits = ops.sign_bits_from_rowmajor(bits)
    img = ops.pack_weight(w, K, N, transpose=True)     # image of W^T: "out" = K, reduction = N
    colsum = torch.zeros(K, device="cuda")
    yb, _, _ = ops.linear_bf16(ops.Blocked.from_f32(dz), img, K, N, sign_bits_in=bits, slope=0.0, colsum=colsum)
    torch.cuda.synchronize()
    want = (_bf(dz) @ _bf(w)) * mask
    got = yb.to_f32()
    assert ((got - want).abs() / (want.abs() + 1)).max().item() < 1e-2
    assert (colsum - got.sum(0)).abs().max().item() < 1e-2 * rows ** 0.5


@pytest.mark.parametrize("A,B,tr", [(256, 256, False), (256, 192, False), (256, 128, False), (256, 64, False),
                                    (256, 32, True), (128, 256, False)])
@pytest.mark.parametrize("rows", [128, 64 * 301])
def test_wgrad(A, B, tr, rows):
    from papr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(A + B)
    a = torch.randn(rows, A, device="cuda", generator=g)
    b = torch.randn(rows, B, device="cuda", generator=g)
    out = torch.zeros((B, A) if tr else (A, B), device="cuda")
    out += 1.0    # accumulation semantics
    ops.wgrad_bf16(ops.Blocked.from_f32(a), ops.Blocked.from_f32(b), out, A, B, transpose_out=tr)
    torch.cuda.synchronize()
    want = _bf(a).t() @ _bf(b)
    if tr:
        want = want.t()
    err = (out - 1.0 - want).abs().max().item()
    assert err < 2e-3 * rows ** 0.5, err


@pytest.mark.parametrize("A,rows", [(256, 128 * 301 + 17), (256, 64), (100, 5000)])
def test_wgrad_bias_column_sums(A, rows):
    """papr_wgrad_bias_bf16: the weight gradient unchanged and, on the side, the column sums of A (bias gradient)."""
    from papr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(A + rows)
    a = torch.randn(rows, A, device="cuda", generator=g)
    b = torch.randn(rows, 192, device="cuda", generator=g)
    ab, bb = ops.Blocked.from_f32(a), ops.Blocked.from_f32(b)
    ref = ops.wgrad_bf16(ab, bb, torch.zeros((A, 192), device="cuda"), A, 192)
    cs = torch.full((A,), 2.0, device="cuda")
    out = ops.wgrad_bf16(ab, bb, torch.zeros((A, 192), device="cuda"), A, 192, a_colsum=cs)
    torch.cuda.synchronize()
    assert (out - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    want = _bf(a).double().sum(0)
    assert (cs.double() - 2.0 - want).abs().max().item() <= 1e-4 * rows ** 0.5 + 1e-3


def test_linear_addend_split_k():
    """Skip-connection layers (mlp.py:54-55): [h, inp] W^T = h W1^T + inp W2^T through the fp32 addend."""
    from papr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    rows = 1000
    h = torch.randn(rows, 256, device="cuda", generator=g)
    inp = torch.randn(rows, 142, device="cuda", generator=g)
    w = torch.randn(256, 398, device="cuda", generator=g) / 20
    b = torch.randn(256, device="cuda", generator=g)
    hb, ib = ops.Blocked.from_f32(h), ops.Blocked.from_f32(inp)
    _, part, _ = ops.linear_bf16(ib, ops.pack_weight(w[:, 256:], 256, 144), 256, 144, out_blocked=False, out_f32=True)
    _, y, _ = ops.linear_bf16(hb, ops.pack_weight(w[:, :256], 256, 256), 256, 256, bias=b, act=True, slope=0.2,
                              out_blocked=False, out_f32=True, addend=part)
    pre = _bf(torch.cat([h, inp], 1)) @ _bf(w).t() + b
    want = torch.where(pre > 0, pre, 0.2 * pre)
    assert (y[:rows] - want).abs().max().item() < 3e-3


# ----------------------------------------------------------------------------- fp32 parity mode on the same kernels
@pytest.mark.parametrize("n_in,n_out,slope", [(117, 256, 0.0), (256, 256, 0.2), (398, 256, None), (256, 32, None),
                                              (39, 256, 0.0), (256, 3, None)])
def test_split_linear_is_fp32_accurate(n_in, n_out, slope):
    """papr_b200.split_gemm: three-way bf16 split through papr_linear_bf16 / papr_wgrad_bf16 == fp32 Linear (forward,
    data gradient, weight gradient, bias gradient), judged against a float64 evaluation and against torch's own fp32."""
    from papr_b200 import split_gemm
    g = torch.Generator(device="cuda").manual_seed(n_in * 3 + n_out)
    rows = 1000
    x = (torch.randn(rows, n_in, device="cuda", generator=g) * 3).requires_grad_(True)
    w = (torch.randn(n_out, n_in, device="cuda", generator=g) / n_in ** 0.5).requires_grad_(True)
    b = torch.randn(n_out, device="cuda", generator=g).requires_grad_(True)
    gy = torch.randn(rows, n_out, device="cuda", generator=g)

    def act(t):
        return t if slope is None else torch.nn.functional.leaky_relu(t, slope)

    y = split_gemm.SplitLinearFn.apply(x, w, b, slope)
    y.backward(gy)
    got = [y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone()]
    outs = {}
    for dt in (torch.float64, torch.float32):
        xd, wd, bd = (t.detach().to(dt).requires_grad_(True) for t in (x, w, b))
        yd = act(xd @ wd.t() + bd)
        yd.backward(gy.to(dt))
        outs[dt] = [yd.detach(), xd.grad, wd.grad, bd.grad]
    for name, a, t64, t32 in zip(("y", "gx", "gw", "gb"), got, outs[torch.float64], outs[torch.float32]):
        scale = float(t64.abs().max())
        e_ours = float((a.double() - t64).abs().max()) / scale
        e_torch = float((t32.double() - t64).abs().max()) / scale
        print(f"split {n_in}->{n_out} {name}: ours {e_ours:.2e}  torch fp32 {e_torch:.2e}")
        assert e_ours <= max(2e-6, 4 * e_torch), (name, e_ours, e_torch)
