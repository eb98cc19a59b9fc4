"""GPU: the tensor-core (tcgen05) gradient kernels of the training step -- zs_gemm_tn_tc, zs_conv2d_nhwc_wgrad_tc,
zs_conv2d_nhwc_dgrad_tc, the data-gradient GEMM through zs_gemm_tc_f32 -- against fp64 torch on the same seeded inputs
(what torch autograd computes for nn.Linear / nn.Conv2d in the reference's `loss.backward()`, model/shape_engine.py:268-272),
and the two-pass zs_mha_bwd_f32 at the ViT-B head geometry.

Tolerances: "bf16x3" (split operands, ~2^-16 per product) is held to 5e-5 normwise; "bf16" (single pass, the mixed-precision
mode of BASELINE config 3) to 1e-2 normwise."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {"bf16x3": 5e-5, "bf16": 1e-2}


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture()
def tc_ops(cuda):
    from zeroshape_b200 import ops
    if ops.device_cc() != 100:
        pytest.skip("tcgen05 kernels need sm_100")
    saved = (ops.TRAIN_ENGINE, ops.TRAIN_PRECISION, ops.TN_LAYOUT)
    ops.TRAIN_ENGINE = "tc"
    yield ops
    ops.TRAIN_ENGINE, ops.TRAIN_PRECISION, ops.TN_LAYOUT = saved


def test_tn_exact_on_integer_operands(tc_ops, cuda):
    """The TN kernel's MN-major shared-memory descriptors on an exact-arithmetic case (small integers: bf16 products and fp32
    sums are exact, so a wrong descriptor or swizzle shows up as a large error, not as rounding).  During bring-up the same case
    also ran through a K-major variant with transposing producers (exact as well; since removed)."""
    ops = tc_ops
    g = torch.Generator().manual_seed(1)
    for M, N, K in ((448, 256, 512), (1000, 130, 300), (31, 128, 256)):
        a = torch.randint(-3, 4, (M, N), generator=g).float()
        b = torch.randint(-3, 4, (M, K), generator=g).float()
        out = ops.gemm_tn(a.to(cuda), b.to(cuda), tc=True)
        torch.cuda.synchronize()
        assert (out.double().cpu() - a.double().T @ b.double()).abs().max().item() == 0.0, (M, N, K)


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("M,N,K", [(1000, 70, 259), (6304, 768, 768), (130, 256, 1024), (70000, 64, 515), (4096, 3072, 768)])
def test_gemm_tn_tc(tc_ops, cuda, precision, M, N, K):
    ops = tc_ops
    ops.TRAIN_PRECISION = precision
    g = torch.Generator().manual_seed(M + N + K)
    a, b = torch.randn(M, N, generator=g), torch.randn(M, K, generator=g)
    ref = a.double().T @ b.double()
    out = ops.gemm_tn(a.to(cuda), b.to(cuda))
    assert _rel(out, ref) < TOL[precision], _rel(out, ref)
    acc = torch.ones(N, K, device=cuda)
    ops.gemm_tn(a.to(cuda), b.to(cuda), out=acc, accumulate=True)
    assert _rel(acc, ref + 1) < TOL[precision]
    # row-strided views (column slices of wider matrices), as the decoder backward passes them
    wide_a = torch.randn(M, N + 24, generator=g).to(cuda)
    wide_b = torch.randn(M, K + 8, generator=g).to(cuda)
    out2 = ops.gemm_tn(wide_a[:, 8:8 + N], wide_b[:, 4:4 + K])
    assert _rel(out2, wide_a[:, 8:8 + N].double().T @ wide_b[:, 4:4 + K].double()) < TOL[precision]


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("cfg", [
    dict(B=2, H=28, W=28, Cin=64, Cout=128, k=3, s=1, pad=(1, 1, 1, 1)),
    dict(B=3, H=15, W=17, Cin=128, Cout=64, k=3, s=2, pad=(1, 1, 1, 1)),
    dict(B=2, H=14, W=14, Cin=256, Cout=96, k=1, s=2, pad=(0, 0, 0, 0)),
    dict(B=1, H=30, W=30, Cin=32, Cout=32, k=3, s=1, pad=(1, 1, 1, 1)),
    dict(B=2, H=23, W=23, Cin=64, Cout=64, k=3, s=2, pad=(0, 1, 0, 1)),          # timm "SAME" padding of an odd map
    dict(B=2, H=32, W=30, Cin=3, Cout=64, k=7, s=2, pad=(3, 3, 3, 3)),           # XYZ stem: FFMA forward / wgrad, warp-per-pixel dgrad
])
def test_conv_gradients_tc(tc_ops, cuda, precision, cfg):
    ops = tc_ops
    ops.TRAIN_PRECISION = precision
    B, H, W, Cin, Cout, k, s, pad = (cfg[n] for n in ("B", "H", "W", "Cin", "Cout", "k", "s", "pad"))
    g = torch.Generator().manual_seed(H * 100 + Cin)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    y = F.conv2d(F.pad(xr, (pad[2], pad[3], pad[0], pad[1])), wr, stride=s)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy.double())
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(cuda)
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous().to(cuda)
    w_ohwi = w.permute(0, 2, 3, 1).contiguous().to(cuda)
    # forward on the same engine (im2col gather), for completeness of the triple
    y_tc = ops.conv2d_nhwc(x_nhwc, w_ohwi, None, s, pad, tc=True, precision=precision)
    assert _rel(y_tc.permute(0, 3, 1, 2), y) < TOL[precision]
    dx = ops.conv2d_nhwc_dgrad(dy_nhwc, w_ohwi, x_nhwc.shape, s, pad)
    assert _rel(dx.permute(0, 3, 1, 2), xr.grad) < TOL[precision], _rel(dx.permute(0, 3, 1, 2), xr.grad)
    dw = ops.conv2d_nhwc_wgrad(x_nhwc, dy_nhwc, k, k, s, pad)
    assert _rel(dw.permute(0, 3, 1, 2), wr.grad) < TOL[precision], _rel(dw.permute(0, 3, 1, 2), wr.grad)
    # and the FFMA kernels they replace agree with them
    dx32 = ops.conv2d_nhwc_dgrad(dy_nhwc, w_ohwi, x_nhwc.shape, s, pad, tc=False)
    dw32 = ops.conv2d_nhwc_wgrad(x_nhwc, dy_nhwc, k, k, s, pad, tc=False)
    assert _rel(dx, dx32) < TOL[precision] and _rel(dw, dw32) < TOL[precision]


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_linear_triple_tc(tc_ops, cuda, precision):
    """nn.Linear forward / dX / dW on the tensor cores (ViT-B fc1 geometry, 2 images of 197 tokens)."""
    ops = tc_ops
    ops.TRAIN_PRECISION = precision
    g = torch.Generator().manual_seed(7)
    x, w, bias = torch.randn(394, 768, generator=g), torch.randn(3072, 768, generator=g) * 0.03, torch.randn(3072, generator=g)
    dy = torch.randn(394, 3072, generator=g)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    y = F.linear(xr, wr, bias.double())
    y.backward(dy.double())
    assert _rel(ops.train_linear(x.to(cuda), w.to(cuda), bias.to(cuda)), y) < TOL[precision]
    assert _rel(ops.train_dgrad(dy.to(cuda), w.to(cuda)), xr.grad) < TOL[precision]
    assert _rel(ops.gemm_tn(dy.to(cuda), x.to(cuda)), wr.grad) < TOL[precision]


@pytest.mark.parametrize("B,T,heads,hd", [(2, 197, 12, 64), (3, 197, 8, 32), (1, 50, 4, 16)])
def test_mha_bwd_two_pass(cuda, B, T, heads, hd):
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(T + hd)
    C = heads * hd
    qkv = torch.randn(B, T, 3 * C, generator=g) * 0.7
    do = torch.randn(B, T, C, generator=g)
    qr = qkv.double().requires_grad_(True)
    q, k, v = qr.reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    o = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1) @ v
    o.transpose(1, 2).reshape(B, T, C).backward(do.double())
    got = ops.mha_bwd(qkv.to(cuda), do.to(cuda), heads)
    assert _rel(got, qr.grad) < 2e-5, _rel(got, qr.grad)


@pytest.mark.parametrize("B,T,heads,hd", [(2, 197, 12, 64), (3, 197, 8, 32), (1, 50, 4, 64), (2, 208, 2, 32), (1, 129, 1, 64), (1, 1, 1, 32)])
def test_mha_bwd_on_the_tensor_cores(cuda, B, T, heads, hd):
    """zs_mha_bwd_tc_f32 (S / dP and their transposes, dQ / dK / dV as tcgen05 MMAs, single fp16 pass): the precision class of the
    bf16 training GEMMs, against fp64 autograd and next to the fp32 FFMA kernel."""
    from zeroshape_b200 import ops
    if ops.device_cc() != 100:
        pytest.skip("tcgen05 needs sm_100")
    g = torch.Generator().manual_seed(T * 3 + hd + heads)
    C = heads * hd
    qkv = torch.randn(B, T, 3 * C, generator=g) * 0.7
    do = torch.randn(B, T, C, generator=g)
    qr = qkv.double().requires_grad_(True)
    q, k, v = qr.reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    o = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1) @ v
    o.transpose(1, 2).reshape(B, T, C).backward(do.double())
    got = ops.mha_bwd(qkv.to(cuda), do.to(cuda), heads, tc=True)
    ffma = ops.mha_bwd(qkv.to(cuda), do.to(cuda), heads, tc=False)
    parts = {n: _rel(got[..., i * C:(i + 1) * C], qr.grad[..., i * C:(i + 1) * C]) for i, n in enumerate("qkv")}
    print(f"mha_bwd_tc B={B} T={T} heads={heads} hd={hd}: rel err dq {parts['q']:.2e} dk {parts['k']:.2e} dv {parts['v']:.2e}; "
          f"FFMA kernel {_rel(ffma, qr.grad):.2e}")
    assert max(parts.values()) < 4e-3, parts
    assert torch.equal(got, ops.mha_bwd(qkv.to(cuda), do.to(cuda), heads, tc=True))        # deterministic


@pytest.mark.parametrize("B,P,L", [(2, 4096, 197), (1, 1000, 197), (3, 130, 50), (1, 1, 197), (1, 300, 208)])
def test_point_attention_bwd_on_the_tensor_cores(cuda, B, P, L):
    """zs_point_attention_bwd_tc_f32 (decoder training: points -> latents + the point's own key) against fp64 autograd of
    ImplFuncAttention's point rows, next to the fp32 FFMA kernel; row-strided latent K / V views as the decoder passes them."""
    from zeroshape_b200 import ops
    if ops.device_cc() != 100:
        pytest.skip("tcgen05 needs sm_100")
    H, hd = 8, 32
    C = H * hd
    g = torch.Generator().manual_seed(P + L)
    qkv = torch.randn(B, P, 3 * C, generator=g) * 0.7
    lat = torch.randn(B, L, 3 * C, generator=g) * 0.7            # latent qkv buffer: K / V are column views (row stride 3C)
    do = torch.randn(B, P, C, generator=g)
    qr, lr = qkv.double().requires_grad_(True), lat.double().requires_grad_(True)
    q, k, v = [t.reshape(B, P, H, hd).permute(0, 2, 1, 3) for t in (qr[..., :C], qr[..., C:2 * C], qr[..., 2 * C:])]
    kl = lr[..., C:2 * C].reshape(B, L, H, hd).permute(0, 2, 1, 3)
    vl = lr[..., 2 * C:].reshape(B, L, H, hd).permute(0, 2, 1, 3)
    sc = torch.cat([q @ kl.transpose(-2, -1), (q * k).sum(-1, keepdim=True)], -1) * hd ** -0.5
    a = sc.softmax(-1)
    o = (a[..., :L] @ vl + a[..., L:] * v).permute(0, 2, 1, 3).reshape(B, P, C)
    o.backward(do.double())
    lat_d = lat.to(cuda)
    out = o.detach().float().to(cuda)
    res = {}
    for tc in (True, False):
        dqkv, dk, dv = ops.point_attention_bwd(qkv.to(cuda), lat_d[..., C:2 * C], lat_d[..., 2 * C:], out, do.to(cuda), H, tc=tc)
        res[tc] = (_rel(dqkv[..., :C], qr.grad[..., :C]), _rel(dqkv[..., C:2 * C], qr.grad[..., C:2 * C]),
                   _rel(dqkv[..., 2 * C:], qr.grad[..., 2 * C:]), _rel(dk, lr.grad[..., C:2 * C]), _rel(dv, lr.grad[..., 2 * C:]))
    print(f"point_attention_bwd B={B} P={P} L={L}: rel err (dq, dk_self, dv_self, dk_lat, dv_lat) tcgen05 "
          + " ".join(f"{e:.1e}" for e in res[True]) + " | FFMA " + " ".join(f"{e:.1e}" for e in res[False]))
    assert max(res[True]) < 4e-3 and max(res[False]) < 2e-5



@pytest.mark.parametrize("B,P,L", [(2, 300, 197), (1, 128, 208), (3, 77, 50), (1, 4096, 197)])
def test_point_attention_fwd_on_the_tensor_cores(cuda, B, P, L):
    """zs_point_attention_tc_f32 (decoder training forward: points -> latents + the point's own key) against the fp64 formula of
    ImplFuncAttention's point rows: split fp16 (fp32-grade, like the FFMA kernel) and the single fp16 pass of the bf16 training mode."""
    from zeroshape_b200 import ops
    if ops.device_cc() != 100:
        pytest.skip("tcgen05 needs sm_100")
    H, hd = 8, 32
    C = H * hd
    g = torch.Generator().manual_seed(P + L)
    qkv = torch.randn(B, P, 3 * C, generator=g) * 0.7
    lat = torch.randn(B, L, 3 * C, generator=g) * 0.7
    q, k, v = [t.double().reshape(B, P, H, hd).permute(0, 2, 1, 3) for t in (qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:])]
    kl = lat[..., C:2 * C].double().reshape(B, L, H, hd).permute(0, 2, 1, 3)
    vl = lat[..., 2 * C:].double().reshape(B, L, H, hd).permute(0, 2, 1, 3)
    a = (torch.cat([q @ kl.transpose(-2, -1), (q * k).sum(-1, keepdim=True)], -1) * hd ** -0.5).softmax(-1)
    ref = (a[..., :L] @ vl + a[..., L:] * v).permute(0, 2, 1, 3).reshape(B, P, C)
    lat_d, qkv_d = lat.to(cuda), qkv.to(cuda)
    ffma = ops.point_attention(qkv_d, lat_d[..., C:2 * C], lat_d[..., 2 * C:], H, tc=False)
    split = ops.point_attention(qkv_d, lat_d[..., C:2 * C], lat_d[..., 2 * C:], H, tc=True)
    old = ops.TRAIN_ENGINE, ops.TRAIN_PRECISION
    ops.TRAIN_ENGINE, ops.TRAIN_PRECISION = "tc", "bf16"
    try:
        single = ops.point_attention(qkv_d, lat_d[..., C:2 * C], lat_d[..., 2 * C:], H)
    finally:
        ops.TRAIN_ENGINE, ops.TRAIN_PRECISION = old
    errs = [_rel(t, ref) for t in (ffma, split, single)]
    print(f"point_attention fwd B={B} P={P} L={L}: rel err FFMA {errs[0]:.1e} | tcgen05 split fp16 {errs[1]:.1e} | single fp16 {errs[2]:.1e}")
    assert errs[0] < 2e-6 and errs[1] < 5e-6 and errs[2] < 2e-3
