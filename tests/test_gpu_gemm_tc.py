"""GPU: tcgen05 split-bf16 GEMM (zs_gemm_tc_f32) against fp64 CPU matmul, and the decoder's "tc" engine
against the oracle.  Tolerances: bf16x3 must behave like fp32 (<= 2e-5 of the output scale per GEMM,
<= 1e-3 relative on decoder logits -- north_star bar); single-pass bf16 is only sanity-checked."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _need_sm100():
    from zeroshape_b200 import ops
    if ops.device_cc() != 100:
        pytest.skip("tcgen05 path needs sm_100")


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (1000, 768, 256), (4096, 256, 259), (300, 256, 515),
                                   (20000, 1024, 256), (777, 256, 1024), (5, 300, 70)])
def test_gemm_tc_bf16x3_is_fp32_grade(cuda, M, N, K):
    _need_sm100()
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(M * 7 + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    ref = (a.double() @ w.double().T + b.double())
    pw = ops.PackedWeight(w.to(cuda))
    out = ops.gemm_tc(a.to(cuda), pw, b.to(cuda))
    scale = ref.abs().max().item()
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 2e-5 * scale, (err, scale)
    # fused epilogues
    out = ops.gemm_tc(a.to(cuda), pw, b.to(cuda), res=r.to(cuda), act=ops.ACT_GELU)
    want = F.gelu(ref) + r.double()
    assert (out.cpu().double() - want).abs().max().item() < 3e-5 * scale
    out = ops.gemm_tc(a.to(cuda), pw, b.to(cuda), res=r.to(cuda), res_mode=ops.RES_BEFORE_ACT, act=ops.ACT_RELU)
    assert (out.cpu().double() - F.relu(ref + r.double())).abs().max().item() < 3e-5 * scale
    out = ops.gemm_tc(a.to(cuda), pw, None, act=ops.ACT_SOFTPLUS100)
    want = F.softplus(a.double() @ w.double().T, beta=100)
    assert (out.cpu().double() - want).abs().max().item() < 3e-5 * scale
    # single-pass mode: bf16-grade
    out = ops.gemm_tc(a.to(cuda), pw, b.to(cuda), precision="bf16")
    assert (out.cpu().double() - ref).abs().max().item() < 3e-2 * scale


@pytest.mark.parametrize("M,N,K", [(196, 1024, 2304), (49, 2048, 512), (392, 256, 1024), (130, 70, 4608)])
def test_gemm_tc_split_k(cuda, M, N, K):
    """Few-tile layers take the split-K path (partial tiles in a workspace + deterministic finalize): same accuracy as the
    one-pass kernel, bit-identical from run to run, every epilogue form."""
    _need_sm100()
    from zeroshape_b200 import ops
    from zeroshape_b200._native import lib
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(cuda)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b, r = torch.randn(N, generator=g).to(cuda), torch.randn(M, N, generator=g).to(cuda)
    ref = a.cpu().double() @ w.double().T + b.cpu().double()
    pw = ops.PackedWeight(w.to(cuda))
    scale = ref.abs().max().item()
    try:
        outs = {}
        for on in (1, 0):
            lib.zs_debug_gemm_splitk(on)
            o1 = ops.gemm_tc(a, pw, b)
            o2 = ops.gemm_tc(a, pw, b, res=r, res_mode=ops.RES_BEFORE_ACT, act=ops.ACT_RELU)
            o3 = ops.gemm_tc(a, pw, b, res=r, act=ops.ACT_GELU)
            assert (o1.cpu().double() - ref).abs().max().item() < 2e-5 * scale
            assert (o2.cpu().double() - F.relu(ref + r.cpu().double())).abs().max().item() < 3e-5 * scale
            assert (o3.cpu().double() - (F.gelu(ref) + r.cpu().double())).abs().max().item() < 3e-5 * scale
            assert torch.equal(o1, ops.gemm_tc(a, pw, b))                 # deterministic (no atomics)
            outs[on] = o1
        assert (outs[1] - outs[0]).abs().max().item() < 2e-5 * scale
        wide = torch.zeros(M, N + 8, device=cuda)                          # row-strided output
        lib.zs_debug_gemm_splitk(1)
        ops.gemm_tc(a, pw, b, out=wide[:, 4:N + 4])
        assert torch.equal(wide[:, 4:N + 4], outs[1]) and wide[:, :4].abs().max().item() == 0 and wide[:, N + 4:].abs().max().item() == 0
    finally:
        lib.zs_debug_gemm_splitk(1)


def test_gemm_tc_exact_on_small_integers(cuda):
    """Integers < 256 are exact in bf16 and their products/sums exact in fp32: any layout bug shows as a
    wrong integer, not as noise."""
    _need_sm100()
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(0)
    a = torch.randint(-8, 9, (256, 128), generator=g).float()
    w = torch.randint(-8, 9, (512, 128), generator=g).float()
    for prec in ("bf16", "bf16x3"):
        out = ops.gemm_tc(a.to(cuda), ops.PackedWeight(w.to(cuda)), precision=prec)
        assert torch.equal(out.cpu(), a @ w.T), prec


def test_gemm_tc_repack_on_weight_update(cuda):
    _need_sm100()
    from zeroshape_b200 import ops
    w = torch.randn(256, 64).to(cuda)
    a = torch.randn(64, 64).to(cuda)
    pw = ops.PackedWeight(w)
    o1 = ops.gemm_tc(a, pw)
    w.mul_(2.0)
    o2 = ops.gemm_tc(a, pw)
    assert torch.allclose(o2, 2 * o1, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3)])
def test_decoder_tc_engine_meets_parity_bar(cuda, precision, tol):
    _need_sm100()
    from oracle.implicit import implicit_forward, implicit_init
    from oracle import eval3d as E
    from zeroshape_b200.model.shape.implicit import Implicit
    sd = implicit_init(seed=3)
    m = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
                 pos_perlayer=False)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    m.engine, m.precision = "tc", precision
    g = torch.Generator().manual_seed(1)
    lat, pts = torch.randn(2, 197, 256, generator=g), torch.rand(2, 3000, 3, generator=g) * 3 - 1.5
    with torch.no_grad():
        ref, _ = implicit_forward(sd, lat, pts)
    out, _ = m(lat.to(cuda), None, pts.to(cuda), need_attn=False)
    from parity import parity_rel, normwise
    rel, nw = parity_rel(out, ref), normwise(out, ref)
    print(f"tc[{precision}] max abs {(out.cpu() - ref).abs().max().item():.3e} parity rel {rel:.3e} normwise {nw:.3e}")
    assert rel < tol and nw < 1e-4
    # thresholded voxel grid identical outside the error band
    n = 21
    occ_ref = E.level_grid(sd, lat[:1], n, -1.5, 1.5)
    occ = m.grid_occupancy(lat[:1].to(cuda), n, -1.5, 1.5).cpu()
    band = (occ_ref - 0.5).abs() > 2.5e-4
    assert torch.equal((occ > 0.5)[band], (occ_ref > 0.5)[band])
    assert band.float().mean() > 0.99


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride,pad", [
    (1, 14, 14, 64, 256, 3, 1, (1, 1, 1, 1)),      # chunk == one filter tap
    (2, 28, 28, 128, 128, 3, 2, (0, 1, 0, 1)),     # timm SAME padding of a stride-2 3x3 (asymmetric)
    (1, 56, 56, 256, 64, 1, 1, (0, 0, 0, 0)),      # 1x1, Cout < one N tile
    (2, 7, 7, 768, 768, 3, 1, (1, 1, 1, 1)),       # intrinsics head: M = 98 < 128, K = 6912
    (1, 30, 30, 32, 1, 1, 1, (0, 0, 0, 0)),        # DPT head last layer: K = 32 (generic tap decode), Cout = 1
    (1, 20, 20, 96, 40, 3, 1, (1, 1, 1, 1)),       # Cin % 64 != 0: chunks straddle taps
    (1, 1, 1, 2048, 2048, 1, 1, (0, 0, 0, 0)),     # CoordEncRes global token (a [B,C] vector as a 1x1 image)
])
def test_conv2d_tc_matches_fp64_conv(cuda, B, H, W, Cin, Cout, k, stride, pad):
    """zs_conv2d_nhwc_tc (tcgen05 implicit GEMM, im2col in the A producer) vs an fp64 convolution."""
    _need_sm100()
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + H + Cin + Cout + k)
    x = torch.randn(B, H, W, Cin, generator=g)
    w = torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5
    b = torch.randn(Cout, generator=g)
    pt, pb, pl, pr = pad
    xp = F.pad(x.permute(0, 3, 1, 2).double(), (pl, pr, pt, pb))
    ref = F.conv2d(xp, w.permute(0, 3, 1, 2).double(), b.double(), stride=stride).permute(0, 2, 3, 1)
    scale = ref.abs().max().item()
    old = ops.ENCODER_ENGINE
    try:
        ops.ENCODER_ENGINE = "tc"
        xd, wd, bd = x.to(cuda), w.to(cuda).contiguous(), b.to(cuda)
        out = ops.conv2d_nhwc(xd, wd, bd, stride, pad)
        assert out.shape == ref.shape
        assert (out.cpu().double() - ref).abs().max().item() < 2e-5 * scale
        # ReLU on load + residual + activation epilogue (ResidualConvUnit_custom / Bottleneck_Conv forms)
        r = torch.randn(ref.shape, generator=g)
        xpr = F.pad(F.relu(x).permute(0, 3, 1, 2).double(), (pl, pr, pt, pb))
        ref2 = F.conv2d(xpr, w.permute(0, 3, 1, 2).double(), b.double(), stride=stride).permute(0, 2, 3, 1)
        out = ops.conv2d_nhwc(xd, wd, bd, stride, pad, pre_relu=True, res=r.to(cuda))
        assert (out.cpu().double() - (ref2 + r.double())).abs().max().item() < 3e-5 * max(scale, 1.0)
        out = ops.conv2d_nhwc(xd, wd, bd, stride, pad, act=ops.ACT_RELU, res=r.to(cuda), res_mode=ops.RES_BEFORE_ACT)
        assert (out.cpu().double() - F.relu(ref + r.double())).abs().max().item() < 3e-5 * max(scale, 1.0)
        # and it agrees with the FFMA kernel it replaces
        ops.ENCODER_ENGINE = "f32"
        out32 = ops.conv2d_nhwc(xd, wd, bd, stride, pad)
        assert (out32.cpu().double() - ref).abs().max().item() < 2e-5 * scale
    finally:
        ops.ENCODER_ENGINE = old
