"""GPU: every dense building block of the C ABI against the PyTorch fp32 CPU op it replaces."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(a, b, rtol=2e-5, atol=2e-5):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err = (a - b).abs().max().item()
    scale = b.abs().max().item() + 1e-30
    assert err <= atol + rtol * scale, f"max err {err:.3e} (scale {scale:.3e})"


@pytest.mark.parametrize("M,N,K", [(197, 768, 256), (1, 3, 768), (4096, 256, 259), (3000, 256, 515), (8195, 1024, 256),
                                   (50000, 256, 3), (130, 130, 17)])
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_gemm(cuda, M, N, K, act):
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(M + N + K + act)
    a, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g)
    ref = F.linear(a, w, b)
    ref = [lambda x: x, F.relu, F.gelu, lambda x: F.softplus(x, beta=100)][act](ref)
    out = ops.gemm(a.to(cuda), w.to(cuda), b.to(cuda), act=act)
    _close(out, ref)
    out = ops.gemm(a.to(cuda), w.to(cuda), b.to(cuda), res=r.to(cuda), res_mode=ops.RES_AFTER_ACT, act=act)
    _close(out, ref + r)
    if act == 1:
        out = ops.gemm(a.to(cuda), w.to(cuda), b.to(cuda), res=r.to(cuda), res_mode=ops.RES_BEFORE_ACT, act=act)
        _close(out, F.relu(F.linear(a, w, b) + r))


def test_gemm_strided_views(cuda):
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(0)
    big = torch.randn(300, 768, generator=g).to(cuda)
    w = torch.randn(64, 256, generator=g).to(cuda)
    _close(ops.gemm(big[:, 256:512], w), F.linear(big[:, 256:512].cpu(), w.cpu()))


CONVS = [  # B,H,W,Cin,Cout,K,stride,pad(t,b,l,r)
    (2, 224, 224, 3, 64, 7, 2, (2, 3, 2, 3)),      # ResNetV2 stem, TF-SAME asymmetric
    (1, 56, 56, 64, 64, 3, 1, (1, 1, 1, 1)),
    (2, 56, 56, 128, 128, 3, 2, (0, 1, 0, 1)),     # SAME stride-2 on even size
    (1, 14, 14, 768, 768, 3, 2, (1, 1, 1, 1)),     # act_postprocess4 conv
    (2, 28, 28, 512, 256, 1, 1, (0, 0, 0, 0)),
    (1, 56, 56, 256, 512, 1, 2, (0, 0, 0, 0)),     # strided 1x1 downsample
    (1, 7, 7, 768, 768, 3, 1, (1, 1, 1, 1)),
    (1, 17, 19, 5, 7, 3, 1, (1, 1, 1, 1)),          # ragged channels
]


@pytest.mark.parametrize("cfg", CONVS)
def test_conv2d_nhwc(cuda, cfg):
    from zeroshape_b200 import ops
    B, H, W, Cin, Cout, K, s, pad = cfg
    g = torch.Generator().manual_seed(sum(cfg[:7]))
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, K, K, generator=g) / (Cin * K * K) ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(F.pad(x, (pad[2], pad[3], pad[0], pad[1])), w, b, stride=s)
    xh = x.permute(0, 2, 3, 1).contiguous().to(cuda)
    wh = w.permute(0, 2, 3, 1).contiguous().to(cuda)
    out = ops.conv2d_nhwc(xh, wh, b.to(cuda), stride=s, pad=pad)
    _close(out.permute(0, 3, 1, 2), ref)
    # pre-ReLU + residual + ReLU epilogue
    ref2 = F.relu(F.conv2d(F.pad(F.relu(x), (pad[2], pad[3], pad[0], pad[1])), w, b, stride=s) + ref)
    out2 = ops.conv2d_nhwc(xh, wh, b.to(cuda), stride=s, pad=pad, act=ops.ACT_RELU, res=out, res_mode=ops.RES_BEFORE_ACT,
                           pre_relu=True)
    _close(out2.permute(0, 3, 1, 2), ref2, rtol=5e-5, atol=5e-5)


def test_layernorm_groupnorm_affine(cuda):
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 197, 768, generator=g) * 3 + 1
    w, b = torch.randn(768, generator=g), torch.randn(768, generator=g)
    _close(ops.layernorm(x.to(cuda), w.to(cuda), b.to(cuda), 1e-6), F.layer_norm(x, (768,), w, b, 1e-6))
    x = torch.randn(2, 256, 28, 28, generator=g) * 2 + 0.5
    w, b = torch.randn(256, generator=g), torch.randn(256, generator=g)
    ref = F.relu(F.group_norm(x, 32, w, b, 1e-5))
    out = ops.groupnorm_nhwc(x.permute(0, 2, 3, 1).contiguous().to(cuda), w.to(cuda), b.to(cuda), 32, 1e-5, True)
    _close(out.permute(0, 3, 1, 2), ref)
    sc, sh = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g)
    xh = x.permute(0, 2, 3, 1).contiguous()
    _close(ops.channel_affine(xh.to(cuda), sc.to(cuda), sh.to(cuda), act=ops.ACT_RELU), F.relu(xh * sc + sh))


@pytest.mark.parametrize("cfg", [(1, 64, 112, 112, 0.0), (2, 256, 56, 56, 0.5), (3, 1024, 14, 14, 0.0), (2, 512, 28, 28, 30.0),
                                 (1, 128, 5, 7, 1.0), (2, 96, 9, 9, 0.0), (1, 256, 3, 3, 0.0)])
def test_groupnorm_tiled_and_fallback(cuda, cfg):
    """timm GroupNormAct(32) of the ResNetV2 stem / stages on NHWC: the tiled kernels (C / 4 a power of two: group size 2 ... 32,
    a mean 15 standard deviations off zero for the shifted one-pass statistics, a ragged last chunk) and the generic kernel
    (C = 96, tiny maps) against torch in double, with and without the residual / ReLU."""
    from zeroshape_b200 import ops
    B, C, H, W, off = cfg
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, C, H, W, generator=g) * 2 + off
    w, b = torch.randn(C, generator=g), torch.randn(C, generator=g)
    r = torch.randn(B, C, H, W, generator=g)
    ref = F.group_norm(x.double(), 32, w.double(), b.double(), 1e-5)
    xh, rh = x.permute(0, 2, 3, 1).contiguous().to(cuda), r.permute(0, 2, 3, 1).contiguous().to(cuda)
    out = ops.groupnorm_nhwc(xh, w.to(cuda), b.to(cuda), 32, 1e-5, False).permute(0, 3, 1, 2).cpu().double()
    assert (out - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item()), (out - ref).abs().max().item()
    out = ops.groupnorm_nhwc(xh, w.to(cuda), b.to(cuda), 32, 1e-5, True, rh).permute(0, 3, 1, 2).cpu().double()
    ref2 = F.relu(ref + r.double())
    assert (out - ref2).abs().max().item() < 2e-5 * max(1.0, ref2.abs().max().item())


def test_pooling_resize_layout(cuda):
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 64, 112, 112, generator=g)
    xh = x.permute(0, 2, 3, 1).contiguous().to(cuda)
    # timm MaxPool2dSame(3, stride 2) on 112 -> pad (0,1) with -inf ; torchvision maxpool pad 1
    ref = F.max_pool2d(F.pad(x, (0, 1, 0, 1), value=float("-inf")), 3, 2)
    _close(ops.maxpool3x3s2_nhwc(xh, 0, 0, 56, 56).permute(0, 3, 1, 2), ref, 0, 0)
    ref = F.max_pool2d(x, 3, 2, padding=1)
    _close(ops.maxpool3x3s2_nhwc(xh, 1, 1, 56, 56).permute(0, 3, 1, 2), ref, 0, 0)
    _close(ops.avgpool_nhwc(xh), x.mean(dim=(2, 3)))
    small = torch.randn(2, 8, 14, 14, generator=g)
    sh_ = small.permute(0, 2, 3, 1).contiguous().to(cuda)
    _close(ops.bilinear_nhwc(sh_, 28, 28, True).permute(0, 3, 1, 2),
           F.interpolate(small, scale_factor=2, mode="bilinear", align_corners=True), 1e-6, 1e-6)
    pe = torch.randn(1, 8, 24, 24, generator=g)
    _close(ops.bilinear_nhwc(pe.permute(0, 2, 3, 1).contiguous().to(cuda), 14, 14, False).permute(0, 3, 1, 2),
           F.interpolate(pe, size=(14, 14), mode="bilinear", align_corners=False), 1e-6, 1e-6)
    _close(ops.nhwc_to_nchw(ops.nchw_to_nhwc(x.to(cuda), 2.0, -1.0)), x * 2 - 1, 0, 1e-7)
    a, b = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    _close(ops.axpby(a.to(cuda), 2.0, b.to(cuda), -0.5, ops.ACT_SIGMOID), torch.sigmoid(2 * a - 0.5 * b), 1e-6, 1e-6)


@pytest.mark.parametrize("B,T,heads,hd", [(2, 197, 12, 64), (1, 197, 8, 32), (3, 50, 4, 16)])
def test_mha(cuda, B, T, heads, hd):
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(T + heads)
    C = heads * hd
    qkv = torch.randn(B, T, 3 * C, generator=g)
    q, k, v = qkv.reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4).unbind(0)
    ref = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1) @ v
    _close(ops.mha(qkv.to(cuda), heads, tc=False), ref.transpose(1, 2).reshape(B, T, C))


@pytest.mark.parametrize("B,T,heads,hd", [(2, 197, 12, 64), (1, 197, 8, 32), (3, 50, 4, 64), (1, 1, 2, 32), (2, 208, 3, 64), (1, 129, 1, 32)])
def test_mha_on_the_tensor_cores(cuda, B, T, heads, hd):
    """zs_mha_tc_f32 (tcgen05, split-fp16 operands, probabilities in tensor memory) vs fp64 and vs the FFMA kernel."""
    if torch.cuda.get_device_capability(0)[0] != 10:
        pytest.skip("tcgen05 needs sm_100")
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(T * 7 + heads)
    C = heads * hd
    qkv = torch.randn(B, T, 3 * C, generator=g) * 1.3
    q, k, v = qkv.double().reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4).unbind(0)
    ref = (((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1) @ v).transpose(1, 2).reshape(B, T, C)
    out = ops.mha(qkv.to(cuda), heads, tc=True)
    err = (out.cpu().double() - ref).abs().max().item()
    ffma = (ops.mha(qkv.to(cuda), heads, tc=False).cpu().double() - ref).abs().max().item()
    one = (ops.mha(qkv.to(cuda), heads, tc=True, precision="fp16").cpu().double() - ref).abs().max().item()
    print(f"mha_tc B={B} T={T} heads={heads} hd={hd}: max err fp16x3 {err:.2e}, fp16 {one:.2e}, FFMA kernel {ffma:.2e} (scale {ref.abs().max().item():.2f})")
    assert err < 4e-6 * ref.abs().max().item() and one < 4e-3 * ref.abs().max().item()


def test_geometry_glue(cuda):
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, H, W = 3, 224, 224
    params = torch.randn(B, 3, generator=g) * 0.3
    K = ops.intr_param2mtx(params.to(cuda), H, W).cpu()
    f = 1.3875
    sf = torch.pow(4., torch.tanh(params[:, 0]))
    ref = torch.zeros(B, 3, 3)
    ref[:, 2, 2] = 1
    ref[:, 0, 0] = f * W * sf
    ref[:, 1, 1] = f * H * sf
    ref[:, 0, 2] = W / 2 + torch.tanh(params[:, 1]) * W / 2
    ref[:, 1, 2] = H / 2 + torch.tanh(params[:, 2]) * H / 2
    _close(K, ref, 1e-6, 1e-6)
    depth = 1.2 + 0.5 * torch.rand(B, 1, H, W, generator=g)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    mask = (((yy - 112) ** 2 + (xx - 112) ** 2) < 80 ** 2).float().view(1, 1, H, W).repeat(B, 1, 1, 1)
    pts, mean, scale = ops.unproject_normalize(depth.to(cuda), mask.to(cuda), K.to(cuda))
    grid = torch.stack([xx.float(), yy.float(), torch.ones(H, W)], -1).view(-1, 3)
    raw = (torch.linalg.inv(K) @ grid.T.unsqueeze(0)).permute(0, 2, 1) * depth.view(B, -1, 1)
    _close(ops.unproject(depth.to(cuda), K.to(cuda)), raw, 1e-5, 1e-6)
    m = mask.view(B, -1) > 0.5
    for b in range(B):
        valid = raw[b][m[b]]
        mu = valid.mean(0)
        sc = (valid - mu).norm(dim=1).max()
        _close(mean[b], mu, 1e-5, 1e-6)
        _close(scale[b], sc, 1e-5, 1e-6)
        want = (raw[b] - mu) / sc
        want[~m[b]] = 0
        _close(pts[b], want, 1e-5, 2e-6)
