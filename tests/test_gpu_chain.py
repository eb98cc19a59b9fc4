"""GPU: chained tcgen05 kernels (csrc/chain_tc.cu) against fp64 CPU math and the decoder oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _need_sm100():
    from zeroshape_b200 import ops
    if ops.device_cc() != 100:
        pytest.skip("tcgen05 path needs sm_100")


@pytest.mark.parametrize("M", [128, 1000, 19000])
def test_chain_mlp_matches_fp64(cuda, M):
    _need_sm100()
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, 256, generator=g) * 1.5 + 0.3
    w1, b1 = torch.randn(1024, 256, generator=g) / 16, torch.randn(1024, generator=g) * 0.1
    w2, b2 = torch.randn(256, 1024, generator=g) / 32, torch.randn(256, generator=g) * 0.1
    lw, lb = 1 + 0.1 * torch.randn(256, generator=g), 0.1 * torch.randn(256, generator=g)
    xd = x.double()
    h = F.layer_norm(xd, (256,), lw.double(), lb.double(), 1e-6)
    ref = xd + F.linear(F.gelu(F.linear(h, w1.double(), b1.double())), w2.double(), b2.double())
    mats = []
    for gi in range(4):
        mats += [w1[256 * gi:256 * (gi + 1), :].to(cuda), w2[:, 256 * gi:256 * (gi + 1)].to(cuda)]
    blob = ops.pack_tiles(mats)
    from zeroshape_b200._native import lib
    try:
        for variant in (1, 0):      # 1 = activations in tensor memory (chain_mlp2_kernel, default), 0 = shared-memory ring E
            lib.zs_debug_chain_variant(variant)
            for prec, tol in (("fp16x3", 4e-6), ("fp16", 4e-3)):
                xc = x.clone().to(cuda)
                ops.chain_mlp(xc, lw.to(cuda), lb.to(cuda), 1e-6, blob, b1.to(cuda), b2.to(cuda), prec)
                err = (xc.cpu().double() - ref).abs().max().item()
                print(f"chain_mlp M={M} variant {variant} {prec}: max err {err:.3e} / scale {ref.abs().max().item():.2f}")
                assert err < tol * ref.abs().max().item(), (variant, prec, err)
    finally:
        lib.zs_debug_chain_variant(1)


@pytest.mark.parametrize("M", [1, 128, 1000, 40000])
def test_chain_pmlp_matches_fp64_and_the_two_kernel_path(cuda, M):
    """zs_chain_pmlp_fwd: x' = x + A Wp^T + bp from the tile-blocked attention output, then the MLP, in one kernel."""
    _need_sm100()
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(M + 11)
    x = torch.randn(M, 256, generator=g) * 1.5 + 0.3
    a = torch.randn(M, 256, generator=g)
    wp, bp = torch.randn(256, 256, generator=g) / 16, torch.randn(256, generator=g) * 0.1
    w1, b1 = torch.randn(1024, 256, generator=g) / 16, torch.randn(1024, generator=g) * 0.1
    w2, b2 = torch.randn(256, 1024, generator=g) / 32, torch.randn(256, generator=g) * 0.1
    x1 = x.double() + F.linear(a.double(), wp.double(), bp.double())
    ref = x1 + F.linear(F.gelu(F.linear(F.layer_norm(x1, (256,), None, None, 1e-6), w1.double(), b1.double())), w2.double(), b2.double())
    mats = []
    for gi in range(4):
        mats += [w1[256 * gi:256 * (gi + 1), :].to(cuda), w2[:, 256 * gi:256 * (gi + 1)].to(cuda)]
    blob, pblob = ops.pack_tiles(mats), ops.pack_generic(wp.to(cuda))
    tiles = (M + 127) // 128
    pad = torch.zeros(tiles * 128, 256)
    pad[:M] = a
    a_blk = pad.view(tiles, 128, 64, 4).permute(0, 2, 1, 3).contiguous().to(cuda)
    for prec, tol in (("fp16x3", 6e-6), ("fp16", 6e-3)):
        xc = x.clone().to(cuda)
        ops.chain_pmlp(xc, a_blk, pblob, bp.to(cuda), 1e-6, blob, b1.to(cuda), b2.to(cuda), prec)
        err = (xc.cpu().double() - ref).abs().max().item()
        x2 = x.clone().to(cuda)
        ops.chain_lin(a_blk, pblob, bp.to(cuda), 1, res=x2, out=x2, precision=prec)
        ops.chain_mlp(x2, None, None, 1e-6, blob, b1.to(cuda), b2.to(cuda), prec)
        d2 = (xc - x2).abs().max().item()
        print(f"chain_pmlp M={M} {prec}: max err {err:.3e} / scale {ref.abs().max().item():.2f}; vs chain_lin + chain_mlp {d2:.3e}")
        assert err < tol * ref.abs().max().item(), (prec, err)
        assert d2 < tol * ref.abs().max().item()


@pytest.mark.parametrize("M", [1, 128, 1000, 40000])
def test_chain_lin_matches_fp64(cuda, M):
    """LN + qkv (3 n-tiles) and proj + in-place residual (1 n-tile) of zs_chain_lin_fwd; zs_point_proj_f32."""
    _need_sm100()
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(M + 5)
    x = torch.randn(M, 256, generator=g) * 1.5 + 0.3
    w, b = torch.randn(768, 256, generator=g) / 16, torch.randn(768, generator=g) * 0.1
    ref = F.linear(F.layer_norm(x.double(), (256,), None, None, 1e-6), w.double(), b.double())
    blob = ops.pack_generic(w.to(cuda))
    for prec, tol in (("fp16x3", 4e-6), ("fp16", 4e-3)):
        out = ops.chain_lin(x.to(cuda), blob, b.to(cuda), 3, do_ln=True, ln_eps=1e-6, precision=prec)
        assert (out.cpu().double() - ref).abs().max().item() < tol * ref.abs().max().item(), prec
    # proj: no LN, residual read and written in place, row-strided input view
    wp, bp = torch.randn(256, 256, generator=g) / 16, torch.randn(256, generator=g) * 0.1
    a = torch.randn(M, 300, generator=g)
    ref2 = x.double() + F.linear(a[:, 8:264].double(), wp.double(), bp.double())
    xc = x.clone().to(cuda)
    ops.chain_lin(a.to(cuda)[:, 8:264], ops.pack_generic(wp.to(cuda)), bp.to(cuda), 1, res=xc, out=xc)
    assert (xc.cpu().double() - ref2).abs().max().item() < 4e-6 * ref2.abs().max().item()
    # the same proj from the tile-blocked layout the register-softmax attention kernel writes (do_ln = 2)
    tiles = (M + 127) // 128
    pad = torch.zeros(tiles * 128, 256)
    pad[:M] = a[:, 8:264]
    a_blk = pad.view(tiles, 128, 64, 4).permute(0, 2, 1, 3).contiguous().to(cuda)
    assert torch.equal(ops.unblock_rows(a_blk, M).cpu(), a[:, 8:264])
    xb = x.clone().to(cuda)
    ops.chain_lin(a_blk, ops.pack_generic(wp.to(cuda)), bp.to(cuda), 1, res=xb, out=xb)
    assert torch.equal(xb, xc)                          # same arithmetic, only the operand fetch differs
    # LinearProj3D
    pts = torch.rand(M, 3, generator=g) * 3 - 1.5
    w3, b3 = torch.randn(256, 3, generator=g), torch.randn(256, generator=g)
    out = ops.point_proj(pts.to(cuda), w3.to(cuda), b3.to(cuda))
    assert (out.cpu().double() - F.linear(pts.double(), w3.double(), b3.double())).abs().max().item() < 2e-6


@pytest.mark.parametrize("P", [1, 130, 5000])
def test_chain_occ_matches_oracle_mlp(cuda, P):
    _need_sm100()
    from oracle.implicit import implicit_init, _occupancy_mlp, _ln
    from zeroshape_b200.model.shape.implicit import Implicit
    sd = implicit_init(seed=11)
    m = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
                 pos_perlayer=False)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(P)
    x = torch.randn(P, 256, generator=g)
    pts = torch.rand(P, 3, generator=g) * 3 - 1.5
    with torch.no_grad():
        sdd = {k: v.double() for k, v in sd.items()}
        ref = _occupancy_mlp(pts.double(), _ln(x.double(), sdd, "norm"), sdd, "impl_mlp").squeeze(-1)
    _, _, occ_blob, biases, w8, b8, _, _ = m._chain_blobs()
    from parity import parity_rel
    from zeroshape_b200._native import lib
    try:
        for variant in (1, 0):
            lib.zs_debug_chain_variant(variant)
            out = ops.chain_occ(x.to(cuda), pts.to(cuda), None, None, m.norm.eps, occ_blob, biases, w8, b8)
            print(f"chain_occ P={P} variant {variant}: parity_rel {parity_rel(out, ref):.3e}")
            assert parity_rel(out, ref) < 2e-4
            sig = ops.chain_occ(x.to(cuda), pts.to(cuda), None, None, m.norm.eps, occ_blob, biases, w8, b8, sigmoid=True)
            assert (sig.cpu().double() - torch.sigmoid(ref)).abs().max().item() < 1e-5
    finally:
        lib.zs_debug_chain_variant(1)


def test_decoder_chain_engine_parity_and_voxels(cuda):
    _need_sm100()
    from oracle.implicit import implicit_forward, implicit_init
    from oracle import eval3d as E
    from zeroshape_b200.model.shape.implicit import Implicit
    sd = implicit_init(seed=12)
    m = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
                 pos_perlayer=False)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    m.engine = "chain"
    g = torch.Generator().manual_seed(2)
    lat, pts = torch.randn(2, 197, 256, generator=g), torch.rand(2, 3000, 3, generator=g) * 3 - 1.5
    with torch.no_grad():
        ref, _ = implicit_forward(sd, lat, pts)
    out, _ = m(lat.to(cuda), None, pts.to(cuda), need_attn=False)
    from parity import parity_rel, normwise
    rel, nw = parity_rel(out, ref), normwise(out, ref)
    print(f"chain max abs {(out.cpu() - ref).abs().max().item():.3e} parity rel {rel:.3e} normwise {nw:.3e}")
    assert rel < 1e-3 and nw < 1e-4
    n = 21
    occ_ref = E.level_grid(sd, lat[:1], n, -1.5, 1.5)
    occ = m.grid_occupancy(lat[:1].to(cuda), n, -1.5, 1.5).cpu()
    band = (occ_ref - 0.5).abs() > 2.5e-5
    assert torch.equal((occ > 0.5)[band], (occ_ref > 0.5)[band]) and band.float().mean() > 0.99


@pytest.mark.parametrize("M", [128, 1000, 4224])
def test_tensor_core_attention_matches_f32_kernel_and_reference_math(cuda, M):
    _need_sm100()
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(M)
    L, C, H = 197, 256, 8
    qkv = torch.randn(M, 3 * C, generator=g)
    lat_qkv = torch.randn(1, L, 3 * C, generator=g)
    k_lat, v_lat = lat_qkv[..., C:2 * C], lat_qkv[..., 2 * C:]
    # reference math (model/shape/implicit.py:38-57) in fp64
    q, k, v = [t.double().reshape(M, H, 32).permute(1, 0, 2) for t in (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:])]
    kl = k_lat[0].double().reshape(L, H, 32).permute(1, 0, 2)
    vl = v_lat[0].double().reshape(L, H, 32).permute(1, 0, 2)
    s = torch.cat([q @ kl.transpose(1, 2), (q * k).sum(-1, keepdim=True)], -1) * 32 ** -0.5
    a = s.softmax(-1)
    ref = (a[..., :L] @ vl + a[..., L:] * v).permute(1, 0, 2).reshape(M, C)
    lat_dev = lat_qkv.to(cuda)
    kp, vp = ops.attn_pack_kv(lat_dev[0, :, C:2 * C], lat_dev[0, :, 2 * C:], H)
    out = ops.attn_tc(qkv.to(cuda), kp, vp, L, 32 ** -0.5)
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 2e-5 * ref.abs().max().item(), err
    f32 = ops.point_attention(qkv.to(cuda).view(1, M, 3 * C), lat_dev[..., C:2 * C], lat_dev[..., 2 * C:], H)
    assert (f32.view(M, C).cpu().double() - ref).abs().max().item() < 1e-5 * ref.abs().max().item()


def test_decoder_with_tensor_core_attention(cuda):
    _need_sm100()
    from oracle.implicit import implicit_forward, implicit_init
    from zeroshape_b200.model.shape.implicit import Implicit
    sd = implicit_init(seed=13)
    m = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
                 pos_perlayer=False)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    m.engine, m.attention = "chain", "tc"
    g = torch.Generator().manual_seed(3)
    lat, pts = torch.randn(2, 197, 256, generator=g), torch.rand(2, 2000, 3, generator=g) * 3 - 1.5
    with torch.no_grad():
        ref, _ = implicit_forward(sd, lat, pts)
    out, _ = m(lat.to(cuda), None, pts.to(cuda), need_attn=False)
    from parity import parity_rel, normwise
    rel, nw = parity_rel(out, ref), normwise(out, ref)
    print(f"chain+tc-attn max abs {(out.cpu() - ref).abs().max().item():.3e} parity rel {rel:.3e} normwise {nw:.3e}")
    assert rel < 1e-3 and nw < 1e-4


@pytest.mark.parametrize("M", [128, 1000, 4224])
def test_fused_attention_kernel_matches_reference_math(cuda, M):
    _need_sm100()
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(M + 1)
    L, C, H = 197, 256, 8
    qkv = torch.randn(M, 3 * C, generator=g)
    lat_qkv = torch.randn(1, L, 3 * C, generator=g)
    q, k, v = [t.double().reshape(M, H, 32).permute(1, 0, 2) for t in (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:])]
    kl = lat_qkv[0, :, C:2 * C].double().reshape(L, H, 32).permute(1, 0, 2)
    vl = lat_qkv[0, :, 2 * C:].double().reshape(L, H, 32).permute(1, 0, 2)
    s = torch.cat([q @ kl.transpose(1, 2), (q * k).sum(-1, keepdim=True)], -1) * 32 ** -0.5
    a = s.softmax(-1)
    ref = (a[..., :L] @ vl + a[..., L:] * v).permute(1, 0, 2).reshape(M, C)
    lat_dev = lat_qkv.to(cuda)
    kb, vb = ops.attn_pack_fused(lat_dev[0, :, C:2 * C], lat_dev[0, :, 2 * C:], H)
    out = ops.attn_fused(qkv.to(cuda), kb, vb, L, 32 ** -0.5)
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 4e-6 * ref.abs().max().item(), err
    fast = ops.attn_fused(qkv.to(cuda), kb, vb, L, 32 ** -0.5, precision="fp16")
    assert (fast.cpu().double() - ref).abs().max().item() < 4e-3 * ref.abs().max().item()


def test_decoder_with_fused_attention(cuda):
    _need_sm100()
    from oracle.implicit import implicit_forward, implicit_init
    from parity import parity_rel, normwise
    from zeroshape_b200.model.shape.implicit import Implicit
    sd = implicit_init(seed=14)
    m = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
                 pos_perlayer=False)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    m.engine, m.attention = "chain", "fused"
    g = torch.Generator().manual_seed(4)
    lat, pts = torch.randn(2, 197, 256, generator=g), torch.rand(2, 2000, 3, generator=g) * 3 - 1.5
    with torch.no_grad():
        ref, _ = implicit_forward(sd, lat, pts)
    out, _ = m(lat.to(cuda), None, pts.to(cuda), need_attn=False)
    rel, nw = parity_rel(out, ref), normwise(out, ref)
    print(f"chain+fused-attn parity rel {rel:.3e} normwise {nw:.3e}")
    assert rel < 1e-3 and nw < 1e-4


@pytest.mark.parametrize("M", [1, 128, 1000, 19000])
def test_qkvattn_kernel_matches_reference_math(cuda, M):
    """zs_chain_qkvattn_fwd: LayerNorm + qkv + point->latent attention of one image in one kernel (implicit.py:105, 30-57)."""
    _need_sm100()
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(M + 7)
    L, C, H = 197, 256, 8
    x = torch.randn(M, C, generator=g) * 1.5 + 0.3
    w, b = torch.randn(3 * C, C, generator=g) / 16, torch.randn(3 * C, generator=g) * 0.1
    lat_qkv = torch.randn(1, L, 3 * C, generator=g)
    qkv = F.linear(F.layer_norm(x.double(), (C,), None, None, 1e-6), w.double(), b.double())
    q, k, v = [t.reshape(M, H, 32).permute(1, 0, 2) for t in (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:])]
    kl = lat_qkv[0, :, C:2 * C].double().reshape(L, H, 32).permute(1, 0, 2)
    vl = lat_qkv[0, :, 2 * C:].double().reshape(L, H, 32).permute(1, 0, 2)
    s = torch.cat([q @ kl.transpose(1, 2), (q * k).sum(-1, keepdim=True)], -1) * 32 ** -0.5
    a = s.softmax(-1)
    ref = (a[..., :L] @ vl + a[..., L:] * v).permute(1, 0, 2).reshape(M, C)
    lat_dev = lat_qkv.to(cuda)
    kb, vb = ops.attn_pack_fused(lat_dev[0, :, C:2 * C], lat_dev[0, :, 2 * C:], H)
    wb = ops.qkvattn_pack(w.to(cuda))
    scale = ref.abs().max().item()
    for prec, flags, tol in (("fp16x3", 0, 8e-6), ("fp16x3", 8, 8e-6), ("fp16x3", 24, 8e-6), ("fp16x3", 25, 8e-4), ("fp16x3", 1, 8e-4), ("fp16x3", 9, 8e-4), ("fp16x3", 7, 2e-3),
                             ("fp16x3", 15, 2e-3), ("fp16", 0, 4e-3), ("fp16", 8, 4e-3)):
        out = ops.chain_qkvattn(x.to(cuda), wb, b.to(cuda), kb, vb, L, 32 ** -0.5, ln_eps=1e-6, precision=prec, flags=flags)
        if flags & 16:
            out = ops.unblock_rows(out, M)
        err = (out.cpu().double() - ref).abs().max().item()
        print(f"qkvattn M={M} {prec} flags={flags}: max err {err:.3e} (scale {scale:.3f})")
        assert err < tol * scale, (prec, flags, err)
    # points mode (first decoder block): x = LinearProj3D(points) recomputed inside the kernel, LayerNorm factors in closed form
    pts = torch.rand(M, 3, generator=g) * 3 - 1.5
    w3, b3 = torch.randn(256, 3, generator=g), torch.randn(256, generator=g)
    x3 = ops.point_proj(pts.to(cuda), w3.to(cuda), b3.to(cuda))
    pp, pp_stat = ops.point_proj_tables(w3.to(cuda), b3.to(cuda))
    o_ref = ops.unblock_rows(ops.chain_qkvattn(x3, wb, b.to(cuda), kb, vb, L, 32 ** -0.5, ln_eps=1e-6, flags=24), M)
    o_pts = ops.unblock_rows(ops.chain_qkvattn_pts(pts.to(cuda), pp, pp_stat, wb, b.to(cuda), kb, vb, L, 32 ** -0.5, ln_eps=1e-6), M)
    d = (o_pts - o_ref).abs().max().item()
    print(f"qkvattn points mode M={M}: max diff to the x-reading kernel {d:.3e} (scale {o_ref.abs().max().item():.3f})")
    assert d < 8e-6 * o_ref.abs().max().item()
    # row-strided output view
    wide = torch.zeros(M, 300, device=cuda)
    ops.chain_qkvattn(x.to(cuda), wb, b.to(cuda), kb, vb, L, 32 ** -0.5, out=wide[:, 4:260])
    assert (wide[:, 4:260].cpu().double() - ref).abs().max().item() < 8e-6 * scale and wide[:, :4].abs().max().item() == 0


def test_decoder_chain_points_mode_equals_the_materialised_x(cuda):
    """fold_point_proj (no point_proj launch, x recomputed in the first block's kernels) vs the path that writes x first."""
    _need_sm100()
    from oracle.implicit import implicit_init
    from zeroshape_b200.model.shape.implicit import Implicit
    sd = implicit_init(seed=16)
    m = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
                 pos_perlayer=False)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    m.engine = "chain"
    g = torch.Generator().manual_seed(9)
    lat, pts = torch.randn(2, 197, 256, generator=g).to(cuda), (torch.rand(2, 3001, 3, generator=g) * 3 - 1.5).to(cuda)
    outs = {}
    for fold in (True, False):
        m.fold_point_proj = fold
        outs[fold], _ = m(lat, None, pts, need_attn=False)
    d = (outs[True] - outs[False]).abs().max().item()
    print(f"points mode vs materialised x: max |diff| of the logits {d:.3e} (logit scale {outs[False].abs().max().item():.3f})")
    assert d < 2e-5 * outs[False].abs().max().item()


@pytest.mark.parametrize("attention,flags", [("qkv", 0), ("qkv", 8), ("qkv", 24), ("qkv", 1), ("qkv", 9), ("qkv", 7), ("fused", 0)])
def test_decoder_chain_attention_variants(cuda, attention, flags):
    _need_sm100()
    from oracle.implicit import implicit_forward, implicit_init
    from parity import parity_rel, normwise
    from zeroshape_b200.model.shape.implicit import Implicit
    sd = implicit_init(seed=15)
    m = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
                 pos_perlayer=False)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    m.engine, m.attention, m.attn_flags = "chain", attention, flags
    g = torch.Generator().manual_seed(8)
    lat, pts = torch.randn(2, 197, 256, generator=g), torch.rand(2, 3000, 3, generator=g) * 3 - 1.5
    with torch.no_grad():
        ref, _ = implicit_forward(sd, lat, pts)
    out, _ = m(lat.to(cuda), None, pts.to(cuda), need_attn=False)
    rel, nw = parity_rel(out, ref), normwise(out, ref)
    print(f"chain attention={attention} flags={flags}: parity rel {rel:.3e} normwise {nw:.3e}")
    # flags 0 (every contraction three fp16 passes) is the shipped policy; 1 (k, v single-pass) stays inside the 5e-4 budget of
    # the precision study, 7 (scores and P V two-pass as well) does not and is only kept as a fast mode
    assert rel < {0: 3e-4, 1: 6e-4, 7: 5e-3}[flags & 7] and nw < {0: 4e-5, 1: 6e-5, 7: 5e-4}[flags & 7]
