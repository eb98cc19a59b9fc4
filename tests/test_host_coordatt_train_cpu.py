"""CPU: the HOST logic of the transformer seen-surface encoder's training tape (model/shape/seen_coord_att_train.py: op order, closures,
gradient routing through the window selection / cls concatenation / front end) with the kernels replaced by per-op torch stand-ins, against
torch autograd over the oracle restatement (pinned to the reference module).  The kernels themselves are checked on the GPU
(tests/test_gpu_coordatt_train.py)."""
import torch
import torch.nn.functional as F

import fake_ops
from oracle import backbone as BB
from oracle.graph_params import seeded_state_dict


def _install(monkeypatch):
    from zeroshape_b200 import ops
    acts = fake_ops.ACTS

    def axpby(a, alpha=1.0, b=None, beta=1.0, act=0):
        return acts[act](alpha * a + (beta * b if b is not None else 0)).contiguous()

    def via_autograd(fn, x, dy):
        xx = x.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            fn(xx).backward(dy)
        return xx.grad

    def ln_bwd(dy, x, gamma, eps, dgamma=None, dbeta=None):
        gg, bb = gamma.detach().clone().requires_grad_(True), torch.zeros_like(gamma).requires_grad_(True)
        xx = x.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            F.layer_norm(xx, (x.shape[-1],), gg, bb, eps).backward(dy)
        if dgamma is not None:
            dgamma += gg.grad
            dbeta += bb.grad
        return xx.grad

    def mha(qkv, heads):
        B, T, C3 = qkv.shape
        hd = C3 // 3 // heads
        q, k, v = qkv.reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
        return (((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1) @ v).transpose(1, 2).reshape(B, T, C3 // 3).contiguous()

    def coord_embed_windows(coord, mask, w, bias, invalid, pos, cls, ws):
        B, H, W, _ = coord.shape
        C = w.shape[0]
        emb = torch.where(mask.unsqueeze(-1) > 0.5, F.linear(coord, w, bias), invalid.expand(B, H, W, C))
        emb = emb.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws * ws, C) + pos[1:].unsqueeze(0)
        return torch.cat([(cls + pos[0]).view(1, 1, C).expand(emb.shape[0], -1, -1), emb], 1).contiguous()

    def gemm_tn(a, b, out=None, accumulate=False, tc=None):
        r = a.T @ b
        if out is not None:
            out.copy_(out + r if accumulate else r)
            return out
        return r.contiguous()

    def colsum(a, out=None, accumulate=False):
        r = a.sum(0)
        if out is not None:
            out.copy_(out + r if accumulate else r)
            return out
        return r

    for name, fn in dict(axpby=axpby, act_bwd=lambda dy, z, act: via_autograd(acts[act], z, dy), layernorm_bwd_generic=ln_bwd, mha=mha,
                         mha_bwd=lambda qkv, dout, heads: via_autograd(lambda t: mha(t, heads), qkv, dout),
                         coord_embed_windows=coord_embed_windows, layernorm=fake_ops.layernorm, gemm=fake_ops.gemm,
                         train_linear=lambda x2, w, bias=None, res=None, res_mode=0, act=0: fake_ops.gemm(x2, w, bias, res, res_mode, act),
                         train_dgrad=lambda dy, w: (dy @ w).contiguous(), gemm_tn=gemm_tn, colsum=colsum).items():
        monkeypatch.setattr(ops, name, fn)


def test_coord_att_tape_routes_every_gradient(monkeypatch):
    _install(monkeypatch)
    from zeroshape_b200.model.depth import dpt_train as T
    from zeroshape_b200.model.shape import seen_coord_att_train as CAT
    from zeroshape_b200.model.shape.seen_coord_enc import CoordEncAtt
    mod = CoordEncAtt(embed_dim=64, n_blocks=2, num_heads=4, win_size=4, drop_path=0.0)
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    sd = seeded_state_dict(shapes, seed=41, implicit_prefix=None)
    sd["coord_embed.two_d_pos_embed"] = mod.state_dict()["coord_embed.two_d_pos_embed"].clone()
    mod.load_state_dict(sd, strict=True)
    mod.train()
    g = torch.Generator().manual_seed(42)
    B, H, W = 2, 12, 8
    mask = torch.rand(B, H, W, generator=g) < 0.7
    coord = torch.randn(B, H, W, 3, generator=g) * 0.4 * mask.unsqueeze(-1)
    wgt = torch.randn(B, 1 + (H // 4) * (W // 4), 64, generator=g)
    sd_ref = {"coord_encoder." + k: v.clone().requires_grad_(k != "coord_embed.two_d_pos_embed") for k, v in sd.items()}
    coord_r = coord.clone().requires_grad_(True)
    out_ref = BB.coord_enc_att_forward(sd_ref, coord_r, mask, heads=4, ws=4)
    (out_ref * wgt).sum().backward()
    with torch.no_grad():
        tp = T.Tape()
        cd = coord.contiguous()
        out = CAT.train_forward(tp, mod, cd, mask.float().contiguous())
        assert (out - out_ref).abs().max().item() < 1e-5
        tp.add(out, wgt)
        tp.backward()
        dcoord = tp.pop(cd)
    for name, p in mod.named_parameters():
        if name == "coord_embed.two_d_pos_embed":
            continue
        assert id(p) in tp.pgrads, name
        gref = sd_ref["coord_encoder." + name].grad
        assert ((tp.pgrads[id(p)] - gref).norm() / gref.norm()).item() < 1e-4, name
    m3 = mask.unsqueeze(-1)
    assert ((dcoord * m3 - coord_r.grad * m3).norm() / (coord_r.grad * m3).norm()).item() < 1e-4
    # DropPath in train mode: a zero keep-probability draw drops a whole sample's residual branch but the output stays finite
    mod.drop_path = 0.5
    with torch.no_grad():
        out_dp = CAT.train_forward(T.Tape(), mod, cd, mask.float().contiguous())
    assert torch.isfinite(out_dp).all() and not torch.allclose(out_dp, out)
