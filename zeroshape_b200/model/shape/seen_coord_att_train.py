"""Training step of the transformer seen-surface encoder (CoordEncAtt + CoordEmb, model/shape/seen_coord_enc.py:13-139) on the tape
of model/depth/dpt_train.py: forward with saved activations, hand-written backward into every parameter and into the XYZ map.

Replaces torch autograd over `CoordEncAtt.forward` for `arch.depth.encoder != resnet` (graph_shape.py:44-46, 150).  Every Block is the
tape composition LayerNorm -> qkv linear -> token attention -> proj (+x) -> LayerNorm -> fc1 -> GELU -> fc2 (+x), i.e. the kernels the ViT
of the depth estimator trains on (tcgen05 GEMMs per ops.TRAIN_ENGINE, zs_layernorm_bwd_generic_f32, two-pass zs_mha_bwd_f32, zs_act_bwd_f32);
timm's DropPath (0.1 in the main blocks) scales the residual branches per sample.  The window front end runs zs_coord_embed_windows_f32
forward; its backward is three reductions over the un-windowed gradient (dW = g_valid^T xyz, db, d invalid token) and one small GEMM
(d xyz = g_valid W).
"""
import torch

from ... import ops
from ..depth import dpt_train as T


def _mul_const(tp, x, s):
    """y = x * s for a constant (broadcastable) factor s."""
    y = (x * s).contiguous()

    def bwd():
        dy = tp.pop(y)
        if dy is not None:
            tp.add(x, dy * s)
    tp.record(bwd)
    return y


def _block(tp, x, blk, heads, scales=None):
    """timm Block on the tape; `scales` = (s_attn, s_mlp) per-sample DropPath factors [N,1,1] or None."""
    a = T.mha(tp, T.linear(tp, T.layernorm(tp, x, blk.norm1), blk.attn.qkv), heads)
    br = T.linear(tp, a, blk.attn.proj)
    x = T.add(tp, x, _mul_const(tp, br, scales[0]) if scales is not None else br)
    h = T.act(tp, T.linear(tp, T.layernorm(tp, x, blk.norm2), blk.mlp.fc1), ops.ACT_GELU)
    br = T.linear(tp, h, blk.mlp.fc2)
    return T.add(tp, x, _mul_const(tp, br, scales[1]) if scales is not None else br)


def _droppath(mod, n, device):
    p = float(getattr(mod, "drop_path", 0.0))
    if not mod.training or p <= 0.0:
        return None
    keep = 1.0 - p
    return tuple(((torch.rand(n, 1, 1, device=device) < keep).float() / keep) for _ in range(2))


def coord_embed_forward(tp, emb_mod, coord, mask_f, need_dcoord=True):
    """CoordEmb.forward on the tape: coord [B,H,W,3] (tape tensor), mask_f float [B,H,W] -> window tokens [B, nw, C]."""
    B, H, W, _ = coord.shape
    ws, C = emb_mod.win_size, emb_mod.embed_dim
    lin = emb_mod.pos_embed
    pos = emb_mod.two_d_pos_embed.detach().reshape(ws * ws + 1, C).contiguous()
    x0 = ops.coord_embed_windows(coord.contiguous(), mask_f, lin.weight.detach(), lin.bias.detach(), emb_mod.invalid_coord_token.detach(),
                                 pos, emb_mod.cls_token.detach().reshape(-1).contiguous(), ws)

    def bwd_front():
        g = tp.pop(x0)
        if g is None:
            return
        tp.padd(emb_mod.cls_token, ops.colsum(g[:, 0, :].contiguous()).view(1, 1, C))
        # un-window: [B*nwy*nwx, ws*ws, C] -> pixel order [B,H,W,C] (the inverse of seen_coord_enc.py:56-59)
        gp = g[:, 1:, :].reshape(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B * H * W, C)
        m = mask_f.reshape(-1, 1)
        gv = (gp * m).contiguous()
        tp.padd(lin.weight, ops.gemm_tn(gv, coord.reshape(-1, 3).contiguous()))
        tp.padd(lin.bias, ops.colsum(gv))
        tp.padd(emb_mod.invalid_coord_token, ops.colsum((gp - gv).contiguous()))
        if need_dcoord:
            tp.add(coord, ops.gemm(gv, lin.weight.detach().t().clone(memory_format=torch.contiguous_format)).view(coord.shape))
    tp.record(bwd_front)
    x = x0
    for blk in emb_mod.blocks:
        x = _block(tp, x, blk, emb_mod.num_heads)
    tok = x[:, 0].contiguous()

    def bwd_select(x=x):
        d = tp.pop(tok)
        if d is not None:
            dx = torch.zeros_like(x)
            dx[:, 0] = d
            tp.add(x, dx)
    tp.record(bwd_select)
    return T.view(tp, tok, B, (H // ws) * (W // ws), C)


def train_forward(tp, mod, coord, mask_f, need_dcoord=True):
    """CoordEncAtt.forward in train mode on the tape -> latent [B, 1 + nw, C]."""
    tok = coord_embed_forward(tp, mod.coord_embed, coord, mask_f, need_dcoord)
    B, _, C = tok.shape
    x = torch.cat([mod.cls_token.detach().expand(B, -1, -1), tok], dim=1).contiguous()

    def bwd_cat(x=x):                      # bind now: `x` is re-assigned by the block loop below
        dx = tp.pop(x)
        if dx is None:
            return
        tp.padd(mod.cls_token, ops.colsum(dx[:, 0].contiguous()).view(1, 1, C))
        tp.add(tok, dx[:, 1:].contiguous())
    tp.record(bwd_cat)
    for blk in mod.blocks:
        x = _block(tp, x, blk, mod.num_heads, _droppath(mod, B, x.device))
    return T.layernorm(tp, x, mod.norm)


def resample_on_tape(tp, seen_nhwc, mask_nchw, h, w):
    """interpolate_coordmap (utils/util.py:336-345) of the already-masked XYZ map on the tape -> (coord [B,h,w,3], mask float [B,h,w])."""
    m = (mask_nchw > 0.5).float()
    mk = ops.bilinear_nhwc(ops.nchw_to_nhwc(m.contiguous()), h, w, False)           # [B,h,w,1]
    mb = (mk > 0.5).float()
    cv = T.bilinear(tp, seen_nhwc, h, w, False)
    return _mul_const(tp, cv, mb / (mk + 1.e-6)), mb.squeeze(-1).contiguous()


class CoordAttTrainFn(torch.autograd.Function):
    """latent = CoordAttTrainFn.apply(module, coord [B,H,W,3], mask_f [B,H,W], *module_parameters)  (the optim.fix_dpt configuration)"""

    @staticmethod
    def forward(ctx, mod, coord, mask_f, *params):
        with torch.no_grad():
            tp = T.Tape()
            out = train_forward(tp, mod, coord.detach().float().contiguous(), mask_f.float().contiguous(), need_dcoord=False)
        ctx.tp, ctx.out, ctx.params = tp, out, params
        return out

    @staticmethod
    def backward(ctx, dout):
        tp = ctx.tp
        with torch.no_grad():
            tp.add(ctx.out, dout)
            tp.backward()
        grads = tuple(tp.pgrads.get(id(p)) if p.requires_grad else None for p in ctx.params)
        ctx.tp = None
        return (None, None, None) + grads
