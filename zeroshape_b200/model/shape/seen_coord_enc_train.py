"""Training step of the ResNet-50 seen-surface encoder (CoordEncRes): forward with batch-statistics BatchNorm and saved
activations, hand-written backward.

Replaces torch autograd over model/shape/seen_coord_enc.py:180-194 (torchvision resnet50 with `fc` = 2 x Bottleneck_Conv +
Linear, `depth_feat_proj` on the layer3 hook) for the `optim.fix_dpt` training configuration of options/shape.yaml, in which
the trainable modules are `coord_encoder` and `impl_network`.  Every conv -> BN (-> +residual) (-> ReLU) unit keeps
(input, conv output, output, batch mean, rstd) on the tape; its backward is ReLU mask -> zs_bn_bwd_f32 -> wgrad / dgrad
(zs_gemm_tn_f32 + zs_gemm_f32 for 1x1 stride-1 convolutions, zs_conv2d_nhwc_wgrad_f32 / _dgrad_f32 otherwise).
BatchNorm running statistics are updated as nn.BatchNorm2d does (momentum 0.1, unbiased variance, num_batches_tracked).
"""
import torch

from ... import ops
from .implicit_train import _Grads


def _ohwi(conv):
    return conv.weight.detach().float().permute(0, 2, 3, 1).contiguous()


def _unit_fwd(tape, x, conv, bn, stride=1, pad=0, relu=True, res=None, update_stats=True):
    """x NHWC -> act(BN_batchstats(conv(x)) + res)."""
    w = _ohwi(conv)
    # ops.TRAIN_ENGINE picks the tcgen05 implicit GEMM (ops.TRAIN_PRECISION) or the plain-fp32 convolution; note that
    # batch-statistics BatchNorm over few samples (the 1x1 global branch) amplifies forward rounding in the gradients
    z = ops.conv2d_nhwc(x, w, None, stride, (pad, pad, pad, pad), tc=ops.train_tc(), precision=ops.TRAIN_PRECISION)
    C = z.shape[-1]
    mean, var, rstd = ops.bn_stats(z.view(-1, C), bn.eps)
    gamma, beta = bn.weight.detach().float(), bn.bias.detach().float()
    scale = gamma * rstd                      # [C] vectors: host glue on the statistics
    shift = beta - mean * scale
    y = ops.channel_affine(z, scale.contiguous(), shift.contiguous(), act=ops.ACT_RELU if relu else ops.ACT_NONE, res=res)
    if update_stats and bn.track_running_stats:
        M = z.numel() // C
        mom = bn.momentum if bn.momentum is not None else 0.1
        bn.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
        bn.running_var.mul_(1 - mom).add_(var * (M / max(M - 1, 1)), alpha=mom)
        bn.num_batches_tracked += 1
    tape.append({"x": x, "w": w, "z": z, "y": y, "mean": mean, "rstd": rstd, "conv": conv, "bn": bn, "stride": stride, "pad": pad,
                 "relu": relu, "has_res": res is not None})
    return y


def _unit_bwd(e, dy, G, need_dx=True):
    """-> (dx or None, dres or None)."""
    conv, bn = e["conv"], e["bn"]
    dy = dy.contiguous()
    if e["relu"]:
        dy = ops.act_bwd(dy, e["y"], ops.ACT_RELU)        # the post-ReLU output has the same sign pattern as the pre-activation
    dres = dy if e["has_res"] else None
    z = e["z"]
    C = z.shape[-1]
    dz = ops.bn_bwd(dy.view(-1, C), z.view(-1, C), e["mean"], e["rstd"], bn.weight.detach().float().contiguous(),
                    G.buf(bn.weight), G.buf(bn.bias)).view(z.shape)
    x, w = e["x"], e["w"]
    Cout, KH, KW, Cin = w.shape
    gw = G.buf(conv.weight)                                # OIHW, the parameter's own layout
    if KH == 1 and KW == 1 and e["stride"] == 1 and e["pad"] == 0:
        ops.gemm_tn(dz.view(-1, Cout), x.view(-1, Cin), out=gw.view(Cout, Cin), accumulate=True)
        dx = ops.train_dgrad(dz.view(-1, Cout), w.view(Cout, Cin)).view(x.shape) if need_dx else None
    else:
        p = e["pad"]
        dw = ops.conv2d_nhwc_wgrad(x, dz, KH, KW, e["stride"], (p, p, p, p))          # OHWI
        gw.add_(dw.permute(0, 3, 1, 2))                                              # -> OIHW (weights-sized; host glue)
        dx = ops.conv2d_nhwc_dgrad(dz, w, x.shape, e["stride"], (p, p, p, p)) if need_dx else None
    return dx, dres


def _bneck_conv_fwd(tape, x, m):
    """utils/layers.py:76-100 Bottleneck_Conv on an NHWC tensor, batch-statistics BatchNorm."""
    p = m.kernel_size // 2
    y = _unit_fwd(tape, x, m.linear1, m.bn1, 1, p, relu=True)
    return _unit_fwd(tape, y, m.linear2, m.bn2, 1, p, relu=True, res=x)


def _bneck_conv_bwd(tape, dy, G):
    e2, e1 = tape.pop(), tape.pop()
    d, dres = _unit_bwd(e2, dy, G)
    d, _ = _unit_bwd(e1, d, G)
    return ops.axpby(d, 1.0, dres, 1.0)


def train_forward(mod, coord_nhwc):
    """coord_nhwc [B,H,W,3] (already masked) -> (latent [B,197,latent], tape)."""
    enc = mod.encoder
    B = coord_nhwc.shape[0]
    T = {"units": [], "blocks": []}
    U = T["units"]
    x = _unit_fwd(U, coord_nhwc, enc.conv1, enc.bn1, 2, 3, relu=True)
    T["pool_in"] = x
    x = ops.maxpool3x3s2_nhwc(x, 1, 1, (x.shape[1] + 2 - 3) // 2 + 1, (x.shape[2] + 2 - 3) // 2 + 1)
    feats = {}
    for li in range(1, 5):
        for blk in getattr(enc, f"layer{li}"):
            down = hasattr(blk, "downsample")
            idt = _unit_fwd(U, x, blk.downsample[0], blk.downsample[1], blk.stride, 0, relu=False) if down else x
            y = _unit_fwd(U, x, blk.conv1, blk.bn1, relu=True)
            y = _unit_fwd(U, y, blk.conv2, blk.bn2, blk.stride, 1, relu=True)
            x = _unit_fwd(U, y, blk.conv3, blk.bn3, relu=True, res=idt)
            T["blocks"].append((li, down))
        feats[li] = x
    T["feat4_shape"] = feats[4].shape
    g = ops.avgpool_nhwc(feats[4]).view(B, 1, 1, -1)
    Ug = T["global"] = []
    g = _bneck_conv_fwd(Ug, g, enc.fc[0])
    g = _bneck_conv_fwd(Ug, g, enc.fc[1])
    T["g_in"] = g.view(B, -1)
    gl = ops.train_linear(T["g_in"], enc.fc[2].weight, enc.fc[2].bias)                      # [B, latent]
    Ul = T["local"] = []
    y = _bneck_conv_fwd(Ul, feats[3], mod.depth_feat_proj[0])
    y = _bneck_conv_fwd(Ul, y, mod.depth_feat_proj[1])
    T["l_in"] = y
    pc = mod.depth_feat_proj[2]
    yl = ops.train_linear(y.view(-1, y.shape[-1]), pc.weight.detach().view(pc.weight.shape[0], -1), pc.bias)   # 1x1 conv with bias
    out = torch.cat([gl.view(B, 1, -1), yl.view(B, -1, yl.shape[-1])], dim=1).contiguous()
    return out, T


def train_backward(mod, T, dout, need_dcoord=False):
    """dout [B,197,latent] -> _Grads over the encoder parameters (and, on request, the gradient w.r.t. the input coordinates)."""
    enc = mod.encoder
    G = _Grads()
    B = dout.shape[0]
    dout = dout.detach().float().contiguous()
    dgl = dout[:, 0].contiguous()                               # [B, latent]
    dyl = dout[:, 1:].contiguous().view(-1, dout.shape[-1])     # [B*196, latent]
    # local branch: 1x1 conv (bias) <- 2 x Bottleneck_Conv(1024) <- layer3 output
    pc = mod.depth_feat_proj[2]
    y = T["l_in"]
    Cl = y.shape[-1]
    w2 = pc.weight.detach().view(pc.weight.shape[0], -1)
    ops.gemm_tn(dyl, y.view(-1, Cl), out=G.buf(pc.weight).view(w2.shape), accumulate=True)
    ops.colsum(dyl, out=G.buf(pc.bias), accumulate=True)
    d = ops.train_dgrad(dyl, w2).view(y.shape)
    d = _bneck_conv_bwd(T["local"], d, G)
    dfeat3_local = _bneck_conv_bwd(T["local"], d, G)
    # global branch: Linear <- 2 x Bottleneck_Conv(2048) <- average pool <- layer4 output
    fc = enc.fc[2]
    ops.gemm_tn(dgl, T["g_in"], out=G.buf(fc.weight), accumulate=True)
    ops.colsum(dgl, out=G.buf(fc.bias), accumulate=True)
    d = ops.train_dgrad(dgl, fc.weight).view(B, 1, 1, -1)
    d = _bneck_conv_bwd(T["global"], d, G)
    d = _bneck_conv_bwd(T["global"], d, G)
    _, H4, W4, C4 = T["feat4_shape"]
    dx = ops.avgpool_bwd_nhwc(d.view(B, C4), H4, W4)
    # residual stages, last block first
    U = T["units"]
    for li, down in reversed(T["blocks"]):
        e3, e2, e1 = U.pop(), U.pop(), U.pop()
        d, dres = _unit_bwd(e3, dx, G)
        d, _ = _unit_bwd(e2, d, G)
        d, _ = _unit_bwd(e1, d, G)
        if down:
            ed = U.pop()
            dd, _ = _unit_bwd(ed, dres, G)
            dx = ops.axpby(d, 1.0, dd, 1.0)
        else:
            dx = ops.axpby(d, 1.0, dres, 1.0)
        # the layer3 hook feeds the local branch: its gradient joins when layer4's first block (the one with the downsample)
        # has been crossed, i.e. dx is now the gradient w.r.t. layer3's output
        if li == 4 and down:
            dx = ops.axpby(dx, 1.0, dfeat3_local, 1.0)
    dx = ops.maxpool3x3s2_bwd_nhwc(T["pool_in"], dx, 1, 1)
    dcoord, _ = _unit_bwd(U.pop(), dx, G, need_dx=need_dcoord)
    return (G, dcoord) if need_dcoord else G


class CoordEncTrainFn(torch.autograd.Function):
    """latent = CoordEncTrainFn.apply(module, coord_nhwc, *module_parameters)"""

    @staticmethod
    def forward(ctx, mod, coord_nhwc, *params):
        with torch.no_grad():
            out, tape = train_forward(mod, coord_nhwc.detach().float().contiguous())
        ctx.mod, ctx.tape, ctx.params = mod, tape, params
        return out

    @staticmethod
    def backward(ctx, dout):
        with torch.no_grad():
            G = train_backward(ctx.mod, ctx.tape, dout)
        grads = tuple((G.get(p) if p.requires_grad else None) for p in ctx.params)
        ctx.tape = None
        return (None, None) + grads
