"""Training step of the encoder side: DPT-hybrid depth estimator, intrinsics head, unproject / normalise glue and the
seen-surface encoder, forward with saved activations and hand-written backward.

Replaces torch autograd over `Graph.forward` up to `latent_depth` (model/compute_graph/graph_shape.py:115-150:
model/depth/dpt_depth.py:68-123, model/depth/vit.py:57-154, model/depth/blocks.py:264-342, timm `vit_base_resnet50_384`,
utils/camera.py:52-108, utils/layers.py:76-100, model/shape/seen_coord_enc.py:180-194) for the default
`options/shape.yaml` training configuration (`fix_dpt: false`), where the shape loss reaches the depth estimator through
the unprojected, normalised seen surface.

Mechanics: a small tape.  Every op runs its forward on the library's fp32 kernels and records a closure that pops the
gradient of its output, calls the matching backward kernel(s) of csrc/train.cu / csrc/gemm_simt.cu and accumulates into the
gradients of its inputs (fan-out = accumulation, so residual / skip connections need no special handling).  Parameter-sized
preprocessing (weight standardisation, OIHW -> OHWI, the 24x24 -> 14x14 position-embedding resize) is differentiated with
torch autograd on the parameters themselves -- the same "weights-only glue" the inference path uses for packing.
"""
import math

import torch

from ... import ops
from ...model.shape import seen_coord_enc_train as cet


class Tape:
    def __init__(self):
        self.fns = []
        self.grads = {}
        self.pgrads = {}          # id(param) -> accumulated gradient (parameter layout)

    def add(self, t, g):
        k = id(t)
        g = g.contiguous()
        self.grads[k] = ops.axpby(self.grads[k], 1.0, g, 1.0) if k in self.grads else g

    def pop(self, t):
        return self.grads.pop(id(t), None)

    def padd(self, p, g):
        k = id(p)
        if k in self.pgrads:
            self.pgrads[k] = ops.axpby(self.pgrads[k], 1.0, g.contiguous(), 1.0)
        else:
            self.pgrads[k] = g.contiguous().clone()

    def record(self, fn):
        self.fns.append(fn)

    def backward(self):
        for fn in reversed(self.fns):
            fn()
        self.fns = []


class _W:
    """A kernel-ready weight derived from a parameter by weights-only torch glue, differentiable back to the parameter."""

    def __init__(self, param, fn):
        self.param = param
        self.leaf = param.detach().float().requires_grad_(True)
        with torch.enable_grad():
            self.out = fn(self.leaf)
        self.val = self.out.detach().contiguous()

    def push(self, tape, dval):
        (g,) = torch.autograd.grad(self.out, self.leaf, dval.reshape(self.out.shape))
        tape.padd(self.param, g)


def _ohwi_fn(w):
    return w.permute(0, 2, 3, 1)


def _ws_ohwi_fn(w, eps=1e-8):
    wf = w.reshape(w.shape[0], -1)
    var, mean = torch.var_mean(wf, dim=1, keepdim=True, unbiased=False)
    return ((wf - mean) / torch.sqrt(var + eps)).reshape(w.shape).permute(0, 2, 3, 1)


def _inv3x3(m):
    """Inverse of a batch of 3x3 matrices by the adjugate (rows r0, r1, r2: columns r1 x r2, r2 x r0, r0 x r1 over the determinant).
    Elementwise torch ops only: torch.linalg.inv reads its LU status back on the host, which stalls the step and cannot be
    captured in a CUDA graph."""
    r0, r1, r2 = m[:, 0], m[:, 1], m[:, 2]
    c0, c1, c2 = torch.linalg.cross(r1, r2), torch.linalg.cross(r2, r0), torch.linalg.cross(r0, r1)
    det = (r0 * c0).sum(-1)
    return torch.stack([c0, c1, c2], dim=2) / det.view(-1, 1, 1)


def _same_pad(i, k, s):
    return max((math.ceil(i / s) - 1) * s + (k - 1) + 1 - i, 0)


# ---- tape ops ------------------------------------------------------------------------------------------------------
def conv(tp, x, W, bias_param=None, stride=1, pad=(0, 0, 0, 0), need_dx=True):
    """x NHWC, W = _W holding an OHWI filter -> y (engine / precision per ops.TRAIN_ENGINE / ops.TRAIN_PRECISION)."""
    w = W.val
    Cout, KH, KW, Cin = w.shape
    y = ops.conv2d_nhwc(x, w, bias_param.detach().float() if bias_param is not None else None, stride, pad, tc=ops.train_tc(),
                        precision=ops.TRAIN_PRECISION)

    def bwd():
        dy = tp.pop(y)
        if dy is None:
            return
        if bias_param is not None:
            tp.padd(bias_param, ops.colsum(dy.view(-1, Cout)))
        if KH == 1 and KW == 1 and stride == 1 and pad == (0, 0, 0, 0):
            W.push(tp, ops.gemm_tn(dy.view(-1, Cout), x.view(-1, Cin)))
            if need_dx:
                tp.add(x, ops.train_dgrad(dy.view(-1, Cout), w.view(Cout, Cin)).view(x.shape))
        else:
            W.push(tp, ops.conv2d_nhwc_wgrad(x, dy, KH, KW, stride, pad))
            if need_dx:
                tp.add(x, ops.conv2d_nhwc_dgrad(dy, w, x.shape, stride, pad))
    tp.record(bwd)
    return y


def view(tp, x, *shape):
    """A reshape is a new tensor object: route its gradient back to the tensor it views."""
    y = x.view(*shape)

    def bwd():
        dy = tp.pop(y)
        if dy is not None:
            tp.add(x, dy.reshape(x.shape))
    tp.record(bwd)
    return y


def relu(tp, x):
    y = ops.axpby(x, 1.0, act=ops.ACT_RELU)

    def bwd():
        dy = tp.pop(y)
        if dy is not None:
            tp.add(x, ops.act_bwd(dy, y, ops.ACT_RELU))
    tp.record(bwd)
    return y


def act(tp, z, a):
    """y = act(z) for GELU / CLAMP01 (pre-activation kept)."""
    y = ops.axpby(z, 1.0, act=a)

    def bwd():
        dy = tp.pop(y)
        if dy is not None:
            tp.add(z, ops.act_bwd(dy, z, a))
    tp.record(bwd)
    return y


def add(tp, a, b):
    y = ops.axpby(a, 1.0, b, 1.0)

    def bwd():
        dy = tp.pop(y)
        if dy is not None:
            tp.add(a, dy)
            tp.add(b, dy)
    tp.record(bwd)
    return y


def groupnorm(tp, x, norm, do_relu, res=None):
    y = ops.groupnorm_nhwc(x, norm.weight.detach().float(), norm.bias.detach().float(), 32, 1e-5, do_relu, res)

    def bwd():
        dy = tp.pop(y)
        if dy is None:
            return
        if do_relu:
            dy = ops.act_bwd(dy, y, ops.ACT_RELU)
        if res is not None:
            tp.add(res, dy)
        dg, db = torch.zeros_like(norm.weight, dtype=torch.float32), torch.zeros_like(norm.bias, dtype=torch.float32)
        tp.add(x, ops.groupnorm_bwd_nhwc(dy, x, norm.weight.detach().float().contiguous(), 32, 1e-5, dg, db))
        tp.padd(norm.weight, dg)
        tp.padd(norm.bias, db)
    tp.record(bwd)
    return y


def layernorm(tp, x, norm):
    y = ops.layernorm(x, norm.weight.detach().float(), norm.bias.detach().float(), norm.eps)

    def bwd():
        dy = tp.pop(y)
        if dy is None:
            return
        dg, db = torch.zeros_like(norm.weight, dtype=torch.float32), torch.zeros_like(norm.bias, dtype=torch.float32)
        tp.add(x, ops.layernorm_bwd_generic(dy, x, norm.weight.detach().float().contiguous(), norm.eps, dg, db))
        tp.padd(norm.weight, dg)
        tp.padd(norm.bias, db)
    tp.record(bwd)
    return y


def linear(tp, x, lin):
    """x [..., K] -> [..., N] (no activation; compose with act / add)."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1])
    w = lin.weight.detach().float()
    y = ops.train_linear(x2, w, lin.bias.detach().float() if lin.bias is not None else None).view(*shp[:-1], w.shape[0])

    def bwd():
        dy = tp.pop(y)
        if dy is None:
            return
        d2 = dy.reshape(-1, w.shape[0])
        tp.padd(lin.weight, ops.gemm_tn(d2, x2))
        if lin.bias is not None:
            tp.padd(lin.bias, ops.colsum(d2))
        tp.add(x, ops.train_dgrad(d2, w).view(shp))
    tp.record(bwd)
    return y


def mha(tp, qkv, heads):
    y = ops.mha(qkv, heads)

    def bwd():
        dy = tp.pop(y)
        if dy is not None:
            tp.add(qkv, ops.mha_bwd(qkv, dy, heads))
    tp.record(bwd)
    return y


def maxpool(tp, x, pt, pl, OH, OW):
    y = ops.maxpool3x3s2_nhwc(x, pt, pl, OH, OW)

    def bwd():
        dy = tp.pop(y)
        if dy is not None:
            tp.add(x, ops.maxpool3x3s2_bwd_nhwc(x, dy, pt, pl))
    tp.record(bwd)
    return y


def bilinear(tp, x, OH, OW, align):
    y = ops.bilinear_nhwc(x, OH, OW, align)

    def bwd():
        dy = tp.pop(y)
        if dy is not None:
            tp.add(x, ops.bilinear_bwd_nhwc(dy, x.shape[1], x.shape[2], align))
    tp.record(bwd)
    return y


# ---- DPT-hybrid ------------------------------------------------------------------------------------------------------
def dpt_forward(tp, m, image):
    """DPTDepthModel.forward(image, get_feat=True) on the tape -> (depth [B,1,H,W], layer_4 feature NHWC [B,7,7,768])."""
    vit = m.pretrained.model
    B, _, H, W = image.shape
    gh, gw = H // 16, W // 16
    x = ops.nchw_to_nhwc(image.float().contiguous(), 2.0, -1.0)

    def ws_conv(x, conv_mod, stride, need_dx=True):
        k = conv_mod.weight.shape[-1]
        ph, pw = _same_pad(x.shape[1], k, stride), _same_pad(x.shape[2], k, stride)
        return conv(tp, x, _W(conv_mod.weight, _ws_ohwi_fn), None, stride, (ph // 2, ph - ph // 2, pw // 2, pw - pw // 2), need_dx)

    bb = vit.patch_embed.backbone
    x = groupnorm(tp, ws_conv(x, bb.stem.conv, 2, need_dx=False), bb.stem.norm, True)
    ph, pw = _same_pad(x.shape[1], 3, 2), _same_pad(x.shape[2], 3, 2)
    x = maxpool(tp, x, ph // 2, pw // 2, (x.shape[1] + ph - 3) // 2 + 1, (x.shape[2] + pw - 3) // 2 + 1)
    stages = []
    for s, st in enumerate(bb.stages):
        for b, blk in enumerate(st.blocks):
            stride = 2 if (b == 0 and s > 0) else 1
            short = x
            if b == 0:
                short = groupnorm(tp, ws_conv(x, blk.downsample.conv, stride), blk.downsample.norm, False)
            y = groupnorm(tp, ws_conv(x, blk.conv1, 1), blk.norm1, True)
            y = groupnorm(tp, ws_conv(y, blk.conv2, stride), blk.norm2, True)
            x = groupnorm(tp, ws_conv(y, blk.conv3, 1), blk.norm3, True, res=short)
        stages.append(x)
    s0, s1, s2 = stages
    # tokens: 1x1 projection, cls token, resized position embedding (vit.py:101-154)
    tok = view(tp, conv(tp, s2, _W(vit.patch_embed.proj.weight, _ohwi_fn), vit.patch_embed.proj.bias), B, gh * gw, 768)
    cls = _W(vit.cls_token, lambda t: t * 1.0)
    pos = _W(vit.pos_embed, lambda pe: _resize_pos(pe, gh, gw))
    xt = torch.cat([cls.val.expand(B, -1, -1), tok], dim=1).contiguous()
    x = ops.axpby(xt, 1.0, pos.val.expand(B, -1, -1).contiguous(), 1.0)

    def bwd_tokens(x=x, tok=tok):
        dx = tp.pop(x)
        if dx is None:
            return
        tp.add(tok, dx[:, 1:].contiguous())
        cls.push(tp, ops.colsum(dx[:, 0].contiguous()).view(1, 1, -1))
        pos.push(tp, ops.colsum(dx.view(B, -1)).view(1, dx.shape[1], dx.shape[2]))     # summed over the batch
    tp.record(bwd_tokens)
    taps = {}
    for i, blk in enumerate(vit.blocks):
        h = layernorm(tp, x, blk.norm1)
        a = mha(tp, linear(tp, h, blk.attn.qkv), m.HEADS)
        x = add(tp, x, linear(tp, a, blk.attn.proj))
        h = layernorm(tp, x, blk.norm2)
        h = act(tp, linear(tp, h, blk.mlp.fc1), ops.ACT_GELU)
        x = add(tp, x, linear(tp, h, blk.mlp.fc2))
        if i in m.HOOKS:
            taps[i] = x
        if i == m.HOOKS[-1]:
            break

    def reassemble(tokens, seq):
        T, C = tokens.shape[1], tokens.shape[2]
        lin = getattr(seq, "0").project[0]
        cat = torch.empty(B, T - 1, 2 * C, device=tokens.device, dtype=torch.float32)
        for b in range(B):
            cat[b] = ops.concat2(tokens[b, 1:], tokens[b, :1].expand(T - 1, C), 1.0)

        def bwd_cat():
            dc = tp.pop(cat)
            if dc is None:
                return
            dt = torch.empty_like(tokens)
            dt[:, 1:] = dc[:, :, :C]
            for b in range(B):
                dt[b, 0] = ops.colsum(dc[b, :, C:])
            tp.add(tokens, dt)
        tp.record(bwd_cat)
        y = view(tp, act(tp, linear(tp, cat, lin), ops.ACT_GELU), B, gh, gw, C)
        c3 = getattr(seq, "3")
        y = conv(tp, y, _W(c3.weight, _ohwi_fn), c3.bias)
        if hasattr(seq, "4"):
            c4 = getattr(seq, "4")
            y = conv(tp, y, _W(c4.weight, _ohwi_fn), c4.bias, 2, (1, 1, 1, 1))
        return y
    l3 = reassemble(taps[m.HOOKS[0]], m.pretrained.act_postprocess3)
    l4 = reassemble(taps[m.HOOKS[1]], m.pretrained.act_postprocess4)
    sc = m.scratch
    r = [conv(tp, t, _W(getattr(sc, f"layer{i + 1}_rn").weight, _ohwi_fn), None, 1, (1, 1, 1, 1)) for i, t in enumerate((s0, s1, l3, l4))]

    def rcu(x, unit):
        y = conv(tp, relu(tp, x), _W(unit.conv1.weight, _ohwi_fn), unit.conv1.bias, 1, (1, 1, 1, 1))
        y = conv(tp, relu(tp, y), _W(unit.conv2.weight, _ohwi_fn), unit.conv2.bias, 1, (1, 1, 1, 1))
        return add(tp, y, x)

    def fusion(i, x, skip=None):
        f = getattr(sc, f"refinenet{i}")
        if skip is not None:
            x = add(tp, x, rcu(skip, f.resConfUnit1))
        x = rcu(x, f.resConfUnit2)
        x = bilinear(tp, x, x.shape[1] * 2, x.shape[2] * 2, True)
        return conv(tp, x, _W(f.out_conv.weight, _ohwi_fn), f.out_conv.bias)
    path = fusion(4, r[3])
    path = fusion(3, path, r[2])
    path = fusion(2, path, r[1])
    path = fusion(1, path, r[0])
    oc = sc.output_conv
    c0, c2, c4 = getattr(oc, "0"), getattr(oc, "2"), getattr(oc, "4")
    y = conv(tp, path, _W(c0.weight, _ohwi_fn), c0.bias, 1, (1, 1, 1, 1))
    y = bilinear(tp, y, y.shape[1] * 2, y.shape[2] * 2, True)
    y = relu(tp, conv(tp, y, _W(c2.weight, _ohwi_fn), c2.bias, 1, (1, 1, 1, 1)))
    z = conv(tp, y, _W(c4.weight, _ohwi_fn), c4.bias)
    depth_nhwc = act(tp, z, ops.ACT_CLAMP01)                    # relu then clamp(0, 1) (dpt_depth.py:106,119)
    return depth_nhwc, l4


def _resize_pos(pe, gh, gw):
    """vit.py:101-115: keep the cls slot, resize the 24x24 grid bilinearly (align_corners=False).  Weights-only glue."""
    g_old = int(math.sqrt(pe.shape[1] - 1))
    grid = pe[:, 1:].reshape(1, g_old, g_old, -1).permute(0, 3, 1, 2)
    grid = torch.nn.functional.interpolate(grid, size=(gh, gw), mode="bilinear", align_corners=False)
    return torch.cat([pe[:, :1], grid.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)], dim=1)


# ---- the whole encoder side --------------------------------------------------------------------------------------------
def encoder_forward(graph, opt, rgb, mask, with_intr=True, with_coord=True):
    """-> (tape, outputs dict).  Mirrors Graph.forward up to latent_depth in train mode (graph_shape.py:115-150); with
    `with_coord=False` it is the depth-only graph (graph_depth.py:61-86: depth, intrinsics, normalised seen surface), with
    `with_intr=False` the depth estimator alone."""
    tp = Tape()
    B = rgb.shape[0]
    H, W = opt.H, opt.W
    depth_nhwc, l4 = dpt_forward(tp, graph.dpt_depth, rgb)
    depth = depth_nhwc.view(B, 1, H, W)
    if not with_intr:
        return tp, {"depth": depth, "depth_nhwc": depth_nhwc, "K": None, "seen": None, "mean": None, "scale": None, "latent": None}
    # intrinsics head: 2 x Bottleneck_Conv(768, k=3) with batch-statistics BatchNorm -> avg pool -> Linear(768, 3) -> K
    U = []
    f = cet._bneck_conv_fwd(U, l4, graph.intr_head[0])
    f = cet._bneck_conv_fwd(U, f, graph.intr_head[1])

    def bwd_head(f=f):
        df = tp.pop(f)
        if df is None:
            return
        G = cet._Grads()
        d = cet._bneck_conv_bwd(U, df, G)
        d = cet._bneck_conv_bwd(U, d, G)
        tp.add(l4, d)
        for p in graph.intr_head.parameters():
            if G.get(p) is not None:
                tp.padd(p, G.get(p))
    tp.record(bwd_head)
    pooled = ops.avgpool_nhwc(f)

    def bwd_pool():
        dpool = tp.pop(pooled)
        if dpool is not None:
            tp.add(f, ops.avgpool_bwd_nhwc(dpool, f.shape[1], f.shape[2]))
    tp.record(bwd_pool)
    params = linear(tp, pooled, graph.intr_proj)
    K = ops.intr_param2mtx(params.contiguous(), H, W)
    mask = mask.float().contiguous()
    seen, mean, scale = ops.unproject_normalize(depth, mask, K)

    def bwd_geom():
        dseen = tp.pop(seen)
        if dseen is None:
            return
        dd, dkinv = ops.unproject_normalize_bwd(depth, mask, K, seen, scale, dseen)
        tp.add(depth_nhwc, dd.view(depth_nhwc.shape))
        # K^-1 -> K -> the three intrinsics parameters (graph_shape.py:98-112): 3x3 / 3-vector algebra per image (host glue)
        kinv = _inv3x3(K)
        dK = -(kinv.transpose(1, 2) @ dkinv @ kinv.transpose(1, 2))
        t = torch.tanh(params)
        dt = 1.0 - t * t
        ln4 = math.log(4.0)
        dp = torch.stack([(dK[:, 0, 0] * K[:, 0, 0] + dK[:, 1, 1] * K[:, 1, 1]) * ln4 * dt[:, 0],
                          dK[:, 0, 2] * (W / 2.0) * dt[:, 1], dK[:, 1, 2] * (H / 2.0) * dt[:, 2]], dim=1)
        tp.add(params, dp)
    tp.record(bwd_geom)
    out = {"depth": depth, "depth_nhwc": depth_nhwc, "K": K, "seen": seen, "mean": mean, "scale": scale, "latent": None}
    if not with_coord:
        return tp, out
    enc = graph.coord_encoder
    if hasattr(enc, "coord_embed"):
        # transformer seen-surface encoder: mask-aware resampling to H/dsp x W/dsp, window tokens, transformer blocks -- all on the tape
        from ...model.shape import seen_coord_att_train as cat
        dsp = opt.arch.depth.dsp
        coord_map, mb = cat.resample_on_tape(tp, view(tp, seen, B, H, W, 3), mask.view(B, 1, H, W), H // dsp, W // dsp)
        out["latent"] = cat.train_forward(tp, enc, coord_map, mb)
        return tp, out
    coord = ops.axpby(seen.view(B, H, W, 3), 1.0 / (1.0 + 1.e-6))
    if enc.training:
        latent, T = cet.train_forward(enc, coord)

        def bwd_enc():
            dl = tp.pop(latent)
            if dl is None:
                return
            G, dcoord = cet.train_backward(enc, T, dl, need_dcoord=True)
            for p in enc.parameters():
                if G.get(p) is not None:
                    tp.padd(p, G.get(p))
            tp.add(seen, ops.axpby(dcoord.view(seen.shape), 1.0 / (1.0 + 1.e-6)))
        tp.record(bwd_enc)
    else:
        raise NotImplementedError("encoder_forward: eval-mode seen-surface encoder inside a training step")
    out["latent"] = latent
    return tp, out


class EncoderTrainFn(torch.autograd.Function):
    """depth, K, seen_points, latent = EncoderTrainFn.apply(graph, opt, rgb, mask, *encoder_parameters)"""

    @staticmethod
    def forward(ctx, graph, opt, rgb, mask, *params):
        ctx.set_materialize_grads(False)      # an output no loss uses arrives as None: no zero tensors, no host test for "all zero"
        with torch.no_grad():
            tp, out = encoder_forward(graph, opt, rgb, mask)
        ctx.tp, ctx.out, ctx.params = tp, out, params
        graph.last_mean, graph.last_scale = out["mean"], out["scale"]
        return out["depth"], out["K"], out["seen"], out["latent"]

    @staticmethod
    def backward(ctx, ddepth, dK, dseen, dlatent):
        tp, out = ctx.tp, ctx.out
        with torch.no_grad():
            if dlatent is not None:
                tp.add(out["latent"], dlatent)
            _seed_depth_and_seen(tp, out, ddepth, dseen)
            tp.backward()
        grads = tuple(tp.pgrads.get(id(p)) if p.requires_grad else None for p in ctx.params)
        ctx.tp = None
        return (None, None, None, None) + grads


def _seed_depth_and_seen(tp, out, ddepth, dseen):
    """Gradients arriving at depth_pred (MiDaS depth loss) and at the normalised seen surface (intrinsics loss); None (the
    Functions do not materialise the gradients of unused outputs) when no loss uses the output.  No host synchronisation:
    the whole step stays stream-ordered (and can be captured in a CUDA graph, zeroshape_b200/graphed.py)."""
    if dseen is not None and out["seen"] is not None:
        tp.add(out["seen"], dseen)
    if ddepth is not None:
        tp.add(out["depth_nhwc"], ddepth.contiguous().view(out["depth_nhwc"].shape))      # [B,1,H,W] and [B,H,W,1] share the memory order


class DepthGraphTrainFn(torch.autograd.Function):
    """depth[, K, seen_points] = DepthGraphTrainFn.apply(graph, opt, rgb, mask, with_intr, *parameters): the depth-only compute
    graph (model/compute_graph/graph_depth.py:61-86) on the tape, for `train.py options/depth.yaml`."""

    @staticmethod
    def forward(ctx, graph, opt, rgb, mask, with_intr, *params):
        ctx.set_materialize_grads(False)
        with torch.no_grad():
            tp, out = encoder_forward(graph, opt, rgb, mask, with_intr=with_intr, with_coord=False)
        ctx.tp, ctx.out, ctx.params, ctx.with_intr = tp, out, params, with_intr
        if not with_intr:
            return out["depth"]
        return out["depth"], out["K"], out["seen"]

    @staticmethod
    def backward(ctx, ddepth, dK=None, dseen=None):
        tp, out = ctx.tp, ctx.out
        with torch.no_grad():
            _seed_depth_and_seen(tp, out, ddepth, dseen)
            tp.backward()
        grads = tuple(tp.pgrads.get(id(p)) if p.requires_grad else None for p in ctx.params)
        ctx.tp = None
        return (None, None, None, None, None) + grads
