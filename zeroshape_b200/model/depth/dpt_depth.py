"""Host-side mirror of the reference's DPT-hybrid depth model with every layer running in the CUDA
library (NHWC, fused epilogues).  Same class name, state_dict keys and forward contract:

    DPTDepthModel(backbone='vitb_rn50_384').forward(image [B,3,H,W] in [0,1], get_feat=True)
        -> depth [B,1,H,W] clamped to [0,1], layer_4 feature [B,768,H/32,W/32]
    (reference: model/depth/dpt_depth.py:96-123, DPT.forward :68-94; glue model/depth/vit.py:57-154,344-476;
     blocks model/depth/blocks.py:50-76,232-342; backbone = timm 0.6.12 `vit_base_resnet50_384`,
     third-party: ResNetV2 (3,4,9) with weight-standardised SAME convs + GroupNorm(32) -> 1x1 proj -> ViT-B)

B200-first choices: activations NHWC fp32 so every conv is an implicit GEMM with a K-contiguous
operand; weight standardisation and the 24x24 -> 14x14 position-embedding resize (done on EVERY forward
by the reference, vit.py:101-123) are weight-only work and are cached per weight version; ReLU-conv,
conv-bias-residual, GroupNorm-add-ReLU are single fused launches; no module-global hook dictionary
(vit.py:157-164) -- the four tapped activations are plain locals.
"""
import math

import torch
import torch.nn as nn

from ... import ops
from ...packing import PackCache, ohwi, ws_ohwi


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container")


def _same_pad(i, k, s):
    return max((math.ceil(i / s) - 1) * s + (k - 1) + 1 - i, 0)


def _gn_params(c):
    m = _Holder()
    m.weight = nn.Parameter(torch.ones(c))
    m.bias = nn.Parameter(torch.zeros(c))
    return m


def _conv_w(cout, cin, k):
    m = _Holder()
    w = torch.empty(cout, cin, k, k)
    nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
    m.weight = nn.Parameter(w)
    return m


def _make_resnetv2():
    bb = _Holder()
    bb.stem = _Holder()
    bb.stem.conv = _conv_w(64, 3, 7)
    bb.stem.norm = _gn_params(64)
    stages = []
    cin = 64
    for depth, cout in zip((3, 4, 9), (256, 512, 1024)):
        st = _Holder()
        blocks = []
        mid = cout // 4
        for b in range(depth):
            blk = _Holder()
            if b == 0:
                blk.downsample = _Holder()
                blk.downsample.conv = _conv_w(cout, cin, 1)
                blk.downsample.norm = _gn_params(cout)
            blk.conv1, blk.norm1 = _conv_w(mid, cin, 1), _gn_params(mid)
            blk.conv2, blk.norm2 = _conv_w(mid, mid, 3), _gn_params(mid)
            blk.conv3, blk.norm3 = _conv_w(cout, mid, 1), _gn_params(cout)
            blocks.append(blk)
            cin = cout
        st.blocks = nn.ModuleList(blocks)
        stages.append(st)
    bb.stages = nn.ModuleList(stages)
    return bb


def _make_vit_block(dim=768, mlp=3072):
    blk = _Holder()
    blk.norm1 = nn.LayerNorm(dim, eps=1e-6)
    blk.attn = _Holder()
    blk.attn.qkv = nn.Linear(dim, dim * 3)
    blk.attn.proj = nn.Linear(dim, dim)
    blk.norm2 = nn.LayerNorm(dim, eps=1e-6)
    blk.mlp = _Holder()
    blk.mlp.fc1 = nn.Linear(dim, mlp)
    blk.mlp.fc2 = nn.Linear(mlp, dim)
    return blk


def _make_readout(features_out, extra_conv):
    """act_postprocess{3,4}: indices 0 (ProjectReadout), 3 (1x1 conv), 4 (3x3 s2 conv) hold parameters."""
    seq = _Holder()
    ro = _Holder()
    ro.project = nn.Sequential(nn.Linear(2 * 768, 768), nn.GELU())
    seq.add_module("0", ro)
    seq.add_module("3", nn.Conv2d(768, features_out, 1))
    if extra_conv:
        seq.add_module("4", nn.Conv2d(features_out, features_out, 3, stride=2, padding=1))
    return seq


def _make_fusion(features=256):
    f = _Holder()
    f.out_conv = nn.Conv2d(features, features, 1)
    for name in ("resConfUnit1", "resConfUnit2"):
        u = _Holder()
        u.conv1 = nn.Conv2d(features, features, 3, padding=1)
        u.conv2 = nn.Conv2d(features, features, 3, padding=1)
        setattr(f, name, u)
    return f


class DPTDepthModel(nn.Module):
    """DPT-hybrid monocular depth (reference: DPTDepthModel, model/depth/dpt_depth.py:96-123)."""

    HOOKS = (8, 11)   # ViT blocks tapped for layer_3 / layer_4 (dpt_depth.py:42)
    HEADS = 12

    def __init__(self, path=None, non_negative=True, num_channels=1, backbone="vitb_rn50_384", features=256, **kwargs):
        super().__init__()
        if backbone != "vitb_rn50_384":
            raise NotImplementedError("only the vitb_rn50_384 backbone is used by ZeroShape (graph_shape.py:31)")
        assert non_negative and num_channels == 1 and features == 256
        self.pretrained = _Holder()
        vit = _Holder()
        vit.cls_token = nn.Parameter(torch.zeros(1, 1, 768))
        vit.pos_embed = nn.Parameter(torch.randn(1, 577, 768) * 0.02)
        vit.patch_embed = _Holder()
        vit.patch_embed.backbone = _make_resnetv2()
        vit.patch_embed.proj = nn.Conv2d(1024, 768, 1)
        vit.blocks = nn.ModuleList([_make_vit_block() for _ in range(12)])
        vit.norm = nn.LayerNorm(768, eps=1e-6)
        vit.head = nn.Linear(768, 1000)            # unused ImageNet classifier, present in checkpoints
        self.pretrained.model = vit
        self.pretrained.act_postprocess3 = _make_readout(768, False)
        self.pretrained.act_postprocess4 = _make_readout(768, True)
        self.scratch = _Holder()
        for i, c in enumerate((256, 512, 768, 768)):
            setattr(self.scratch, f"layer{i + 1}_rn", nn.Conv2d(c, features, 3, padding=1, bias=False))
        for i in range(1, 5):
            setattr(self.scratch, f"refinenet{i}", _make_fusion(features))
        head = _Holder()
        head.add_module("0", nn.Conv2d(features, features // 2, 3, padding=1))
        head.add_module("2", nn.Conv2d(features // 2, 32, 3, padding=1))
        head.add_module("4", nn.Conv2d(32, 1, 1))
        nn.init.constant_(getattr(head, "4").bias, 0.05)       # dpt_depth.py:109
        self.scratch.output_conv = head
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                nn.init.zeros_(m.bias)
        self._cache = PackCache(self)
        if path is not None:
            self.load_state_dict(torch.load(path, map_location="cpu"))

    # -- packed weights ----------------------------------------------------------------------------
    def _w(self, name, conv, ws=False):
        return self._cache.get(name, lambda: ws_ohwi(conv.weight) if ws else ohwi(conv.weight))

    def _pos_embed(self, gh, gw):
        def build():
            pe = self.pretrained.model.pos_embed.detach().float()
            g_old = int(math.sqrt(pe.shape[1] - 1))
            grid = pe[:, 1:].reshape(1, g_old, g_old, -1).contiguous()          # already NHWC
            grid = ops.bilinear_nhwc(grid, gh, gw, False).reshape(1, gh * gw, -1)
            return torch.cat([pe[:, :1], grid], dim=1).contiguous()
        return self._cache.get(f"pos.{gh}x{gw}", build)

    # -- stages ------------------------------------------------------------------------------------
    def _gn(self, x, norm, relu, res=None):
        return ops.groupnorm_nhwc(x, norm.weight, norm.bias, 32, 1e-5, relu, res)

    def _ws_conv(self, x, name, conv, stride):
        k = conv.weight.shape[-1]
        ph, pw = _same_pad(x.shape[1], k, stride), _same_pad(x.shape[2], k, stride)
        return ops.conv2d_nhwc(x, self._w(name, conv, ws=True), None, stride, (ph // 2, ph - ph // 2, pw // 2, pw - pw // 2))

    def _resnetv2(self, x):
        bb = self.pretrained.model.patch_embed.backbone
        x = self._gn(self._ws_conv(x, "stem", bb.stem.conv, 2), bb.stem.norm, True)
        ph, pw = _same_pad(x.shape[1], 3, 2), _same_pad(x.shape[2], 3, 2)
        x = ops.maxpool3x3s2_nhwc(x, ph // 2, pw // 2, (x.shape[1] + ph - 3) // 2 + 1, (x.shape[2] + pw - 3) // 2 + 1)
        outs = []
        for s, st in enumerate(bb.stages):
            for b, blk in enumerate(st.blocks):
                stride = 2 if (b == 0 and s > 0) else 1
                t = f"s{s}b{b}"
                short = x
                if b == 0:
                    short = self._gn(self._ws_conv(x, t + "d", blk.downsample.conv, stride), blk.downsample.norm, False)
                y = self._gn(self._ws_conv(x, t + "c1", blk.conv1, 1), blk.norm1, True)
                y = self._gn(self._ws_conv(y, t + "c2", blk.conv2, stride), blk.norm2, True)
                x = self._gn(self._ws_conv(y, t + "c3", blk.conv3, 1), blk.norm3, True, res=short)   # relu(gn(y) + shortcut)
            outs.append(x)
        return outs

    def _vit(self, feat, B, gh, gw):
        vit = self.pretrained.model
        tok = ops.conv2d_nhwc(feat, self._w("proj", vit.patch_embed.proj), vit.patch_embed.proj.bias).view(B, gh * gw, 768)
        x = torch.cat([vit.cls_token.detach().expand(B, -1, -1), tok], dim=1).contiguous()
        x = ops.axpby(x, 1.0, self._pos_embed(gh, gw).expand(B, -1, -1).contiguous(), 1.0)
        taps = {}
        for i, blk in enumerate(vit.blocks):
            h = ops.layernorm(x, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
            a = ops.mha(ops.linear(h, blk.attn.qkv.weight, blk.attn.qkv.bias), self.HEADS)
            x = ops.linear(a, blk.attn.proj.weight, blk.attn.proj.bias, res=x)
            h = ops.layernorm(x, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
            h = ops.linear(h, blk.mlp.fc1.weight, blk.mlp.fc1.bias, act=ops.ACT_GELU)
            x = ops.linear(h, blk.mlp.fc2.weight, blk.mlp.fc2.bias, res=x)
            if i in self.HOOKS:
                taps[i] = x
            if i == self.HOOKS[-1]:
                break          # blocks after the last tap and the final norm do not reach any output (vit.py:152)
        return taps

    def _reassemble(self, tokens, seq, tag, gh, gw):
        """ProjectReadout + (token-major == NHWC) + 1x1 conv (+3x3 s2 conv): vit.py:32-43,66-96,430-461."""
        B, T, C = tokens.shape
        lin = getattr(seq, "0").project[0]
        cat = torch.empty(B, T - 1, 2 * C, device=tokens.device, dtype=torch.float32)
        for b in range(B):   # [patch tokens | cls broadcast]; ldb = 0 broadcasts the cls row
            cat[b] = ops.concat2(tokens[b, 1:], tokens[b, :1].expand(T - 1, C), 1.0)
        y = ops.linear(cat, lin.weight, lin.bias, act=ops.ACT_GELU).view(B, gh, gw, C)
        c3 = getattr(seq, "3")
        y = ops.conv2d_nhwc(y, self._w(tag + ".3", c3), c3.bias)
        if hasattr(seq, "4"):
            c4 = getattr(seq, "4")
            y = ops.conv2d_nhwc(y, self._w(tag + ".4", c4), c4.bias, 2, (1, 1, 1, 1))
        return y

    def _rcu(self, x, unit, tag, extra=None):
        """ResidualConvUnit_custom (blocks.py:264-287): conv2(relu(conv1(relu(x)))) + x  [+ extra]."""
        y = ops.conv2d_nhwc(x, self._w(tag + "1", unit.conv1), unit.conv1.bias, 1, (1, 1, 1, 1), pre_relu=True)
        y = ops.conv2d_nhwc(y, self._w(tag + "2", unit.conv2), unit.conv2.bias, 1, (1, 1, 1, 1), pre_relu=True, res=x)
        return ops.axpby(y, 1.0, extra, 1.0) if extra is not None else y

    def _fusion(self, i, x, skip=None):
        f = getattr(self.scratch, f"refinenet{i}")
        if skip is not None:
            x = self._rcu(skip, f.resConfUnit1, f"rn{i}u1", extra=x)
        x = self._rcu(x, f.resConfUnit2, f"rn{i}u2")
        x = ops.bilinear_nhwc(x, x.shape[1] * 2, x.shape[2] * 2, True)
        return ops.conv2d_nhwc(x, self._w(f"rn{i}o", f.out_conv), f.out_conv.bias)

    # -- forward -----------------------------------------------------------------------------------
    def forward(self, image, get_feat=False):
        if torch.is_grad_enabled() and (image.requires_grad or any(p.requires_grad for p in self.parameters()) and self.training):
            raise NotImplementedError("zeroshape_b200.DPTDepthModel: backward is not implemented in this revision; "
                                      "call under torch.no_grad() / eval()")
        with torch.no_grad():
            self._cache.refresh()
            B, _, H, W = image.shape
            gh, gw = H // 16, W // 16
            x = ops.nchw_to_nhwc(image.float().contiguous(), 2.0, -1.0)            # image * 2 - 1 (dpt_depth.py:116)
            s0, s1, s2 = self._resnetv2(x)
            taps = self._vit(s2, B, gh, gw)
            l3 = self._reassemble(taps[self.HOOKS[0]], self.pretrained.act_postprocess3, "pp3", gh, gw)
            l4 = self._reassemble(taps[self.HOOKS[1]], self.pretrained.act_postprocess4, "pp4", gh, gw)
            sc = self.scratch
            r = [ops.conv2d_nhwc(t, self._w(f"rn_in{i}", getattr(sc, f"layer{i + 1}_rn")), None, 1, (1, 1, 1, 1))
                 for i, t in enumerate((s0, s1, l3, l4))]
            path = self._fusion(4, r[3])
            path = self._fusion(3, path, r[2])
            path = self._fusion(2, path, r[1])
            path = self._fusion(1, path, r[0])
            oc = sc.output_conv
            c0, c2, c4 = getattr(oc, "0"), getattr(oc, "2"), getattr(oc, "4")
            y = ops.conv2d_nhwc(path, self._w("h0", c0), c0.bias, 1, (1, 1, 1, 1))
            y = ops.bilinear_nhwc(y, y.shape[1] * 2, y.shape[2] * 2, True)
            y = ops.conv2d_nhwc(y, self._w("h2", c2), c2.bias, 1, (1, 1, 1, 1), act=ops.ACT_RELU)
            y = ops.conv2d_nhwc(y, self._w("h4", c4), c4.bias, act=ops.ACT_CLAMP01)     # relu + clamp(0,1) (:106,119)
            depth = y.view(B, 1, y.shape[1], y.shape[2])                                 # C == 1: NHWC == NCHW
            self.last_feat_nhwc = l4
            if get_feat:
                return depth, ops.nhwc_to_nchw(l4)
            return depth
