"""Host-side mirror of the reference's ResNet-50 seen-surface encoder (the configured one:
options/shape.yaml:26 `arch.depth.encoder: resnet`), every layer in the CUDA library.

    CoordEncRes(opt).forward(coord_obj [B,3,H,W], mask_obj [B,1,H,W]) -> latent [B, 1+H/16*W/16, latent_dim]
    (reference: model/shape/seen_coord_enc.py:141-194 over torchvision resnet50)

state_dict keys = torchvision's (`encoder.conv1/bn1/layerN.M.{conv,bn}K/downsample.{0,1}`) plus
`encoder.fc.{0,1}` Bottleneck_Conv(2048), `encoder.fc.2` Linear and `depth_feat_proj.{0,1,2}`.
Eval-mode BatchNorm is folded into the (OHWI) filters once per weight version; conv+BN+ReLU and
conv+BN+add+ReLU are single launches.  In train mode the forward uses batch statistics and is differentiable
(seen_coord_enc_train.py: the `optim.fix_dpt` training configuration).

The transformer variant `CoordEncAtt` (+ `CoordEmb`; model/shape/seen_coord_enc.py:13-139, `arch.depth.encoder != resnet`) is mirrored for
inference: the window front end is one launch (zs_coord_embed_windows_f32), every Block runs on the LayerNorm / tcgen05 linear / attention
kernels of the ViT path; in train mode it runs on the tape of seen_coord_att_train.py (DropPath, hand-written backward).
"""
import torch
import torch.nn as nn

from ... import ops
from ...packing import PackCache, fold_bn_ohwi, ohwi
from ...utils.layers import Bottleneck_Conv


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container")


def _tv_bottleneck(cin, width, stride, down):
    b = _Holder()
    b.conv1, b.bn1 = nn.Conv2d(cin, width, 1, bias=False), nn.BatchNorm2d(width)
    b.conv2, b.bn2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, bias=False), nn.BatchNorm2d(width)
    b.conv3, b.bn3 = nn.Conv2d(width, width * 4, 1, bias=False), nn.BatchNorm2d(width * 4)
    if down:
        b.downsample = nn.Sequential(nn.Conv2d(cin, width * 4, 1, stride=stride, bias=False), nn.BatchNorm2d(width * 4))
    b.stride = stride
    return b


def _tv_resnet50():
    r = _Holder()
    r.conv1, r.bn1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), nn.BatchNorm2d(64)
    cin = 64
    for li, (depth, width, stride) in enumerate(((3, 64, 1), (4, 128, 2), (6, 256, 2), (3, 512, 2)), start=1):
        blocks = []
        for b in range(depth):
            blocks.append(_tv_bottleneck(cin, width, stride if b == 0 else 1, b == 0))
            cin = width * 4
        setattr(r, f"layer{li}", nn.ModuleList(blocks))
    for m in r.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    return r


class CoordEncRes(nn.Module):
    def __init__(self, opt):
        super().__init__()
        assert opt.arch.depth.dsp == 1
        if opt.arch.win_size != 16:
            raise NotImplementedError("win_size 32 variant is not the shipped configuration (options/shape.yaml:23)")
        latent = opt.arch.latent_dim
        self.encoder = _tv_resnet50()
        # seen_coord_enc.py:148 starts the trunk from torchvision's ImageNet weights (`resnet50(pretrained=True)`, a download).
        # No network here: the same state_dict is taken from a local file -- `opt.arch.depth.resnet50_weights` or the
        # environment variable ZEROSHAPE_RESNET50_WEIGHTS (torchvision's resnet50-*.pth) -- loaded before `fc` is replaced,
        # exactly like the reference.  Checkpoint-driven inference (demo.py / evaluate.py) overwrites every weight anyway.
        import os
        path = getattr(opt.arch.depth, "resnet50_weights", None) if hasattr(opt.arch, "depth") else None
        path = path or os.environ.get("ZEROSHAPE_RESNET50_WEIGHTS")
        self.imagenet_init = False
        if path:
            sd = torch.load(path, map_location="cpu")
            sd = sd.get("state_dict", sd) if isinstance(sd, dict) else sd
            missing, unexpected = self.encoder.load_state_dict({k: v for k, v in sd.items() if not k.startswith("fc.")}, strict=False)
            if missing:
                raise RuntimeError(f"CoordEncRes: {path} lacks ResNet-50 trunk tensors: {missing[:4]} ...")
            self.imagenet_init = True
        self._warned_init = False
        self.encoder.fc = nn.Sequential(Bottleneck_Conv(2048), Bottleneck_Conv(2048), nn.Linear(2048, latent))
        self.depth_feat_proj = nn.Sequential(Bottleneck_Conv(1024), Bottleneck_Conv(1024), nn.Conv2d(1024, latent, 1))
        self._cache = PackCache(self)

    def _load_from_state_dict(self, *args, **kwargs):
        self._loaded_checkpoint = True          # weights come from a checkpoint: the initialisation no longer matters
        return super()._load_from_state_dict(*args, **kwargs)

    def _cbr(self, x, conv, bn, tag, act, stride=1, pad=0, res=None):
        w, b = self._cache.get(tag, lambda: fold_bn_ohwi(conv.weight, bn))
        return ops.conv2d_nhwc(x, w, b, stride, (pad, pad, pad, pad), act=act, res=res,
                               res_mode=ops.RES_BEFORE_ACT if res is not None else ops.RES_NONE)

    def forward_nhwc(self, coord_nhwc):
        """coord_nhwc [B,H,W,3] (already multiplied by the mask) -> [B,197,latent]."""
        if self.training:
            # batch-statistics BatchNorm + saved activations + hand-written backward (seen_coord_enc_train.py)
            from .seen_coord_enc_train import CoordEncTrainFn, train_forward
            params = list(self.parameters())
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                if not self.imagenet_init and not self._warned_init and not getattr(self, "_loaded_checkpoint", False):
                    import warnings
                    warnings.warn("CoordEncRes: training starts with a ResNet-50 trunk that was neither initialised from ImageNet "
                                  "weights (the reference: torchvision resnet50(pretrained=True), seen_coord_enc.py:148) nor loaded "
                                  "from a checkpoint -- set opt.arch.depth.resnet50_weights or ZEROSHAPE_RESNET50_WEIGHTS",
                                  stacklevel=2)
                    self._warned_init = True
                return CoordEncTrainFn.apply(self, coord_nhwc, *params)
            with torch.no_grad():
                return train_forward(self, coord_nhwc.float().contiguous())[0]
        self._cache.refresh()
        enc = self.encoder
        B = coord_nhwc.shape[0]
        x = self._cbr(coord_nhwc, enc.conv1, enc.bn1, "stem", ops.ACT_RELU, 2, 3)
        x = ops.maxpool3x3s2_nhwc(x, 1, 1, (x.shape[1] + 2 - 3) // 2 + 1, (x.shape[2] + 2 - 3) // 2 + 1)
        feats = {}
        for li in range(1, 5):
            for bi, blk in enumerate(getattr(enc, f"layer{li}")):
                t = f"l{li}b{bi}"
                idt = x
                if hasattr(blk, "downsample"):
                    idt = self._cbr(x, blk.downsample[0], blk.downsample[1], t + "d", ops.ACT_NONE, blk.stride)
                y = self._cbr(x, blk.conv1, blk.bn1, t + "c1", ops.ACT_RELU)
                y = self._cbr(y, blk.conv2, blk.bn2, t + "c2", ops.ACT_RELU, blk.stride, 1)
                x = self._cbr(y, blk.conv3, blk.bn3, t + "c3", ops.ACT_RELU, res=idt)
            feats[li] = x
        g = ops.avgpool_nhwc(feats[4])                                    # [B,2048]
        g = enc.fc[0].run_nhwc(g, self._cache, "fc0")
        g = enc.fc[1].run_nhwc(g, self._cache, "fc1")
        g = ops.linear(g, enc.fc[2].weight, enc.fc[2].bias).unsqueeze(1)  # [B,1,latent]
        y = self.depth_feat_proj[0].run_nhwc(feats[3], self._cache, "p0")
        y = self.depth_feat_proj[1].run_nhwc(y, self._cache, "p1")
        pc = self.depth_feat_proj[2]
        y = ops.conv2d_nhwc(y, self._cache.get("p2", lambda: ohwi(pc.weight)), pc.bias)
        return torch.cat([g, y.view(B, -1, y.shape[-1])], dim=1).contiguous()

    def forward(self, coord_obj, mask_obj):
        assert coord_obj.dim() == 4 and mask_obj.dim() == 4
        with torch.no_grad():
            x = ops.nchw_to_nhwc((coord_obj * mask_obj.float()).float().contiguous())
        if self.training:
            return self.forward_nhwc(x)
        with torch.no_grad():
            return self.forward_nhwc(x)


# ---- transformer seen-surface encoder (SURVEY.md section 8a row a7') ---------------------------------------------------
def _sincos_1d(dim, pos):
    """utils/pos_embed.py:53-70."""
    import numpy as np
    omega = np.arange(dim // 2, dtype=np.float32)
    omega /= dim / 2.
    omega = 1. / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    """utils/pos_embed.py:21-50 (MAE-style fixed embedding; w goes first in the meshgrid)."""
    import numpy as np
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape([2, 1, grid_size, grid_size])
    emb = np.concatenate([_sincos_1d(embed_dim // 2, grid[0]), _sincos_1d(embed_dim // 2, grid[1])], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb


def _make_block(dim, mlp_ratio):
    """Parameter container with timm Block's sub-module names (norm1, attn.qkv, attn.proj, norm2, mlp.fc1, mlp.fc2)."""
    blk = _Holder()
    blk.norm1 = nn.LayerNorm(dim, eps=1e-6)
    blk.attn = _Holder()
    blk.attn.qkv = nn.Linear(dim, dim * 3)
    blk.attn.proj = nn.Linear(dim, dim)
    blk.norm2 = nn.LayerNorm(dim, eps=1e-6)
    blk.mlp = _Holder()
    blk.mlp.fc1 = nn.Linear(dim, int(dim * mlp_ratio))
    blk.mlp.fc2 = nn.Linear(int(dim * mlp_ratio), dim)
    return blk


def _run_block(x, blk, heads):
    """timm Block forward (eval: DropPath is the identity) on [N, T, C]."""
    h = ops.layernorm(x, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
    a = ops.mha(ops.linear(h, blk.attn.qkv.weight, blk.attn.qkv.bias), heads)
    x = ops.linear(a, blk.attn.proj.weight, blk.attn.proj.bias, res=x)
    h = ops.layernorm(x, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
    h = ops.linear(h, blk.mlp.fc1.weight, blk.mlp.fc1.bias, act=ops.ACT_GELU)
    return ops.linear(h, blk.mlp.fc2.weight, blk.mlp.fc2.bias, res=x)


class CoordEmb(nn.Module):
    """model/shape/seen_coord_enc.py:13-78: every ws x ws window of the XYZ map -> one token (cls row of a single Block)."""

    def __init__(self, embed_dim, win_size=8, num_heads=8):
        super().__init__()
        self.embed_dim, self.win_size, self.num_heads = embed_dim, win_size, num_heads
        self.two_d_pos_embed = nn.Parameter(torch.zeros(1, win_size * win_size + 1, embed_dim), requires_grad=False)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Linear(3, embed_dim)
        self.blocks = nn.ModuleList([_make_block(embed_dim, 2.0)])
        self.invalid_coord_token = nn.Parameter(torch.zeros(embed_dim,))
        nn.init.normal_(self.cls_token, std=.02)
        self.two_d_pos_embed.data.copy_(torch.from_numpy(get_2d_sincos_pos_embed(embed_dim, win_size, cls_token=True)).float().unsqueeze(0))
        nn.init.normal_(self.invalid_coord_token, std=.02)

    def forward(self, coord_obj, mask_obj):
        """coord_obj [B,H,W,3], mask_obj bool [B,H,W] -> [B, (H/ws)*(W/ws), C]."""
        B, H, W, _ = coord_obj.shape
        ws = self.win_size
        emb = ops.coord_embed_windows(coord_obj.float().contiguous(), mask_obj.float().contiguous(), self.pos_embed.weight.detach(),
                                      self.pos_embed.bias.detach(), self.invalid_coord_token.detach(),
                                      self.two_d_pos_embed.detach().reshape(ws * ws + 1, -1).contiguous(),
                                      self.cls_token.detach().reshape(-1).contiguous(), ws)
        for blk in self.blocks:
            emb = _run_block(emb, blk, self.num_heads)
        return emb[:, 0].reshape(B, (H // ws) * (W // ws), -1)


class CoordEncAtt(nn.Module):
    """model/shape/seen_coord_enc.py:80-139: window tokens + a global cls token through n_blocks transformer blocks."""

    def __init__(self, embed_dim=768, n_blocks=12, num_heads=12, win_size=8, mlp_ratio=4., drop_path=0.1):
        super().__init__()
        self.num_heads = num_heads
        self.drop_path = drop_path                               # timm DropPath of the main blocks (train mode only)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.coord_embed = CoordEmb(embed_dim, win_size, num_heads)
        self.blocks = nn.ModuleList([_make_block(embed_dim, mlp_ratio) for _ in range(n_blocks)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        nn.init.normal_(self.cls_token, std=.02)
        for m in self.modules():                                 # seen_coord_enc.py:109-117
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    def forward(self, coord_obj, mask_obj):
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .seen_coord_att_train import CoordAttTrainFn      # DropPath + hand-written backward on the tape
            return CoordAttTrainFn.apply(self, coord_obj, mask_obj.float(), *list(self.parameters()))
        with torch.no_grad():
            x = self.coord_embed(coord_obj, mask_obj)
            x = torch.cat([self.cls_token.detach().expand(x.shape[0], -1, -1), x], dim=1).contiguous()
            for blk in self.blocks:
                x = _run_block(x, blk, self.num_heads)
            return ops.layernorm(x, self.norm.weight, self.norm.bias, self.norm.eps)
