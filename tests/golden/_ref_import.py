"""Import shims that let the real reference modules under /root/reference be imported in the
build container (no GPU, no timm / mcubes / trimesh).  TEST INFRASTRUCTURE for golden generation
(`make_golden.py`) and for the optional direct-reference tests; never used on the GPU box
(/root/reference does not exist there) and never imported by the product.

timm==0.6.12 is absent, so the four layer classes the reference imports from
``timm.models.vision_transformer`` (model/shape/implicit.py:8, seen_coord_enc.py:9, utils/layers.py:5)
are restated here from the published timm 0.6.12 definitions (names of sub-modules kept, because
they become state_dict keys).
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("ZEROSHAPE_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "model", "shape"))


class _Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


class _Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        x = (self.attn_drop(attn) @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x))


class _Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0,
                 init_values=None, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.ls1 = nn.Identity()
        self.drop_path1 = _DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.ls2 = nn.Identity()
        self.drop_path2 = _DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x):
        x = x + self.drop_path1(self.ls1(self.attn(self.norm1(x))))
        return x + self.drop_path2(self.ls2(self.mlp(self.norm2(x))))


class _PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


def install_shims():
    """Put /root/reference and stand-ins for its missing third-party imports on sys.path/modules."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if "timm" not in sys.modules:
        import importlib.machinery
        timm = types.ModuleType("timm")
        timm.models = types.ModuleType("timm.models")
        vt = types.ModuleType("timm.models.vision_transformer")
        for m_ in (timm, timm.models, vt):      # importlib.util.find_spec() rejects modules without a spec
            m_.__spec__ = importlib.machinery.ModuleSpec(m_.__name__, None)
        vt.Mlp, vt.DropPath, vt.Block, vt.PatchEmbed, vt.Attention = _Mlp, _DropPath, _Block, _PatchEmbed, _Attention
        timm.models.vision_transformer = vt

        def _no_create(*a, **k):
            raise RuntimeError("timm.create_model is not available in the shim")
        timm.create_model = _no_create
        sys.modules.update({"timm": timm, "timm.models": timm.models, "timm.models.vision_transformer": vt})
    for missing in ("mcubes", "trimesh"):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except ImportError:
                sys.modules[missing] = types.ModuleType(missing)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # torchvision.resnet50(pretrained=True) would hit the network (seen_coord_enc.py:148)
    import torchvision
    if not getattr(torchvision.models, "_zs_patched", False):
        orig = torchvision.models.resnet50

        def resnet50_offline(*a, **k):
            k.pop("pretrained", None)
            k["weights"] = None
            return orig(**k)
        torchvision.models.resnet50 = resnet50_offline
        torchvision.models._zs_patched = True


def fill_deterministic(module_or_sd, seed):
    """Overwrite every float tensor of a state_dict with reproducible values (independent of the
    module's own init RNG order): keys sorted, one generator, fan-in scaled normal for >=2-D,
    small noise around 1 for norm weights / running_var, small noise for biases and means."""
    sd = module_or_sd.state_dict() if isinstance(module_or_sd, nn.Module) else module_or_sd
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(sd.keys()):
        v = sd[k]
        if not torch.is_floating_point(v):
            out[k] = v.clone()
            continue
        if k.endswith("pos_embed") and v.ndim == 3 and v.shape[0] == 1:
            out[k] = v.clone()           # keep the fixed sincos table
        elif v.ndim >= 2:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) * (1.0 / fan_in) ** 0.5
        elif k.endswith("running_var"):
            out[k] = 1.0 + 0.2 * torch.rand(v.shape, generator=g)
        elif k.endswith(".weight"):
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
    if isinstance(module_or_sd, nn.Module):
        module_or_sd.load_state_dict(out, strict=True)
    return out


# ---------------------------------------------------------------------------------------------------
# Stand-in for timm.create_model("vit_base_resnet50_384") so that the reference's OWN glue code
# (model/depth/vit.py forward_flex + hooks + act_postprocess, blocks.py, dpt_depth.py, graph_shape.py)
# can be executed for real on CPU.  Sub-module names reproduce timm 0.6.12's, so the state_dict keys
# under `dpt_depth.pretrained.model.*` are the ones SURVEY.md section 8(b) lists.
import math as _math
import torch.nn.functional as _F


class _StdConv2dSame(nn.Conv2d):
    def __init__(self, cin, cout, k, stride=1, eps=1e-8):
        super().__init__(cin, cout, k, stride=stride, padding=0, bias=False)
        self.eps = eps

    def forward(self, x):
        def pad(i, k, s):
            return max((_math.ceil(i / s) - 1) * s + (k - 1) + 1 - i, 0)
        ph, pw = pad(x.shape[-2], self.kernel_size[0], self.stride[0]), pad(x.shape[-1], self.kernel_size[1], self.stride[1])
        x = _F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2])
        w = _F.batch_norm(self.weight.reshape(1, self.out_channels, -1), None, None, training=True, momentum=0.,
                          eps=self.eps).reshape_as(self.weight)
        return _F.conv2d(x, w, None, self.stride)


class _GroupNormAct(nn.GroupNorm):
    def __init__(self, c, act=True):
        super().__init__(32, c, eps=1e-5)
        self.apply_act = act

    def forward(self, x):
        x = _F.group_norm(x, self.num_groups, self.weight, self.bias, self.eps)
        return _F.relu(x) if self.apply_act else x


class _MaxPoolSame(nn.Module):
    def forward(self, x):
        def pad(i):
            return max((_math.ceil(i / 2) - 1) * 2 + 2 + 1 - i, 0)
        ph, pw = pad(x.shape[-2]), pad(x.shape[-1])
        x = _F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2], value=float("-inf"))
        return _F.max_pool2d(x, 3, 2)


class _DownsampleConv(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv = _StdConv2dSame(cin, cout, 1, stride)
        self.norm = _GroupNormAct(cout, act=False)

    def forward(self, x):
        return self.norm(self.conv(x))


class _BottleneckV2(nn.Module):
    def __init__(self, cin, cout, stride, first):
        super().__init__()
        mid = cout // 4
        self.downsample = _DownsampleConv(cin, cout, stride) if first else None
        self.conv1, self.norm1 = _StdConv2dSame(cin, mid, 1), _GroupNormAct(mid)
        self.conv2, self.norm2 = _StdConv2dSame(mid, mid, 3, stride), _GroupNormAct(mid)
        self.conv3, self.norm3 = _StdConv2dSame(mid, cout, 1), _GroupNormAct(cout, act=False)

    def forward(self, x):
        s = x if self.downsample is None else self.downsample(x)
        x = self.norm3(self.conv3(self.norm2(self.conv2(self.norm1(self.conv1(x))))))
        return _F.relu(x + s)


class _Stage(nn.Module):
    def __init__(self, cin, cout, stride, depth):
        super().__init__()
        self.blocks = nn.Sequential(*[_BottleneckV2(cin if b == 0 else cout, cout, stride if b == 0 else 1, b == 0)
                                      for b in range(depth)])

    def forward(self, x):
        return self.blocks(x)


class _ResNetV2(nn.Module):
    def __init__(self):
        super().__init__()
        self.stem = nn.Sequential()
        self.stem.add_module("conv", _StdConv2dSame(3, 64, 7, 2))
        self.stem.add_module("norm", _GroupNormAct(64))
        self.stem.add_module("pool", _MaxPoolSame())
        self.stages = nn.Sequential(_Stage(64, 256, 1, 3), _Stage(256, 512, 2, 4), _Stage(512, 1024, 2, 9))
        self.norm = nn.Identity()

    def forward(self, x):
        return self.norm(self.stages(self.stem(x)))


class _HybridEmbed(nn.Module):
    def __init__(self):
        super().__init__()
        self.backbone = _ResNetV2()
        self.proj = nn.Conv2d(1024, 768, 1)


class FakeTimmHybridViT(nn.Module):
    """Attribute surface the reference touches: patch_embed.{backbone,proj}, cls_token, pos_embed,
    pos_drop, blocks, norm (+ unused head)."""

    def __init__(self):
        super().__init__()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, 768))
        self.pos_embed = nn.Parameter(torch.zeros(1, 577, 768))
        self.pos_drop = nn.Dropout(0.0)
        self.patch_embed = _HybridEmbed()
        self.blocks = nn.Sequential(*[_Block(768, 12, 4.0, qkv_bias=True, norm_layer=lambda d: nn.LayerNorm(d, eps=1e-6))
                                      for _ in range(12)])
        self.norm = nn.LayerNorm(768, eps=1e-6)
        self.head = nn.Linear(768, 1000)


def install_fake_timm_factory():
    install_shims()
    sys.modules["timm"].create_model = lambda name, pretrained=False, **k: FakeTimmHybridViT()


def reference_opt(device="cpu"):
    """EasyDict with the fields graph_shape.Graph / Loss read (options/shape.yaml defaults)."""
    from utils.util import EasyDict as edict
    return edict(reference_opt_dict(device))


def reference_opt_dict(device="cpu"):
    return (dict(
        device=device, H=224, W=224,
        pretrain=dict(depth=None),
        arch=dict(num_heads=8, latent_dim=256, win_size=16,
                  depth=dict(encoder="resnet", n_blocks=12, dsp=2, pretrained=None), rgb=dict(encoder=None, n_blocks=12),
                  impl=dict(n_channels=256, att_blocks=2, mlp_ratio=4., posenc_perlayer=False, mlp_layers=8, posenc_3D=0,
                            skip_in=[2, 4, 6])),
        optim=dict(fix_dpt=False),
        training=dict(depth_loss=dict(grad_reg=0.1, depth_inv=True, mask_shrink=False),
                      shape_loss=dict(impt_weight=1, impt_thres=0.01)),
        loss_weight=dict(shape=1, depth=None, intr=None),
        eval=dict(vox_res=64, range=[-1.5, 1.5], num_points=10000, brute_force=False, icp=False,
                  f_thresholds=[0.005, 0.01, 0.02, 0.05, 0.1, 0.2]),
        data=dict(dataset_test="synthetic"),
    ))
