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
        timm = types.ModuleType("timm")
        timm.models = types.ModuleType("timm.models")
        vt = types.ModuleType("timm.models.vision_transformer")
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
