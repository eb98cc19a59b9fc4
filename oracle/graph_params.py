"""Oracle (TEST INFRASTRUCTURE): name -> shape table of the reference `graph_shape.Graph` state_dict
(shipped config: options/shape.yaml) and a seeded initialiser.  Key names follow SURVEY.md section 8(b):
timm 0.6.12 hybrid ViT under `dpt_depth.pretrained.model.*`, DPT reassemble/scratch
(model/depth/vit.py:430-461, blocks.py:50-76,290-319, dpt_depth.py:100-108), torchvision resnet50 under
`coord_encoder.encoder.*` (model/shape/seen_coord_enc.py:145-178), Bottleneck_Conv (utils/layers.py:76-82),
intrinsics head (graph_shape.py:19-28) and the Implicit decoder (model/shape/implicit.py)."""
import math

import torch

from .implicit import implicit_param_shapes, implicit_init


def _bn(s, p, c):
    s[p + ".weight"] = (c,); s[p + ".bias"] = (c,)
    s[p + ".running_mean"] = (c,); s[p + ".running_var"] = (c,); s[p + ".num_batches_tracked"] = ()


def _bottleneck_conv(s, p, c, k):
    s[p + ".linear1.weight"] = (c, c, k, k); _bn(s, p + ".bn1", c)
    s[p + ".linear2.weight"] = (c, c, k, k); _bn(s, p + ".bn2", c)


def dpt_param_shapes(pre="dpt_depth."):
    s = {}
    m = pre + "pretrained.model."
    s[m + "cls_token"] = (1, 1, 768); s[m + "pos_embed"] = (1, 577, 768)
    bb = m + "patch_embed.backbone."
    s[bb + "stem.conv.weight"] = (64, 3, 7, 7); s[bb + "stem.norm.weight"] = (64,); s[bb + "stem.norm.bias"] = (64,)
    cin = 64
    for st, (depth, out) in enumerate(zip((3, 4, 9), (256, 512, 1024))):
        mid = out // 4
        for b in range(depth):
            p = f"{bb}stages.{st}.blocks.{b}."
            if b == 0:
                s[p + "downsample.conv.weight"] = (out, cin, 1, 1)
                s[p + "downsample.norm.weight"] = (out,); s[p + "downsample.norm.bias"] = (out,)
            s[p + "conv1.weight"] = (mid, cin, 1, 1); s[p + "norm1.weight"] = (mid,); s[p + "norm1.bias"] = (mid,)
            s[p + "conv2.weight"] = (mid, mid, 3, 3); s[p + "norm2.weight"] = (mid,); s[p + "norm2.bias"] = (mid,)
            s[p + "conv3.weight"] = (out, mid, 1, 1); s[p + "norm3.weight"] = (out,); s[p + "norm3.bias"] = (out,)
            cin = out
    s[m + "patch_embed.proj.weight"] = (768, 1024, 1, 1); s[m + "patch_embed.proj.bias"] = (768,)
    for i in range(12):
        p = f"{m}blocks.{i}."
        for n in ("norm1", "norm2"):
            s[p + n + ".weight"] = (768,); s[p + n + ".bias"] = (768,)
        s[p + "attn.qkv.weight"] = (2304, 768); s[p + "attn.qkv.bias"] = (2304,)
        s[p + "attn.proj.weight"] = (768, 768); s[p + "attn.proj.bias"] = (768,)
        s[p + "mlp.fc1.weight"] = (3072, 768); s[p + "mlp.fc1.bias"] = (3072,)
        s[p + "mlp.fc2.weight"] = (768, 3072); s[p + "mlp.fc2.bias"] = (768,)
    s[m + "norm.weight"] = (768,); s[m + "norm.bias"] = (768,)
    s[m + "head.weight"] = (1000, 768); s[m + "head.bias"] = (1000,)
    for n in ("3", "4"):
        p = f"{pre}pretrained.act_postprocess{n}."
        s[p + "0.project.0.weight"] = (768, 1536); s[p + "0.project.0.bias"] = (768,)
        s[p + "3.weight"] = (768, 768, 1, 1); s[p + "3.bias"] = (768,)
    s[pre + "pretrained.act_postprocess4.4.weight"] = (768, 768, 3, 3); s[pre + "pretrained.act_postprocess4.4.bias"] = (768,)
    sc = pre + "scratch."
    for i, c in enumerate((256, 512, 768, 768)):
        s[f"{sc}layer{i + 1}_rn.weight"] = (256, c, 3, 3)
    for i in range(1, 5):
        p = f"{sc}refinenet{i}."
        s[p + "out_conv.weight"] = (256, 256, 1, 1); s[p + "out_conv.bias"] = (256,)
        for u in ("resConfUnit1", "resConfUnit2"):
            for c in ("conv1", "conv2"):
                s[f"{p}{u}.{c}.weight"] = (256, 256, 3, 3); s[f"{p}{u}.{c}.bias"] = (256,)
    s[sc + "output_conv.0.weight"] = (128, 256, 3, 3); s[sc + "output_conv.0.bias"] = (128,)
    s[sc + "output_conv.2.weight"] = (32, 128, 3, 3); s[sc + "output_conv.2.bias"] = (32,)
    s[sc + "output_conv.4.weight"] = (1, 32, 1, 1); s[sc + "output_conv.4.bias"] = (1,)
    return s


def resnet50_param_shapes(pre):
    s = {}
    s[pre + "conv1.weight"] = (64, 3, 7, 7); _bn(s, pre + "bn1", 64)
    cin = 64
    for li, (depth, width) in enumerate(((3, 64), (4, 128), (6, 256), (3, 512)), start=1):
        for b in range(depth):
            p = f"{pre}layer{li}.{b}."
            s[p + "conv1.weight"] = (width, cin, 1, 1); _bn(s, p + "bn1", width)
            s[p + "conv2.weight"] = (width, width, 3, 3); _bn(s, p + "bn2", width)
            s[p + "conv3.weight"] = (width * 4, width, 1, 1); _bn(s, p + "bn3", width * 4)
            if b == 0:
                s[p + "downsample.0.weight"] = (width * 4, cin, 1, 1); _bn(s, p + "downsample.1", width * 4)
            cin = width * 4
    return s


def coord_enc_res_param_shapes(pre="coord_encoder.", latent=256):
    s = resnet50_param_shapes(pre + "encoder.")
    _bottleneck_conv(s, pre + "encoder.fc.0", 2048, 1); _bottleneck_conv(s, pre + "encoder.fc.1", 2048, 1)
    s[pre + "encoder.fc.2.weight"] = (latent, 2048); s[pre + "encoder.fc.2.bias"] = (latent,)
    _bottleneck_conv(s, pre + "depth_feat_proj.0", 1024, 1); _bottleneck_conv(s, pre + "depth_feat_proj.1", 1024, 1)
    s[pre + "depth_feat_proj.2.weight"] = (latent, 1024, 1, 1); s[pre + "depth_feat_proj.2.bias"] = (latent,)
    return s


def graph_shape_param_shapes():
    s = {}
    _bottleneck_conv(s, "intr_head.0", 768, 3); _bottleneck_conv(s, "intr_head.1", 768, 3)
    s["intr_proj.weight"] = (3, 768); s["intr_proj.bias"] = (3,)
    s.update(dpt_param_shapes("dpt_depth."))
    s.update(coord_enc_res_param_shapes("coord_encoder."))
    s.update({"impl_network." + k: v for k, v in implicit_param_shapes().items()})
    return s


def seeded_state_dict(shapes, seed, implicit_prefix="impl_network."):
    """Deterministic 'realistic' weights: fan-in-scaled normals for matrices/filters, norms near (1, 0),
    BN running stats near (0, 1); Implicit part from implicit_init (reference init + re-centred field)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k in sorted(shapes):
        shp = shapes[k]
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_var"):
            sd[k] = 1.0 + 0.2 * torch.rand(shp, generator=g)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        elif len(shp) >= 2 and not k.endswith(("cls_token", "pos_embed")):
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[k] = torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_in) * (0.5 if "fc2" in k or "proj" in k else 1.0)
        elif k.endswith(("cls_token", "pos_embed")):
            sd[k] = 0.02 * torch.randn(shp, generator=g)
        elif k.endswith(".weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            sd[k] = 0.05 * torch.randn(shp, generator=g)
    # keep the synthetic model in a sane operating range: depth inside (0,1) (head bias like
    # dpt_depth.py:109 but larger), intrinsics near the focal prior (graph_shape.py:27-28 zero-inits them)
    for k in sd:
        if ".scratch.refinenet" in k and k.endswith(".weight"):
            sd[k] = sd[k] * 0.4
        elif k.endswith("scratch.output_conv.4.weight"):
            sd[k] = sd[k] * 0.05
        elif k.endswith("scratch.output_conv.4.bias"):
            sd[k] = sd[k] * 0 + 0.5
        elif k == "intr_proj.weight":
            sd[k] = sd[k] * 0.01
        elif k in ("coord_encoder.encoder.fc.2.weight", "coord_encoder.depth_feat_proj.2.weight"):
            sd[k] = sd[k] * 0.02          # latents of O(1), like a trained encoder feeding latent_proj
    if implicit_prefix:
        for k, v in implicit_init(seed).items():
            sd[implicit_prefix + k] = v
    return sd
