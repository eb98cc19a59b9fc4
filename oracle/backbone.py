"""Oracle (TEST INFRASTRUCTURE): encoder side of the hot path, functional over a state_dict with the
reference's parameter names.

  dpt_depth_forward   <- DPTDepthModel.forward / DPT.forward     (model/depth/dpt_depth.py:68-94,115-123)
  hybrid_vit_forward  <- forward_flex + hooks                     (model/depth/vit.py:101-154,157-164)
                         over timm==0.6.12 `vit_base_resnet50_384` (third-party, NOT in /root/reference:
                         restated from the published architecture -- ResNetV2 (3,4,9) non-preact
                         bottlenecks, weight-standardised SAME convs (eps 1e-8), GroupNorm(32, eps 1e-5),
                         1x1 proj, ViT-B/12 heads, LN eps 1e-6).  Cross-checked against
                         the independent `transformers` DPT-hybrid port in tests/test_oracle_backbone.py.
  reassemble          <- forward_vit / ProjectReadout / act_postprocess{3,4} (vit.py:32-43,57-98,430-461)
  fusion / head       <- FeatureFusionBlock_custom, ResidualConvUnit_custom (blocks.py:264-342),
                         head (dpt_depth.py:100-108)
  bottleneck_conv     <- Bottleneck_Conv                           (utils/layers.py:76-100)
  coord_enc_res       <- CoordEncRes.forward                       (model/shape/seen_coord_enc.py:180-194)
                         over torchvision resnet50 (eval-mode BatchNorm)
  intr_param2mtx, unproj_depth, valid_norm_fac, interpolate_coordmap, graph_shape_forward
                      <- model/compute_graph/graph_shape.py:89-148, utils/camera.py:52-108, utils/util.py:336-345
All eval-mode (BatchNorm running statistics, no DropPath).
"""
import math

import torch
import torch.nn.functional as F

GN_GROUPS, GN_EPS, WS_EPS, VIT_LN_EPS, BN_EPS = 32, 1e-5, 1e-8, 1e-6, 1e-5
STAGE_DEPTHS = (3, 4, 9)
STAGE_CH = (256, 512, 1024)


# ---- timm ResNetV2 pieces ------------------------------------------------------------------------
def same_pad(i, k, s):
    """TF 'SAME' total padding for one dim (timm padding.get_same_padding)."""
    return max((math.ceil(i / s) - 1) * s + (k - 1) + 1 - i, 0)


def ws_conv_same(x, w, stride):
    """timm StdConv2dSame: dynamic SAME zero pad, weight standardisation over (Cin,kh,kw), no bias."""
    ph, pw = same_pad(x.shape[2], w.shape[2], stride), same_pad(x.shape[3], w.shape[3], stride)
    x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    wf = w.reshape(w.shape[0], -1)
    var, mean = torch.var_mean(wf, dim=1, keepdim=True, unbiased=False)
    ws = ((wf - mean) / torch.sqrt(var + WS_EPS)).reshape(w.shape)
    return F.conv2d(x, ws, None, stride)


def _gn(x, sd, pre, relu):
    y = F.group_norm(x, GN_GROUPS, sd[pre + ".weight"], sd[pre + ".bias"], GN_EPS)
    return F.relu(y) if relu else y


def resnetv2_stem_stages(sd, x, pre):
    """-> (stage0 out [B,256,H/4,W/4], stage1 out [B,512,H/8,W/8], stage2 out [B,1024,H/16,W/16])"""
    x = _gn(ws_conv_same(x, sd[pre + "stem.conv.weight"], 2), sd, pre + "stem.norm", True)
    ph, pw = same_pad(x.shape[2], 3, 2), same_pad(x.shape[3], 3, 2)
    x = F.max_pool2d(F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float("-inf")), 3, 2)
    outs = []
    for s, depth in enumerate(STAGE_DEPTHS):
        for b in range(depth):
            p = f"{pre}stages.{s}.blocks.{b}."
            stride = 2 if (b == 0 and s > 0) else 1
            short = x
            if b == 0:
                short = _gn(ws_conv_same(x, sd[p + "downsample.conv.weight"], stride), sd, p + "downsample.norm", False)
            y = _gn(ws_conv_same(x, sd[p + "conv1.weight"], 1), sd, p + "norm1", True)
            y = _gn(ws_conv_same(y, sd[p + "conv2.weight"], stride), sd, p + "norm2", True)
            y = _gn(ws_conv_same(y, sd[p + "conv3.weight"], 1), sd, p + "norm3", False)
            x = F.relu(y + short)
        outs.append(x)
    return outs


def vit_block(sd, x, pre, heads):
    B, T, C = x.shape
    hd = C // heads
    h = F.layer_norm(x, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], VIT_LN_EPS)
    qkv = F.linear(h, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"]).reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1)
    h = (a @ v).transpose(1, 2).reshape(B, T, C)
    x = x + F.linear(h, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    h = F.layer_norm(x, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], VIT_LN_EPS)
    h = F.linear(F.gelu(F.linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])),
                 sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return x + h


def resize_pos_embed(posemb, gh, gw):
    """model/depth/vit.py:101-115 (start_index 1)."""
    tok, grid = posemb[:, :1], posemb[0, 1:]
    g_old = int(math.sqrt(len(grid)))
    grid = grid.reshape(1, g_old, g_old, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(gh, gw), mode="bilinear", align_corners=False)
    return torch.cat([tok, grid.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)], dim=1)


def hybrid_vit_forward(sd, x, pre="pretrained.model.", heads=12, hooks=(8, 11)):
    """-> dict with the four hooked activations '1','2' (ResNet stages 0,1), '3','4' (ViT blocks 8, 11)."""
    B, _, H, W = x.shape
    pos = resize_pos_embed(sd[pre + "pos_embed"], H // 16, W // 16)
    s0, s1, s2 = resnetv2_stem_stages(sd, x, pre + "patch_embed.backbone.")
    t = F.conv2d(s2, sd[pre + "patch_embed.proj.weight"], sd[pre + "patch_embed.proj.bias"]).flatten(2).transpose(1, 2)
    t = torch.cat([sd[pre + "cls_token"].expand(B, -1, -1), t], dim=1) + pos
    acts = {"1": s0, "2": s1}
    n_blocks = sum(1 for k in sd if k.startswith(pre + "blocks.") and k.endswith(".norm1.weight"))
    for i in range(n_blocks):
        t = vit_block(sd, t, f"{pre}blocks.{i}.", heads)
        if i == hooks[0]:
            acts["3"] = t
        if i == hooks[1]:
            acts["4"] = t
    return acts


def _readout_project(sd, t, pre):
    """ProjectReadout (vit.py:32-43) + Transpose + dynamic Unflatten -> [B,768,h,w] (h*w = T-1)."""
    cls = t[:, 0].unsqueeze(1).expand_as(t[:, 1:])
    y = F.gelu(F.linear(torch.cat([t[:, 1:], cls], -1), sd[pre + "0.project.0.weight"], sd[pre + "0.project.0.bias"]))
    return y.transpose(1, 2)


def rcu(sd, x, pre):
    """ResidualConvUnit_custom (blocks.py:264-287), bn=False, activation ReLU (not in place)."""
    y = F.conv2d(F.relu(x), sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    y = F.conv2d(F.relu(y), sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    return y + x


def fusion(sd, pre, x, skip=None):
    """FeatureFusionBlock_custom.forward (blocks.py:321-342), align_corners=True."""
    if skip is not None:
        x = x + rcu(sd, skip, pre + "resConfUnit1.")
    x = rcu(sd, x, pre + "resConfUnit2.")
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    return F.conv2d(x, sd[pre + "out_conv.weight"], sd[pre + "out_conv.bias"])


def dpt_depth_forward(sd, image, pre="", get_feat=True):
    """image [B,3,H,W] in [0,1] -> (depth [B,1,H,W] clamped to [0,1], layer_4 feature [B,768,H/32,W/32])."""
    x = image * 2 - 1                                                           # dpt_depth.py:116
    B, _, H, W = x.shape
    acts = hybrid_vit_forward(sd, x, pre + "pretrained.model.")
    gh, gw = H // 16, W // 16
    l1, l2 = acts["1"], acts["2"]
    p3, p4 = pre + "pretrained.act_postprocess3.", pre + "pretrained.act_postprocess4."
    l3 = _readout_project(sd, acts["3"], p3).reshape(B, -1, gh, gw)
    l3 = F.conv2d(l3, sd[p3 + "3.weight"], sd[p3 + "3.bias"])
    l4 = _readout_project(sd, acts["4"], p4).reshape(B, -1, gh, gw)
    l4 = F.conv2d(l4, sd[p4 + "3.weight"], sd[p4 + "3.bias"])
    l4 = F.conv2d(l4, sd[p4 + "4.weight"], sd[p4 + "4.bias"], stride=2, padding=1)
    sc = pre + "scratch."
    r = [F.conv2d(t, sd[f"{sc}layer{i + 1}_rn.weight"], None, padding=1) for i, t in enumerate((l1, l2, l3, l4))]
    path = fusion(sd, sc + "refinenet4.", r[3])
    path = fusion(sd, sc + "refinenet3.", path, r[2])
    path = fusion(sd, sc + "refinenet2.", path, r[1])
    path = fusion(sd, sc + "refinenet1.", path, r[0])
    oc = sc + "output_conv."
    y = F.conv2d(path, sd[oc + "0.weight"], sd[oc + "0.bias"], padding=1)
    y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True)
    y = F.relu(F.conv2d(y, sd[oc + "2.weight"], sd[oc + "2.bias"], padding=1))
    y = F.relu(F.conv2d(y, sd[oc + "4.weight"], sd[oc + "4.bias"]))
    y = y.clamp(min=0, max=1)
    return (y, l4) if get_feat else y


# ---- Bottleneck_Conv / torchvision ResNet-50 / CoordEncRes ------------------------------------------
BN_TRAINING = False   # tests of the training path flip this: nn.BatchNorm2d in train mode normalises with batch statistics


def _bn(x, sd, pre):
    if BN_TRAINING:
        return F.batch_norm(x, None, None, sd[pre + ".weight"], sd[pre + ".bias"], True, 0.0, BN_EPS)
    return F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"], sd[pre + ".weight"], sd[pre + ".bias"],
                        False, 0.0, BN_EPS)


def bottleneck_conv(sd, x, pre, k):
    """utils/layers.py:76-100."""
    squeeze = x.dim() == 2
    if squeeze:
        x = x.unsqueeze(-1).unsqueeze(-1)
    y = F.relu(_bn(F.conv2d(x, sd[pre + "linear1.weight"], None, padding=k // 2), sd, pre + "bn1"))
    y = _bn(F.conv2d(y, sd[pre + "linear2.weight"], None, padding=k // 2), sd, pre + "bn2")
    y = F.relu(y + x)
    return y.squeeze(-1).squeeze(-1) if squeeze else y


def resnet50_features(sd, x, pre):
    """torchvision resnet50 trunk -> (layer3 out [B,1024,H/16,W/16], pooled layer4 [B,2048])."""
    x = F.relu(_bn(F.conv2d(x, sd[pre + "conv1.weight"], None, 2, 3), sd, pre + "bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = {}
    for li, (depth, stride) in enumerate(((3, 1), (4, 2), (6, 2), (3, 2)), start=1):
        for b in range(depth):
            p = f"{pre}layer{li}.{b}."
            s = stride if b == 0 else 1
            idt = x
            y = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"]), sd, p + "bn1"))
            y = F.relu(_bn(F.conv2d(y, sd[p + "conv2.weight"], None, s, 1), sd, p + "bn2"))
            y = _bn(F.conv2d(y, sd[p + "conv3.weight"]), sd, p + "bn3")
            if (p + "downsample.0.weight") in sd:
                idt = _bn(F.conv2d(x, sd[p + "downsample.0.weight"], None, s), sd, p + "downsample.1")
            x = F.relu(y + idt)
        feats[li] = x
    return feats[3], torch.flatten(F.adaptive_avg_pool2d(feats[4], 1), 1)


def coord_enc_res(sd, coord, mask, pre="coord_encoder."):
    """CoordEncRes.forward (seen_coord_enc.py:180-194), win_size 16."""
    B = coord.shape[0]
    l3, pooled = resnet50_features(sd, coord * mask.float(), pre + "encoder.")
    g = bottleneck_conv(sd, pooled, pre + "encoder.fc.0.", 1)
    g = bottleneck_conv(sd, g, pre + "encoder.fc.1.", 1)
    g = F.linear(g, sd[pre + "encoder.fc.2.weight"], sd[pre + "encoder.fc.2.bias"]).unsqueeze(1)
    y = bottleneck_conv(sd, l3, pre + "depth_feat_proj.0.", 1)
    y = bottleneck_conv(sd, y, pre + "depth_feat_proj.1.", 1)
    y = F.conv2d(y, sd[pre + "depth_feat_proj.2.weight"], sd[pre + "depth_feat_proj.2.bias"])
    return torch.cat([g, y.view(B, g.shape[-1], -1).permute(0, 2, 1)], dim=1)


# ---- geometry glue ----------------------------------------------------------------------------------
def intr_param2mtx(params, H, W):
    """graph_shape.py:89-113."""
    B = len(params)
    f = 1.3875
    K = torch.zeros(3, 3, device=params.device).float().unsqueeze(0).repeat(B, 1, 1)
    K[:, 2, 2] += 1
    sf = torch.pow(4., torch.tanh(params[:, 0]))
    K[:, 0, 0] += f * W * sf
    K[:, 1, 1] += f * H * sf
    K[:, 0, 2] += W / 2 + torch.tanh(params[:, 1]) * W / 2
    K[:, 1, 2] += H / 2 + torch.tanh(params[:, 2]) * H / 2
    return K


def unproj_depth(depth, intr):
    """utils/camera.py:88-108."""
    B, _, H, W = depth.shape
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=depth.device), torch.arange(W, dtype=torch.float32, device=depth.device),
                            indexing="ij")
    grid = torch.stack([xx, yy, torch.ones_like(yy)], dim=-1).view(-1, 3).unsqueeze(0).repeat(B, 1, 1)
    rays = torch.linalg.inv(intr).float() @ grid.permute(0, 2, 1)
    return rays.permute(0, 2, 1) * depth.view(B, H * W, 1)


def valid_norm_fac(pts, mask):
    """utils/camera.py:52-78."""
    B = pts.shape[0]
    mask = mask.view(B, pts.shape[1])
    means, dists = [], []
    for b in range(B):
        v = pts[b][mask[b]]
        mu = v.mean(dim=0)
        means.append(mu)
        dists.append((v - mu).norm(dim=1).max())
    return torch.stack(means), torch.stack(dists)


def interpolate_coordmap(coord_map, mask_input, size, bg_coord=0):
    """utils/util.py:336-345."""
    mask = (mask_input > 0.5).float()
    cv = F.interpolate(coord_map * mask, size, mode="bilinear", align_corners=False)
    mask = F.interpolate(mask, size, mode="bilinear", align_corners=False)
    out = cv / (mask + 1.e-6)
    mb = (mask > 0.5).float()
    return out * mb + bg_coord * (1 - mb), mb


def graph_shape_encode(sd, rgb, mask_map, H=224, W=224):
    """Graph.forward up to var.latent_depth (graph_shape.py:115-148), resnet encoder, dsp=1.
    -> dict(depth_pred, intr_pred, seen_points, latent_depth, validity_mask)."""
    B = rgb.shape[0]
    depth, feat = dpt_depth_forward(sd, rgb, "dpt_depth.")
    f = bottleneck_conv(sd, feat, "intr_head.0.", 3)
    f = bottleneck_conv(sd, f, "intr_head.1.", 3)
    params = F.linear(F.adaptive_avg_pool2d(f, 1).squeeze(-1).squeeze(-1), sd["intr_proj.weight"], sd["intr_proj.bias"])
    K = intr_param2mtx(params, H, W)
    pts = unproj_depth(depth, K)
    mean, scale = valid_norm_fac(pts, mask_map > 0.5)
    seen = (pts - mean.unsqueeze(1)) / scale.unsqueeze(-1).unsqueeze(-1)
    seen[(mask_map <= 0.5).view(B, -1)] = 0
    seen_map = seen.view(B, H, W, 3).permute(0, 3, 1, 2).contiguous()
    dsp, mask_dsp = interpolate_coordmap(seen_map, mask_map, (H, W))
    latent = coord_enc_res(sd, dsp, mask_dsp)
    return dict(depth_pred=depth, intr_pred=K, seen_points=seen, latent_depth=latent,
                validity_mask=(mask_map > 0.5).float().view(B, -1), intr_feat=feat)


# ---- transformer seen-surface encoder (model/shape/seen_coord_enc.py:13-139; SURVEY.md section 8a row a7') ----------------
def coord_emb_forward(sd, coord, mask, pre, heads, ws):
    """CoordEmb.forward :49-78.  coord [B,H,W,3], mask bool [B,H,W] -> [B, (H/ws)*(W/ws), C]."""
    emb = F.linear(coord, sd[pre + "pos_embed.weight"], sd[pre + "pos_embed.bias"])
    emb = torch.where(mask.unsqueeze(-1), emb, sd[pre + "invalid_coord_token"].expand_as(emb))
    B, H, W, C = emb.shape
    emb = emb.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws * ws, C)
    pos = sd[pre + "two_d_pos_embed"]
    emb = emb + pos[:, 1:, :]
    cls = (sd[pre + "cls_token"] + pos[:, :1, :]).expand(emb.shape[0], -1, -1)
    emb = vit_block(sd, torch.cat((cls, emb), dim=1), pre + "blocks.0.", heads)
    return emb[:, 0].view(B, (H // ws) * (W // ws), -1)


def coord_enc_att_forward(sd, coord, mask, pre="coord_encoder.", heads=8, ws=8):
    """CoordEncAtt.forward :119-139 (eval mode) -> [B, 1 + (H/ws)*(W/ws), C]."""
    x = coord_emb_forward(sd, coord, mask, pre + "coord_embed.", heads, ws)
    x = torch.cat((sd[pre + "cls_token"].expand(x.shape[0], -1, -1), x), dim=1)
    n_blocks = 1 + max(int(k[len(pre + "blocks."):].split(".")[0]) for k in sd if k.startswith(pre + "blocks."))
    for i in range(n_blocks):
        x = vit_block(sd, x, f"{pre}blocks.{i}.", heads)
    C = x.shape[-1]
    return F.layer_norm(x, (C,), sd[pre + "norm.weight"], sd[pre + "norm.bias"], VIT_LN_EPS)
