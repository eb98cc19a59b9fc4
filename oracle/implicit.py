"""Oracle (TEST INFRASTRUCTURE): implicit occupancy decoder, functional over a state_dict.

Restates /root/reference/model/shape/implicit.py:
  * ``implicit_forward``        <- Implicit.forward            (implicit.py:251-288)
  * ``_attn_block``             <- ImplFuncBlock.forward       (implicit.py:99-109)
                                   ImplFuncAttention.forward   (implicit.py:25-79)
  * ``_occupancy_mlp``          <- MLPBlocks.forward           (implicit.py:168-184)
  * ``sincos_pos_embed_2d``     <- utils/pos_embed.py:21-68
Configuration is the shipped one (options/shape.yaml:19-44): C=256, 8 heads, 2 attention blocks
(the 2nd is "last_layer"), pos-embed added before block 0 only, 8 hidden MLP layers with skips at
2,4,6, no NeRF posenc.  All math is fp32 PyTorch on whatever device the tensors live on.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-6          # implicit.py:191 (partial(nn.LayerNorm, eps=1e-6))
SOFTPLUS_BETA = 100.0  # implicit.py:166
SKIP_IN = (2, 4, 6)    # options/shape.yaml:44


def sincos_pos_embed_2d(dim, grid, cls_token=True):
    """Fixed 2D sin-cos table [1(+cls)+grid*grid, dim] (utils/pos_embed.py:21-68), float64 numpy."""
    def one_axis(d, pos):
        omega = np.arange(d // 2, dtype=np.float32)
        omega /= d / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)
    gh = np.arange(grid, dtype=np.float32)
    gw = np.arange(grid, dtype=np.float32)
    mesh = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid, grid)
    emb = np.concatenate([one_axis(dim // 2, mesh[0]), one_axis(dim // 2, mesh[1])], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, dim]), emb], axis=0)
    return emb


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], LN_EPS)


def _lin(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _attention(x, n_points, sd, pre, heads, last):
    """ImplFuncAttention.forward (implicit.py:25-79). x: [B, L+P, C] (already normed)."""
    B, N, C = x.shape
    L = N - n_points
    hd = C // heads
    scale = hd ** -0.5
    qkv = _lin(x, sd, pre + ".qkv").reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    ql, kl, vl = q[:, :, :L], k[:, :, :L], v[:, :, :L]
    qp, kp, vp = q[:, :, L:], k[:, :, L:], v[:, :, L:]
    # each query point sees the L latents and itself (implicit.py:38-46)
    s_cross = (qp @ kl.transpose(-2, -1)) * scale
    s_self = (qp * kp).sum(-1, keepdim=True) * scale
    a = torch.cat([s_cross, s_self], dim=-1).softmax(dim=-1)
    o_cross = (a[..., :L] @ vl).transpose(1, 2).reshape(B, n_points, C)
    o_self = (a[..., L:] * vp).transpose(1, 2).reshape(B, n_points, C)
    out_p = o_cross + o_self
    attn_vis = a[..., :-1].mean(dim=1)
    if last:
        return _lin(out_p, sd, pre + ".proj"), attn_vis
    # latents only see latents (implicit.py:65-71)
    al = ((ql @ kl.transpose(-2, -1)) * scale).softmax(dim=-1)
    out_l = (al @ vl).transpose(1, 2).reshape(B, L, C)
    return _lin(torch.cat([out_l, out_p], dim=1), sd, pre + ".proj"), attn_vis


def _attn_block(x, n_points, sd, pre, heads, last):
    """ImplFuncBlock.forward (implicit.py:99-109), eval mode (DropPath = identity)."""
    a, vis = _attention(_ln(x, sd, pre + ".norm1"), n_points, sd, pre + ".attn", heads, last)
    x = (x[:, -n_points:] if last else x) + a
    h = _lin(F.gelu(_lin(_ln(x, sd, pre + ".norm2"), sd, pre + ".mlp.fc1")), sd, pre + ".mlp.fc2")
    return x + h, vis


def _occupancy_mlp(points, feat, sd, pre):
    """MLPBlocks.forward (implicit.py:168-184)."""
    n_layers = len([k for k in sd if k.startswith(pre + ".layers.") and k.endswith(".weight")])
    inputs = torch.cat([points, feat], dim=-1)
    x = inputs
    for l in range(n_layers):
        if l in SKIP_IN:
            x = torch.cat([x, inputs], -1) / np.sqrt(2)   # numpy float64 scalar, result stays fp32
        x = _lin(x, sd, f"{pre}.layers.{l}")
        if l < n_layers - 1:
            x = F.softplus(x, beta=SOFTPLUS_BETA)
    return x


def implicit_forward(sd, latent_depth, points_3D, prefix="", heads=8):
    """Implicit.forward (implicit.py:251-288) -> (logits [B,P], attn [B,P,L])."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd
    latent = _lin(latent_depth, sd, "latent_proj")
    L = latent.shape[1]
    pfeat = _lin(points_3D, sd, "point_proj.proj")
    P = pfeat.shape[1]
    x = torch.cat([latent, pfeat], dim=1)
    n_blocks = len({k.split(".")[1] for k in sd if k.startswith("blocks_attn.")})
    vis = []
    for l in range(n_blocks):
        if l == 0:  # posenc_perlayer: false (options/shape.yaml:39)
            x = torch.cat([x[:, :L] + sd["pos_embed"], x[:, L:]], dim=1)
        x, a = _attn_block(x, P, sd, f"blocks_attn.{l}", heads, last=(l == n_blocks - 1))
        vis.append(a)
    x = _ln(x, sd, "norm")
    attn = torch.stack(vis, dim=-1).mean(dim=-1)
    return _occupancy_mlp(points_3D, x, sd, "impl_mlp").squeeze(-1), attn


def implicit_param_shapes(C=256, latent_dim=256, L=197, hidden=8, mlp_ratio=4, n_blocks=2):
    """Name -> shape of the Implicit state_dict (implicit.py:186-249)."""
    s = {"pos_embed": (1, L, C),
         "point_proj.proj.weight": (C, 3), "point_proj.proj.bias": (C,),
         "latent_proj.weight": (C, latent_dim), "latent_proj.bias": (C,)}
    for b in range(n_blocks):
        p = f"blocks_attn.{b}"
        s.update({f"{p}.norm1.weight": (C,), f"{p}.norm1.bias": (C,),
                  f"{p}.attn.qkv.weight": (3 * C, C), f"{p}.attn.qkv.bias": (3 * C,),
                  f"{p}.attn.proj.weight": (C, C), f"{p}.attn.proj.bias": (C,),
                  f"{p}.norm2.weight": (C,), f"{p}.norm2.bias": (C,),
                  f"{p}.mlp.fc1.weight": (int(C * mlp_ratio), C), f"{p}.mlp.fc1.bias": (int(C * mlp_ratio),),
                  f"{p}.mlp.fc2.weight": (C, int(C * mlp_ratio)), f"{p}.mlp.fc2.bias": (C,)})
    s.update({"norm.weight": (C,), "norm.bias": (C,)})
    dims = [3 + C] + [C] * hidden + [1]
    for l in range(len(dims) - 1):
        din = dims[l] + (dims[0] if l in SKIP_IN else 0)
        s[f"impl_mlp.layers.{l}.weight"] = (dims[l + 1], din)
        s[f"impl_mlp.layers.{l}.bias"] = (dims[l + 1],)
    return s


def implicit_init(seed=0, recentre=True, **kw):
    """Reference init scheme (implicit.py:235-249): xavier-uniform Linear, zero bias, LN 1/0,
    fixed sincos pos_embed.  With ``recentre`` the last bias is shifted so the synthetic field has
    both signs (random init gives all-positive logits)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shp in implicit_param_shapes(**kw).items():
        if name == "pos_embed":
            grid = int(math.isqrt(shp[1] - 1))
            sd[name] = torch.from_numpy(sincos_pos_embed_2d(shp[2], grid)).float().unsqueeze(0)
        elif name.endswith(".bias"):
            sd[name] = torch.zeros(shp)
        elif "norm" in name:
            sd[name] = torch.ones(shp)
        else:
            bound = math.sqrt(6.0 / (shp[0] + shp[1]))
            sd[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    if recentre:
        lat = torch.randn(1, sd["pos_embed"].shape[1], sd["latent_proj.weight"].shape[1], generator=g)
        pts = (torch.rand(1, 4096, 3, generator=g) * 2 - 1) * 1.5
        with torch.no_grad():
            lg, _ = implicit_forward(sd, lat, pts)
        last = max(int(k.split(".")[2]) for k in sd if k.startswith("impl_mlp.layers."))
        sd[f"impl_mlp.layers.{last}.bias"] = sd[f"impl_mlp.layers.{last}.bias"] - lg.median()
    return sd
