"""CPU stand-ins for the C-ABI wrappers of zeroshape_b200.ops (TEST INFRASTRUCTURE): same signatures and
layout conventions (NHWC, OHWI filters), implemented with PyTorch CPU ops.  They let the `-m "not gpu"`
suite execute the product's HOST orchestration (layer order, packing, strides, residual placement) against
the oracle without a GPU.  Never used by the product."""
import math

import torch
import torch.nn.functional as F

ACTS = {0: lambda x: x, 1: F.relu, 2: F.gelu, 3: lambda x: F.softplus(x, beta=100), 4: torch.sigmoid,
        5: lambda x: x.clamp(0, 1)}


def _epi(y, bias, res, res_mode, act):
    if bias is not None:
        y = y + bias
    if res is not None and res_mode == 0:
        res_mode = 2
    if res is not None and res_mode == 1:
        y = y + res
    y = ACTS[act](y)
    if res is not None and res_mode == 2:
        y = y + res
    return y


def gemm(a, w, bias=None, res=None, res_mode=0, act=0, out=None):
    y = _epi(a @ w.T, bias, res, res_mode, act)
    if out is not None:
        out.copy_(y)
        return out
    return y.contiguous()


def linear(x, w, bias=None, act=0, res=None, res_mode=0, tc=None):
    return _epi(F.linear(x, w), bias, res, res_mode, act).contiguous()


class PackedWeight:
    def __init__(self, w):
        self.src = w


def gemm_tc(a, pw, bias=None, res=None, res_mode=0, act=0, out=None, precision="bf16x3"):
    return gemm(a, pw.src, bias, res, res_mode, act, out)


def conv2d_nhwc(x, w, bias=None, stride=1, pad=(0, 0, 0, 0), act=0, res=None, res_mode=0, pre_relu=False, tc=None, precision=None):
    xn = x.permute(0, 3, 1, 2)
    if pre_relu:
        xn = F.relu(xn)
    xn = F.pad(xn, (pad[2], pad[3], pad[0], pad[1]))
    y = F.conv2d(xn, w.permute(0, 3, 1, 2), None, stride).permute(0, 2, 3, 1)
    return _epi(y, bias, res, res_mode, act).contiguous()


def layernorm(x, g, b, eps):
    return F.layer_norm(x, (x.shape[-1],), g, b, eps)


def groupnorm_nhwc(x, g, b, groups, eps, relu, res=None):
    y = F.group_norm(x.permute(0, 3, 1, 2), groups, g, b, eps).permute(0, 2, 3, 1)
    if res is not None:
        y = y + res
    return (F.relu(y) if relu else y).contiguous()


def channel_affine(x, scale, shift, act=0, res=None):
    y = x * scale + shift
    if res is not None:
        y = y + res
    return ACTS[act](y).contiguous()


def axpby(a, alpha=1.0, b=None, beta=1.0, act=0):
    y = a * alpha
    if b is not None:
        y = y + b * beta
    return ACTS[act](y).contiguous()


def maxpool3x3s2_nhwc(x, pt, pl, OH, OW):
    H, W = x.shape[1], x.shape[2]
    pb, pr = (OH - 1) * 2 + 3 - H - pt, (OW - 1) * 2 + 3 - W - pl
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, max(pr, 0), pt, max(pb, 0)), value=float("-inf"))
    return F.max_pool2d(xn, 3, 2)[:, :, :OH, :OW].permute(0, 2, 3, 1).contiguous()


def avgpool_nhwc(x):
    return x.mean(dim=(1, 2))


def bilinear_nhwc(x, OH, OW, align):
    return F.interpolate(x.permute(0, 3, 1, 2), size=(OH, OW), mode="bilinear", align_corners=bool(align)).permute(0, 2, 3, 1).contiguous()


def nchw_to_nhwc(x, scale=1.0, shift=0.0):
    return (x * scale + shift).permute(0, 2, 3, 1).contiguous()


def nhwc_to_nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def mha(qkv, heads, tc=None, precision="fp16x3"):
    B, T, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    q, k, v = qkv.reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4).unbind(0)
    return (((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1) @ v).transpose(1, 2).reshape(B, T, C).contiguous()


def point_attention(qkv_p, k_lat, v_lat, heads, attn=None, attn_scale=1.0, attn_accumulate=False, tc=None):
    B, P, C3 = qkv_p.shape
    C = C3 // 3
    hd = C // heads
    q, k, v = qkv_p.reshape(B, P, 3, heads, hd).permute(2, 0, 3, 1, 4).unbind(0)
    L = k_lat.shape[1]
    kl = k_lat.reshape(B, L, heads, hd).permute(0, 2, 1, 3)
    vl = v_lat.reshape(B, L, heads, hd).permute(0, 2, 1, 3)
    s = torch.cat([(q @ kl.transpose(-2, -1)), (q * k).sum(-1, keepdim=True)], -1) * hd ** -0.5
    a = s.softmax(-1)
    out = (a[..., :L] @ vl + a[..., L:] * v).transpose(1, 2).reshape(B, P, C)
    if attn is not None:
        vis = a[..., :L].mean(1) * attn_scale
        if attn_accumulate:
            attn += vis
        else:
            attn.copy_(vis)
    return out


def dense_grid(n, rmin, rmax, x0, x1, device):
    g = torch.linspace(rmin, rmax, n)
    return torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1)[x0:x1].contiguous()


def concat2(a, b, s=1.0):
    return torch.cat([a, b.expand(a.shape[0], -1) if b.shape[0] != a.shape[0] else b], -1) / s


def intr_param2mtx(params, H, W):
    from oracle.backbone import intr_param2mtx as ref
    return ref(params, H, W)


def unproject(depth, K):
    from oracle.backbone import unproj_depth
    return unproj_depth(depth, K)


def unproject_normalize(depth, mask, K):
    from oracle.backbone import unproj_depth, valid_norm_fac
    B = depth.shape[0]
    pts = unproj_depth(depth, K)
    mean, scale = valid_norm_fac(pts, mask > 0.5)
    seen = (pts - mean.unsqueeze(1)) / scale.view(B, 1, 1)
    seen[(mask <= 0.5).view(B, -1)] = 0
    return seen, mean, scale


def device_cc():
    return 0


# ---- evaluation ops (marching cubes, surface sampling, Chamfer, statistics): the oracle's numpy / C restatements ----------------
def marching_cubes(vol, iso, x_offset=0):
    from oracle import eval3d as E
    v, f = E.marching_cubes(vol.detach().cpu().numpy(), float(iso))
    v = v.copy()
    v[:, 0] += x_offset
    return torch.from_numpy(v).float(), torch.from_numpy(f.astype("int32"))


def marching_cubes_count(vol, iso):
    v, f = marching_cubes(vol, iso)
    return (v, f), torch.tensor([v.shape[0], f.shape[0]], dtype=torch.int32)


def marching_cubes_emit(vol, iso, ws, V, F, x_offset=0):
    v, f = ws
    v = v.clone()
    v[:, 0] += x_offset
    return v, f


class MeshFuture:
    def __init__(self, vol, iso, x_offset=0):
        self.vol, self.iso, self.x_offset = vol, iso, x_offset

    def result(self):
        return marching_cubes(self.vol, self.iso, self.x_offset)


def mesh_sample(verts, faces, num, vscale=1.0, voffset=0.0, seed=0):
    import numpy as np
    from oracle import eval3d as E
    if faces.shape[0] == 0:
        return torch.zeros(num, 3)
    v = verts.double().numpy() * vscale + voffset
    return torch.from_numpy(E.sample_surface(v, faces.numpy().astype(np.int64), num, np.random.RandomState(seed))).float()


def chamfer_nn(xyz1, xyz2):
    from oracle import eval3d as E
    d1, d2, i1, i2 = E.chamfer_nn(xyz1.numpy(), xyz2.numpy())
    return torch.from_numpy(d1), torch.from_numpy(d2), torch.from_numpy(i1), torch.from_numpy(i2)


def chamfer_stats(sq1, sq2, thresholds, squared=True):
    d1, d2 = (sq1.sqrt(), sq2.sqrt()) if squared else (sq1, sq2)
    th = torch.tensor(list(thresholds), dtype=torch.float32)
    p = (d1.unsqueeze(-1) < th).float().mean(dim=1)
    r = (d2.unsqueeze(-1) < th).float().mean(dim=1)
    return d1.mean(dim=1), d2.mean(dim=1), p, r


def mean_axis1(x):
    return x.mean(dim=1)


EVAL_OPS = ("marching_cubes", "MeshFuture", "marching_cubes_count", "marching_cubes_emit", "mesh_sample", "chamfer_nn", "chamfer_stats", "mean_axis1")


def install_eval(monkeypatch):
    import zeroshape_b200.ops as ops
    g = globals()
    for name in EVAL_OPS:
        monkeypatch.setattr(ops, name, g[name])


def install(monkeypatch):
    """Patch zeroshape_b200.ops in place (pytest monkeypatch restores it)."""
    import zeroshape_b200.ops as ops
    g = globals()
    for name in ("gemm", "linear", "PackedWeight", "gemm_tc", "conv2d_nhwc", "layernorm", "groupnorm_nhwc", "channel_affine",
                 "axpby", "maxpool3x3s2_nhwc", "avgpool_nhwc", "bilinear_nhwc", "nchw_to_nhwc", "nhwc_to_nchw", "mha",
                 "point_attention", "dense_grid", "concat2", "intr_param2mtx", "unproject", "unproject_normalize", "device_cc"):
        monkeypatch.setattr(ops, name, g[name])


# ---- per-op stand-ins for the TRAINING kernels (tapes of model/shape/implicit_train.py, seen_coord_att_train.py, dpt_train.py):
# forward ops as above, backward ops through torch autograd of the matching forward stand-in ------------------------------------
def _via_autograd(fn, x, dy):
    xx = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        fn(xx).backward(dy)
    return xx.grad


def act_bwd(dy, z, act):
    return _via_autograd(ACTS[act], z, dy)


def layernorm_bwd(dy, x, gamma, eps, dgamma=None, dbeta=None):
    gg, bb = gamma.detach().clone().requires_grad_(True), torch.zeros_like(gamma).requires_grad_(True)
    xx = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        F.layer_norm(xx, (x.shape[-1],), gg, bb, eps).backward(dy)
    if dgamma is not None:
        dgamma += gg.grad
        dbeta += bb.grad
    return xx.grad


layernorm_bwd_generic = layernorm_bwd


def mha_bwd(qkv, dout, heads, tc=None):
    return _via_autograd(lambda t: mha(t, heads), qkv, dout)


def point_attention_bwd(qkv_p, k_lat, v_lat, out, dout, heads, tc=None):
    qq, kk, vv = (t.detach().clone().requires_grad_(True) for t in (qkv_p, k_lat, v_lat))
    with torch.enable_grad():
        point_attention(qq, kk, vv, heads).backward(dout)
    return qq.grad, kk.grad, vv.grad


def train_linear(x2, w, bias=None, res=None, res_mode=0, act=0):
    return gemm(x2, w, bias, res, res_mode, act)


def train_dgrad(dy, w):
    return (dy @ w).contiguous()


def gemm_tn(a, b, out=None, accumulate=False, tc=None):
    r = a.T @ b
    if out is not None:
        out.copy_(out + r if accumulate else r)
        return out
    return r.contiguous()


def colsum(a, out=None, accumulate=False):
    r = a.sum(0)
    if out is not None:
        out.copy_(out + r if accumulate else r)
        return out
    return r


def coord_embed_windows(coord, mask, w, bias, invalid, pos, cls, ws):
    B, H, W, _ = coord.shape
    C = w.shape[0]
    emb = torch.where(mask.unsqueeze(-1) > 0.5, F.linear(coord, w, bias), invalid.expand(B, H, W, C))
    emb = emb.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws * ws, C) + pos[1:].unsqueeze(0)
    return torch.cat([(cls + pos[0]).view(1, 1, C).expand(emb.shape[0], -1, -1), emb], 1).contiguous()


def train_tc():
    return False


def conv2d_nhwc_dgrad(dy, w_ohwi, in_shape, stride, pad, tc=None):
    x = torch.zeros(in_shape)
    return _via_autograd(lambda t: conv2d_nhwc(t, w_ohwi, None, stride, pad), x, dy)


def conv2d_nhwc_wgrad(x, dy, kh, kw, stride, pad, out=None, accumulate=False, tc=None):
    w = torch.zeros(dy.shape[-1], kh, kw, x.shape[-1])
    g = _via_autograd(lambda t: conv2d_nhwc(x, t, None, stride, pad), w, dy)
    if out is not None:
        out.copy_(out + g if accumulate else g)
        return out
    return g


def groupnorm_bwd_nhwc(dy, x, gamma, groups, eps, dgamma, dbeta):
    gg, bb = gamma.detach().clone().requires_grad_(True), torch.zeros_like(gamma).requires_grad_(True)
    xx = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        groupnorm_nhwc(xx, gg, bb, groups, eps, False).backward(dy)
    dgamma += gg.grad
    dbeta += bb.grad
    return xx.grad


def maxpool3x3s2_bwd_nhwc(x, dy, pt, pl):
    return _via_autograd(lambda t: maxpool3x3s2_nhwc(t, pt, pl, dy.shape[1], dy.shape[2]), x, dy)


def bilinear_bwd_nhwc(dy, H, W, align):
    x = torch.zeros(dy.shape[0], H, W, dy.shape[3])
    return _via_autograd(lambda t: bilinear_nhwc(t, dy.shape[1], dy.shape[2], align), x, dy)


def bn_stats(x2d, eps):
    mean = x2d.mean(0)
    var = x2d.var(0, unbiased=False)
    return mean, var, 1.0 / torch.sqrt(var + eps)


def bn_bwd(dy2d, x2d, mean, rstd, gamma, dgamma, dbeta):
    """Batch-statistics BatchNorm backward (the statistics depend on x)."""
    xx, gg, bb = x2d.detach().clone().requires_grad_(True), gamma.detach().clone().requires_grad_(True), torch.zeros_like(gamma).requires_grad_(True)
    with torch.enable_grad():
        m, v = xx.mean(0), xx.var(0, unbiased=False)
        eps = (1.0 / (rstd * rstd) - x2d.var(0, unbiased=False)).clamp_min(0).mean()      # recover eps from rstd (uniform over channels)
        ((xx - m) / torch.sqrt(v + eps) * gg + bb).backward(dy2d)
    dgamma += gg.grad
    dbeta += bb.grad
    return xx.grad


def avgpool_bwd_nhwc(dy, H, W):
    return (dy / (H * W)).view(dy.shape[0], 1, 1, dy.shape[1]).expand(-1, H, W, -1).contiguous()


TRAIN_OPS = ("bn_stats", "bn_bwd", "avgpool_nhwc", "avgpool_bwd_nhwc", "channel_affine", "train_tc", "conv2d_nhwc", "conv2d_nhwc_dgrad", "conv2d_nhwc_wgrad", "groupnorm_nhwc", "groupnorm_bwd_nhwc", "maxpool3x3s2_nhwc",
             "maxpool3x3s2_bwd_nhwc", "bilinear_nhwc", "bilinear_bwd_nhwc", "nchw_to_nhwc",
             "axpby", "act_bwd", "layernorm_bwd", "layernorm_bwd_generic", "mha", "mha_bwd", "point_attention", "point_attention_bwd",
             "train_linear", "train_dgrad", "gemm_tn", "colsum", "coord_embed_windows", "layernorm", "gemm", "concat2")


def install_train(monkeypatch):
    """Replace the training-step wrappers of zeroshape_b200.ops by the CPU stand-ins above (host-logic tests of the tapes)."""
    from zeroshape_b200 import ops
    g = globals()
    for name in TRAIN_OPS:
        monkeypatch.setattr(ops, name, g[name])
