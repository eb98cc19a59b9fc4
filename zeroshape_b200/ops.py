"""Thin torch-tensor wrappers over the C ABI (device memory + streams come from PyTorch; the math
is in libzeroshape_b200.so).  Every wrapper launches on torch's current CUDA stream.

Image tensors are NHWC fp32 contiguous unless noted.
"""
import math

import os

import torch

from . import _native
from ._native import lib, check

ACT_NONE, ACT_RELU, ACT_GELU, ACT_SOFTPLUS100, ACT_SIGMOID, ACT_CLAMP01 = 0, 1, 2, 3, 4, 5
RES_NONE, RES_BEFORE_ACT, RES_AFTER_ACT = 0, 1, 2


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t, name, dtype=torch.float32):
    if t is None:
        return
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError(f"{name}: expected a CUDA tensor (zeroshape_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")


def device_cc():
    return lib.zs_device_cc()


def gemm(a, w, bias=None, res=None, res_mode=RES_NONE, act=ACT_NONE, out=None):
    """out[M,N] = epi(a[M,K] @ w[N,K]^T).  `a`/`out`/`res` may be 2-D row-strided views (last dim contiguous)."""
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    for t, n in ((a, "a"), (w, "w"), (res, "res"), (out, "out")):
        if t is not None:
            if not t.is_cuda or t.dtype != torch.float32 or t.stride(-1) != 1:
                raise TypeError(f"gemm: bad tensor {n}")
    _chk(bias, "bias")
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    if res is None:
        res_mode = RES_NONE
    elif res_mode == RES_NONE:
        res_mode = RES_AFTER_ACT
    check(lib.zs_gemm_f32(_p(a), a.stride(0), _p(w), w.stride(0), _p(bias), _p(res),
                          res.stride(0) if res is not None else 0, res_mode, _p(out), out.stride(0),
                          M, N, K, act, _stream()), "zs_gemm_f32")
    return out


def _no_cache():
    return None


class _LiveMatrix:
    """The CURRENT value of a weight tensor as a matrix (weak reference: the cache object hangs off the tensor itself)."""

    def __init__(self, w):
        import weakref
        self.ref = weakref.ref(w)

    def __call__(self):
        t = self.ref().detach()
        return t.reshape(t.shape[0], -1) if t.dim() != 2 else t


class PackedWeight:
    """fp32 W[N,K] packed for the tcgen05 GEMM (hi/lo bf16, UMMA swizzled tiles).  `w` may be a tensor or a zero-argument
    callable returning the CURRENT weight (e.g. `lambda: param.detach()`): the blob is re-packed whenever the live tensor's
    storage, version, shape or device changed (in-place optimizer steps, load_state_dict(assign=True), module.to(device),
    EMA swaps that re-assign param.data)."""

    def __init__(self, w):
        self._get = w if callable(w) else (lambda: w)
        self.key = None
        self.blob = None

    def __reduce_ex__(self, protocol):
        # a cache attached to a parameter (`_packed_of`) must not travel with copy.deepcopy / torch.save of the module: the copy
        # would keep reading the ORIGINAL tensor through the weak reference.  It unpickles as None and is rebuilt on first use.
        if isinstance(self._get, _LiveMatrix):
            return (_no_cache, ())
        return super().__reduce_ex__(protocol)

    @property
    def src(self):
        return self._get()

    def get(self):
        w = self._get()
        key = (w.data_ptr(), w._version, tuple(w.shape), w.device)
        if self.key != key:
            wd = w.detach()
            if wd.dtype != torch.float32 or wd.stride(-1) != 1 or not wd.is_cuda:
                raise TypeError("PackedWeight: expected a CUDA fp32 matrix with contiguous rows")
            N, K = wd.shape
            self.blob = torch.empty(lib.zs_gemm_tc_packed_bytes(N, K), device=wd.device, dtype=torch.uint8)
            check(lib.zs_gemm_tc_pack(_p(wd), wd.stride(0), N, K, _p(self.blob), _stream()), "zs_gemm_tc_pack")
            self.key = key
        return self.blob


PRECISIONS = {"bf16x3": 0, "bf16": 1, "fp16x3": 0, "fp16": 1}     # zs_gemm_tc_f32 (bf16 images): parity / single-pass mode
# decoder chain kernels (csrc/chain_tc.cu): fp16 operand images; "bf16x3" / "bf16" are accepted as the names of the
# parity / single-pass modes for callers written against round 1
CHAIN_PRECISIONS = {"fp16x3": 0, "fp16": 1, "bf16x3": 0, "bf16": 1}
FMT_BF16, FMT_FP16 = 0, 1


def gemm_tc(a, pw, bias=None, res=None, res_mode=RES_NONE, act=ACT_NONE, out=None, precision="bf16x3"):
    """Tensor-core variant of `gemm`: `pw` is a PackedWeight of W[N,K]."""
    N, K = pw.src.shape
    assert a.dim() == 2 and a.shape[1] == K
    for t, n in ((a, "a"), (res, "res"), (out, "out")):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32 or t.stride(-1) != 1):
            raise TypeError(f"gemm_tc: bad tensor {n}")
    _chk(bias, "bias")
    M = a.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    if res is None:
        res_mode = RES_NONE
    elif res_mode == RES_NONE:
        res_mode = RES_AFTER_ACT
    check(lib.zs_gemm_tc_f32(_p(a), a.stride(0), _p(pw.get()), _p(bias), _p(res), res.stride(0) if res is not None else 0,
                             res_mode, _p(out), out.stride(0), M, N, K, act, PRECISIONS[precision], _stream()),
          "zs_gemm_tc_f32")
    return out


def pack_tiles(mats, fmt=FMT_FP16):
    """Concatenate zs_gemm_tc_pack_fmt images of fp32 matrices [256, K] (K padded to 64) -> uint8 blob (chain kernels: fp16)."""
    blobs = []
    for w in mats:
        w = w.detach().float().contiguous()
        assert w.shape[0] == 256, w.shape
        b = torch.empty(lib.zs_gemm_tc_packed_bytes(256, w.shape[1]), device=w.device, dtype=torch.uint8)
        check(lib.zs_gemm_tc_pack_fmt(_p(w), w.stride(0), 256, w.shape[1], _p(b), fmt, _stream()), "zs_gemm_tc_pack_fmt")
        blobs.append(b)
    return torch.cat(blobs)


def pack_generic(w, fmt=FMT_FP16):
    """zs_gemm_tc_pack_fmt image of one fp32 matrix [N, K] (N tiles of 256 rows, K padded to 64); chain kernels: fp16."""
    w = w.detach().float().contiguous()
    b = torch.empty(lib.zs_gemm_tc_packed_bytes(w.shape[0], w.shape[1]), device=w.device, dtype=torch.uint8)
    check(lib.zs_gemm_tc_pack_fmt(_p(w), w.stride(0), w.shape[0], w.shape[1], _p(b), fmt, _stream()), "zs_gemm_tc_pack_fmt")
    return b


def attn_pack_kv(k_lat, v_lat, heads=8):
    """Per-image latent keys/values [L, C] (row-strided views) -> (Kpacked, Vpacked) tensor-core operand images."""
    L, C = k_lat.shape
    hd = C // heads
    assert hd == 32 and L <= 208
    kp = torch.zeros(heads, 208, hd, device=k_lat.device, dtype=torch.float32)
    kp[:, :L] = k_lat.reshape(L, heads, hd).permute(1, 0, 2)
    vp = torch.zeros(heads, hd, 208, device=v_lat.device, dtype=torch.float32)
    vp[:, :, :L] = v_lat.reshape(L, heads, hd).permute(1, 2, 0)
    return (torch.cat([pack_generic(kp[h], FMT_BF16) for h in range(heads)]),
            torch.cat([pack_generic(vp[h], FMT_BF16) for h in range(heads)]))


def attn_pack_fused(k_lat, v_lat, heads=8):
    """Per-image latent keys/values [L, C] -> (Kblob, Vblob) of zs_chain_attn_fwd.
    Kblob: 4 head-pair tiles (keys along the rows, the two heads' 32 dims side by side), Vblob: 8 heads x 32 KB."""
    L, C = k_lat.shape
    hd = C // heads
    assert hd == 32 and heads == 8 and L <= 208
    kp = torch.zeros(256, C, device=k_lat.device, dtype=torch.float32)
    kp[:L] = k_lat
    kblob = torch.cat([pack_generic(kp[:, 64 * p:64 * p + 64]) for p in range(heads // 2)])
    vp = torch.zeros(heads, hd, 208, device=v_lat.device, dtype=torch.float32)
    vp[:, :, :L] = v_lat.reshape(L, heads, hd).permute(1, 2, 0)
    vpacked = torch.cat([pack_generic(vp[h]) for h in range(heads)])
    vblob = vpacked.view(heads, 4, 2, 32768)[:, :, :, :4096].contiguous().view(-1)
    return kblob, vblob


def attn_fused(qkv, kblob, vblob, n_keys, scale, precision="bf16x3", out=None):
    """qkv [M,768] -> attention output [M,256]: scores, softmax and P.V in one tcgen05 kernel (one image)."""
    assert qkv.dim() == 2 and qkv.shape[1] == 768 and qkv.stride(1) == 1 and qkv.is_cuda and qkv.dtype == torch.float32
    M = qkv.shape[0]
    O = out if out is not None else torch.empty(M, 256, device=qkv.device, dtype=torch.float32)
    assert O.stride(0) == 256 and O.stride(1) == 1
    assert kblob.numel() == 4 * 65536 and vblob.numel() == 8 * 32768
    check(lib.zs_chain_attn_fwd(_p(qkv), qkv.stride(0), M, _p(kblob), _p(vblob), n_keys, scale, _p(O),
                                CHAIN_PRECISIONS[precision], _stream()), "zs_chain_attn_fwd")
    return O


QKVATTN_FLAGS = {"kv1": 1, "s2": 2, "pv2": 4, "tmem_p": 8, "regs_blocked": 16}


def unblock_rows(O_blk, M):
    """tile-blocked attention output [tiles, 64, 128, 4] -> row-major [M, 256] (tests / debugging only)."""
    t = O_blk.shape[0]
    return O_blk.permute(0, 2, 1, 3).reshape(t * 128, 256)[:M]


def qkvattn_pack(w_qkv_folded):
    """qkv weight [768, 256] (norm1 affine already folded) -> Wblob of zs_chain_qkvattn_fwd: per head pair the rows
    [q of heads 2p, 2p+1 | k | v | 64 zero rows] as one 256-row fp16 operand image."""
    w = w_qkv_folded.detach().float()
    assert w.shape == (768, 256)
    blobs = []
    for p in range(4):
        wp = torch.zeros(256, 256, device=w.device, dtype=torch.float32)
        for j in range(3):
            wp[64 * j:64 * j + 64] = w[256 * j + 64 * p:256 * j + 64 * p + 64]
        blobs.append(pack_generic(wp))
    blob = torch.cat(blobs)
    assert blob.numel() == lib.zs_chain_qkvattn_blob_bytes()
    return blob


def chain_qkvattn(x, wblob, bias_qkv, kblob, vblob, n_keys, scale, ln_eps=1e-6, precision="fp16x3", flags=0, out=None):
    """x [M,256] -> attention output [M,256] of one image: LayerNorm + qkv + point->latent attention in one tcgen05 kernel."""
    assert x.dim() == 2 and x.shape[1] == 256 and x.is_cuda and x.dtype == torch.float32 and x.stride(1) == 1
    _chk(bias_qkv, "bias_qkv")
    assert bias_qkv.numel() == 768 and kblob.numel() == 4 * 65536 and vblob.numel() == 8 * 32768
    assert wblob.numel() == lib.zs_chain_qkvattn_blob_bytes()
    M = x.shape[0]
    if int(flags) & 16:     # tile-blocked output (see include/zeroshape_b200.h): [ceil(M/128), 64 chunks, 128 rows, 4]
        assert int(flags) & 8, "flags 16 (scores in registers) is a mode of the TMEM-probability kernel (flags 8)"
        tiles = (M + 127) // 128
        O = out if out is not None else torch.empty(tiles, 64, 128, 4, device=x.device, dtype=torch.float32)
        assert O.is_contiguous() and O.numel() == tiles * 32768 and O.dtype == torch.float32
        check(lib.zs_chain_qkvattn_fwd(_p(x), x.stride(0), M, ln_eps, _p(wblob), _p(bias_qkv), _p(kblob), _p(vblob), n_keys, scale,
                                       _p(O), 256, CHAIN_PRECISIONS[precision], int(flags), _stream()), "zs_chain_qkvattn_fwd")
        return O
    O = out if out is not None else torch.empty(M, 256, device=x.device, dtype=torch.float32)
    assert O.shape == (M, 256) and O.stride(1) == 1 and O.dtype == torch.float32
    check(lib.zs_chain_qkvattn_fwd(_p(x), x.stride(0), M, ln_eps, _p(wblob), _p(bias_qkv), _p(kblob), _p(vblob), n_keys, scale,
                                   _p(O), O.stride(0), CHAIN_PRECISIONS[precision], int(flags), _stream()), "zs_chain_qkvattn_fwd")
    return O


ATTN_SUBCHUNK = 1 << 18   # rows per scores -> P.V round trip of the two-kernel variant (attn_tc)


def attn_tc(qkv, kpacked, vpacked, n_keys, scale, precision="bf16x3", out=None):
    """qkv [M,768] (q|k|v of the query points) -> attention output [M,256] on the tensor cores (one image)."""
    assert qkv.dim() == 2 and qkv.shape[1] == 768 and qkv.stride(1) == 1 and qkv.is_cuda and qkv.dtype == torch.float32
    M = qkv.shape[0]
    dev = qkv.device
    O = out if out is not None else torch.empty(M, 256, device=dev, dtype=torch.float32)
    sub = min(M, ATTN_SUBCHUNK)
    P = torch.empty(sub, 8 * 208, device=dev, dtype=torch.float32)
    R = torch.empty(sub, 256, device=dev, dtype=torch.float32)
    Rinv = torch.empty(sub, 8, device=dev, dtype=torch.float32)
    prec = PRECISIONS[precision]
    for s in range(0, M, sub):
        m = min(sub, M - s)
        q = qkv[s:s + m]
        check(lib.zs_attn_scores_tc(_p(q), q.stride(0), _p(kpacked), m, n_keys, scale, _p(P), _p(R), _p(Rinv), prec, _stream()),
              "zs_attn_scores_tc")
        check(lib.zs_attn_pv_tc(_p(P), _p(vpacked), _p(R), _p(Rinv), _p(O[s:s + m]), m, prec, _stream()), "zs_attn_pv_tc")
    return O


def chain_mlp(x, ln_w, ln_b, ln_eps, blob, b1, b2, precision="bf16x3"):
    """In place: x[M,256] <- x + fc2(GELU(fc1(LayerNorm(x)))) on the chained tcgen05 kernel."""
    assert x.dim() == 2 and x.shape[1] == 256 and x.is_cuda and x.dtype == torch.float32 and x.stride(1) == 1
    assert blob.numel() == lib.zs_chain_mlp_blob_bytes()
    check(lib.zs_chain_mlp_fwd(_p(x), x.stride(0), x.shape[0], _p(ln_w), _p(ln_b), ln_eps, _p(blob), _p(b1), _p(b2),
                               CHAIN_PRECISIONS[precision], _stream()), "zs_chain_mlp_fwd")
    return x


def point_proj_tables(w, bias):
    """LinearProj3D weights [256,3] / bias [256] -> (pp [4,256] = w[:,0] | w[:,1] | w[:,2] | bias, pp_stat [16]) for the points
    mode of chain_qkvattn / chain_pmlp (include/zeroshape_b200.h: zs_chain_qkvattn_pts_fwd)."""
    rows = torch.cat([w.detach().t().float(), bias.detach().float().view(1, -1)], dim=0).contiguous()        # [4, 256]
    r = rows.double()
    m = r.mean(dim=1)
    c = (r - m[:, None]) @ (r - m[:, None]).t() / r.shape[1]
    st = torch.stack([m[0], m[1], m[2], m[3], c[0, 0], c[1, 1], c[2, 2], c[3, 3], c[0, 1], c[0, 2], c[1, 2], c[0, 3], c[1, 3], c[2, 3],
                      m[0] * 0, m[0] * 0]).float().contiguous()
    return rows, st


def chain_qkvattn_pts(points, pp, pp_stat, wblob, bias_qkv, kblob, vblob, n_keys, scale, ln_eps=1e-6, precision="fp16x3", flags=24):
    """chain_qkvattn of the FIRST decoder block with x = LinearProj3D(points) recomputed in the kernel -> tile-blocked output."""
    _chk(points, "points"); _chk(bias_qkv, "bias_qkv"); _chk(pp, "pp"); _chk(pp_stat, "pp_stat")
    assert points.dim() == 2 and points.shape[1] == 3 and pp.shape == (4, 256) and pp_stat.numel() == 16
    M = points.shape[0]
    O = torch.empty((M + 127) // 128, 64, 128, 4, device=points.device, dtype=torch.float32)
    check(lib.zs_chain_qkvattn_pts_fwd(_p(points), M, _p(pp), _p(pp_stat), ln_eps, _p(wblob), _p(bias_qkv), _p(kblob), _p(vblob), n_keys,
                                       scale, _p(O), CHAIN_PRECISIONS[precision], int(flags), _stream()), "zs_chain_qkvattn_pts_fwd")
    return O


def chain_pmlp(x, a_blk, proj_blob, proj_bias, ln_eps, mlp_blob, b1, b2, precision="fp16x3", points=None, pp=None):
    """In place: x[M,256] <- x' + fc2(GELU(fc1(LayerNorm(x')))), x' = x + unblock(a_blk) proj^T + proj_bias (one tcgen05 kernel).
    With points / pp the incoming x is LinearProj3D(points), recomputed in the kernel (x is only written)."""
    assert x.dim() == 2 and x.shape[1] == 256 and x.is_cuda and x.dtype == torch.float32 and x.stride(1) == 1
    M = x.shape[0]
    assert a_blk.is_contiguous() and a_blk.dtype == torch.float32 and a_blk.numel() == ((M + 127) // 128) * 32768
    assert mlp_blob.numel() == lib.zs_chain_mlp_blob_bytes() and proj_blob.numel() == lib.zs_gemm_tc_packed_bytes(256, 256)
    _chk(proj_bias, "proj_bias"); _chk(b1, "b1"); _chk(b2, "b2")
    if points is not None:
        _chk(points, "points"); _chk(pp, "pp")
        assert points.shape == (M, 3) and pp.shape == (4, 256)
    check(lib.zs_chain_pmlp_fwd(_p(x), x.stride(0), M, _p(a_blk), _p(proj_blob), _p(proj_bias), ln_eps, _p(mlp_blob), _p(b1), _p(b2),
                                _p(points), _p(pp), CHAIN_PRECISIONS[precision], _stream()), "zs_chain_pmlp_fwd")
    return x


def point_proj(points, w, bias):
    """points [M,3] -> [M,C] = points @ w[C,3]^T + bias (LinearProj3D), one output-bandwidth-bound launch."""
    _chk(points, "points"); _chk(w, "w"); _chk(bias, "bias")
    assert points.dim() == 2 and points.shape[1] == 3 and w.shape[1] == 3
    out = torch.empty(points.shape[0], w.shape[0], device=points.device, dtype=torch.float32)
    check(lib.zs_point_proj_f32(_p(points), points.shape[0], _p(w), _p(bias), _p(out), w.shape[0], _stream()), "zs_point_proj_f32")
    return out


def chain_lin(x, blob, bias, n_tiles, do_ln=False, ln_eps=1e-6, res=None, out=None, precision="bf16x3"):
    """out[M, 256*n_tiles] = LN?(x)[M,256] W^T + bias (+ res) on the chained tcgen05 kernel; `out` may alias `res`."""
    blocked = x.dim() == 4        # tile-blocked attention output of chain_qkvattn(flags & 16): rows come from `res` / `out`
    if blocked:
        assert x.shape[1:] == (64, 128, 4) and x.is_contiguous() and x.is_cuda and x.dtype == torch.float32 and not do_ln and n_tiles == 1
        M = (res if res is not None else out).shape[0]
        assert (M + 127) // 128 == x.shape[0]
        do_ln, ldx = 2, 256
    else:
        assert x.dim() == 2 and x.shape[1] == 256 and x.is_cuda and x.dtype == torch.float32 and x.stride(1) == 1
        M, ldx = x.shape[0], x.stride(0)
    assert blob.numel() == lib.zs_gemm_tc_packed_bytes(256 * n_tiles, 256)
    _chk(bias, "bias")
    if out is None:
        out = torch.empty(M, 256 * n_tiles, device=x.device, dtype=torch.float32)
    assert out.shape == (M, 256 * n_tiles) and out.stride(1) == 1
    if res is not None:
        assert res.shape == out.shape and res.stride(1) == 1 and res.dtype == torch.float32
    check(lib.zs_chain_lin_fwd(_p(x), ldx, M, int(do_ln), ln_eps, _p(blob), n_tiles, _p(bias), _p(res),
                               res.stride(0) if res is not None else 0, _p(out), out.stride(0), CHAIN_PRECISIONS[precision], _stream()),
          "zs_chain_lin_fwd")
    return out


def chain_occ(x, points, ln_w, ln_b, ln_eps, blob, biases, w8, b8, sigmoid=False, precision="bf16x3"):
    """logits[M] = MLPBlocks([points, LayerNorm(x)]) on the chained tcgen05 kernel."""
    assert x.dim() == 2 and x.shape[1] == 256 and x.stride(1) == 1 and points.shape == (x.shape[0], 3)
    _chk(points, "points"); _chk(biases, "biases"); _chk(w8, "w8")
    assert blob.numel() == lib.zs_chain_occ_blob_bytes()
    out = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
    check(lib.zs_chain_occ_fwd(_p(x), x.stride(0), _p(points), x.shape[0], _p(ln_w), _p(ln_b), ln_eps, _p(blob), _p(biases),
                               _p(w8), float(b8), _p(out), int(sigmoid), CHAIN_PRECISIONS[precision], _stream()), "zs_chain_occ_fwd")
    return out


# Dense layers of the encoder side run on the tcgen05 kernel when the device has it ("auto"); "f32" forces the
# bit-faithful FFMA kernels (parity pinning), "tc" requires the tensor-core path.
if os.environ.get("ZS_GEMM_SPLITK"):              # A/B switch of the split-K path of the tcgen05 GEMM / convolution (few-tile layers)
    lib.zs_debug_gemm_splitk(int(os.environ["ZS_GEMM_SPLITK"]))

ENCODER_ENGINE = "auto"
ENCODER_PRECISION = "bf16x3"
# Graph.forward in eval mode replays the image -> latents encoder from a CUDA graph (model/compute_graph/graph_shape.py); False =
# launch it op by op (also forced while an OpTimer is active: it needs the individual launches)
ENCODER_CUDA_GRAPH = os.environ.get("ZS_ENCODER_GRAPH", "1") != "0"


def _encoder_tc():
    if ENCODER_ENGINE == "f32":
        return False
    ok = device_cc() == 100
    if ENCODER_ENGINE == "tc" and not ok:
        raise RuntimeError("ENCODER_ENGINE='tc' needs an sm_100 device")
    return ok


def _packed_of(w):
    """tcgen05 operand image of a weight matrix / OHWI filter, cached on the tensor object.  The cache reads the LIVE tensor
    on every use (PackedWeight above), so re-assigning `param.data` after a first forward cannot leave a stale image."""
    pw = getattr(w, "_zs_packed", None)
    if pw is None:
        pw = PackedWeight(_LiveMatrix(w))
        w._zs_packed = pw
    return pw


def linear(x, w, bias=None, act=ACT_NONE, res=None, res_mode=RES_NONE, tc=None):
    """F.linear on the last dim of a contiguous tensor (`tc`: None = ENCODER_ENGINE policy, False = force FFMA)."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1])
    r2 = res.reshape(-1, w.shape[0]) if res is not None else None
    if w.shape[0] >= 64 and w.shape[1] % 4 == 0 and w.shape[1] >= 64 and (tc if tc is not None else _encoder_tc()):
        y = gemm_tc(x2, _packed_of(w), bias, r2, res_mode, act, precision=ENCODER_PRECISION)
    else:
        y = gemm(x2, w, bias, r2, res_mode, act)
    return y.reshape(*shp[:-1], w.shape[0])


def conv2d_nhwc(x, w, bias=None, stride=1, pad=(0, 0, 0, 0), act=ACT_NONE, res=None, res_mode=RES_NONE,
                pre_relu=False, tc=None, precision=None):
    """x [B,H,W,Cin]; w [Cout,KH,KW,Cin]; pad = (top, bottom, left, right).  `tc`: None = ENCODER_ENGINE policy, False = FFMA."""
    _chk(x, "x"); _chk(w, "w"); _chk(bias, "bias"); _chk(res, "res")
    B, H, W, Cin = x.shape
    Cout, KH, KW, Cin2 = w.shape
    assert Cin == Cin2, (x.shape, w.shape)
    pt, pb, pl, pr = pad
    OH = (H + pt + pb - KH) // stride + 1
    OW = (W + pl + pr - KW) // stride + 1
    y = torch.empty(B, OH, OW, Cout, device=x.device, dtype=torch.float32)
    if res is None:
        res_mode = RES_NONE
    elif res_mode == RES_NONE:
        res_mode = RES_AFTER_ACT
    if Cin % 4 == 0 and KH * KW * Cin >= 32 and (tc if tc is not None else _encoder_tc()):
        check(lib.zs_conv2d_nhwc_tc(_p(x), B, H, W, Cin, _p(_packed_of(w).get()), _p(bias), _p(res), res_mode, _p(y), Cout,
                                    KH, KW, stride, pt, pl, OH, OW, act, int(pre_relu), PRECISIONS[precision or ENCODER_PRECISION],
                                    _stream()), "zs_conv2d_nhwc_tc")
        return y
    check(lib.zs_conv2d_nhwc_f32(_p(x), B, H, W, Cin, _p(w), _p(bias), _p(res), res_mode, _p(y), Cout, KH, KW,
                                 stride, pt, pl, OH, OW, act, int(pre_relu), _stream()), "zs_conv2d_nhwc_f32")
    return y


def layernorm(x, gamma, beta, eps):
    _chk(x, "x"); _chk(gamma, "gamma"); _chk(beta, "beta")
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty_like(x)
    check(lib.zs_layernorm_f32(_p(x), C, _p(gamma), _p(beta), _p(y), C, rows, C, eps, _stream()), "zs_layernorm_f32")
    return y


def groupnorm_nhwc(x, gamma, beta, groups, eps, relu, res=None):
    _chk(x, "x"); _chk(res, "res")
    B, H, W, C = x.shape
    y = torch.empty_like(x)
    ws = torch.empty(lib.zs_groupnorm_ws_bytes(B, H * W, C, groups) // 8, device=x.device, dtype=torch.float64)
    check(lib.zs_groupnorm_nhwc_ws_f32(_p(x), _p(gamma), _p(beta), _p(res), _p(y), B, H * W, C, groups, eps, int(relu), _p(ws),
                                       _stream()), "zs_groupnorm_nhwc_ws_f32")
    return y


def channel_affine(x, scale, shift, act=ACT_NONE, res=None):
    _chk(x, "x"); _chk(scale, "scale"); _chk(shift, "shift"); _chk(res, "res")
    C = x.shape[-1]
    y = torch.empty_like(x)
    check(lib.zs_channel_affine_f32(_p(x), _p(scale), _p(shift), _p(res), _p(y), x.numel() // C, C, act, _stream()),
          "zs_channel_affine_f32")
    return y


def axpby(a, alpha=1.0, b=None, beta=1.0, act=ACT_NONE):
    _chk(a, "a"); _chk(b, "b")
    y = torch.empty_like(a)
    check(lib.zs_axpby_f32(_p(a), alpha, _p(b), beta, _p(y), a.numel(), act, _stream()), "zs_axpby_f32")
    return y


def maxpool3x3s2_nhwc(x, pad_top, pad_left, OH, OW):
    _chk(x, "x")
    B, H, W, C = x.shape
    y = torch.empty(B, OH, OW, C, device=x.device, dtype=torch.float32)
    check(lib.zs_maxpool3x3s2_nhwc_f32(_p(x), _p(y), B, H, W, C, pad_top, pad_left, OH, OW, _stream()),
          "zs_maxpool3x3s2_nhwc_f32")
    return y


def avgpool_nhwc(x):
    _chk(x, "x")
    B, H, W, C = x.shape
    y = torch.empty(B, C, device=x.device, dtype=torch.float32)
    check(lib.zs_avgpool_nhwc_f32(_p(x), _p(y), B, H * W, C, _stream()), "zs_avgpool_nhwc_f32")
    return y


def bilinear_nhwc(x, OH, OW, align_corners):
    _chk(x, "x")
    B, H, W, C = x.shape
    y = torch.empty(B, OH, OW, C, device=x.device, dtype=torch.float32)
    check(lib.zs_bilinear_nhwc_f32(_p(x), _p(y), B, H, W, C, OH, OW, int(align_corners), _stream()),
          "zs_bilinear_nhwc_f32")
    return y


def nchw_to_nhwc(x, scale=1.0, shift=0.0):
    _chk(x, "x")
    B, C, H, W = x.shape
    y = torch.empty(B, H, W, C, device=x.device, dtype=torch.float32)
    check(lib.zs_nchw_to_nhwc_f32(_p(x), _p(y), B, C, H, W, scale, shift, _stream()), "zs_nchw_to_nhwc_f32")
    return y


def nhwc_to_nchw(x):
    _chk(x, "x")
    B, H, W, C = x.shape
    y = torch.empty(B, C, H, W, device=x.device, dtype=torch.float32)
    check(lib.zs_nhwc_to_nchw_f32(_p(x), _p(y), B, C, H, W, _stream()), "zs_nhwc_to_nchw_f32")
    return y


def mha(qkv, heads, tc=None, precision="fp16x3"):
    """qkv [B,T,3*C] -> [B,T,C] (timm Attention core, softmax(q k^T / sqrt(hd)) v).  `tc`: None = ENCODER_ENGINE policy
    (tcgen05 kernel zs_mha_tc_f32 when the shape fits: T <= 208, head dim 32 / 64), False = the fp32 FFMA kernel."""
    _chk(qkv, "qkv")
    B, T, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    out = torch.empty(B, T, C, device=qkv.device, dtype=torch.float32)
    if tc is None:
        tc = _encoder_tc()
    if tc and T <= 208 and hd in (32, 64):
        check(lib.zs_mha_tc_f32(_p(qkv), _p(out), B, T, heads, hd, hd ** -0.5, CHAIN_PRECISIONS[precision], _stream()), "zs_mha_tc_f32")
        return out
    check(lib.zs_mha_f32(_p(qkv), _p(out), B, T, heads, hd, hd ** -0.5, _stream()), "zs_mha_f32")
    return out


def point_attention(qkv_p, k_lat, v_lat, heads, attn=None, attn_scale=1.0, attn_accumulate=False, tc=None):
    """qkv_p [B,P,3C]; k_lat/v_lat [B,L,C] views with row stride; returns [B,P,C].  `tc`: None = tensor cores
    (zs_point_attention_tc_f32, one fp16 pass) when the training engine is on them in its single-pass mode and no attention
    map is asked for, else the FFMA kernel; True = tensor cores in split-fp16 (fp32-grade)."""
    _chk(qkv_p, "qkv_p")
    B, Pn, C3 = qkv_p.shape
    C = C3 // 3
    L = k_lat.shape[1]
    assert k_lat.stride(2) == 1 and v_lat.stride(2) == 1 and k_lat.stride(1) == v_lat.stride(1)
    assert k_lat.stride(0) == L * k_lat.stride(1) and v_lat.stride(0) == L * v_lat.stride(1)
    out = torch.empty(B, Pn, C, device=qkv_p.device, dtype=torch.float32)
    _chk(attn, "attn")
    single = tc is None
    if tc is None:
        tc = train_tc() and TRAIN_PRECISION == "bf16"
    if tc and attn is None and L <= 208 and C // heads == 32 and k_lat.stride(1) % 4 == 0 and k_lat.data_ptr() % 16 == 0 and v_lat.data_ptr() % 16 == 0:
        check(lib.zs_point_attention_tc_f32(_p(qkv_p), _p(k_lat), _p(v_lat), k_lat.stride(1), _p(out), B, Pn, L, heads, C // heads,
                                            (C // heads) ** -0.5, 1 if single else 0, _stream()), "zs_point_attention_tc_f32")
        return out
    check(lib.zs_point_attention_f32(_p(qkv_p), _p(k_lat), _p(v_lat), k_lat.stride(1), _p(out), _p(attn), attn_scale,
                                     int(attn_accumulate), B, Pn, L, heads, C // heads, (C // heads) ** -0.5, _stream()),
          "zs_point_attention_f32")
    return out


def coord_embed_windows(coord, mask, w, bias, invalid, pos, cls, ws):
    """CoordEmb front end: coord [B,H,W,3], mask [B,H,W] (float 0/1) -> window tokens [B*(H/ws)*(W/ws), ws*ws+1, C]."""
    for t, n in ((coord, "coord"), (mask, "mask"), (w, "w"), (bias, "bias"), (invalid, "invalid"), (pos, "pos"), (cls, "cls")):
        _chk(t, n)
    B, H, W, _ = coord.shape
    C = w.shape[0]
    assert coord.shape[3] == 3 and mask.shape == (B, H, W) and w.shape == (C, 3) and pos.numel() == (ws * ws + 1) * C
    out = torch.empty(B * (H // ws) * (W // ws), ws * ws + 1, C, device=coord.device, dtype=torch.float32)
    check(lib.zs_coord_embed_windows_f32(_p(coord), _p(mask), _p(w), _p(bias), _p(invalid), _p(pos), _p(cls), _p(out), B, H, W, C, ws,
                                         _stream()), "zs_coord_embed_windows_f32")
    return out


def dense_grid(n, rmin, rmax, x0, x1, device):
    out = torch.empty(x1 - x0, n, n, 3, device=device, dtype=torch.float32)
    check(lib.zs_dense_grid_f32(_p(out), n, rmin, rmax, x0, x1, _stream()), "zs_dense_grid_f32")
    return out


def concat2(a, b, s=1.0):
    """cat([a, b], -1) / s for 2-D row-strided inputs (the skip connections of MLPBlocks divide by sqrt(2), implicit.py:179-180)."""
    assert a.dim() == 2 and b.dim() == 2 and a.shape[0] == b.shape[0]
    rows = a.shape[0]
    y = torch.empty(rows, a.shape[1] + b.shape[1], device=a.device, dtype=torch.float32)
    check(lib.zs_concat2_f32(_p(a), a.stride(0), a.shape[1], _p(b), b.stride(0), b.shape[1], s, _p(y), y.stride(0),
                             rows, _stream()), "zs_concat2_f32")
    return y


def intr_param2mtx(params, H, W):
    _chk(params, "params")
    B = params.shape[0]
    K = torch.empty(B, 3, 3, device=params.device, dtype=torch.float32)
    check(lib.zs_intr_param2mtx_f32(_p(params), _p(K), B, H, W, _stream()), "zs_intr_param2mtx_f32")
    return K


def unproject_normalize(depth, mask, K):
    """depth, mask [B,1,H,W] (or [B,H,W]); K [B,3,3] -> seen_points [B,HW,3], mean [B,3], scale [B]."""
    depth = depth.contiguous(); mask = mask.contiguous().float(); K = K.contiguous()
    _chk(depth, "depth"); _chk(mask, "mask"); _chk(K, "K")
    B = depth.shape[0]
    H, W = depth.shape[-2:]
    pts = torch.empty(B, H * W, 3, device=depth.device, dtype=torch.float32)
    mean = torch.empty(B, 3, device=depth.device, dtype=torch.float32)
    scale = torch.empty(B, device=depth.device, dtype=torch.float32)
    check(lib.zs_unproject_normalize_f32(_p(depth), _p(mask), _p(K), _p(pts), _p(mean), _p(scale), B, H, W, None,
                                         _stream()), "zs_unproject_normalize_f32")
    return pts, mean, scale


def unproject(depth, K):
    """Raw unprojection: depth [B,1,H,W], K [B,3,3] -> [B,HW,3] (utils/camera.py:88-108)."""
    depth = depth.contiguous(); K = K.contiguous()
    _chk(depth, "depth"); _chk(K, "K")
    B = depth.shape[0]
    H, W = depth.shape[-2:]
    pts = torch.empty(B, H * W, 3, device=depth.device, dtype=torch.float32)
    check(lib.zs_unproject_normalize_f32(_p(depth), None, _p(K), _p(pts), None, None, B, H, W, None, _stream()),
          "zs_unproject_normalize_f32")
    return pts


def chamfer_nn(xyz1, xyz2):
    """-> dist1 [b,n], dist2 [b,m] (squared), idx1, idx2 (int32)."""
    _chk(xyz1, "xyz1"); _chk(xyz2, "xyz2")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    d1 = torch.empty(b, n, device=dev); d2 = torch.empty(b, m, device=dev)
    i1 = torch.empty(b, n, device=dev, dtype=torch.int32); i2 = torch.empty(b, m, device=dev, dtype=torch.int32)
    ws = torch.empty(lib.zs_chamfer_ws_bytes(b, n, m), device=dev, dtype=torch.uint8)
    check(lib.zs_chamfer_nn_fwd(_p(xyz1), _p(xyz2), b, n, m, _p(d1), _p(d2), _p(i1), _p(i2), _p(ws), _stream()),
          "zs_chamfer_nn_fwd")
    return d1, d2, i1, i2


def chamfer_nn_bwd(xyz1, xyz2, g1, g2, i1, i2):
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1 = torch.zeros_like(xyz1); gx2 = torch.zeros_like(xyz2)
    check(lib.zs_chamfer_nn_bwd(_p(xyz1), _p(xyz2), _p(g1), _p(g2), _p(i1), _p(i2), b, n, m, _p(gx1), _p(gx2),
                                _stream()), "zs_chamfer_nn_bwd")
    return gx1, gx2


BVH_QUERY_VARIANT = 0      # zs_nn_bvh_query kernel: 0 = one thread per query, 1 = warp-cooperative (identical results)


class NNBvh:
    """Flat box hierarchy over `sets` point clouds [sets, n, 3] (zs_nn_bvh_build): exact NN queries at ~1/20 of the
    brute-force pair evaluations, same distances / tie-breaking as `chamfer_nn`."""

    def __init__(self, pts):
        _chk(pts, "pts")
        assert pts.dim() == 3 and pts.shape[2] == 3
        self.sets, self.n = pts.shape[0], pts.shape[1]
        self.blob = torch.empty(lib.zs_nn_bvh_bytes(self.sets, self.n), device=pts.device, dtype=torch.uint8)
        check(lib.zs_nn_bvh_build(_p(pts), self.sets, self.n, _p(self.blob), _stream()), "zs_nn_bvh_build")

    def morton_order(self, s=0):
        """Original indices of set `s` in Morton order (int32 [n]): a cache-friendly processing order for queries made of
        the same points under any rigid / similarity transform."""
        npad = (self.n + 31) // 32 * 32
        per_set = self.blob.numel() // self.sets
        pts = self.blob[s * per_set: s * per_set + npad * 16].view(torch.int32).view(npad, 4)
        return pts[:self.n, 3].contiguous()

    def query(self, q, batch=None, q_order=None, variant=None):
        """q [sets_q, nq, 3] (sets_q = 1: shared by the batch) -> dist [batch, nq] (squared), idx [batch, nq] int32."""
        _chk(q, "q")
        sets_q, nq = q.shape[0], q.shape[1]
        batch = batch or max(self.sets, sets_q)
        dist = torch.empty(batch, nq, device=q.device, dtype=torch.float32)
        idx = torch.empty(batch, nq, device=q.device, dtype=torch.int32)
        _chk(q_order, "q_order", torch.int32)
        check(lib.zs_nn_bvh_query(_p(self.blob), self.sets, self.n, _p(q), sets_q, nq, batch, _p(q_order), _p(dist), _p(idx),
                                  BVH_QUERY_VARIANT if variant is None else int(variant), _stream()), "zs_nn_bvh_query")
        return dist, idx


def chamfer_stats(sq1, sq2, thresholds, squared=True):
    _chk(sq1, "sq1"); _chk(sq2, "sq2")
    b, n = sq1.shape
    m = sq2.shape[1]
    dev = sq1.device
    thr = torch.tensor(list(thresholds), device=dev, dtype=torch.float32)
    T = thr.numel()
    mean1 = torch.empty(b, device=dev); mean2 = torch.empty(b, device=dev)
    f1 = torch.empty(b, T, device=dev); f2 = torch.empty(b, T, device=dev)
    check(lib.zs_chamfer_stats(_p(sq1), _p(sq2), b, n, m, _p(thr), T, int(squared), _p(mean1), _p(mean2), _p(f1), _p(f2), _stream()),
          "zs_chamfer_stats")
    return mean1, mean2, f1, f2


def marching_cubes_count(vol, iso):
    """Pass 1 of marching cubes on vol [nx,n,n] (nx == n: the whole grid; nx < n: an x-slab): -> (workspace, counts [2] int32 on
    the device = #vertices, #faces).  No host sync: batch several volumes, read all counts at once, then `marching_cubes_emit`."""
    _chk(vol, "vol")
    nx, n = vol.shape[0], vol.shape[1]
    assert vol.shape == (nx, n, n) and 1 <= nx <= n
    ws = torch.empty(lib.zs_mc_slab_ws_bytes(nx, n), device=vol.device, dtype=torch.uint8)
    counts = torch.empty(2, device=vol.device, dtype=torch.int32)
    check(lib.zs_mc_slab_count(_p(vol), nx, n, float(iso), _p(ws), _p(counts), _stream()), "zs_mc_slab_count")
    return ws, counts


def marching_cubes_emit(vol, iso, ws, V, F, x_offset=0):
    """Pass 2: -> (verts [V,3] fp32 in global index units, faces [F,3] int32)."""
    nx, n = vol.shape[0], vol.shape[1]
    verts = torch.empty(V, 3, device=vol.device, dtype=torch.float32)
    faces = torch.empty(F, 3, device=vol.device, dtype=torch.int32)
    if V > 0:
        check(lib.zs_mc_slab_emit(_p(vol), nx, n, float(iso), _p(ws), _p(verts), _p(faces), int(x_offset), _stream()),
              "zs_mc_slab_emit")
    return verts, faces


class MeshFuture:
    """Marching cubes with the size read-back taken off the critical path: pass 1 (classify + scans) and an asynchronous copy of
    the two counts to pinned host memory are queued at construction; `result()` waits for THAT copy only (an event, not the
    stream), allocates the outputs and queues pass 2.  Work queued between the two calls -- the next shape's decoder -- keeps
    the GPU busy while the host learns the sizes."""

    def __init__(self, vol, iso, x_offset=0):
        self.vol, self.iso, self.x_offset = vol, iso, x_offset
        self.ws, counts = marching_cubes_count(vol, iso)
        self.host = torch.empty(2, dtype=torch.int32, pin_memory=True)
        self.host.copy_(counts, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()

    def result(self):
        self.event.synchronize()
        V, F = self.host.tolist()
        return marching_cubes_emit(self.vol, self.iso, self.ws, V, F, self.x_offset)


def marching_cubes(vol, iso, x_offset=0):
    """vol [n,n,n] (or an x-slab [nx,n,n] whose first slice is global slice `x_offset`) fp32 CUDA ->
    (verts [V,3] fp32 in index units, faces [F,3] int32), both on device."""
    ws, counts = marching_cubes_count(vol, iso)
    V, F = counts.tolist()   # the one host sync of the mesh path (sizes of the outputs); MeshFuture defers it
    return marching_cubes_emit(vol, iso, ws, V, F, x_offset)


def mesh_sample(verts, faces, num, vscale=1.0, voffset=0.0, seed=0):
    dev = verts.device
    F = faces.shape[0]
    pts = torch.empty(num, 3, device=dev, dtype=torch.float32)
    ws = torch.empty(lib.zs_mesh_sample_ws_bytes(F), device=dev, dtype=torch.uint8)
    check(lib.zs_mesh_sample(_p(verts) if F else None, _p(faces) if F else None, verts.shape[0], F, vscale, voffset,
                             num, seed, _p(ws), _p(pts), _stream()), "zs_mesh_sample")
    return pts


# ---------------------------------------------------------------------------------------------------------------------
# Per-op device timing (bench.py's per-kernel roofline table, tools/diag_decoder.py).  Not used on the product path.
class OpTimer:
    """Context manager: wraps the module-level op wrappers with CUDA event pairs on the current stream.

        with ops.OpTimer() as t:  ...run the hot path...
        t.summary() -> {op name: (launch groups, total ms)}   (synchronises)
    """
    NAMES = ("point_proj", "chain_lin", "chain_qkvattn", "attn_fused", "attn_tc", "point_attention", "chain_mlp", "chain_pmlp", "chain_qkvattn_pts", "chain_occ", "gemm_tc", "gemm",
             "conv2d_nhwc", "layernorm", "groupnorm_nhwc", "mha", "bilinear_nhwc", "dense_grid", "axpby", "marching_cubes",
             "mesh_sample", "unproject_normalize", "concat2", "maxpool3x3s2_nhwc", "avgpool_nhwc", "chamfer_nn")

    def __init__(self, clock_probe=False):
        self.events = []
        self.saved = {}
        self.clock_probe = clock_probe
        self.clocks = []

    def _probe(self, name):
        if self.clock_probe:
            buf = torch.empty(1, device="cuda", dtype=torch.float32)
            check(lib.zs_debug_clock_mhz(_p(buf), _stream()), "zs_debug_clock_mhz")
            self.clocks.append((name, buf))

    def __enter__(self):
        g = globals()
        self.saved["ENCODER_CUDA_GRAPH"] = g["ENCODER_CUDA_GRAPH"]        # the timer needs the individual launches
        g["ENCODER_CUDA_GRAPH"] = False
        for name in self.NAMES:
            fn = g.get(name)
            if fn is None:
                continue
            self.saved[name] = fn

            def wrapped(*a, __fn=fn, __name=name, **k):
                key = __name
                if __name == "chain_lin":
                    key = "chain_lin[qkv]" if (a[3] if len(a) > 3 else k.get("n_tiles")) == 3 else "chain_lin[proj]"
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = __fn(*a, **k)
                e1.record()
                self.events.append((key, e0, e1))
                self._probe(key)
                return r
            g[name] = wrapped
        return self

    def __exit__(self, *exc):
        globals().update(self.saved)
        return False

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for key, e0, e1 in self.events:
            c, ms = out.get(key, (0, 0.0))
            out[key] = (c + 1, ms + e0.elapsed_time(e1))
        return out

    def clock_trace(self):
        torch.cuda.synchronize()
        return [(n, float(b.item())) for n, b in self.clocks]


# ---------------------------------------------------------------------------------------------------------------------
# Training-step kernels (csrc/train.cu, csrc/gemm_tn_tc.cu): backward of every layer, BCE shape loss, AdamW.
#
# The GEMM-shaped work of the step (forward, data gradients, weight gradients of every nn.Linear / nn.Conv2d) runs on the
# tcgen05 kernels when TRAIN_ENGINE allows it: "auto" = tensor cores on an sm_100 device, "tc" = require them, "f32" = the
# FFMA kernels (bit-faithful gradients for parity pinning).  TRAIN_PRECISION: "bf16x3" (fp32-grade, split operands) or
# "bf16" (single pass with fp32 accumulation -- the mixed-precision mode of BASELINE config 3; master weights, activations and
# gradients stay fp32 in memory).
TRAIN_ENGINE = "auto"
TRAIN_PRECISION = "bf16x3"
TN_LAYOUT = 0          # operand layout of the TN (weight-gradient) kernel: MN-major tiles


def train_tc():
    if TRAIN_ENGINE == "f32":
        return False
    ok = device_cc() == 100
    if TRAIN_ENGINE == "tc" and not ok:
        raise RuntimeError("TRAIN_ENGINE='tc' needs an sm_100 device")
    return ok


def train_linear(x2, w, bias=None, res=None, res_mode=RES_NONE, act=ACT_NONE):
    """Forward of nn.Linear inside a training step: x2 [M,K] @ w[N,K]^T (+ bias ...)."""
    N, K = w.shape
    if train_tc() and N >= 64 and K >= 64:
        return gemm_tc(x2, PackedWeight(w.detach()), bias, res, res_mode, act, precision=TRAIN_PRECISION)
    return gemm(x2, w, bias, res, res_mode, act)


def train_dgrad(dy, w):
    """dX[M,K] = dY[M,N] @ w[N,K]  (data gradient of nn.Linear / a 1x1 convolution); weights-only transpose + pack."""
    N, K = w.shape
    wt = w.detach().t().clone(memory_format=torch.contiguous_format)      # clone() normalises the strides of a [K, 1] transpose
    if train_tc() and N >= 64 and K >= 64:
        return gemm_tc(dy, PackedWeight(wt), precision=TRAIN_PRECISION)
    return gemm(dy, wt)


def gemm_tn(a, b, out=None, accumulate=False, tc=None):
    """out[N,K] (+)= a[M,N]^T @ b[M,K]  (weight gradient dW = dY^T X); 2-D row-strided inputs."""
    assert a.dim() == 2 and b.dim() == 2 and a.shape[0] == b.shape[0] and a.stride(1) == 1 and b.stride(1) == 1
    M, N = a.shape
    K = b.shape[1]
    if out is None:
        out = torch.empty(N, K, device=a.device, dtype=torch.float32)
        accumulate = False
    assert out.shape == (N, K) and out.stride(1) == 1
    if (tc if tc is not None else train_tc()) and N >= 32 and K >= 64 and M >= 128:
        check(lib.zs_gemm_tn_tc(_p(a), a.stride(0), _p(b), b.stride(0), _p(out), out.stride(0), M, N, K, int(accumulate),
                                PRECISIONS[TRAIN_PRECISION], TN_LAYOUT, _stream()), "zs_gemm_tn_tc")
        return out
    check(lib.zs_gemm_tn_f32(_p(a), a.stride(0), _p(b), b.stride(0), _p(out), out.stride(0), M, N, K, int(accumulate), _stream()),
          "zs_gemm_tn_f32")
    return out


def colsum(a, out=None, accumulate=False):
    """out[N] (+)= a[M,N].sum(0)  (bias gradient)."""
    assert a.dim() == 2 and a.stride(1) == 1
    M, N = a.shape
    if out is None:
        out = torch.empty(N, device=a.device, dtype=torch.float32)
        accumulate = False
    check(lib.zs_colsum_f32(_p(a), a.stride(0), M, N, _p(out), int(accumulate), _stream()), "zs_colsum_f32")
    return out


def act_bwd(dy, z, act):
    _chk(dy, "dy"); _chk(z, "z")
    dx = torch.empty_like(dy)
    check(lib.zs_act_bwd_f32(_p(dy), _p(z), _p(dx), dy.numel(), act, _stream()), "zs_act_bwd_f32")
    return dx


def layernorm_bwd(dy, x, gamma, eps, dgamma=None, dbeta=None):
    """dx of LayerNorm over the last dim (256); dgamma / dbeta are accumulated into when given."""
    _chk(dy, "dy"); _chk(x, "x"); _chk(gamma, "gamma"); _chk(dgamma, "dgamma"); _chk(dbeta, "dbeta")
    C = x.shape[-1]
    dx = torch.empty_like(x)
    check(lib.zs_layernorm_bwd_f32(_p(dy), _p(x), _p(gamma), eps, _p(dx), _p(dgamma), _p(dbeta), x.numel() // C, C, _stream()),
          "zs_layernorm_bwd_f32")
    return dx


def point_attention_bwd(qkv_p, k_lat, v_lat, out, dout, heads, tc=None):
    """Backward of `point_attention`: -> dqkv_p [B,P,3C], dk_lat [B,L,C], dv_lat [B,L,C].  `tc`: None = tensor cores
    (zs_point_attention_bwd_tc_f32, one fp16 pass) when the training engine is on them in its single-pass mode, else FFMA."""
    _chk(qkv_p, "qkv_p"); _chk(out, "out"); _chk(dout, "dout")
    B, Pn, C3 = qkv_p.shape
    C = C3 // 3
    L = k_lat.shape[1]
    assert k_lat.stride(2) == 1 and v_lat.stride(2) == 1 and k_lat.stride(1) == v_lat.stride(1)
    assert k_lat.stride(0) == L * k_lat.stride(1) and v_lat.stride(0) == L * v_lat.stride(1)
    dqkv = torch.empty_like(qkv_p)
    dk = torch.empty(B, L, C, device=qkv_p.device, dtype=torch.float32)
    dv = torch.empty(B, L, C, device=qkv_p.device, dtype=torch.float32)
    if tc is None:
        tc = train_tc() and TRAIN_PRECISION == "bf16"
    if tc and L <= 208 and C // heads == 32 and k_lat.stride(1) % 4 == 0:
        ws = torch.empty(lib.zs_point_attention_bwd_tc_ws_bytes(B, Pn, heads), device=qkv_p.device, dtype=torch.uint8)
        check(lib.zs_point_attention_bwd_tc_f32(_p(qkv_p), _p(k_lat), _p(v_lat), k_lat.stride(1), _p(dout), _p(dqkv), _p(dk), _p(dv), C,
                                                B, Pn, L, heads, C // heads, (C // heads) ** -0.5, _p(ws), _stream()),
              "zs_point_attention_bwd_tc_f32")
        return dqkv, dk, dv
    check(lib.zs_point_attention_bwd_f32(_p(qkv_p), _p(k_lat), _p(v_lat), k_lat.stride(1), _p(out), _p(dout), _p(dqkv), _p(dk), _p(dv),
                                         C, B, Pn, L, heads, C // heads, (C // heads) ** -0.5, _stream()), "zs_point_attention_bwd_f32")
    return dqkv, dk, dv


def mha_bwd(qkv, dout, heads, tc=None):
    """Backward of `mha`: dqkv [B,T,3C].  `tc`: None = tensor cores (zs_mha_bwd_tc_f32: one fp16 pass, fp32 accumulation) when the
    training engine is on them in its single-pass mode (TRAIN_PRECISION == "bf16") and the shape fits, else the fp32 FFMA kernels."""
    _chk(qkv, "qkv"); _chk(dout, "dout")
    B, T, C3 = qkv.shape
    C = C3 // 3
    dqkv = torch.empty_like(qkv)
    if tc is None:
        tc = train_tc() and TRAIN_PRECISION == "bf16"
    if tc and T <= 208 and C // heads in (32, 64):
        ws = torch.empty(lib.zs_mha_bwd_tc_ws_bytes(B, T, heads), device=qkv.device, dtype=torch.uint8)
        check(lib.zs_mha_bwd_tc_f32(_p(qkv), _p(dout), _p(dqkv), B, T, heads, C // heads, (C // heads) ** -0.5, _p(ws), _stream()),
              "zs_mha_bwd_tc_f32")
        return dqkv
    ws = torch.empty(lib.zs_mha_bwd_ws_bytes(B, T, heads), device=qkv.device, dtype=torch.uint8)
    check(lib.zs_mha_bwd_f32(_p(qkv), _p(dout), _p(dqkv), B, T, heads, C // heads, (C // heads) ** -0.5, _p(ws), _stream()),
          "zs_mha_bwd_f32")
    return dqkv


def bce_logits_loss(logits, sdf, impt_thres, impt_weight):
    """utils/loss.py:18-28 -> scalar device tensor (mean of the weighted BCE-with-logits against sdf < 0)."""
    _chk(logits, "logits"); _chk(sdf, "sdf")
    assert logits.shape == sdf.shape
    ws = torch.empty(1, device=logits.device, dtype=torch.float64)
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    check(lib.zs_bce_logits_fwd(_p(logits), _p(sdf), logits.numel(), impt_thres, impt_weight, _p(ws), _p(loss), _stream()),
          "zs_bce_logits_fwd")
    return loss


def bce_logits_loss_bwd(logits, sdf, impt_thres, impt_weight, grad_scale=1.0):
    d = torch.empty_like(logits)
    check(lib.zs_bce_logits_bwd(_p(logits), _p(sdf), logits.numel(), impt_thres, impt_weight, grad_scale, _p(d), _stream()),
          "zs_bce_logits_bwd")
    return d


def erode_mask(mask, pool=4):
    """MidasLoss.erode_mask: [B,1,H,W] raw mask -> 1.0 where a whole pool x pool block is valid."""
    mask = mask.float().contiguous()
    _chk(mask, "mask")
    B, _, H, W = mask.shape
    out = torch.empty_like(mask)
    check(lib.zs_mask_erode_f32(_p(mask), _p(out), B, H, W, pool, _stream()), "zs_mask_erode_f32")
    return out


def midas_loss(pred, gt, mask, alpha=0.1, inverse_depth=True, need_grad=True, grad_scale=1.0):
    """model/depth/midas_loss.py MidasLoss (image-based reduction, no mask shrinking) -> (loss scalar tensor, d loss / d pred or None).
    pred, gt, mask [B,1,H,W]."""
    pred, gt, mask = pred.contiguous(), gt.contiguous(), mask.float().contiguous()
    _chk(pred, "pred"); _chk(gt, "gt"); _chk(mask, "mask")
    assert pred.dim() == 4 and pred.shape[1] == 1 and pred.shape == gt.shape == mask.shape
    B, _, H, W = pred.shape
    ws = torch.empty((lib.zs_midas_ws_bytes(B, H, W) + 7) // 8, device=pred.device, dtype=torch.float64)
    loss = torch.empty((), device=pred.device, dtype=torch.float32)
    dpred = torch.empty_like(pred) if need_grad else None
    check(lib.zs_midas_loss_f32(_p(pred), _p(gt), _p(mask), B, H, W, float(alpha), int(bool(inverse_depth)), float(grad_scale),
                                _p(ws), _p(loss), _p(dpred), _stream()), "zs_midas_loss_f32")
    return loss, dpred


def adamw_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step):
    for t, n in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _chk(t, n)
    check(lib.zs_adamw_f32(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), lr, beta1, beta2, eps, weight_decay, step,
                           _stream()), "zs_adamw_f32")


def adamw_step_multi(params, grads, exp_avgs, exp_avg_sqs, lr, beta1, beta2, eps, weight_decay, step):
    """AdamW for a whole parameter list in one launch (zs_adamw_multi_f32); every tensor fp32 contiguous on one device."""
    rows = []
    for p, g, m, v in zip(params, grads, exp_avgs, exp_avg_sqs):
        for t, n in ((p, "param"), (g, "grad"), (m, "exp_avg"), (v, "exp_avg_sq")):
            _chk(t, n)
        rows.append((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()))
    if not rows:
        return
    table = torch.tensor(rows, dtype=torch.int64).to(params[0].device, non_blocking=False)
    check(lib.zs_adamw_multi_f32(_p(table), len(rows), lr, beta1, beta2, eps, weight_decay, step, _stream()), "zs_adamw_multi_f32")


def adamw_step_multi_dev(params, grads, exp_avgs, exp_avg_sqs, hyper):
    """The same launch with the step-dependent scalars in device memory (`hyper` = 7 fp32: lr, beta1, beta2, eps, weight decay,
    1 - beta1^step, sqrt(1 - beta2^step)) and the tensor table passed by value: capturable in a CUDA graph (zs_adamw_multi_dev_f32)."""
    import ctypes
    flat = []
    for p, g, m, v in zip(params, grads, exp_avgs, exp_avg_sqs):
        for t, n in ((p, "param"), (g, "grad"), (m, "exp_avg"), (v, "exp_avg_sq")):
            _chk(t, n)
        flat += [p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()]
    if not flat:
        return
    _chk(hyper, "hyper")
    if hyper.numel() < 7:
        raise ValueError("adamw_step_multi_dev: hyper holds 7 fp32 values")
    table = (ctypes.c_uint64 * len(flat))(*flat)
    check(lib.zs_adamw_multi_dev_f32(ctypes.cast(table, ctypes.c_void_p), len(flat) // 5, _p(hyper), _stream()), "zs_adamw_multi_dev_f32")


def conv2d_nhwc_dgrad(dy, w_ohwi, in_shape, stride, pad, tc=None):
    """dx [B,H,W,Cin] of conv2d_nhwc given dy [B,OH,OW,Cout]; w_ohwi [Cout,KH,KW,Cin]; pad = (top, bottom, left, right)."""
    _chk(dy, "dy"); _chk(w_ohwi, "w")
    B, H, W, Cin = in_shape
    Cout, KH, KW, _ = w_ohwi.shape
    OH, OW = dy.shape[1], dy.shape[2]
    wd = w_ohwi.permute(3, 1, 2, 0).contiguous().view(Cin, KH * KW * Cout)       # weights only: [ci][(kh,kw,co)]
    dx = torch.empty(B, H, W, Cin, device=dy.device, dtype=torch.float32)
    if (tc if tc is not None else train_tc()) and Cout % 4 == 0 and Cin >= 32 and KH * KW * Cout >= 64:
        check(lib.zs_conv2d_nhwc_dgrad_tc(_p(dy), B, H, W, Cin, _p(PackedWeight(wd).get()), _p(dx), Cout, KH, KW, stride, pad[0], pad[2],
                                          OH, OW, PRECISIONS[TRAIN_PRECISION], _stream()), "zs_conv2d_nhwc_dgrad_tc")
        return dx
    check(lib.zs_conv2d_nhwc_dgrad_f32(_p(dy), B, H, W, Cin, _p(wd), _p(dx), Cout, KH, KW, stride, pad[0], pad[2], OH, OW, _stream()),
          "zs_conv2d_nhwc_dgrad_f32")
    return dx


def conv2d_nhwc_wgrad(x, dy, kh, kw, stride, pad, out=None, accumulate=False, tc=None):
    """dw [Cout,KH,KW,Cin] (+)= wgrad of conv2d_nhwc."""
    _chk(x, "x"); _chk(dy, "dy")
    B, H, W, Cin = x.shape
    OH, OW, Cout = dy.shape[1], dy.shape[2], dy.shape[3]
    if out is None:
        out = torch.empty(Cout, kh, kw, Cin, device=x.device, dtype=torch.float32)
        accumulate = False
    _chk(out, "out")
    if (tc if tc is not None else train_tc()) and Cin % 8 == 0 and Cout >= 32:
        check(lib.zs_conv2d_nhwc_wgrad_tc(_p(x), B, H, W, Cin, _p(dy), _p(out), Cout, kh, kw, stride, pad[0], pad[2], OH, OW,
                                          int(accumulate), PRECISIONS[TRAIN_PRECISION], TN_LAYOUT, _stream()), "zs_conv2d_nhwc_wgrad_tc")
        return out
    check(lib.zs_conv2d_nhwc_wgrad_f32(_p(x), B, H, W, Cin, _p(dy), _p(out), Cout, kh, kw, stride, pad[0], pad[2], OH, OW,
                                       int(accumulate), _stream()), "zs_conv2d_nhwc_wgrad_f32")
    return out


def bn_stats(x2d, eps):
    """Per-channel batch statistics of x [M,C] -> mean, biased var, rstd."""
    _chk(x2d, "x")
    M, C = x2d.shape
    ws = torch.empty(2 * C, device=x2d.device, dtype=torch.float64)
    mean, var, rstd = (torch.empty(C, device=x2d.device, dtype=torch.float32) for _ in range(3))
    check(lib.zs_bn_stats_f32(_p(x2d), M, C, eps, _p(ws), _p(mean), _p(var), _p(rstd), _stream()), "zs_bn_stats_f32")
    return mean, var, rstd


def bn_bwd(dy2d, x2d, mean, rstd, gamma, dgamma, dbeta):
    """BatchNorm (batch statistics) backward; dgamma / dbeta are accumulated into."""
    for t, n in ((dy2d, "dy"), (x2d, "x"), (mean, "mean"), (rstd, "rstd"), (gamma, "gamma"), (dgamma, "dgamma"), (dbeta, "dbeta")):
        _chk(t, n)
    M, C = x2d.shape
    ws = torch.empty(2 * C, device=x2d.device, dtype=torch.float64)
    dx = torch.empty_like(x2d)
    check(lib.zs_bn_bwd_f32(_p(dy2d), _p(x2d), _p(mean), _p(rstd), _p(gamma), M, C, _p(ws), _p(dx), _p(dgamma), _p(dbeta), _stream()),
          "zs_bn_bwd_f32")
    return dx


def maxpool3x3s2_bwd_nhwc(x, dy, pad_top, pad_left):
    _chk(x, "x"); _chk(dy, "dy")
    B, H, W, C = x.shape
    dx = torch.empty_like(x)
    check(lib.zs_maxpool3x3s2_bwd_nhwc_f32(_p(x), _p(dy), _p(dx), B, H, W, C, pad_top, pad_left, dy.shape[1], dy.shape[2], _stream()),
          "zs_maxpool3x3s2_bwd_nhwc_f32")
    return dx


def avgpool_bwd_nhwc(dy, H, W):
    _chk(dy, "dy")
    B, C = dy.shape
    dx = torch.empty(B, H, W, C, device=dy.device, dtype=torch.float32)
    check(lib.zs_avgpool_bwd_nhwc_f32(_p(dy), _p(dx), B, H * W, C, _stream()), "zs_avgpool_bwd_nhwc_f32")
    return dx


def coldot(a, b, out=None, accumulate=False):
    """out[N] (+)= sum_m a[m,n] * b[m,n]."""
    assert a.shape == b.shape and a.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, N = a.shape
    if out is None:
        out = torch.empty(N, device=a.device, dtype=torch.float32)
        accumulate = False
    check(lib.zs_coldot_f32(_p(a), a.stride(0), _p(b), b.stride(0), M, N, _p(out), int(accumulate), _stream()), "zs_coldot_f32")
    return out


def layernorm_bwd_generic(dy, x, gamma, eps, dgamma=None, dbeta=None):
    """LayerNorm backward over the last dim of any width; dgamma / dbeta accumulated into when given."""
    _chk(dy, "dy"); _chk(x, "x"); _chk(gamma, "gamma")
    C = x.shape[-1]
    rows = x.numel() // C
    dx = torch.empty_like(x)
    xhat = torch.empty_like(x) if dgamma is not None else None
    check(lib.zs_layernorm_bwd_generic_f32(_p(dy), _p(x), _p(gamma), eps, _p(dx), _p(xhat), rows, C, _stream()),
          "zs_layernorm_bwd_generic_f32")
    if dgamma is not None:
        coldot(dy.view(rows, C), xhat.view(rows, C), out=dgamma, accumulate=True)
        colsum(dy.view(rows, C), out=dbeta, accumulate=True)
    return dx


def groupnorm_bwd_nhwc(dy, x, gamma, groups, eps, dgamma, dbeta):
    for t, n in ((dy, "dy"), (x, "x"), (gamma, "gamma"), (dgamma, "dgamma"), (dbeta, "dbeta")):
        _chk(t, n)
    B, H, W, C = x.shape
    dx = torch.empty_like(x)
    check(lib.zs_groupnorm_bwd_nhwc_f32(_p(dy), _p(x), _p(gamma), _p(dx), _p(dgamma), _p(dbeta), B, H * W, C, groups, eps, _stream()),
          "zs_groupnorm_bwd_nhwc_f32")
    return dx


def bilinear_bwd_nhwc(dy, H, W, align_corners):
    _chk(dy, "dy")
    B, OH, OW, C = dy.shape
    dx = torch.empty(B, H, W, C, device=dy.device, dtype=torch.float32)
    check(lib.zs_bilinear_bwd_nhwc_f32(_p(dy), _p(dx), B, H, W, C, OH, OW, int(align_corners), _stream()), "zs_bilinear_bwd_nhwc_f32")
    return dx


def unproject_normalize_bwd(depth, mask, K, seen_points, scale, dseen):
    """-> ddepth [B,1,H,W], dKinv [B,3,3] (gradient w.r.t. inverse(K))."""
    depth = depth.contiguous(); mask = mask.contiguous().float(); K = K.contiguous(); dseen = dseen.contiguous()
    for t, n in ((depth, "depth"), (mask, "mask"), (K, "K"), (seen_points, "seen_points"), (scale, "scale"), (dseen, "dseen")):
        _chk(t, n)
    B = depth.shape[0]
    H, W = depth.shape[-2:]
    dd = torch.empty(B, 1, H, W, device=depth.device, dtype=torch.float32)
    dk = torch.empty(B, 3, 3, device=depth.device, dtype=torch.float32)
    check(lib.zs_unproject_normalize_bwd_f32(_p(depth), _p(mask), _p(K), _p(seen_points), _p(scale), _p(dseen), _p(dd), _p(dk), B, H, W,
                                             _stream()), "zs_unproject_normalize_bwd_f32")
    return dd, dk


def mean_axis1(x):
    """x [A,M,N] -> [A,N] mean over the middle axis."""
    _chk(x, "x")
    A, M, N = x.shape
    out = torch.empty(A, N, device=x.device, dtype=torch.float32)
    check(lib.zs_mean_axis1_f32(_p(x), _p(out), A, M, N, _stream()), "zs_mean_axis1_f32")
    return out


# ---- input pipeline (csrc/preprocess.cu) --------------------------------------------------------------
def rgba_crop_resize(img, left, top, cw, ch, H, W, xbounds, xcoef, ybounds, ycoef):
    """uint8 RGBA [H0,W0,4] -> uint8 RGBA [H,W,4]: PIL crop (zero outside) + PIL BICUBIC resize of the RGBA image, byte-exact."""
    assert img.is_cuda and img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 4 and img.is_contiguous()
    for t in (xbounds, xcoef, ybounds, ycoef):
        assert t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()
    assert xbounds.shape == (W, 2) and ybounds.shape == (H, 2) and xcoef.shape[0] == W and ycoef.shape[0] == H
    out = torch.empty(H, W, 4, dtype=torch.uint8, device=img.device)
    check(lib.zs_rgba_crop_resize_u8(_p(img), img.shape[0], img.shape[1], int(left), int(top), int(cw), int(ch), _p(out), H, W,
                                     _p(xbounds), _p(xcoef), xcoef.shape[1], _p(ybounds), _p(ycoef), ycoef.shape[1], _stream()),
          "zs_rgba_crop_resize_u8")
    return out


def rgba_composite(img, bgcolor):
    """uint8 RGBA [H,W,4] -> (rgb [3,H,W], mask [1,H,W]) fp32: to_tensor, then `rgb * mask + bgcolor * (1 - mask)` and
    `mask > 0.5` when bgcolor is not None (demo.py:46-52)."""
    assert img.is_cuda and img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 4 and img.is_contiguous()
    H, W = img.shape[:2]
    rgb = torch.empty(3, H, W, device=img.device, dtype=torch.float32)
    mask = torch.empty(1, H, W, device=img.device, dtype=torch.float32)
    check(lib.zs_rgba_composite_f32(_p(img), H, W, int(bgcolor is not None), float(bgcolor or 0.0), _p(rgb), _p(mask), _stream()),
          "zs_rgba_composite_f32")
    return rgb, mask


def erode_square(mask, radius):
    """mask [B,H,W] fp32 -> minimum over the (2 radius + 1)^2 window clipped to the image (cv2.erode 3x3, `radius` iterations)."""
    _chk(mask, "mask")
    assert mask.dim() == 3
    out = torch.empty_like(mask)
    check(lib.zs_erode_square_f32(_p(mask), _p(out), mask.shape[0], mask.shape[1], mask.shape[2], int(radius), _stream()),
          "zs_erode_square_f32")
    return out

