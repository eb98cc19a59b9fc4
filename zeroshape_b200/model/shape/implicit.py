"""Host-side mirror of the reference's `model/shape/implicit.py` (same class name, constructor
arguments, parameter names/shapes and forward contract), with all math in the CUDA library.

    Implicit.forward(latent_depth, latent_semantic, points_3D) -> (logits [B,P], attn [B,P,L])
                                                   (reference: model/shape/implicit.py:251-288)

B200-first restructuring (results identical to the reference's):
  * the latent tokens never attend to the query points (implicit.py:65-71), so everything on the
    latent side -- latent_proj, +pos_embed, block-0 self-attention/MLP over the 197 latents and the
    K/V of both blocks -- is computed ONCE per image (`prepare_latents`) instead of once per grid
    slice (utils/eval_3D.py:37-43 calls the whole module 129 times);
  * the query-point side runs through the chained tcgen05 kernels of csrc/chain_tc.cu (`engine="auto"`: per block one
    LayerNorm + qkv + attention kernel, proj + residual, the fused MLP; then the fused occupancy MLP) or through the chain
    of plain-fp32 kernels (`engine="f32"`, the bit-faithful path used for parity pinning);
  * `grid_occupancy` generates the (N+1)^3 query points on the fly (no [B,N,N,N,3] tensor).

Training (decoder slice): when autograd is enabled and a decoder parameter or the
latents require grad, `forward` routes through `implicit_train.ImplicitTrainFn` -- fp32 forward with saved
activations and the hand-written backward kernels of csrc/train.cu -- so `loss.backward()` fills `param.grad`.
"""
import math
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from ... import ops

SQRT2 = float(np.sqrt(2))


def _sincos_2d(dim, grid, cls_token=True):
    """Fixed 2-D sin-cos position table (reference: utils/pos_embed.py:21-68)."""
    def axis(d, pos):
        omega = 1.0 / 10000 ** (np.arange(d // 2, dtype=np.float32) / (d / 2.0))
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)
    g = np.arange(grid, dtype=np.float32)
    mesh = np.stack(np.meshgrid(g, g), axis=0).reshape(2, 1, grid, grid)
    emb = np.concatenate([axis(dim // 2, mesh[0]), axis(dim // 2, mesh[1])], axis=1)
    return np.concatenate([np.zeros([1, dim]), emb], axis=0) if cls_token else emb


class _Holder(nn.Module):
    """Parameter container; never called."""
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container")


def _block_params(dim, mlp_ratio, norm_layer):
    blk = _Holder()
    blk.norm1 = norm_layer(dim)
    blk.attn = _Holder()
    blk.attn.qkv = nn.Linear(dim, dim * 3, bias=True)
    blk.attn.proj = nn.Linear(dim, dim)
    blk.norm2 = norm_layer(dim)
    blk.mlp = _Holder()
    blk.mlp.fc1 = nn.Linear(dim, int(dim * mlp_ratio))
    blk.mlp.fc2 = nn.Linear(int(dim * mlp_ratio), dim)
    return blk


class Implicit(nn.Module):
    """Implicit occupancy decoder conditioned on depth latents (reference class of the same name,
    model/shape/implicit.py:186-288).  Constructor signature kept."""

    def __init__(self, num_patches, latent_dim=768, semantic=False, n_channels=512, n_blocks_attn=2,
                 n_layers_mlp=6, num_heads=16, posenc_3D=0, mlp_ratio=4.,
                 norm_layer=partial(nn.LayerNorm, eps=1e-6), drop_path=0.1, skip_in=[], pos_perlayer=True):
        super().__init__()
        if semantic:
            raise NotImplementedError("semantic (RGB) branch is disabled in the shipped config (options/shape.yaml:32)")
        if posenc_3D != 0:
            raise NotImplementedError("posenc_3D != 0 is not used by the shipped config (options/shape.yaml:43)")
        if n_layers_mlp <= 0:
            raise NotImplementedError("pred_head variant (mlp_layers=0) is not used by the shipped config")
        self.num_patches, self.pos_perlayer, self.semantic = num_patches, pos_perlayer, semantic
        self.num_heads, self.n_channels, self.skip_in = num_heads, n_channels, list(skip_in)
        self.drop_path = drop_path
        self.point_proj = _Holder()
        self.point_proj.proj = nn.Linear(3, n_channels)
        self.latent_proj = nn.Linear(latent_dim, n_channels, bias=True)
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, n_channels), requires_grad=False)
        self.blocks_attn = nn.ModuleList([_block_params(n_channels, mlp_ratio, norm_layer) for _ in range(n_blocks_attn)])
        self.norm = norm_layer(n_channels)
        dims = [3 + n_channels] + [n_channels] * n_layers_mlp + [1]
        self.impl_mlp = _Holder()
        self.impl_mlp.layers = nn.ModuleList([
            nn.Linear(dims[l] + (dims[0] if l in self.skip_in else 0), dims[l + 1]) for l in range(len(dims) - 1)])
        self.engine = "auto"          # "auto" / "chain" (chained tcgen05 kernels) | "tc" (per-layer tcgen05 GEMMs) | "f32" (FFMA, bit-faithful)
        self.precision = "fp16x3"     # tensor-core operand precision of the chain kernels: "fp16x3" (parity) | "fp16" (one pass)
        # point attention: "qkv" (LayerNorm + qkv + flash-style attention in ONE tcgen05 kernel, q/k/v never reach HBM) |
        # "fused" (qkv by zs_chain_lin_fwd, then one flash-style attention kernel) | "tc" (two grouped launches) | "f32" (FFMA)
        self.attention = "qkv"
        # zs_chain_qkvattn_fwd policy: 8 = probabilities in tensor memory (the faster kernel), every contraction three fp16
        # passes; +1 = k, v single-pass (inside the 5e-4 parity budget of profiles/r2_precision_study.md, no measured speed-up)
        self.fold_point_proj = True   # first block recomputes LinearProj3D(points) in its kernels instead of reading x (needs fuse_proj)
        self.fuse_proj = True      # attention output projection + residual inside the MLP kernel (zs_chain_pmlp_fwd; needs flags & 16)
        self.attn_flags = 24       # zs_chain_qkvattn_fwd: 8 = probabilities in tensor memory, 16 = scores in registers + tile-blocked output
        self.lin_fused = True         # chain engine: LN+qkv and proj+residual on zs_chain_lin_fwd (False: layernorm + zs_gemm_tc_f32)
        self._pw = {}                 # id(nn.Linear) -> ops.PackedWeight (tcgen05 operand images of the weights)
        # query points per pass of the per-layer / chained engines (bounds scratch memory: ~5 KB per point).  One pass over a whole
        # 129^3 grid (2.15 M points, 11 GB) instead of 9 slice-aligned passes: persistent kernels lose the partial last wave and the
        # pipeline fill/drain once per LAUNCH (1950 tiles = 13.2 waves per pass cost 14; a single pass of 16771 tiles = 113.3 costs 114)
        self.point_chunk = 1 << 22
        self.initialize_weights()

    # -- init (reference: implicit.py:235-249) ---------------------------------------------------
    def initialize_weights(self):
        grid = int(self.num_patches ** .5)
        self.pos_embed.data.copy_(torch.from_numpy(_sincos_2d(self.pos_embed.shape[-1], grid)).float().unsqueeze(0))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    # -- helpers ---------------------------------------------------------------------------------
    @staticmethod
    def _ln(x, m):
        return ops.layernorm(x, m.weight, m.bias, m.eps)

    def _wants_grad(self, latent_depth, points_3D):
        if not torch.is_grad_enabled():
            return False
        if points_3D.requires_grad:
            raise NotImplementedError("zeroshape_b200.Implicit: gradients w.r.t. the query points are not implemented "
                                      "(the reference never asks for them: graph_shape.py:160-185 builds them under no_grad)")
        # eval-mode calls with parameters that merely *could* take gradients (the nn.Module default) stay on the inference
        # kernels, like every demo / evaluation call of the reference under torch.no_grad()
        return latent_depth.requires_grad or (self.training and any(p.requires_grad for p in self.parameters()))

    def prepare_latents(self, latent_depth):
        """Per-image latent-side work -> dict with K/V views of both blocks ([B,L,C], row stride 3C)."""
        C = self.n_channels
        # the latent side is a per-image constant (0.02 % of the work): only the per-layer "tc" engine runs it on the tensor
        # cores; the chained engine keeps it on the bit-faithful FFMA kernels so that K / V carry no split-precision error
        tc = self._use_tc() and self.engine == "tc"
        lat = ops.linear(latent_depth.float().contiguous(), self.latent_proj.weight, self.latent_proj.bias, tc=tc)
        kv = []
        nb = len(self.blocks_attn)
        for l, blk in enumerate(self.blocks_attn):
            if self.pos_perlayer or l == 0:
                lat = ops.axpby(lat, 1.0, self.pos_embed.expand_as(lat).contiguous(), 1.0)
            qkv = ops.linear(self._ln(lat, blk.norm1), blk.attn.qkv.weight, blk.attn.qkv.bias, tc=tc)
            kv.append((qkv[..., C:2 * C], qkv[..., 2 * C:]))
            if l == nb - 1:
                break
            att = ops.mha(qkv, self.num_heads, tc=self._use_tc())      # tcgen05, fp16x3: the same 1.5e-6 error as the FFMA kernel
            lat = ops.linear(att, blk.attn.proj.weight, blk.attn.proj.bias, res=lat, tc=tc)
            h = ops.linear(self._ln(lat, blk.norm2), blk.mlp.fc1.weight, blk.mlp.fc1.bias, act=ops.ACT_GELU, tc=tc)
            lat = ops.linear(h, blk.mlp.fc2.weight, blk.mlp.fc2.bias, res=lat, tc=tc)
        return {"kv": kv, "B": latent_depth.shape[0], "L": latent_depth.shape[1]}

    # -- per-layer engines: chain of kernels over a chunk of points --------------------------------
    #    "f32": plain-fp32 FFMA GEMMs (bit-faithful);  "tc": tcgen05 split-bf16 GEMMs (zs_gemm_tc_f32)
    def _lin(self, x2d, mod, tc, act=ops.ACT_NONE, res=None):
        if tc and mod.weight.shape[0] >= 64:
            pw = self._pw.get(id(mod))
            if pw is None or pw.src is not mod.weight:
                pw = self._pw[id(mod)] = ops.PackedWeight(mod.weight)
            return ops.gemm_tc(x2d, pw, mod.bias, res=res, act=act, precision=self.precision)
        return ops.gemm(x2d, mod.weight, mod.bias, res=res, act=act)

    # chained tcgen05 kernels (csrc/chain_tc.cu): weight blobs in MMA consumption order, rebuilt on weight change
    def _chain_ok(self):
        return self.n_channels == 256 and len(self.impl_mlp.layers) == 9 and self.skip_in == [2, 4, 6] and \
            all(b.mlp.fc1.weight.shape == (1024, 256) for b in self.blocks_attn)

    def _chain_blobs(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if getattr(self, "_chain_cache", None) is None or self._chain_cache[0] != key:
            # The LayerNorm affine is folded into the consumer GEMM (W.(g*xhat + b) = (W*g).xhat + W.b), so the loader
            # warps only normalise: fc1' = fc1*gamma2, b1' = b1 + fc1.beta2; same for the `feat` columns of impl_mlp.
            mlp, mlp_b1 = [], []
            lin_blobs = []    # per block: (qkv blob with norm1 folded, qkv bias', proj blob)
            for blk in self.blocks_attn:
                wq = blk.attn.qkv.weight.detach().double()
                g1, be1 = blk.norm1.weight.detach().double(), blk.norm1.bias.detach().double()
                wq_f = (wq * g1[None, :]).float()
                lin_blobs.append((ops.pack_generic(wq_f),
                            (blk.attn.qkv.bias.detach().double() + wq @ be1).float().contiguous(),
                            ops.pack_generic(blk.attn.proj.weight.detach().float()),
                            ops.qkvattn_pack(wq_f)))
            for blk in self.blocks_attn:
                w1, w2 = blk.mlp.fc1.weight.detach().double(), blk.mlp.fc2.weight.detach()
                g2, be2 = blk.norm2.weight.detach().double(), blk.norm2.bias.detach().double()
                mlp_b1.append((blk.mlp.fc1.bias.detach().double() + w1 @ be2).float().contiguous())
                w1 = (w1 * g2[None, :]).float()
                mats = []
                for g in range(4):
                    mats += [w1[256 * g:256 * (g + 1), :], w2[:, 256 * g:256 * (g + 1)]]
                mlp.append(ops.pack_tiles(mats))
            L = [lin.weight.detach().double() for lin in self.impl_mlp.layers]
            gn, bn = self.norm.weight.detach().double(), self.norm.bias.detach().double()
            s = 1.0 / float(np.sqrt(2))
            biases = [lin.bias.detach().double().clone() for lin in self.impl_mlp.layers[:8]]
            biases[0] += L[0][:, 3:259] @ bn
            mats = [torch.cat([L[0][:, 3:259] * gn[None, :], L[0][:, 0:3]], dim=1)]       # K order [feat | xyz]
            for l in range(1, 8):
                if l in self.skip_in:                                                      # cat([h, xyz, feat]) / sqrt(2)
                    biases[l] += (L[l][:, 259:515] @ bn) * s
                    mats += [torch.cat([L[l][:, 259:515] * gn[None, :], L[l][:, 256:259]], dim=1) * s, L[l][:, 0:256] * s]
                else:
                    mats.append(L[l])
            occ = ops.pack_tiles([m.float() for m in mats])
            biases = torch.stack(biases).float().contiguous()
            w8 = self.impl_mlp.layers[8].weight.detach().reshape(-1).contiguous()
            self._chain_cache = (key, mlp, occ, biases, w8, float(self.impl_mlp.layers[8].bias.detach()), mlp_b1, lin_blobs)
        return self._chain_cache

    def _pp_tables(self):
        w, b = self.point_proj.proj.weight, self.point_proj.proj.bias
        key = (w.data_ptr(), w._version, b.data_ptr(), b._version, str(w.device))
        if getattr(self, "_pp_cache", None) is None or self._pp_cache[0] != key:
            self._pp_cache = (key,) + ops.point_proj_tables(w, b)
        return self._pp_cache[1], self._pp_cache[2]

    def _points_chain(self, lat, pts, attn_out=None, tc=False, sigmoid=False):
        """pts [B,P,3] contiguous -> logits [B,P]; optionally fills attn_out [B,P,L]."""
        B, P, _ = pts.shape
        C = self.n_channels
        nb = len(self.blocks_attn)
        chain = tc and self.engine != "tc" and self._chain_ok()
        fold_pp = False
        if chain:
            _, mlp_blobs, occ_blob, occ_biases, w8, b8, mlp_b1, lin_blobs = self._chain_blobs()
            # points mode: the first block recomputes x = LinearProj3D(points) inside its two kernels (no point_proj launch, x is never
            # read before the first block has written it)
            fold_pp = (self.fold_point_proj and attn_out is None and self.attention == "qkv" and self.lin_fused and self.fuse_proj
                       and (self.attn_flags & 24) == 24 and nb >= 1)
            if fold_pp:
                pp, pp_stat = self._pp_tables()
                x = torch.empty(B * P, C, device=pts.device, dtype=torch.float32)
            else:
                x = ops.point_proj(pts.reshape(B * P, 3), self.point_proj.proj.weight, self.point_proj.proj.bias)
        else:
            x = ops.gemm(pts.reshape(B * P, 3), self.point_proj.proj.weight, self.point_proj.proj.bias)   # K=3: FFMA
        for l, blk in enumerate(self.blocks_attn):
            k_lat, v_lat = lat["kv"][l]
            if chain and attn_out is None and self.attention == "qkv" and self.lin_fused:
                # LayerNorm + qkv + attention of a tile in one kernel; then x += proj(a) and the MLP as below
                packs = lat.setdefault("kv_fused", {})
                blocked = bool(self.attn_flags & 16)     # the attention output stays tile-blocked between the two kernels
                a = None if blocked else torch.empty(B * P, C, device=x.device, dtype=torch.float32)
                for b in range(B):
                    if (l, b) not in packs:
                        packs[(l, b)] = ops.attn_pack_fused(k_lat[b], v_lat[b], self.num_heads)
                    first = fold_pp and l == 0
                    pts_b = pts[b] if first else None
                    if first:
                        ab = ops.chain_qkvattn_pts(pts_b, pp, pp_stat, lin_blobs[l][3], lin_blobs[l][1], packs[(l, b)][0], packs[(l, b)][1],
                                                   lat["L"], (C // self.num_heads) ** -0.5, ln_eps=blk.norm1.eps, precision=self.precision,
                                                   flags=self.attn_flags)
                    else:
                        ab = ops.chain_qkvattn(x[b * P:(b + 1) * P], lin_blobs[l][3], lin_blobs[l][1], packs[(l, b)][0], packs[(l, b)][1],
                                               lat["L"], (C // self.num_heads) ** -0.5, ln_eps=blk.norm1.eps, precision=self.precision,
                                               flags=self.attn_flags, out=None if blocked else a[b * P:(b + 1) * P])
                    if blocked and self.fuse_proj:       # x += proj(a) and the MLP in one kernel
                        ops.chain_pmlp(x[b * P:(b + 1) * P], ab, lin_blobs[l][2], blk.attn.proj.bias, blk.norm2.eps, mlp_blobs[l],
                                       mlp_b1[l], blk.mlp.fc2.bias, self.precision, points=pts_b, pp=pp if first else None)
                    elif blocked:
                        xb = x[b * P:(b + 1) * P]
                        ops.chain_lin(ab, lin_blobs[l][2], blk.attn.proj.bias, 1, res=xb, out=xb, precision=self.precision)
                if blocked and self.fuse_proj:
                    continue
                if not blocked:
                    ops.chain_lin(a, lin_blobs[l][2], blk.attn.proj.bias, 1, res=x, out=x, precision=self.precision)   # x += proj(a)
                ops.chain_mlp(x, None, None, blk.norm2.eps, mlp_blobs[l], mlp_b1[l], blk.mlp.fc2.bias, self.precision)
                continue
            if chain and self.lin_fused:    # LayerNorm statistics + qkv GEMM in one launch (norm1 affine folded into the packed weights)
                qkv = ops.chain_lin(x, lin_blobs[l][0], lin_blobs[l][1], 3, do_ln=True, ln_eps=blk.norm1.eps, precision=self.precision)
            else:
                qkv = self._lin(self._ln(x, blk.norm1), blk.attn.qkv, tc)
            if chain and attn_out is None and self.attention in ("fused", "qkv"):
                # flash-style tensor-core attention: scores, softmax and P.V of a tile never leave the SM
                packs = lat.setdefault("kv_fused", {})
                a = torch.empty(B * P, C, device=x.device, dtype=torch.float32)
                for b in range(B):
                    if (l, b) not in packs:
                        packs[(l, b)] = ops.attn_pack_fused(k_lat[b], v_lat[b], self.num_heads)
                    ops.attn_fused(qkv[b * P:(b + 1) * P], packs[(l, b)][0], packs[(l, b)][1], lat["L"],
                                   (C // self.num_heads) ** -0.5, self.precision, out=a[b * P:(b + 1) * P])
            elif chain and attn_out is None and self.attention == "tc":
                # tensor-core attention (two grouped tcgen05 launches per image; K/V operand images cached per image)
                packs = lat.setdefault("kv_tc", {})
                a = torch.empty(B * P, C, device=x.device, dtype=torch.float32)
                for b in range(B):
                    if (l, b) not in packs:
                        packs[(l, b)] = ops.attn_pack_kv(k_lat[b], v_lat[b], self.num_heads)
                    ops.attn_tc(qkv[b * P:(b + 1) * P], packs[(l, b)][0], packs[(l, b)][1], lat["L"],
                                (C // self.num_heads) ** -0.5, self.precision, out=a[b * P:(b + 1) * P])
            else:
                a = ops.point_attention(qkv.view(B, P, 3 * C), k_lat, v_lat, self.num_heads, attn=attn_out,
                                        attn_scale=1.0 / nb, attn_accumulate=(l > 0), tc=False).view(B * P, C)
            del qkv
            if chain:
                if self.lin_fused:
                    ops.chain_lin(a, lin_blobs[l][2], blk.attn.proj.bias, 1, res=x, out=x, precision=self.precision)   # x += proj(a)
                else:
                    x = self._lin(a, blk.attn.proj, tc, res=x)
                ops.chain_mlp(x, None, None, blk.norm2.eps, mlp_blobs[l], mlp_b1[l], blk.mlp.fc2.bias, self.precision)
                continue
            x = self._lin(a, blk.attn.proj, tc, res=x)
            h = self._lin(self._ln(x, blk.norm2), blk.mlp.fc1, tc, act=ops.ACT_GELU)
            x = self._lin(h, blk.mlp.fc2, tc, res=x)
            del h
        if chain:
            out = ops.chain_occ(x, pts.reshape(B * P, 3), None, None, self.norm.eps, occ_blob,
                                occ_biases, w8, b8, sigmoid=sigmoid, precision=self.precision)
            return out.reshape(B, P)
        feat = self._ln(x, self.norm)
        inputs = ops.concat2(pts.reshape(B * P, 3), feat, 1.0)
        h = inputs
        n_layers = len(self.impl_mlp.layers)
        for l, lin in enumerate(self.impl_mlp.layers):
            if l in self.skip_in:
                h = ops.concat2(h, inputs, SQRT2)
            h = self._lin(h, lin, tc, act=ops.ACT_SOFTPLUS100 if l < n_layers - 1 else ops.ACT_NONE)
        return h.reshape(B, P)

    def _points_f32(self, lat, pts, attn_out=None, sigmoid=False):
        """Logits [B,P] of a chunk of points; with `sigmoid` the occupancy of compute_level_grid (eval_3D.py:46): fused into
        the occupancy-MLP kernel's epilogue on the chained engine, one elementwise launch otherwise."""
        tc = self._use_tc()
        if sigmoid and not (tc and attn_out is None and self.engine != "tc" and self._chain_ok()):
            return ops.axpby(self._points_chain(lat, pts, attn_out, tc=tc), 1.0, act=ops.ACT_SIGMOID)
        return self._points_chain(lat, pts, attn_out, tc=tc, sigmoid=sigmoid)

    def _use_tc(self):
        if self.engine == "f32":
            return False
        ok = ops.device_cc() == 100
        if self.engine == "tc" and not ok:
            raise RuntimeError("engine='tc' needs an sm_100 device")
        return ok

    # -- public API --------------------------------------------------------------------------------
    def forward(self, latent_depth, latent_semantic, points_3D, need_attn=True):
        assert latent_semantic is None
        if self._wants_grad(latent_depth, points_3D):
            # training call (graph_shape.py:185): differentiable logits; the attention map the reference also returns there
            # is never used by the training loop (graph_shape.py:185, shape_engine.py:248-277) and is not produced
            from .implicit_train import ImplicitTrainFn
            params = [p for p in self.parameters()]
            return ImplicitTrainFn.apply(self, latent_depth, points_3D, *params), None
        with torch.no_grad():
            pts = points_3D.float().contiguous()
            B, P, _ = pts.shape
            lat = self.prepare_latents(latent_depth)
            L = lat["L"]
            attn = torch.empty(B, P, L, device=pts.device, dtype=torch.float32) if need_attn else None
            logits = torch.empty(B, P, device=pts.device, dtype=torch.float32)
            for s in range(0, P, self.point_chunk):
                e = min(P, s + self.point_chunk)
                pc = pts[:, s:e].contiguous()
                ac = torch.empty(B, e - s, L, device=pts.device, dtype=torch.float32) if need_attn else None
                logits[:, s:e] = self._points_f32(lat, pc, ac)
                if need_attn:
                    attn[:, s:e] = ac
            return logits, attn

    @torch.no_grad()
    def grid_occupancy(self, latent_depth, n, rmin, rmax, x0=0, x1=None, sigmoid=True, lat=None):
        """Occupancy over x-slices [x0,x1) of the dense (n)^3 grid of utils/eval_3D.py:10-20:
        returns [B, x1-x0, n, n] (sigmoid(logit) like compute_level_grid, eval_3D.py:46)."""
        x1 = n if x1 is None else x1
        lat = lat if lat is not None else self.prepare_latents(latent_depth)
        B = lat["B"]
        out = torch.empty(B, x1 - x0, n, n, device=latent_depth.device, dtype=torch.float32)
        slices = max(1, self.point_chunk // (n * n * max(B, 1)))
        for s in range(x0, x1, slices):
            e = min(x1, s + slices)
            pts = ops.dense_grid(n, rmin, rmax, s, e, latent_depth.device).view(1, -1, 3).expand(B, -1, -1).contiguous()
            out[:, s - x0:e - x0] = self._points_f32(lat, pts, sigmoid=sigmoid).view(B, e - s, n, n)
        return out
