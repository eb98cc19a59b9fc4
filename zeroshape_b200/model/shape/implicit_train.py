"""Training step of the implicit decoder: forward with saved activations and hand-written backward.

Replaces torch autograd over the reference's `Implicit.forward` (model/shape/implicit.py:251-288) for the training call
`self.impl_network(var.latent_depth, var.latent_semantic, var.gt_points_cam)` (model/compute_graph/graph_shape.py:185):
every layer's forward runs on the fp32 kernels of the library and every gradient on the kernels of csrc/train.cu
(zs_gemm_f32 for dX = dY W, zs_gemm_tn_f32 for dW = dY^T X, zs_colsum_f32 for bias gradients, zs_layernorm_bwd_f32,
zs_act_bwd_f32, zs_point_attention_bwd_f32, zs_mha_bwd_f32).  `ImplicitTrainFn` exposes it as a torch.autograd.Function so the
reference's `loss.backward(); optim.step()` (model/shape_engine.py:268-277) keeps working: gradients arrive in `param.grad`
of the `nn.Parameter`s with the reference's names, and `dlatent_depth` flows on to whatever produced the latents.

Structure of the computation (B images, L = 197 latents, P points per image, C = 256):
  latents : lat0 = latent_proj(z) + pos_embed ; block 0 in full (self-attention over the latents + MLP) ; block 1 only as
            the source of keys / values (the last block does not update the latents, implicit.py:59-63)
  points  : x = point_proj(p) ; per block x += proj(attn(LN1 x | K_lat, V_lat, self)) ; x += fc2(GELU(fc1(LN2 x)))
  head    : MLPBlocks([p, LN(x)]) with skips at 2, 4, 6 and Softplus(beta = 100)
DropPath (timm, per image, implicit.py:103-104) is applied to both residual branches of a block in train mode.
"""
import numpy as np
import torch

from ... import ops

SQRT2 = float(np.sqrt(2))


def _lin_fwd(x2d, lin, act=ops.ACT_NONE, res=None):
    return ops.train_linear(x2d, lin.weight, lin.bias, res=res, act=act)


class _Grads:
    """name -> accumulated gradient tensor (zero-initialised lazily), keyed by the nn.Parameter object."""

    def __init__(self):
        self.g = {}

    def buf(self, p):
        if id(p) not in self.g:
            self.g[id(p)] = torch.zeros_like(p, dtype=torch.float32)
        return self.g[id(p)]

    def get(self, p):
        return self.g.get(id(p))


def _lin_bwd(dy, x2d, lin, grads, need_dx=True):
    """dy [M,N], x2d [M,K] (row-strided views allowed) -> dx [M,K]; accumulates dW, db."""
    ops.gemm_tn(dy, x2d, out=grads.buf(lin.weight), accumulate=True)
    if lin.bias is not None:
        ops.colsum(dy, out=grads.buf(lin.bias), accumulate=True)
    if not need_dx:
        return None
    # dX = dY @ W = gemm(dY, (W^T)): a weights-only transpose (+ tensor-core packing), then the forward GEMM kernel
    return ops.train_dgrad(dy, lin.weight)


def _ln_bwd(dy, x, norm, grads):
    return ops.layernorm_bwd(dy.contiguous(), x, norm.weight, norm.eps, grads.buf(norm.weight), grads.buf(norm.bias))


def _droppath_scales(net, B, device):
    """timm DropPath: per-image keep mask / keep_prob; one draw per residual branch.  None when inactive."""
    if not net.training or net.drop_path <= 0.0:
        return None
    keep = 1.0 - net.drop_path
    nb = len(net.blocks_attn)
    m = (torch.rand(nb, 2, B) < keep).float() / keep
    return m.tolist()


def _scaled_residual(x, branch, scales, B):
    """x [B*R, C] + scales[b] * branch, per image (DropPath); scales None -> plain add."""
    if scales is None:
        return ops.axpby(x, 1.0, branch, 1.0)
    R = x.shape[0] // B
    out = torch.empty_like(x)
    for b in range(B):
        out[b * R:(b + 1) * R] = ops.axpby(x[b * R:(b + 1) * R].contiguous(), 1.0, branch[b * R:(b + 1) * R].contiguous(), scales[b])
    return out


def _scale_rows(t, scales, B):
    if scales is None:
        return t
    R = t.shape[0] // B
    out = torch.empty_like(t)
    for b in range(B):
        out[b * R:(b + 1) * R] = ops.axpby(t[b * R:(b + 1) * R].contiguous(), scales[b])
    return out


def train_forward(net, latent_depth, points):
    """-> (logits [B,P], tape).  fp32 kernels, every tensor the backward needs is kept on the tape."""
    B, L, _ = latent_depth.shape
    P = points.shape[1]
    C, H = net.n_channels, net.num_heads
    nb = len(net.blocks_attn)
    z = latent_depth.detach().float().contiguous()
    pts = points.detach().float().contiguous().view(B * P, 3)
    dp = _droppath_scales(net, B, z.device)
    T = {"B": B, "L": L, "P": P, "z": z.view(B * L, -1), "pts": pts, "dp": dp, "blocks": []}
    # ---- latent side ----
    lat = ops.axpby(_lin_fwd(T["z"], net.latent_proj), 1.0, net.pos_embed.detach().expand(B, -1, -1).contiguous().view(B * L, C), 1.0)
    x = ops.gemm(pts, net.point_proj.proj.weight, net.point_proj.proj.bias)
    for l, blk in enumerate(net.blocks_attn):
        last = l == nb - 1
        if l > 0 and net.pos_perlayer:
            lat = ops.axpby(lat, 1.0, net.pos_embed.detach().expand(B, -1, -1).contiguous().view(B * L, C), 1.0)
        t = {"lat_in": lat, "x_in": x}
        s_attn = dp[l][0] if dp else None
        s_mlp = dp[l][1] if dp else None
        # latents: keys / values (and, except in the last block, their own update)
        hl = ops.layernorm(lat, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
        qkv_l = _lin_fwd(hl, blk.attn.qkv).view(B, L, 3 * C)
        k_lat, v_lat = qkv_l[..., C:2 * C], qkv_l[..., 2 * C:]
        t.update(hl=hl, qkv_l=qkv_l)
        # points
        h = ops.layernorm(x, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
        qkv_p = _lin_fwd(h, blk.attn.qkv).view(B, P, 3 * C)
        a = ops.point_attention(qkv_p, k_lat, v_lat, H).view(B * P, C)
        x_mid = _scaled_residual(x, _lin_fwd(a, blk.attn.proj), s_attn, B)
        h2 = ops.layernorm(x_mid, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
        u = _lin_fwd(h2, blk.mlp.fc1)
        g = ops.axpby(u, 1.0, act=ops.ACT_GELU)
        x = _scaled_residual(x_mid, _lin_fwd(g, blk.mlp.fc2), s_mlp, B)
        t.update(h=h, qkv_p=qkv_p, a=a, x_mid=x_mid, h2=h2, u=u, g=g)
        if not last:
            al = ops.mha(qkv_l, H).view(B * L, C)
            lat_mid = _scaled_residual(lat, _lin_fwd(al, blk.attn.proj), s_attn, B)
            hl2 = ops.layernorm(lat_mid, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
            ul = _lin_fwd(hl2, blk.mlp.fc1)
            gl = ops.axpby(ul, 1.0, act=ops.ACT_GELU)
            lat = _scaled_residual(lat_mid, _lin_fwd(gl, blk.mlp.fc2), s_mlp, B)
            t.update(al=al, lat_mid=lat_mid, hl2=hl2, ul=ul, gl=gl)
        T["blocks"].append(t)
    T["x_final"] = x
    feat = ops.layernorm(x, net.norm.weight, net.norm.bias, net.norm.eps)
    inputs = ops.concat2(pts, feat, 1.0)
    hcur = inputs
    T["inputs"] = inputs
    T["mlp_in"], T["mlp_z"] = [], []
    n_layers = len(net.impl_mlp.layers)
    for l, lin in enumerate(net.impl_mlp.layers):
        if l in net.skip_in:
            hcur = ops.concat2(hcur, inputs, SQRT2)
        zl = _lin_fwd(hcur, lin)
        T["mlp_in"].append(hcur)
        T["mlp_z"].append(zl)
        hcur = ops.axpby(zl, 1.0, act=ops.ACT_SOFTPLUS100) if l < n_layers - 1 else zl
    return hcur.view(B, P), T


def train_backward(net, T, dlogits):
    """dlogits [B,P] -> (_Grads over the decoder parameters, dlatent_depth [B,L,latent_dim])."""
    B, L, P = T["B"], T["L"], T["P"]
    C, H = net.n_channels, net.num_heads
    nb = len(net.blocks_attn)
    G = _Grads()
    dp = T["dp"]
    # ---- occupancy MLP ----
    n_layers = len(net.impl_mlp.layers)
    dh = dlogits.detach().float().contiguous().view(B * P, 1)
    dinputs = torch.zeros_like(T["inputs"])
    for l in range(n_layers - 1, -1, -1):
        lin = net.impl_mlp.layers[l]
        dz = dh if l == n_layers - 1 else ops.act_bwd(dh.contiguous(), T["mlp_z"][l], ops.ACT_SOFTPLUS100)
        din = _lin_bwd(dz, T["mlp_in"][l], lin, G)
        if l in net.skip_in:
            k = din.shape[1] - dinputs.shape[1]
            dinputs = ops.axpby(dinputs, 1.0, din[:, k:].contiguous(), 1.0 / SQRT2)
            dh = ops.axpby(din[:, :k].contiguous(), 1.0 / SQRT2)
        elif l == 0:
            dinputs = ops.axpby(dinputs, 1.0, din, 1.0)
        else:
            dh = din
    dx = _ln_bwd(dinputs[:, 3:], T["x_final"], net.norm, G)
    # ---- attention blocks, last to first; latent-side key / value gradients are collected per block ----
    dlat = None                       # gradient w.r.t. the latents ENTERING the block that was processed last
    for l in range(nb - 1, -1, -1):
        blk, t = net.blocks_attn[l], T["blocks"][l]
        last = l == nb - 1
        s_attn = dp[l][0] if dp else None
        s_mlp = dp[l][1] if dp else None
        # points: MLP branch
        dbr = _scale_rows(dx, s_mlp, B)
        dg = _lin_bwd(dbr, t["g"], blk.mlp.fc2, G)
        du = ops.act_bwd(dg, t["u"], ops.ACT_GELU)
        dh2 = _lin_bwd(du, t["h2"], blk.mlp.fc1, G)
        dx_mid = ops.axpby(dx, 1.0, _ln_bwd(dh2, t["x_mid"], blk.norm2, G), 1.0)
        # points: attention branch
        dbr = _scale_rows(dx_mid, s_attn, B)
        da = _lin_bwd(dbr, t["a"], blk.attn.proj, G)
        qkv_l = t["qkv_l"]
        dqkv_p, dk_lat, dv_lat = ops.point_attention_bwd(t["qkv_p"], qkv_l[..., C:2 * C], qkv_l[..., 2 * C:], t["a"].view(B, P, C),
                                                         da.view(B, P, C).contiguous(), H)
        dhp = _lin_bwd(dqkv_p.view(B * P, 3 * C), t["h"], blk.attn.qkv, G)
        dx = ops.axpby(dx_mid, 1.0, _ln_bwd(dhp, t["x_in"], blk.norm1, G), 1.0)
        # latents
        dqkv_l = torch.zeros(B, L, 3 * C, device=dx.device, dtype=torch.float32)
        dqkv_l[..., C:2 * C] = dk_lat
        dqkv_l[..., 2 * C:] = dv_lat
        if last:
            dlat_out = None           # the last block does not update the latents
        else:
            dlat_out = dlat           # gradient w.r.t. this block's latent output (from the blocks after it)
        if dlat_out is not None:
            dbr = _scale_rows(dlat_out, s_mlp, B)
            dgl = _lin_bwd(dbr, t["gl"], blk.mlp.fc2, G)
            dul = ops.act_bwd(dgl, t["ul"], ops.ACT_GELU)
            dhl2 = _lin_bwd(dul, t["hl2"], blk.mlp.fc1, G)
            dlat_mid = ops.axpby(dlat_out, 1.0, _ln_bwd(dhl2, t["lat_mid"], blk.norm2, G), 1.0)
            dbr = _scale_rows(dlat_mid, s_attn, B)
            dal = _lin_bwd(dbr, t["al"], blk.attn.proj, G)
            dqkv_l = ops.axpby(dqkv_l, 1.0, ops.mha_bwd(qkv_l.contiguous(), dal.view(B, L, C).contiguous(), H), 1.0)
            dhl = _lin_bwd(dqkv_l.view(B * L, 3 * C), t["hl"], blk.attn.qkv, G)
            dlat = ops.axpby(dlat_mid, 1.0, _ln_bwd(dhl, t["lat_in"], blk.norm1, G), 1.0)
        else:
            dhl = _lin_bwd(dqkv_l.view(B * L, 3 * C), t["hl"], blk.attn.qkv, G)
            dlat = _ln_bwd(dhl, t["lat_in"], blk.norm1, G)
    # ---- input projections ----
    _lin_bwd(dx, T["pts"], net.point_proj.proj, G, need_dx=False)
    dz = _lin_bwd(dlat, T["z"], net.latent_proj, G)          # pos_embed is a fixed buffer (requires_grad=False)
    return G, dz.view(B, L, -1)


class ImplicitTrainFn(torch.autograd.Function):
    """logits = ImplicitTrainFn.apply(net, latent_depth, points, *decoder_parameters)"""

    @staticmethod
    def forward(ctx, net, latent_depth, points, *params):
        with torch.no_grad():
            logits, tape = train_forward(net, latent_depth, points)
        ctx.net, ctx.tape, ctx.params = net, tape, params
        ctx.need_dz = latent_depth.requires_grad
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        with torch.no_grad():
            G, dz = train_backward(ctx.net, ctx.tape, dlogits)
        grads = tuple((G.get(p) if p.requires_grad else None) for p in ctx.params)
        ctx.tape = None
        return (None, dz if ctx.need_dz else None, None) + grads


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics on the zs_adamw_multi_f32 kernel (model/shape_engine.py:75-136: AdamW, betas (0.9, 0.95),
    four parameter groups with lr / lr_ft and zero weight decay for the no-decay sets).

    A real `torch.optim.Optimizer`: `param_groups` (per-group lr / betas / eps / weight_decay, so CosineAnnealingLR and the
    reference's group construction work), per-parameter `state` {step, exp_avg, exp_avg_sq} (a parameter whose gradient is
    None in early steps gets its own bias correction), `state_dict` / `load_state_dict` (util.save_checkpoint serialises every
    `optim*` attribute).  One launch per (group, step count) bucket.

    `capturable=True`: the step-dependent scalars (lr, bias corrections) live in device memory and the tensor table travels in
    the kernel parameters (zs_adamw_multi_dev_f32), so `step()` can be captured in a CUDA graph (zeroshape_b200/graphed.py).
    Under capture `step()` records the launches without advancing the host-side step counters; every replay is preceded by
    `prepare_replay()` (counters + 1, scalars refreshed with a stream-ordered copy) and followed by `mark_updated()`."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, capturable=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.capturable = bool(capturable)
        self._hyper = {}          # (group index, bucket ordinal) -> device tensor [8]
        self._captured = []       # [(group, [params], hyper tensor)] of the captured step

    @staticmethod
    def _hyper_values(group, step):
        import numpy as np
        b1, b2 = np.float32(group["betas"][0]), np.float32(group["betas"][1])
        bc1 = np.float32(1.0) - np.power(b1, np.float32(step))
        bc2 = np.sqrt(np.float32(1.0) - np.power(b2, np.float32(step)))
        return torch.tensor([float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                             float(bc1), float(bc2), 0.0], dtype=torch.float32)

    def _hyper_buf(self, key, device):
        if key not in self._hyper:
            self._hyper[key] = torch.zeros(8, device=device, dtype=torch.float32)
        return self._hyper[key]

    def _ensure_state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        capturing = self.capturable and torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        if capturing:
            self._captured = []
        for gi, group in enumerate(self.param_groups):
            buckets = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self._ensure_state(p)
                if not capturing:
                    st["step"] = int(st["step"]) + 1
                buckets.setdefault(int(st["step"]) + (1 if capturing else 0), []).append(p)
            for k, (step, ps) in enumerate(buckets.items()):
                grads = [p.grad.contiguous() for p in ps]             # kept alive until the launch is enqueued
                if self.capturable:
                    hyper = self._hyper_buf((gi, k), ps[0].device)
                    if capturing:
                        # (a non-contiguous gradient would be copied into a buffer of the graph's pool: fine, it is re-made per replay)
                        self._captured.append((group, ps, hyper))
                    else:
                        # pageable source: staged by the runtime before the call returns, ordered on the stream, no synchronisation
                        hyper.copy_(self._hyper_values(group, step), non_blocking=True)
                    ops.adamw_step_multi_dev([p.detach() for p in ps], grads, [self.state[p]["exp_avg"] for p in ps],
                                             [self.state[p]["exp_avg_sq"] for p in ps], hyper)
                else:
                    ops.adamw_step_multi([p.detach() for p in ps], grads, [self.state[p]["exp_avg"] for p in ps],
                                         [self.state[p]["exp_avg_sq"] for p in ps], float(group["lr"]), group["betas"][0],
                                         group["betas"][1], group["eps"], group["weight_decay"], step)
                if not capturing:
                    for p in ps:
                        # the kernel wrote through the raw pointer: bump the version counter so that the packed-weight caches
                        # (keyed on data_ptr / _version) re-pack, exactly as after a torch optimizer step
                        torch.autograd.graph.increment_version(p)
        return loss

    def prepare_replay(self):
        """Before each replay of a graph that holds a captured `step()`: advance the step counters of the captured parameters
        and refresh their device-side scalars (current group lr, bias corrections)."""
        if not self._captured:
            raise RuntimeError("FusedAdamW.prepare_replay: no captured step (construct with capturable=True and capture step())")
        for group, ps, hyper in self._captured:
            step = None
            for p in ps:
                st = self.state[p]
                st["step"] = int(st["step"]) + 1
                step = st["step"]
            hyper.copy_(self._hyper_values(group, step), non_blocking=True)

    def mark_updated(self):
        """After a replay: the parameters changed behind autograd's back -> bump their version counters (packed-weight caches)."""
        for _, ps, _ in self._captured:
            for p in ps:
                torch.autograd.graph.increment_version(p)
