"""ORACLE (test infrastructure only -- never imported by the product): CPU restatement of the reference's MiDaS
scale-and-shift-invariant depth loss, model/depth/midas_loss.py (used by utils/loss.py:30-34 `Loss.depth_loss`, i.e. the
`loss_weight.depth` term of options/depth.yaml), and its analytic gradient.

  midas_loss(pred, gt, mask, alpha, inverse_depth)      forward, same op sequence as the reference:
      masked_shift_and_scale :33-62 (nanmedian / mean-absolute-deviation alignment), masked_l1_loss :6-9,
      compute_scale_and_shift :11-30, gradient_loss :88-107 over 4 scales (GradientMatchingTerm :120-143),
      image-based reduction :75-84 (MidasLoss default), total = ssi + alpha * reg :160-185 (mask_shrink = False)
  midas_loss_grad(...)                                  closed-form d loss / d pred (what torch autograd produces for the
      forward above); the CUDA kernels implement exactly these formulas.

Pinned by tests/test_oracle_midas.py against the REAL reference module: forward values and autograd gradients stored in
tests/golden/midas.npz by tests/golden/make_golden_midas.py (run in the build container, where /root/reference exists).
"""
import torch


def _lower_median(vals):
    """torch.nanmedian semantics on the valid values: the element of rank (n - 1) // 2; (value, position in `vals`)."""
    n = vals.numel()
    if n == 0:
        return None, -1
    order = torch.argsort(vals, stable=True)
    k = order[(n - 1) // 2]
    return vals[k], int(k)


def _align(x, valid):
    """masked_shift_and_scale for one map: (x - median_valid) / (mean_abs_dev + 1e-6); -> aligned, t, s, flat index of the median."""
    xv = x[valid]
    n = int(valid.sum())
    t, k = _lower_median(xv)
    if t is None:
        t = x.new_zeros(())
        idx = -1
    else:
        idx = int(torch.nonzero(valid.flatten())[k])
    d = (x - t).abs()
    d = torch.where(valid, d, torch.zeros_like(d))
    s = d.sum() / (n + 1)
    return (x - t) / (s + 1e-6), t, s, idx


def _scale_shift(p, t, m):
    a00, a01, a11 = (m * p * p).sum(), (m * p).sum(), m.sum()
    b0, b1 = (m * p * t).sum(), (m * t).sum()
    det = a00 * a11 - a01 * a01
    if float(det) == 0.0:
        z = p.new_zeros(())
        return z, z, (a00, a01, a11, b0, b1, det)
    return (a11 * b0 - a01 * b1) / (det + 1e-6), (-a01 * b0 + a00 * b1) / (det + 1e-6), (a00, a01, a11, b0, b1, det)


def erode_mask(mask, pool=4):
    """MidasLoss.erode_mask (midas_loss.py:158-167): valid only where a whole pool x pool block of the raw mask is valid."""
    m = 1 - mask.float()
    h, w = m.shape[2], m.shape[3]
    m = torch.nn.functional.max_pool2d(m, kernel_size=pool)
    m = torch.nn.functional.interpolate(m, (h, w), mode="nearest")
    return (m == 0).float()


def midas_loss(pred, gt, mask, alpha=0.1, inverse_depth=True, scales=4):
    """pred, gt, mask [B,1,H,W] -> scalar (fp32 torch tensor).  mask_raw > 0.5 is the valid set (mask_shrink False)."""
    B = pred.shape[0]
    valid = mask > 0.5
    num = pred.new_zeros(())
    for b in range(B):
        ap, _, _, _ = _align(pred[b, 0], valid[b, 0])
        ag, _, _, _ = _align(gt[b, 0], valid[b, 0])
        e = (ap - ag).abs()
        num = num + torch.where(valid[b, 0], e, torch.zeros_like(e)).sum()
    ssi = num / (valid.sum() + 1.e-6)
    if alpha <= 0:
        return ssi
    reg = pred.new_zeros(())
    for b in range(B):
        m = valid[b, 0].float()
        p = 1 / (pred[b, 0] + 1e-6) if inverse_depth else pred[b, 0]
        t = 1 / (gt[b, 0] + 1e-6) if inverse_depth else gt[b, 0]
        x0, x1, _ = _scale_shift(p, t, m)
        q = x0 * p + x1
        for s in range(scales):
            st = 2 ** s
            qs, ts, ms = q[::st, ::st], t[::st, ::st], m[::st, ::st]
            M = ms.sum()
            diff = ms * (qs - ts)
            gx = (diff[:, 1:] - diff[:, :-1]).abs() * ms[:, 1:] * ms[:, :-1]
            gy = (diff[1:, :] - diff[:-1, :]).abs() * ms[1:, :] * ms[:-1, :]
            il = gx.sum() + gy.sum()
            if float(M) != 0.0:
                il = il / M
            reg = reg + il / B
    return ssi + alpha * reg


def midas_loss_grad(pred, gt, mask, alpha=0.1, inverse_depth=True, scales=4):
    """Closed-form d midas_loss / d pred, [B,1,H,W] (float64 arithmetic on the fp32 inputs)."""
    pred, gt = pred.double(), gt.double()
    B, _, H, W = pred.shape
    valid = mask > 0.5
    N = float(valid.sum()) + 1.e-6
    g = torch.zeros_like(pred)
    for b in range(B):
        v = valid[b, 0]
        n = int(v.sum())
        P, T = pred[b, 0], gt[b, 0]
        ap, tp, sp, midx = _align(P, v)
        ag, _, _, _ = _align(T, v)
        c = 1.0 / (sp + 1e-6)
        e = torch.sign(ap - ag) * v
        sg = torch.sign(P - tp) * v
        Gs = -(c * c) / N * (e * (P - tp)).sum()
        Gt = -c / N * e.sum() - Gs * sg.sum() / (n + 1)
        gb = e * c / N + Gs * sg / (n + 1)
        if midx >= 0:
            gb.view(-1)[midx] += Gt
        if alpha > 0:
            m = v.double()
            p = 1 / (P + 1e-6) if inverse_depth else P
            t = 1 / (T + 1e-6) if inverse_depth else T
            x0, x1, (a00, a01, a11, b0, b1, det) = _scale_shift(p, t, m)
            q = x0 * p + x1
            gq = torch.zeros_like(q)
            for s in range(scales):
                st = 2 ** s
                qs, ts, ms = q[::st, ::st], t[::st, ::st], m[::st, ::st]
                M = float(ms.sum())
                w = (1.0 / M if M != 0.0 else 1.0) / B
                diff = ms * (qs - ts)
                dx = torch.sign(diff[:, 1:] - diff[:, :-1]) * ms[:, 1:] * ms[:, :-1] * w
                dy = torch.sign(diff[1:, :] - diff[:-1, :]) * ms[1:, :] * ms[:-1, :] * w
                gs = torch.zeros_like(qs)
                gs[:, 1:] += dx * ms[:, 1:]
                gs[:, :-1] -= dx * ms[:, :-1]
                gs[1:, :] += dy * ms[1:, :]
                gs[:-1, :] -= dy * ms[:-1, :]
                gq[::st, ::st] += gs
            if float(det) != 0.0:
                D = det + 1e-6
                Gx0, Gx1 = (gq * p).sum(), gq.sum()
                dx0_a00, dx0_a01, dx0_b0 = -x0 * a11 / D, (-b1 + 2 * a01 * x0) / D, a11 / D
                dx1_a00, dx1_a01, dx1_b0 = (b1 - x1 * a11) / D, (-b0 + 2 * a01 * x1) / D, -a01 / D
                Ga00 = Gx0 * dx0_a00 + Gx1 * dx1_a00
                Ga01 = Gx0 * dx0_a01 + Gx1 * dx1_a01
                Gb0 = Gx0 * dx0_b0 + Gx1 * dx1_b0
                gp = gq * x0 + m * (2 * p * Ga00 + Ga01 + t * Gb0)
            else:
                gp = torch.zeros_like(q)
            gb = gb + alpha * (gp * (-(p * p)) if inverse_depth else gp)
        g[b, 0] = gb
    return g


def depth_metrics(prediction, target, mask, thresholds=(1.25, 1.25 ** 2, 1.25 ** 3), depth_cap=None, prediction_type="depth"):
    """Restatement of DepthMetric.compute_metrics (utils/eval_depth.py:41-110) -> ({key: [B]}, aligned depth [B,1,H,W])."""
    p, t = prediction.float().squeeze(1), target.float().squeeze(1)
    m = (mask.float().squeeze(1) > 0.5)
    mf = m.float()
    pd = torch.where(m, 1.0 / (p + 1.e-6) if prediction_type == "depth" else p, torch.zeros_like(p))
    td = torch.where(m, 1.0 / t, torch.zeros_like(t))
    a00, a01, a11 = (mf * pd * pd).sum((1, 2)), (mf * pd).sum((1, 2)), mf.sum((1, 2))
    b0, b1 = (mf * pd * td).sum((1, 2)), (mf * td).sum((1, 2))
    det = a00 * a11 - a01 * a01
    ok = det > 0
    safe = torch.where(ok, det, torch.ones_like(det))
    x0 = torch.where(ok, (a11 * b0 - a01 * b1) / safe, torch.zeros_like(det))
    x1 = torch.where(ok, (-a01 * b0 + a00 * b1) / safe, torch.zeros_like(det))
    al = x0.view(-1, 1, 1) * pd + x1.view(-1, 1, 1)
    if depth_cap is not None:
        al = torch.clamp(al, min=1.0 / depth_cap)
    d = 1.0 / al
    out = {}
    n = mf.sum((1, 2))
    ratio = torch.where(m, torch.max(d / t, t / d), torch.zeros_like(d))
    for th in thresholds:
        out["d>{}".format(th)] = ((ratio > th).float() * mf).sum((1, 2)) / n
    diff = torch.where(m, d - t, torch.zeros_like(d))
    out["rmse"] = torch.sqrt((diff ** 2).sum((1, 2)) / n)
    out["l1_err"] = diff.abs().sum((1, 2)) / n
    out["abs_rel"] = torch.where(m, diff.abs() / t, torch.zeros_like(d)).sum((1, 2)) / n
    return out, d.unsqueeze(1)
