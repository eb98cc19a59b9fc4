"""Host-side mirror of the reference's utils/loss.py `Loss` (same method names and arguments).

  shape_loss(pred_occ_raw [B,N], gt_sdf [B,N])   utils/loss.py:18-28   BCE-with-logits vs (sdf < 0), importance weight near
                                                                       the surface -> zs_bce_logits_fwd / _bwd (autograd-aware)
  intr_loss(seen_pred, seen_gt, mask)            utils/loss.py:36-41   masked mean squared distance (autograd-aware)
  depth_loss(pred, gt, mask)                     utils/loss.py:30-34   MiDaS scale-and-shift-invariant loss + gradient matching
                                                                       -> zs_midas_loss_f32 (value and gradient in three launches)
"""
import torch
import torch.nn as nn

from .. import ops


class _ShapeLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, sdf, thres, weight):
        lg, sd = logits.detach().float().contiguous(), sdf.detach().float().contiguous()
        ctx.save_for_backward(lg, sd)
        ctx.thres, ctx.weight = thres, weight
        return ops.bce_logits_loss(lg, sd, thres, weight)

    @staticmethod
    def backward(ctx, gout):
        lg, sd = ctx.saved_tensors
        d = ops.bce_logits_loss_bwd(lg, sd, ctx.thres, ctx.weight, 1.0)
        return d * gout, None, None, None


class _DepthLossFn(torch.autograd.Function):
    """MiDaS loss value + gradient in one pass of three launches (zs_midas_loss_f32); backward = stored gradient x upstream."""

    @staticmethod
    def forward(ctx, pred, gt, mask, alpha, inverse):
        loss, dpred = ops.midas_loss(pred.detach().float(), gt.detach().float(), mask.detach().float(), alpha, inverse,
                                     need_grad=pred.requires_grad)
        ctx.dpred = dpred
        return loss

    @staticmethod
    def backward(ctx, gout):
        return (ctx.dpred * gout if ctx.dpred is not None else None), None, None, None, None


class _IntrLossFn(torch.autograd.Function):
    """utils/loss.py:36-41: masked mean squared distance of the normalised seen surfaces; three [B,HW,3]-sized elementwise
    tensors (host glue, as the reference's own lines)."""

    @staticmethod
    def forward(ctx, seen_pred, seen_gt, mask):
        diff = (seen_pred - seen_gt).detach()
        denom = mask.sum() + 1.e-8
        ctx.save_for_backward(diff, mask, denom)
        return ((diff ** 2).sum(-1) * mask).sum() / denom

    @staticmethod
    def backward(ctx, gout):
        diff, mask, denom = ctx.saved_tensors
        g = diff * (2.0 * gout / denom) * mask.unsqueeze(-1)
        return g, None, None


class Loss(nn.Module):
    def __init__(self, opt):
        super().__init__()
        tr = opt.get("training", None) if isinstance(opt, dict) else getattr(opt, "training", None)
        sl = (tr or {}).get("shape_loss", {}) if tr is not None else {}
        self.impt_thres = float(sl.get("impt_thres", 0.01))       # options/shape.yaml:76-78
        self.impt_weight = float(sl.get("impt_weight", 1.0))
        dl = (tr or {}).get("depth_loss", {}) if tr is not None else {}
        self.depth_alpha = float(dl.get("grad_reg", 0.1))         # options/depth.yaml:44-47 -> MidasLoss(alpha, inverse_depth, shrink_mask)
        self.depth_inv = bool(dl.get("depth_inv", True))
        self.depth_mask_shrink = bool(dl.get("mask_shrink", False))

    def shape_loss(self, pred_occ_raw, gt_sdf):
        assert pred_occ_raw.dim() == 2 and gt_sdf.dim() == 2
        return _ShapeLossFn.apply(pred_occ_raw, gt_sdf, self.impt_thres, self.impt_weight)

    def intr_loss(self, seen_pred, seen_gt, mask):
        assert seen_pred.dim() == 3 and seen_gt.dim() == 3 and mask.dim() == 2
        return _IntrLossFn.apply(seen_pred, seen_gt, mask)

    def depth_loss(self, pred_depth, gt_depth, mask):
        assert pred_depth.dim() == gt_depth.dim() == mask.dim() == 4
        assert pred_depth.shape[1] == gt_depth.shape[1] == mask.shape[1] == 1
        if self.depth_mask_shrink:
            mask = ops.erode_mask(mask.detach(), 4)               # MidasLoss.erode_mask (midas_loss.py:158-167)
        return _DepthLossFn.apply(pred_depth, gt_depth, mask, self.depth_alpha, self.depth_inv)
