"""Host-side mirror of the reference's utils/loss.py `Loss` (same method names and arguments).

  shape_loss(pred_occ_raw [B,N], gt_sdf [B,N])   utils/loss.py:18-28   BCE-with-logits vs (sdf < 0), importance weight near
                                                                       the surface -> zs_bce_logits_fwd / _bwd (autograd-aware)
  intr_loss(seen_pred, seen_gt, mask)            utils/loss.py:36-41   masked mean squared distance (forward value only:
                                                                       the encoders have no backward in this revision)
  depth_loss(...)                                utils/loss.py:30-34   MiDaS SSI loss -- depth-engine row, raises
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


class Loss(nn.Module):
    def __init__(self, opt):
        super().__init__()
        tr = opt.get("training", None) if isinstance(opt, dict) else getattr(opt, "training", None)
        sl = (tr or {}).get("shape_loss", {}) if tr is not None else {}
        self.impt_thres = float(sl.get("impt_thres", 0.01))       # options/shape.yaml:76-78
        self.impt_weight = float(sl.get("impt_weight", 1.0))

    def shape_loss(self, pred_occ_raw, gt_sdf):
        assert pred_occ_raw.dim() == 2 and gt_sdf.dim() == 2
        return _ShapeLossFn.apply(pred_occ_raw, gt_sdf, self.impt_thres, self.impt_weight)

    def intr_loss(self, seen_pred, seen_gt, mask):
        assert seen_pred.dim() == 3 and seen_gt.dim() == 3 and mask.dim() == 2
        with torch.no_grad():     # scalar diagnostics on three small tensors (host glue, like the reference's own line)
            distance = ((seen_pred - seen_gt) ** 2).sum(-1)
            return (distance * mask).sum() / (mask.sum() + 1.e-8)

    def depth_loss(self, pred_depth, gt_depth, mask):
        raise NotImplementedError("MiDaS scale-and-shift-invariant depth loss (model/depth/midas_loss.py) belongs to the "
                                  "depth-engine training row (SURVEY.md section 8f rank 4); set loss_weight.depth = None")
