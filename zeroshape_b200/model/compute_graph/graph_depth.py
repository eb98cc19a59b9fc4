"""Host-side mirror of the reference's depth-only compute graph (model/compute_graph/graph_depth.py:61-97):
DPT depth + intrinsics head -> var.depth_pred, var.intr_pred, var.seen_points_pred.  BASELINE config (1).
`forward(training=True)` runs the same graph on the tape of model/depth/dpt_train.py and returns the reference's losses
(graph_depth.py:99-105: MiDaS depth loss on depth_pred, masked squared distance of the normalised seen surfaces), both
differentiable into every parameter of the depth estimator and the intrinsics head (`train.py options/depth.yaml`)."""
import torch
import torch.nn as nn

from ... import ops
from ...packing import PackCache
from ...utils.layers import Bottleneck_Conv
from ...utils.util import EasyDict as edict
from ..depth.dpt_depth import DPTDepthModel


class Graph(nn.Module):

    def __init__(self, opt):
        super().__init__()
        self.dpt_depth = DPTDepthModel(backbone="vitb_rn50_384")
        # graph_depth.py:16-19: options/depth.yaml:17 points at the omnidata DPT-hybrid checkpoint the depth estimator is
        # fine-tuned from; without it training would silently start from a random initialisation
        pretrained = getattr(getattr(opt.arch, "depth", None), "pretrained", None)
        if pretrained is not None:
            checkpoint = torch.load(pretrained, map_location="cpu")
            self.dpt_depth.load_state_dict(checkpoint["model_state_dict"])
        self.with_intr = opt.loss_weight.intr is not None
        if self.with_intr:
            self.intr_feat_channels = 768
            self.intr_head = nn.Sequential(Bottleneck_Conv(768, kernel_size=3), Bottleneck_Conv(768, kernel_size=3))
            self.intr_pool = nn.AdaptiveAvgPool2d((1, 1))
            self.intr_proj = nn.Linear(768, 3)
            nn.init.zeros_(self.intr_proj.weight)
            nn.init.zeros_(self.intr_proj.bias)
            self._intr_cache = PackCache(self.intr_head)
        from ...utils.loss import Loss
        self.loss_fns = Loss(opt)

    def forward(self, opt, var, training=False, get_loss=True):
        B = len(var.idx)
        if training:
            return self._forward_train(opt, var, get_loss)
        with torch.no_grad():
            var.depth_pred = self.dpt_depth(var.rgb_input_map, get_feat=False)
            if self.with_intr:
                self._intr_cache.refresh()
                f = self.intr_head[0].run_nhwc(self.dpt_depth.last_feat_nhwc, self._intr_cache, "h0")
                f = self.intr_head[1].run_nhwc(f, self._intr_cache, "h1")
                params = ops.linear(ops.avgpool_nhwc(f), self.intr_proj.weight, self.intr_proj.bias)
                var.intr_pred = ops.intr_param2mtx(params, opt.H, opt.W)
                mask = var.mask_input_map.float().contiguous()
                var.seen_points_pred, _, _ = ops.unproject_normalize(var.depth_pred, mask, var.intr_pred)
        return (var, edict()) if get_loss else var

    def _forward_train(self, opt, var, get_loss):
        """graph_depth.py:61-105 in train mode: batch-statistics BatchNorm in the intrinsics head, losses with gradients."""
        from ..depth.dpt_train import DepthGraphTrainFn
        B = len(var.idx)
        mask = var.mask_input_map.float().contiguous()
        params = [p for p in self.parameters()]
        if self.with_intr:
            var.depth_pred, var.intr_pred, var.seen_points_pred = DepthGraphTrainFn.apply(self, opt, var.rgb_input_map, mask, True, *params)
            with torch.no_grad():
                var.seen_points_gt, _, _ = ops.unproject_normalize(var.depth_input_map.float(), mask, var.intr.float())
                var.validity_mask = (var.mask_input_map > 0.5).float().view(B, -1)
        else:
            var.depth_pred = DepthGraphTrainFn.apply(self, opt, var.rgb_input_map, mask, False, *params)
        if not get_loss:
            return var
        loss = edict()
        if opt.loss_weight.depth is not None:
            loss.depth = self.loss_fns.depth_loss(var.depth_pred, var.depth_input_map, var.mask_input_map)
        if opt.loss_weight.intr is not None:
            loss.intr = self.loss_fns.intr_loss(var.seen_points_pred, var.seen_points_gt, var.validity_mask)
        return var, loss
