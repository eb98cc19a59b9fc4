"""Host-side mirror of the reference's utils/eval_depth.py `DepthMetric` (same constructor, `metric_keys`, `compute_metrics`):
scale-and-shift aligned depth metrics of the depth engine's evaluation loop (model/depth_engine.py), one launch per batch
(zs_depth_metrics_f32) instead of ~40 boolean-gather / reduction launches."""
import torch

from .. import ops
from .._native import check, lib


class DepthMetric:
    def __init__(self, thresholds=[1.25, 1.25 ** 2, 1.25 ** 3], depth_cap=None, prediction_type='depth'):
        self.thresholds = thresholds
        self.depth_cap = depth_cap
        self.metric_keys = self.get_metric_keys()
        self.prediction_type = prediction_type

    def get_metric_keys(self):
        return ['d>{}'.format(t) for t in self.thresholds] + ['rmse', 'l1_err', 'abs_rel']

    def compute_metrics(self, prediction, target, mask):
        """prediction, target, mask [B,1,H,W] -> ({key: [B]}, aligned prediction depth [B,1,H,W]) (utils/eval_depth.py:41-110)."""
        prediction, target, mask = prediction.float().contiguous(), target.float().contiguous(), mask.float().contiguous()
        assert prediction.shape == target.shape == mask.shape and prediction.dim() == 4 and prediction.shape[1] == 1
        if self.prediction_type not in ('depth', 'disparity'):
            raise ValueError('Unknown prediction type: {}'.format(self.prediction_type))
        for t, n in ((prediction, "prediction"), (target, "target"), (mask, "mask")):
            ops._chk(t, n)
        B, _, H, W = prediction.shape
        T = len(self.thresholds)
        thr = torch.tensor([float(t) for t in self.thresholds], device=prediction.device, dtype=torch.float32)
        out = torch.empty(B, T + 3, device=prediction.device, dtype=torch.float32)
        depth = torch.empty_like(prediction)
        check(lib.zs_depth_metrics_f32(ops._p(prediction), ops._p(target), ops._p(mask), B, H, W, ops._p(thr), T,
                                       float(self.depth_cap) if self.depth_cap is not None else 0.0,
                                       int(self.prediction_type == 'disparity'), ops._p(out), ops._p(depth), ops._stream()),
              "zs_depth_metrics_f32")
        metrics = {k: out[:, i] for i, k in enumerate(self.metric_keys)}
        return metrics, depth
