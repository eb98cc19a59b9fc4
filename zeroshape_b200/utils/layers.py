"""Host-side mirror of the reference's utils/layers.py pieces that are on the hot path.

  Bottleneck_Conv (utils/layers.py:76-100): conv -> BN -> ReLU -> conv -> BN -> +x -> ReLU, used by the
  intrinsics head (kernel 3, graph_shape.py:20-23) and CoordEncRes (kernel 1, seen_coord_enc.py:149-153).
Parameter names (`linear1`, `bn1`, `linear2`, `bn2`) are the reference's.  Eval mode folds BN into the filters; train mode
uses batch statistics (forward value here, gradients via model/shape/seen_coord_enc_train.py).
"""
import torch
import torch.nn as nn

from .. import ops
from ..packing import fold_bn_ohwi


class Bottleneck_Conv(nn.Module):
    def __init__(self, n_channels, kernel_size=1):
        super().__init__()
        self.kernel_size = kernel_size
        self.linear1 = nn.Conv2d(n_channels, n_channels, kernel_size=kernel_size, padding=kernel_size // 2, bias=False)
        self.bn1 = nn.BatchNorm2d(n_channels)
        self.linear2 = nn.Conv2d(n_channels, n_channels, kernel_size=kernel_size, padding=kernel_size // 2, bias=False)
        self.bn2 = nn.BatchNorm2d(n_channels)

    def run_nhwc(self, x, cache, tag):
        """x [B,H,W,C] NHWC (a [B,C] vector is a 1x1 image) -> same shape.  BN folded via `cache`."""
        squeeze = x.dim() == 2
        if self.training:
            # train mode = batch statistics + running-stat update, exactly what nn.BatchNorm2d does even when the parameters
            # are frozen (the reference leaves intr_head in train mode under optim.fix_dpt, graph_shape.py:35-38).
            # Forward value only here; the differentiable use goes through seen_coord_enc_train.CoordEncTrainFn.
            from ..model.shape.seen_coord_enc_train import _bneck_conv_fwd
            with torch.no_grad():
                y = _bneck_conv_fwd([], x.view(x.shape[0], 1, 1, x.shape[1]) if squeeze else x, self)
            return y.view(y.shape[0], -1) if squeeze else y
        if squeeze:
            x = x.view(x.shape[0], 1, 1, x.shape[1])
        p = self.kernel_size // 2
        w1, b1 = cache.get(tag + ".1", lambda: fold_bn_ohwi(self.linear1.weight, self.bn1))
        w2, b2 = cache.get(tag + ".2", lambda: fold_bn_ohwi(self.linear2.weight, self.bn2))
        y = ops.conv2d_nhwc(x, w1, b1, 1, (p, p, p, p), act=ops.ACT_RELU)
        y = ops.conv2d_nhwc(y, w2, b2, 1, (p, p, p, p), act=ops.ACT_RELU, res=x, res_mode=ops.RES_BEFORE_ACT)
        return y.view(y.shape[0], -1) if squeeze else y
