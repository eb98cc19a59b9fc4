"""Small host utilities mirroring the pieces of the reference's utils/util.py that the hot path uses."""
import torch

from .. import ops


class EasyDict(dict):
    """Attribute-access dict (reference: utils/util.py EasyDict) -- `var` / `opt` containers."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {}, **kwargs)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __delattr__(self, k):
        del self[k]


def interpolate_coordmap(coord_map, mask_input, size, bg_coord=0):
    """Mask-aware bilinear resampling of an XYZ map (reference: utils/util.py:336-345).
    coord_map [B,3,H,W], mask_input [B,1,H,W] -> (coord_out [B,3,h,w], mask_binary [B,1,h,w]).
    Runs on the library's NHWC bilinear kernel."""
    assert coord_map.dim() == 4 and mask_input.dim() == 4
    h, w = size
    mask = (mask_input > 0.5).float()
    cv = ops.nchw_to_nhwc((coord_map * mask).contiguous())
    mk = ops.nchw_to_nhwc(mask.contiguous())
    cv = ops.nhwc_to_nchw(ops.bilinear_nhwc(cv, h, w, False))
    mk = ops.nhwc_to_nchw(ops.bilinear_nhwc(mk, h, w, False))
    coord_out = cv / (mk + 1.e-6)
    mask_binary = (mk > 0.5).float()
    return coord_out * mask_binary + bg_coord * (1 - mask_binary), mask_binary
