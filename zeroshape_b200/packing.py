"""Weight re-packing for the CUDA kernels (once per weight update, never per image).

The reference keeps convolution filters in PyTorch's OIHW layout, applies weight standardisation
(timm StdConv2dSame) and eval-mode BatchNorm at every forward.  The kernels want OHWI filters
(K-contiguous implicit-GEMM operand), already standardised, with BatchNorm folded into a per-output-
channel scale (into the filter) and shift (into the bias).  `PackCache` rebuilds the derived tensors
only when a source parameter changed (`load_state_dict`, optimizer step -> tensor `_version` bump).
These are tiny tensor-algebra preprocessing steps on the weights themselves and run through PyTorch.
"""
import torch


def ohwi(w):
    """[O,I,KH,KW] -> contiguous [O,KH,KW,I]."""
    return w.detach().float().permute(0, 2, 3, 1).contiguous()


def ws_ohwi(w, eps=1e-8):
    """timm StdConv2dSame weight standardisation (biased variance over I*KH*KW, eps inside the sqrt)."""
    w = w.detach().float()
    wf = w.reshape(w.shape[0], -1)
    var, mean = torch.var_mean(wf, dim=1, keepdim=True, unbiased=False)
    return ((wf - mean) / torch.sqrt(var + eps)).reshape(w.shape).permute(0, 2, 3, 1).contiguous()


def fold_bn_ohwi(w, bn, conv_bias=None):
    """conv -> eval BatchNorm == conv with filter*scale and bias shift.  Returns (w_ohwi, bias)."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    if conv_bias is not None:
        shift = shift + conv_bias.detach().float() * scale
    wf = w.detach().float() * scale.view(-1, 1, 1, 1)
    return wf.permute(0, 2, 3, 1).contiguous(), shift.contiguous()


class PackCache:
    """dict of derived tensors, invalidated when any tensor of `module.state_dict()` changes."""

    def __init__(self, module):
        self.module = module
        self.key = None
        self.store = {}

    def _version_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.module.parameters()) + list(self.module.buffers()))

    def refresh(self):
        key = self._version_key()
        if key != self.key:
            self.store = {}
            self.key = key
        return self.store

    def get(self, name, builder):
        if name not in self.store:
            self.store[name] = builder()
        return self.store[name]
