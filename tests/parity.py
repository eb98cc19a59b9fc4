"""The parity metric of DESIGN.md section 3 (used by every decoder / depth parity test).

North-star bar: "within 1e-3 relative of the fp32 reference".  A pure ratio |d|/|ref| is undefined where the
occupancy logit crosses zero -- which is exactly the iso-surface the decoder exists to locate -- so the bar is
applied as an allclose with rtol = 1e-3 and atol = 1e-3 * 0.25 * rms(ref):

    rel(a, ref) = max |a - ref| / max(|ref|, 0.25 * rms(ref))        must be < 1e-3

together with a normwise bound ||a - ref|| / ||ref|| < 1e-4 and equality of the thresholded voxel grid outside the
band the elementwise bound allows.
"""
import torch


def parity_rel(a, ref):
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    floor = 0.25 * ref.pow(2).mean().sqrt()
    return ((a - ref).abs() / ref.abs().clamp_min(floor)).max().item()


def normwise(a, ref):
    a, ref = a.detach().double().cpu(), ref.detach().double().cpu()
    return ((a - ref).pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item()
