"""Camera helpers of the hot path (reference: utils/camera.py:52-108, 156-230)."""
import numpy as np
import torch

from .. import ops


def unproj_depth(opt, depth, intr):
    """depth [B,1,H,W], intr [B,3,3] -> camera-frame points [B,H*W,3] (utils/camera.py:88-108).
    (the Graph uses the fused unproject+normalise variant of the same kernel)"""
    B, _, H, W = depth.shape
    assert opt.H == H == W
    return ops.unproject(depth, intr.float())


def valid_norm_fac(seen_points, mask):
    """Masked mean and max-distance of seen points (utils/camera.py:52-78), without the per-sample
    Python loop / host syncs.  seen_points [B,HW,3], mask [B,1,H,W] bool."""
    B = seen_points.shape[0]
    m = mask.view(B, -1, 1).float()
    cnt = m.sum(dim=1)
    mean = (seen_points * m).sum(dim=1) / cnt
    d = ((seen_points - mean.unsqueeze(1)).norm(dim=2, keepdim=True) * m - (1 - m)).max(dim=1)[0].squeeze(-1)
    return mean, d


def _rot(axis, deg):
    a = torch.tensor(deg, dtype=torch.float32) * np.pi / 180
    c, s = torch.cos(a), torch.sin(a)
    R = torch.eye(3)
    if axis == "y":      # azim_to_rotation_matrix (camera.py:156-171)
        R[0, 0], R[0, 2], R[2, 0], R[2, 2] = c, s, -s, c
    elif axis == "x":    # elev_to_rotation_matrix (camera.py:173-187)
        R[1, 1], R[1, 2], R[2, 1], R[2, 2] = c, -s, s, c
    else:                # roll_to_rotation_matrix (camera.py:189-206)
        R[0, 0], R[0, 1], R[1, 0], R[1, 1] = c, s, -s, c
    return R


_SPHERE_CACHE = {}


def get_rotation_sphere(azim_sample=4, elev_sample=4, roll_sample=4, scales=[1.0], device='cuda'):
    """Rotation table for the brute-force pose search (utils/camera.py:208-230): R = s*Rz@Rx@Ry@P.
    Built once per (sampling) and cached (the reference rebuilds 6912 matrices per sample)."""
    key = (azim_sample, elev_sample, roll_sample, tuple(scales))
    if key not in _SPHERE_CACHE:
        Pm = torch.tensor([[-1, 0, 0], [0, 0, -1], [0, -1, 0]], dtype=torch.float32)
        out = []
        for scale in scales:
            for az in np.linspace(0, 360, num=azim_sample, endpoint=False):
                Ry = _rot("y", az)
                for el in np.linspace(0, 360, num=elev_sample, endpoint=False):
                    Rx = _rot("x", el)
                    for ro in np.linspace(0, 360, num=roll_sample, endpoint=False):
                        out.append(scale * _rot("z", ro) @ Rx @ Ry @ Pm)
        _SPHERE_CACHE[key] = torch.stack(out, dim=0)
    return _SPHERE_CACHE[key].to(device)
