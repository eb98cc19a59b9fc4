"""Drop-in for the reference's utils/camera.py: every public name of that module (the data loaders use `pose` / `Pose`,
data/synthetic.py:139-140; the engines use the projection helpers) with the hot-path functions on the CUDA library.

  Pose / pose            utils/camera.py:6-49      valid_norm_fac        :52-78       get_pixel_grid     :80-86
  unproj_depth           :88-108                   to_hom / world2cam / cam2img / proj_points             :110-154
  azim/elev/roll_to_rotation_matrix  :156-206      get_rotation_sphere   :208-230
"""
import numpy as np
import torch

from .. import ops


class Pose:
    """[..., 3, 4] camera poses [R | t] and their algebra (utils/camera.py:6-47).  Host-side helper of the data loaders."""

    def __call__(self, R=None, t=None):
        assert R is not None or t is not None
        if R is not None and not isinstance(R, torch.Tensor):
            R = torch.tensor(R)
        if t is not None and not isinstance(t, torch.Tensor):
            t = torch.tensor(t)
        if R is None:      # pure translation
            R = torch.eye(3, device=t.device).repeat(*t.shape[:-1], 1, 1)
        if t is None:      # pure rotation
            t = torch.zeros(R.shape[:-1], device=R.device)
        assert R.shape[:-1] == t.shape and R.shape[-2:] == (3, 3)
        out = torch.cat([R.float(), t.float().unsqueeze(-1)], dim=-1)
        assert out.shape[-2:] == (3, 4)
        return out

    def invert(self, pose, use_inverse=False):
        R, t = pose[..., :3], pose[..., 3:]
        Ri = R.inverse() if use_inverse else R.transpose(-1, -2)
        return self(R=Ri, t=(-Ri @ t)[..., 0])

    def compose_pair(self, pose_a, pose_b):
        """x -> pose_b(pose_a(x))"""
        Ra, ta = pose_a[..., :3], pose_a[..., 3:]
        Rb, tb = pose_b[..., :3], pose_b[..., 3:]
        return self(R=Rb @ Ra, t=(Rb @ ta + tb)[..., 0])

    def compose(self, pose_list):
        """x -> poseN(...pose2(pose1(x)))"""
        out = pose_list[0]
        for nxt in pose_list[1:]:
            out = self.compose_pair(out, nxt)
        return out


pose = Pose()


def get_pixel_grid(opt, H, W):
    """[H*W, 3] homogeneous pixel coordinates (x, y, 1), row-major (utils/camera.py:80-86)."""
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=opt.device),
                            torch.arange(W, dtype=torch.float32, device=opt.device), indexing="ij")
    return torch.stack([xs, ys, torch.ones_like(xs)], dim=-1).view(-1, 3)


def to_hom(X):
    """[..., 3] -> [..., 4] (utils/camera.py:110-117)."""
    return torch.cat([X, torch.ones_like(X[..., :1])], dim=-1)


def world2cam(X_world, pose):
    """[B,N,3] world points through [B,3,4] poses (utils/camera.py:119-128)."""
    return to_hom(X_world) @ pose.transpose(-1, -2)


def cam2img(X_cam, cam_intr):
    """(utils/camera.py:130-138)"""
    return X_cam @ cam_intr.transpose(-1, -2)


def proj_points(opt, points, intr, pose):
    """-> (pixel coordinates [B,N,2], camera depth [B,N]) (utils/camera.py:140-154)."""
    cam = world2cam(points, pose)
    img = cam2img(cam, intr)
    return img[..., :2] / img[..., 2:], cam[..., 2]


def _cos_sin(a, representation):
    if representation == "trig":
        return a[:, 0], a[:, 1]
    if representation == "angle":
        a = a * np.pi / 180
    elif representation != "rad":
        raise ValueError(representation)
    return torch.cos(a), torch.sin(a)


def azim_to_rotation_matrix(azim, representation='angle'):
    """Rotation about +Y by the azimuth (utils/camera.py:156-172): [B] (or [B,2] cos/sin) -> [B,3,3]."""
    c, s = _cos_sin(azim, representation)
    R = torch.eye(3, device=azim.device).repeat(len(azim), 1, 1)
    R[:, 0, 0], R[:, 0, 2], R[:, 2, 0], R[:, 2, 2] = c, s, -s, c
    return R


def elev_to_rotation_matrix(elev, representation='angle'):
    """Rotation about +X by the elevation (utils/camera.py:174-189)."""
    c, s = _cos_sin(elev, representation)
    R = torch.eye(3, device=elev.device).repeat(len(elev), 1, 1)
    R[:, 1, 1], R[:, 1, 2], R[:, 2, 1], R[:, 2, 2] = c, -s, s, c
    return R


def roll_to_rotation_matrix(roll, representation='angle'):
    """Rotation about +Z by the roll (utils/camera.py:191-206)."""
    c, s = _cos_sin(roll, representation)
    R = torch.eye(3, device=roll.device).repeat(len(roll), 1, 1)
    R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1] = c, s, -s, c
    return R


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
