"""Oracle (TEST INFRASTRUCTURE): CPU restatement of the reference's evaluation path.

  dense_grid / level_grid   <- utils/eval_3D.py:10-20, 22-46 (slice loop over x, sigmoid)
  marching_cubes            <- PyMCubes 0.1.4 `mcubes.marching_cubes` (call site utils/eval_3D.py:250);
                               third-party, absent here: restated from the published algorithm
                               (inside iff value <= iso, welded vertex per crossed edge, linear
                               interpolation in double, index-unit coordinates).  Parity unpinned by
                               the reference; known-answer tests in tests/test_oracle_eval.py.
  scale_vertices            <- utils/eval_3D.py:252-255 (v / S * (max-min) + min with S = n)
  sample_surface            <- trimesh 4.0.8 `Trimesh.sample` (call site :261): area-weighted faces,
                               uniform barycentric with reflection; distributional parity only.
  normalize_pc, fscore      <- utils/eval_3D.py:93-102, 215-231
  chamfer (C)               <- external/chamfer3D/chamfer3D.cu via oracle/chamfer_ref.c
  rotation_sphere           <- utils/camera.py:156-230
  brute_force_search        <- utils/eval_3D.py:140-170
"""
import ctypes
import os

import numpy as np
import torch

from .implicit import implicit_forward
from .mc_tables import TRI_TABLE, EDGES, CORNERS


# ------------------------------------------------------------------------------------------------
def dense_grid(n, rmin, rmax, batch=1):
    g = torch.linspace(rmin, rmax, n)
    pts = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), dim=-1)
    return pts.repeat(batch, 1, 1, 1, 1)


@torch.no_grad()
def level_grid(sd, latent_depth, n, rmin, rmax, x0=0, x1=None):
    """sigmoid occupancy over slices [x0,x1) of the (n)^3 grid, one decoder call per x-slice."""
    x1 = n if x1 is None else x1
    B = latent_depth.shape[0]
    pts = dense_grid(n, rmin, rmax, B).view(B, n, n * n, 3)
    occ = [implicit_forward(sd, latent_depth.float(), pts[:, i])[0] for i in range(x0, x1)]
    return torch.sigmoid(torch.stack(occ, dim=1).view(B, x1 - x0, n, n))


# ------------------------------------------------------------------------------------------------
_NTRI = np.array([len(r) // 3 for r in TRI_TABLE], dtype=np.int64)
_TRI = np.full((256, 15), -1, dtype=np.int64)
for _c, _r in enumerate(TRI_TABLE):
    _TRI[_c, :len(_r)] = _r
_EDGE_OWNER = np.zeros((12, 4), dtype=np.int64)
for _e, (_a, _b) in enumerate(EDGES):
    _ca, _cb = CORNERS[_a], CORNERS[_b]
    _ax = [i for i in range(3) if _ca[i] != _cb[i]][0]
    _lo = _ca if _ca[_ax] < _cb[_ax] else _cb
    _EDGE_OWNER[_e] = [_lo[0], _lo[1], _lo[2], _ax]


def marching_cubes(vol, iso):
    """vol [nx,ny,nz] float (the reference passes cubes, utils/eval_3D.py:250; x-slabs are used by the multi-GPU tests)
    -> (vertices float64 [V,3] in index units, faces int64 [F,3]).
    Ordering: vertices by (grid point linear index, axis x<y<z); faces by (cell linear index,
    table order) -- the same deterministic order the CUDA kernel uses, so outputs compare exactly."""
    vol = np.asarray(vol)
    assert vol.ndim == 3
    nx, ny, nz = vol.shape
    inside = vol <= iso
    f64 = vol.astype(np.float64)
    flags = np.zeros((nx, ny, nz), dtype=np.int64)
    flags[:-1, :, :] |= (inside[:-1] != inside[1:]).astype(np.int64) * 1
    flags[:, :-1, :] |= (inside[:, :-1] != inside[:, 1:]).astype(np.int64) * 2
    flags[:, :, :-1] |= (inside[:, :, :-1] != inside[:, :, 1:]).astype(np.int64) * 4
    cnt = ((flags & 1) + ((flags >> 1) & 1) + ((flags >> 2) & 1)).reshape(-1)
    vbase = np.concatenate([[0], np.cumsum(cnt)[:-1]]).reshape(nx, ny, nz)
    # vertices
    verts = np.zeros((int(cnt.sum()), 3), dtype=np.float64)
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    for axis in range(3):
        sel = ((flags >> axis) & 1).astype(bool)
        rank = np.zeros_like(flags)
        for a in range(axis):
            rank += (flags >> a) & 1
        ids = (vbase + rank)[sel]
        i0, j0, k0 = ii[sel], jj[sel], kk[sel]
        f1 = f64[i0, j0, k0]
        f2 = f64[i0 + (axis == 0), j0 + (axis == 1), k0 + (axis == 2)]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(f2 == f1, 0.5, (float(iso) - f1) / (f2 - f1))
        p = np.stack([i0, j0, k0], axis=1).astype(np.float64)
        # the product stores fp32 vertices: round the interpolated coordinate like it does
        p[:, axis] = (p[:, axis] + t)
        verts[ids] = p
    # faces
    c = inside
    case = (c[:-1, :-1, :-1] * 1 + c[1:, :-1, :-1] * 2 + c[1:, 1:, :-1] * 4 + c[:-1, 1:, :-1] * 8 +
            c[:-1, :-1, 1:] * 16 + c[1:, :-1, 1:] * 32 + c[1:, 1:, 1:] * 64 + c[:-1, 1:, 1:] * 128).astype(np.int64)
    ci, cj, ck = np.nonzero(_NTRI[case] > 0)          # C-order == cell linear index order
    cs = case[ci, cj, ck]
    faces = []
    for t in range(5):
        m = _NTRI[cs] > t
        if not m.any():
            break
        tri = np.zeros((int(m.sum()), 3), dtype=np.int64)
        for corner in range(3):
            e = _TRI[cs[m], t * 3 + corner]
            oi = ci[m] + _EDGE_OWNER[e, 0]
            oj = cj[m] + _EDGE_OWNER[e, 1]
            ok = ck[m] + _EDGE_OWNER[e, 2]
            ax = _EDGE_OWNER[e, 3]
            fl = flags[oi, oj, ok]
            rank = np.where(ax >= 1, fl & 1, 0) + np.where(ax >= 2, (fl >> 1) & 1, 0)
            tri[:, corner] = vbase[oi, oj, ok] + rank
        order = np.nonzero(m)[0] * 5 + t
        faces.append((order, tri))
    if not faces:
        return verts, np.zeros((0, 3), dtype=np.int64)
    order = np.concatenate([o for o, _ in faces])
    tris = np.concatenate([t for _, t in faces], axis=0)
    return verts, tris[np.argsort(order, kind="stable")]


def scale_vertices(verts, n, rmin, rmax):
    return verts / n * (rmax - rmin) + rmin


def sample_surface(verts, faces, count, rng):
    """trimesh.sample.sample_surface semantics with an explicit numpy Generator/RandomState."""
    if len(faces) == 0:
        return np.zeros([count, 3])
    tri = verts[faces]
    area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    cum = np.cumsum(area)
    pick = np.searchsorted(cum, rng.random(count) * cum[-1])
    o, vec = tri[pick, 0], tri[pick, 1:] - tri[pick, :1]
    r = rng.random((count, 2, 1))
    flip = r.sum(axis=1).reshape(-1) > 1.0
    r[flip] -= 1.0
    r = np.abs(r)
    return (vec * r).sum(axis=1) + o


def mesh_area(verts, faces):
    tri = verts[faces]
    return float((0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)).sum())


# ------------------------------------------------------------------------------------------------
def normalize_pc(pc):
    z = pc - pc.mean(dim=1, keepdim=True)
    lx = z[:, :, 0].max(dim=-1)[0] - z[:, :, 0].min(dim=-1)[0]
    ly = z[:, :, 1].max(dim=-1)[0] - z[:, :, 1].min(dim=-1)[0]
    lm = torch.stack([lx, ly], dim=-1).max(dim=-1)[0].unsqueeze(-1).unsqueeze(-1)
    return z / (lm + 1.e-7)


def fscore(dist1, dist2, thresholds=(0.005, 0.01, 0.02, 0.05, 0.1, 0.2)):
    out = []
    for th in thresholds:
        p = torch.mean((dist1 < th).float(), dim=1)
        r = torch.mean((dist2 < th).float(), dim=1)
        f = 2 * p * r / (p + r)
        f[torch.isnan(f)] = 0
        out.append(f)
    return torch.stack(out, dim=1)


_clib = None


def _lib():
    global _clib
    if _clib is None:
        from .build_oracle import build_c
        _clib = ctypes.CDLL(build_c())
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
        _clib.chamfer_nn_ref.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, ip]
        _clib.chamfer_nn_ref.restype = None
        _clib.chamfer_grad_ref.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, ip, fp, fp]
        _clib.chamfer_grad_ref.restype = None
    return _clib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def chamfer_nn(xyz1, xyz2):
    """numpy [b,n,3],[b,m,3] float32 -> dist1, dist2 (squared, fp32), idx1, idx2 (int32)."""
    a = np.ascontiguousarray(xyz1, dtype=np.float32)
    b_ = np.ascontiguousarray(xyz2, dtype=np.float32)
    b, n, _ = a.shape
    m = b_.shape[1]
    d1 = np.zeros((b, n), np.float32); i1 = np.zeros((b, n), np.int32)
    d2 = np.zeros((b, m), np.float32); i2 = np.zeros((b, m), np.int32)
    if m > 0 and n > 0:
        _lib().chamfer_nn_ref(_fp(a), _fp(b_), b, n, m, _fp(d1), _ip(i1))
        _lib().chamfer_nn_ref(_fp(b_), _fp(a), b, m, n, _fp(d2), _ip(i2))
    return d1, d2, i1, i2


def chamfer_grad(xyz1, xyz2, g1, g2, i1, i2):
    a = np.ascontiguousarray(xyz1, dtype=np.float32)
    b_ = np.ascontiguousarray(xyz2, dtype=np.float32)
    b, n, _ = a.shape
    m = b_.shape[1]
    ga = np.zeros_like(a); gb = np.zeros_like(b_)
    _lib().chamfer_grad_ref(_fp(a), _fp(b_), b, n, m, _fp(np.ascontiguousarray(g1, np.float32)),
                            _ip(np.ascontiguousarray(i1, np.int32)), _fp(ga), _fp(gb))
    _lib().chamfer_grad_ref(_fp(b_), _fp(a), b, m, n, _fp(np.ascontiguousarray(g2, np.float32)),
                            _ip(np.ascontiguousarray(i2, np.int32)), _fp(gb), _fp(ga))
    return ga, gb


# ------------------------------------------------------------------------------------------------
def _axis_rot(kind, deg):
    ang = torch.tensor([deg]) * np.pi / 180
    cos, sin = torch.cos(ang), torch.sin(ang)
    R = torch.eye(3)[None].repeat(1, 1, 1)
    zeros = torch.zeros(1)
    if kind == "azim":      # utils/camera.py:156-171
        R[:, 0, :] = torch.stack([cos, zeros, sin], dim=-1)
        R[:, 2, :] = torch.stack([-sin, zeros, cos], dim=-1)
    elif kind == "elev":    # utils/camera.py:173-187
        R[:, 1, 1:] = torch.stack([cos, -sin], dim=-1)
        R[:, 2, 1:] = torch.stack([sin, cos], dim=-1)
    else:                   # roll, utils/camera.py:189-206
        R[:, 0, :2] = torch.stack([cos, sin], dim=-1)
        R[:, 1, :2] = torch.stack([-sin, cos], dim=-1)
    return R


def rotation_sphere(azim_sample=4, elev_sample=4, roll_sample=4, scales=(1.0,)):
    """utils/camera.py:208-230."""
    perm = torch.tensor([[-1, 0, 0], [0, 0, -1], [0, -1, 0]]).float().unsqueeze(0)
    out = []
    for scale in scales:
        for azim in np.linspace(0, 360, num=azim_sample, endpoint=False):
            for elev in np.linspace(0, 360, num=elev_sample, endpoint=False):
                for roll in np.linspace(0, 360, num=roll_sample, endpoint=False):
                    out.append((scale * _axis_rot("roll", roll) @ _axis_rot("elev", elev) @ _axis_rot("azim", azim) @ perm).float())
    return torch.cat(out, dim=0)


def brute_force_search(pc_pred, pc_gt, thresholds, rotations, batch_size=24):
    """utils/eval_3D.py:140-170 on CPU with the C chamfer oracle.  pc_* [n,3] tensors."""
    pc_pred = pc_pred.unsqueeze(0).float()
    pc_gt = normalize_pc(pc_gt.unsqueeze(0).float())
    best_cd, best = np.inf, None
    for i in range(0, len(rotations), batch_size):
        R = rotations[i:i + batch_size]
        rot = normalize_pc((R @ pc_pred.repeat(R.shape[0], 1, 1).permute(0, 2, 1)).permute(0, 2, 1)).contiguous()
        d1, d2, _, _ = chamfer_nn(rot.numpy(), pc_gt.repeat(R.shape[0], 1, 1).numpy())
        acc, comp = torch.from_numpy(d1).sqrt(), torch.from_numpy(d2).sqrt()
        f = fscore(acc, comp, thresholds)
        acc, comp = acc.mean(dim=1), comp.mean(dim=1)
        cd = (acc + comp) / 2
        for j in range(len(cd)):
            if cd[j] < best_cd:
                best_cd = cd[j]
                best = (acc[j], comp[j], f[j], rot[j].clone(), i + j)
    return best
