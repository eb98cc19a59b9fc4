"""Small invocations of the round-2 kernels for `compute-sanitizer --tool memcheck python tools/sanitize_small.py` (ragged tile counts,
padded tile-blocked buffers, points mode).  Not a test: the numerics are covered by tests/test_gpu_*.py."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zeroshape_b200 import ops  # noqa: E402
from zeroshape_b200.data.preprocess import resize_coeffs  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for M in (1, 130, 1000):
    x = (torch.randn(M, 256, generator=g) * 1.5).to(dev)
    w = (torch.randn(768, 256, generator=g) / 16).to(dev)
    b = (torch.randn(768, generator=g) * 0.1).to(dev)
    lat = torch.randn(197, 512, generator=g).to(dev)
    kb, vb = ops.attn_pack_fused(lat[:, :256], lat[:, 256:], 8)
    wb = ops.qkvattn_pack(w)
    a_blk = ops.chain_qkvattn(x, wb, b, kb, vb, 197, 32 ** -0.5, flags=24)
    pts = (torch.rand(M, 3, generator=g) * 3 - 1.5).to(dev)
    w3, b3 = torch.randn(256, 3, generator=g).to(dev), torch.randn(256, generator=g).to(dev)
    pp, st = ops.point_proj_tables(w3, b3)
    a2 = ops.chain_qkvattn_pts(pts, pp, st, wb, b, kb, vb, 197, 32 ** -0.5)
    w1, b1 = (torch.randn(1024, 256, generator=g) / 16).to(dev), torch.zeros(1024, device=dev)
    w2, b2 = (torch.randn(256, 1024, generator=g) / 32).to(dev), torch.zeros(256, device=dev)
    mats = []
    for gi in range(4):
        mats += [w1[256 * gi:256 * (gi + 1), :], w2[:, 256 * gi:256 * (gi + 1)]]
    blob, pblob = ops.pack_tiles(mats), ops.pack_generic((torch.randn(256, 256, generator=g) / 16).to(dev))
    ops.chain_pmlp(x.clone(), a_blk, pblob, b2, 1e-6, blob, b1, b2)
    xe = torch.empty(M, 256, device=dev)
    ops.chain_pmlp(xe, a2, pblob, b2, 1e-6, blob, b1, b2, points=pts, pp=pp)
    ops.chain_lin(a_blk, pblob, b2, 1, res=x, out=x)
    torch.cuda.synchronize()
for B, T, H, hd in ((1, 197, 12, 64), (2, 50, 3, 32), (1, 1, 1, 64)):
    ops.mha(torch.randn(B, T, 3 * H * hd, generator=g).to(dev), H, tc=True)
img = torch.from_numpy(np.random.RandomState(0).randint(0, 256, (90, 70, 4)).astype(np.uint8)).to(dev)
xb, xk = (torch.from_numpy(a).to(dev) for a in resize_coeffs(110, 224))
out = ops.rgba_crop_resize(img, -20, -15, 110, 110, 224, 224, xb, xk, xb, xk)
ops.rgba_composite(out, 1.0)
ops.erode_square(torch.rand(2, 33, 41, device=dev).round(), 5)
torch.cuda.synchronize()
print("sanitize_small: done")
