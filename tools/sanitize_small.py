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
# round 2, last session: tiled GroupNorm (ragged chunks, group size 2 and 32), stride-2 stem dgrad (partial tiles), 128-bit bilinear
# forward / gather backward (odd sizes, both conventions), capturable AdamW (more than one parameter table per launch group),
# attention backward kernels on ragged token / point counts
for B, C, H, W in ((1, 64, 37, 29), (2, 256, 15, 15), (1, 1024, 5, 4), (2, 96, 9, 9)):
    x = torch.randn(B, H, W, C, device=dev)
    ops.groupnorm_nhwc(x, torch.ones(C, device=dev), torch.zeros(C, device=dev), 32, 1e-5, True, x.clone())
for B, H, W, Cin, Cout, K, pad in ((1, 37, 29, 3, 64, 7, 3), (2, 18, 20, 4, 16, 3, 1), (1, 16, 16, 3, 64, 7, 3)):
    OH, OW = (H + 2 * pad - K) // 2 + 1, (W + 2 * pad - K) // 2 + 1
    ops.conv2d_nhwc_dgrad(torch.randn(B, OH, OW, Cout, device=dev), torch.randn(Cout, K, K, Cin, device=dev), (B, H, W, Cin), 2,
                          (pad, pad, pad, pad), tc=False)
for B, H, W, C, OH, OW, al in ((1, 7, 9, 12, 14, 18, True), (2, 24, 24, 8, 14, 14, False), (1, 13, 11, 4, 29, 23, False), (2, 1, 1, 4, 4, 4, True)):
    ops.bilinear_nhwc(torch.randn(B, H, W, C, device=dev), OH, OW, al)
    ops.bilinear_bwd_nhwc(torch.randn(B, OH, OW, C, device=dev), H, W, al)
ps = [torch.randn(int(n), device=dev) for n in torch.randint(1, 500, (130,), generator=g)]
hyper = torch.tensor([1e-3, 0.9, 0.95, 1e-8, 0.05, 0.1, 0.3, 0.0], device=dev)
ops.adamw_step_multi_dev(ps, [torch.randn_like(p) for p in ps], [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps], hyper)
for B, T, H, hd in ((1, 197, 12, 64), (2, 50, 3, 32), (1, 130, 2, 32)):
    qkv = torch.randn(B, T, 3 * H * hd, device=dev)
    ops.mha_bwd(qkv, torch.randn(B, T, H * hd, device=dev), H, tc=True)
torch.cuda.synchronize()
print("sanitize_small: done")
