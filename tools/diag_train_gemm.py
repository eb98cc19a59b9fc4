#!/usr/bin/env python
"""Device-time table of the training step's GEMM-shaped layers on the tensor-core kernels (forward / data gradient / weight
gradient), batch 32 geometry of BASELINE config 3.  `--once` runs every kernel a single time (for ncu captures)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zeroshape_b200 import ops      # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3          # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    ops.TRAIN_ENGINE, ops.TRAIN_PRECISION = "tc", args.precision
    reps = 1 if args.once else 10
    B = args.batch
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []
    linears = [("ViT qkv 768->2304", B * 197, 768, 2304), ("ViT fc1 768->3072", B * 197, 768, 3072), ("ViT fc2 3072->768", B * 197, 3072, 768),
               ("decoder fc1 256->1024", B * 4096, 256, 1024), ("decoder occ 256->256", B * 4096, 256, 256)]
    for name, M, K, N in linears:
        x = torch.randn(M, K, device=dev, generator=g)
        w = torch.randn(N, K, device=dev, generator=g) * 0.03
        dy = torch.randn(M, N, device=dev, generator=g)
        flop = 2.0 * M * K * N
        for kind, fn in (("fwd", lambda: ops.train_linear(x, w)), ("dgrad", lambda: ops.train_dgrad(dy, w)), ("wgrad", lambda: ops.gemm_tn(dy, x))):
            us = timed(fn, reps)
            rows.append((name, kind, us, flop / us / 1e6))
    convs = [("refinenet 3x3 256->256 @56", 56, 256, 256, 3, 1), ("BiT 1x1 64->256 @56", 56, 64, 256, 1, 1), ("BiT 3x3 64->64 @56", 56, 64, 64, 3, 1),
             ("head 3x3 256->128 @112", 112, 256, 128, 3, 1), ("head 3x3 128->32 @224", 224, 128, 32, 3, 1), ("ResNet 3x3 256->256 @14", 14, 256, 256, 3, 1),
             ("ResNet 1x1 1024->256 @14", 14, 1024, 256, 1, 1)]
    for name, H, Cin, Cout, k, s in convs:
        pad = (k // 2,) * 4
        x = torch.randn(B, H, H, Cin, device=dev, generator=g)
        w = torch.randn(Cout, k, k, Cin, device=dev, generator=g) * 0.03
        dy = torch.randn(B, H, H, Cout, device=dev, generator=g)
        flop = 2.0 * B * H * H * Cin * Cout * k * k
        for kind, fn in (("fwd", lambda: ops.conv2d_nhwc(x, w, None, s, pad, tc=True, precision=args.precision)),
                         ("dgrad", lambda: ops.conv2d_nhwc_dgrad(dy, w, x.shape, s, pad)),
                         ("wgrad", lambda: ops.conv2d_nhwc_wgrad(x, dy, k, k, s, pad))):
            us = timed(fn, reps)
            rows.append((name, kind, us, flop / us / 1e6))
        del x, w, dy
    print(f"| layer (batch {B}) | pass | us | algorithmic TFLOP/s ({args.precision}) |")
    print("|---|---|---|---|")
    for name, kind, us, tf in rows:
        print(f"| {name} | {kind} | {us:.1f} | {tf:.1f} |")


if __name__ == "__main__":
    main()
