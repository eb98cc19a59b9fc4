"""Device time of representative encoder layers (DPT-hybrid at 224 x 224) on the tcgen05 GEMM / implicit-GEMM convolution,
batch 1 and 8, warm caches, CUDA events over 50 launches -- to see which layer classes sit far above their arithmetic / traffic time.
   python tools/diag_encoder_layers.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zeroshape_b200 import ops

dev = torch.device("cuda", 0)


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps * 1e3       # us


def lin(M, N, K, tag):
    a = torch.randn(M, K, device=dev)
    w = ops.PackedWeight(torch.randn(N, K, device=dev) * 0.05)
    b = torch.randn(N, device=dev)
    us = timeit(lambda: ops.gemm_tc(a, w, b))
    fl = 2.0 * M * N * K * 3
    by = 4.0 * (M * K + M * N) + 4.0 * N * K
    print(f"linear {tag:28s} M={M:6d} N={N:5d} K={K:5d}: {us:8.1f} us | {fl / us * 1e-6:7.1f} TFLOP/s executed | traffic floor {by / 7.7e6:6.1f} us")


def conv(B, H, Cin, Cout, k, s, tag):
    x = torch.randn(B, H, H, Cin, device=dev)
    w = torch.randn(Cout, k, k, Cin, device=dev) * 0.05
    p = (k - 1) // 2
    us = timeit(lambda: ops.conv2d_nhwc(x, w, None, s, (p, p, p, p), tc=True))
    OH = (H + 2 * p - k) // s + 1
    M, K = B * OH * OH, k * k * Cin
    fl = 2.0 * M * Cout * K * 3
    by = 4.0 * (B * H * H * Cin + M * Cout) + 4.0 * Cout * K
    print(f"conv   {tag:28s} M={M:6d} N={Cout:5d} K={K:5d}: {us:8.1f} us | {fl / us * 1e-6:7.1f} TFLOP/s executed | traffic floor {by / 7.7e6:6.1f} us")


t = torch.randn(1000, device=dev)
print(f"host floor: axpby on 1000 floats {timeit(lambda: ops.axpby(t, 2.0)):6.1f} us per call; torch add {timeit(lambda: t + 1):6.1f} us")
g = torch.cuda.CUDAGraph()
a_ = torch.randn(197, 768, device=dev); w_ = ops.PackedWeight(torch.randn(768, 768, device=dev) * 0.05); b_ = torch.randn(768, device=dev)
for _ in range(3):
    ops.gemm_tc(a_, w_, b_)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(20):
        ops.gemm_tc(a_, w_, b_)
print(f"ViT proj M=197 replayed from a graph of 20: {timeit(lambda: g.replay(), 10) / 20:6.1f} us per GEMM (device time, no host in between)")
for B in (int(v) for v in os.environ.get("DIAG_BATCHES", "1,8").split(",")):
    print(f"---- batch {B}")
    lin(B * 197, 2304, 768, "ViT qkv")
    lin(B * 197, 768, 768, "ViT proj")
    lin(B * 197, 3072, 768, "ViT fc1")
    lin(B * 197, 768, 3072, "ViT fc2")
    conv(B, 56, 64, 64, 3, 1, "R50 stage0 3x3 64")
    conv(B, 56, 64, 256, 1, 1, "R50 stage0 1x1 64->256")
    conv(B, 28, 128, 128, 3, 1, "R50 stage1 3x3 128")
    conv(B, 14, 256, 256, 3, 1, "R50 stage2 3x3 256")
    conv(B, 14, 256, 1024, 1, 1, "R50 stage2 1x1 256->1024")
    conv(B, 14, 1024, 256, 1, 1, "R50 stage2 1x1 1024->256")
    conv(B, 7, 256, 256, 3, 1, "refinenet4 rcu 3x3 @7")
    conv(B, 14, 256, 256, 3, 1, "refinenet3 rcu 3x3 @14")
    conv(B, 28, 256, 256, 3, 1, "refinenet2 rcu 3x3 @28")
    conv(B, 56, 256, 256, 3, 1, "refinenet1 rcu 3x3 @56")
    conv(B, 112, 256, 128, 3, 1, "head conv 3x3 @112")
    conv(B, 224, 128, 32, 3, 1, "head conv 3x3 @224")
    x = torch.randn(B, 56, 56, 256, device=dev)
    g_, b_ = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    print(f"groupnorm [B,56,56,256]: {timeit(lambda: ops.groupnorm_nhwc(x, g_, b_, 32, 1e-5, True, None)):8.1f} us")
    x = torch.randn(B, 112, 112, 256, device=dev)
    print(f"bilinear 112->224 x256:  {timeit(lambda: ops.bilinear_nhwc(x, 224, 224, True)):8.1f} us")
