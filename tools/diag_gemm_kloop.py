"""Per-chunk cost of the tcgen05 GEMM's K loop for a lone CTA (one 128 x 256 output tile, no split-K): device time from CUDA-graph replays
against the number of 64-wide K-chunks, bf16x3 (three MMA passes, 64 KB of weights per chunk) and bf16 (one pass, 32 KB)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zeroshape_b200 import ops
from zeroshape_b200._native import lib

dev = torch.device("cuda", 0)
lib.zs_debug_gemm_splitk(0)


def dev_time(fn, n=20, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        g.replay()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / (reps * n) * 1e3


for M, N in ((128, 256), (128, 64), (1024, 256), (128 * 148, 256)):
    for prec in ("bf16x3", "bf16"):
        row = []
        for chunks in (2, 8, 32, 128):
            K = 64 * chunks
            a = torch.randn(M, K, device=dev)
            w = ops.PackedWeight(torch.randn(N, K, device=dev) * 0.05)
            row.append((chunks, dev_time(lambda: ops.gemm_tc(a, w, None, precision=prec))))
        slope = (row[-1][1] - row[-2][1]) / (row[-1][0] - row[-2][0])
        print(f"M={M:6d} N={N:4d} {prec:7s}: " + "  ".join(f"{c:3d} chunks {t:7.1f} us" for c, t in row) + f"  | {slope:5.2f} us per chunk", flush=True)
