import os, sys
sys.path.insert(0, os.getcwd())
import torch
from zeroshape_b200 import ops
dev = torch.device("cuda", 0)
def t(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for B, H, C in ((32, 112, 128), (32, 56, 256), (32, 28, 256), (8, 112, 128)):
    x = torch.randn(B, H, H, C, device=dev); y = torch.randn(B, 2 * H, 2 * H, C, device=dev)
    print(f"[{B},{H},{H},{C}] -> x2: fwd {t(lambda: ops.bilinear_nhwc(x, 2 * H, 2 * H, True)):7.1f} us  bwd {t(lambda: ops.bilinear_bwd_nhwc(y, H, H, True)):7.1f} us  (bytes fwd {(x.numel() + y.numel()) * 4 / 1e6:.0f} MB)")
