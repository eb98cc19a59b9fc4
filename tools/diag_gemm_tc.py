"""Diagnostics for the tcgen05 GEMM on a real B200 (run under gpurun; prints structured evidence so a
layout/descriptor bug can be identified from one run).  Not a test; writes gpurun_out/diag_gemm_tc.txt."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zeroshape_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
out_lines = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    out_lines.append(s)


def run(M, N, K, prec, a=None, w=None, tag=""):
    g = torch.Generator().manual_seed(1)
    a = torch.randint(-4, 5, (M, K), generator=g).float() if a is None else a
    w = torch.randint(-4, 5, (N, K), generator=g).float() if w is None else w
    ref = a.double() @ w.double().T
    out = ops.gemm_tc(a.to(dev), ops.PackedWeight(w.to(dev)), precision=prec)
    torch.cuda.synchronize()
    o = out.cpu().double()
    err = (o - ref).abs()
    log(f"[{tag}] M={M} N={N} K={K} prec={prec}: max|err|={err.max().item():.4g} mismatches={(err > 1e-3 * (ref.abs().max().item() + 1)).sum().item()}/{M * N}")
    return o, ref


log("device", torch.cuda.get_device_name(0), "cc", ops.device_cc())
# 1. identity weight: C[m,n] = A[m,n] for n < K -- any permutation of rows/cols/k is directly visible
K = 64
a = (torch.arange(128).view(-1, 1) * 100 + torch.arange(K).view(1, -1)).float()   # value encodes (row, k)
a = a % 251                                                                          # exact in bf16
w = torch.zeros(256, K)
w[torch.arange(K), torch.arange(K)] = 1
o, ref = run(128, 256, K, "bf16", a, w, "identity")
if (o - ref).abs().max() > 0:
    log("row0 got ", o[0, :16].tolist())
    log("row0 want", ref[0, :16].tolist())
    log("row1 got ", o[1, :16].tolist())
    log("row9 got ", o[9, :16].tolist())
    log("col-sum got", o.sum(0)[:8].tolist(), "want", ref.sum(0)[:8].tolist())
for prec in ("bf16", "bf16x3"):
    run(128, 256, 64, prec, tag="ints")
    run(128, 256, 256, prec, tag="ints-k256")
    run(1000, 768, 259, prec, tag="ints-ragged")
g = torch.Generator().manual_seed(2)
a = torch.randn(4096, 256, generator=g); w = torch.randn(1024, 256, generator=g) / 16
for prec in ("bf16", "bf16x3"):
    o, ref = run(4096, 1024, 256, prec, a, w, "randn")
    log(f"   rel err vs fp64: {((o - ref).abs().max() / ref.abs().max()).item():.3e}")
# timing
for (M, N, K) in ((148 * 128 * 8, 256, 256), (148 * 128 * 8, 1024, 256), (148 * 128 * 8, 256, 1024), (148 * 128 * 8, 768, 256)):
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev)
    pw = ops.PackedWeight(w)
    c = torch.empty(M, N, device=dev)
    for prec in ("bf16x3", "bf16"):
        for _ in range(3):
            ops.gemm_tc(a, pw, out=c, precision=prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm_tc(a, pw, out=c, precision=prec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        log(f"[time] M={M} N={N} K={K} {prec}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s (algorithmic)")
    for _ in range(3):
        ops.gemm(a, w, out=c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.gemm(a, w, out=c)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    log(f"[time] M={M} N={N} K={K} f32-ffma: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/diag_gemm_tc.txt", "w").write("\n".join(out_lines) + "\n")

# ---- chained kernels ------------------------------------------------------------------------------------
try:
    M = 148 * 128 * 8
    g = torch.Generator().manual_seed(3)
    w1 = (torch.randn(1024, 256, generator=g) / 16).to(dev); w2 = (torch.randn(256, 1024, generator=g) / 32).to(dev)
    b1 = torch.zeros(1024, device=dev); b2 = torch.zeros(256, device=dev)
    lw, lb = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    mats = []
    for gi in range(4):
        mats += [w1[256 * gi:256 * (gi + 1), :], w2[:, 256 * gi:256 * (gi + 1)]]
    blob = ops.pack_tiles(mats)
    x0 = torch.randn(M, 256, device=dev)
    for prec in ("bf16x3", "bf16"):
        x = x0.clone()
        for _ in range(2):
            ops.chain_mlp(x, lw, lb, 1e-6, blob, b1, b2, prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.chain_mlp(x, lw, lb, 1e-6, blob, b1, b2, prec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        log(f"[time] chain_mlp M={M} {prec}: {ms:.3f} ms  {2 * M * 2 * 256 * 1024 / ms / 1e9:.1f} TFLOP/s (algorithmic)  {M / ms / 1e3:.1f} Mpts/s")
    from zeroshape_b200.model.shape.implicit import Implicit
    net = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
                   pos_perlayer=False).to(dev).eval()
    _, _, occ_blob, biases, w8, b8, _, _ = net._chain_blobs()
    pts = torch.rand(M, 3, device=dev)
    for prec in ("bf16x3", "bf16"):
        for _ in range(2):
            ops.chain_occ(x0, pts, None, None, 1e-6, occ_blob, biases, w8, b8, precision=prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.chain_occ(x0, pts, None, None, 1e-6, occ_blob, biases, w8, b8, precision=prec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        log(f"[time] chain_occ M={M} {prec}: {ms:.3f} ms  {2 * M * 724224 / ms / 1e9:.1f} TFLOP/s (algorithmic)  {M / ms / 1e3:.1f} Mpts/s")
except Exception as ex:
    log("chain diag failed:", repr(ex))
open("gpurun_out/diag_gemm_tc.txt", "w").write("\n".join(out_lines) + "\n")

# ---- attention variants -----------------------------------------------------------------------------------
try:
    M = 148 * 128 * 8
    L, C, H = 197, 256, 8
    qkv = torch.randn(M, 3 * C, device=dev)
    latq = torch.randn(1, L, 3 * C, device=dev)
    kp, vp = ops.attn_pack_kv(latq[0, :, C:2 * C], latq[0, :, 2 * C:], H)
    kb, vf = ops.attn_pack_fused(latq[0, :, C:2 * C], latq[0, :, 2 * C:], H)
    def timeit(fn, n=5):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    for prec in ("bf16x3", "bf16"):
        log(f"[time] attn_fused M={M} {prec}: {timeit(lambda: ops.attn_fused(qkv, kb, vf, L, 32 ** -0.5, prec)):.3f} ms")
    log(f"[time] attn_tc (2 kernels) M={M} bf16x3: {timeit(lambda: ops.attn_tc(qkv, kp, vp, L, 32 ** -0.5)):.3f} ms")
    log(f"[time] point_attention f32 M={M}: {timeit(lambda: ops.point_attention(qkv.view(1, M, 3 * C), latq[..., C:2 * C], latq[..., 2 * C:], H), 2):.3f} ms")
except Exception as ex:
    log("attention diag failed:", repr(ex))
open("gpurun_out/diag_gemm_tc.txt", "w").write("\n".join(out_lines) + "\n")
