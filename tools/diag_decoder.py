"""Per-op device times (CUDA events, warm) and SM-clock probes of one decoder pass on a real B200.
    python tools/diag_decoder.py [points]      -> gpurun_out/diag_decoder.txt
Not a test.  Shows (a) where a pass of `points` query points spends its time, (b) the effective SM clock right after
every kernel (power-cap droop), (c) old (layernorm + gemm_tc) vs fused (chain_lin) qkv / proj."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zeroshape_b200 import ops  # noqa: E402
from zeroshape_b200._native import lib  # noqa: E402
from zeroshape_b200.model.shape.implicit import Implicit  # noqa: E402

dev = torch.device("cuda:0")
ONCE = "--once" in sys.argv          # ncu mode: 2 warm passes + 1 pass, nothing else
_args = [a for a in sys.argv[1:] if not a.startswith("--")]
P = int(_args[0]) if _args else 15 * 129 * 129
lines = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    lines.append(s)


torch.manual_seed(0)
net = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
               pos_perlayer=False).to(dev).eval()
lat_in = torch.randn(1, 197, 256, device=dev)
pts = (torch.rand(1, P, 3, device=dev) * 3 - 1.5).contiguous()
FLOP = {"chain_lin[qkv]": 2 * 196608, "chain_lin[proj]": 2 * 65536, "attn_fused": 2 * 2 * (50432 + 256), "chain_mlp": 2 * 524288,
        "chain_occ": 2 * 724224, "chain_pmlp": 2 * (524288 + 65536), "gemm_tc": 0, "point_proj": 2 * 768, "chain_qkvattn": 2 * (196608 + 2 * (50432 + 256))}
with torch.no_grad():
    lat = net.prepare_latents(lat_in)
    if os.environ.get("ZS_CHAIN_DBG"):
        lib.zs_debug_chain_variant(int(os.environ["ZS_CHAIN_DBG"]))
    if ONCE:
        for _ in range(3):
            net._points_chain(lat, pts, tc=True, sigmoid=True)
        torch.cuda.synchronize()
        sys.exit(0)
    variants = [("qkv", 24, 1), ("qkv", 24, 0), ("qkv", 24, 1), ("qkv", 24, 0)]     # third field: point_proj folded into the first block
    for attention, flags, mlp_variant in variants:
        fused = True
        net.attention, net.attn_flags = attention, flags
        net.fold_point_proj = bool(mlp_variant)
        mlp_variant = 1
        lib.zs_debug_chain_variant(mlp_variant | (int(os.environ.get("ZS_CHAIN_DBG", "0")) & ~1))
        for _ in range(3):
            net._points_chain(lat, pts, tc=True, sigmoid=True)
        torch.cuda.synchronize()
        reps = 5
        with ops.OpTimer(clock_probe=False) as t:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                net._points_chain(lat, pts, tc=True, sigmoid=True)
            e1.record()
        summ = t.summary()
        total = e0.elapsed_time(e1) / reps
        log(f"\n== attention={attention} flags={flags} mlp/occ variant {mlp_variant}: {P} points, pass {total:.3f} ms ({P / total / 1e3:.1f} Mpts/s; x{2146689 / P:.2f} = {total * 2146689 / P:.1f} ms per 129^3 shape)")
        for k, (c, ms) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
            per = ms / c
            tf = FLOP.get(k, 0) * P / per / 1e9 if per > 0 else 0
            log(f"   {k:18s} x{c // reps}  {per * 1e3:8.1f} us/launch  {ms / reps:7.3f} ms/pass  {100 * ms / reps / total:5.1f}%   {tf:7.1f} TFLOP/s algorithmic ({3 * tf:.0f} executed at 3 passes)")
        with ops.OpTimer(clock_probe=True) as t:
            for _ in range(3):
                net._points_chain(lat, pts, tc=True, sigmoid=True)
        tr = t.clock_trace()
        log("   SM MHz after each kernel (3rd pass):", ", ".join(f"{n}:{v:.0f}" for n, v in tr[-len(tr) // 3:]))
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/diag_decoder.txt", "w").write("\n".join(lines) + "\n")
