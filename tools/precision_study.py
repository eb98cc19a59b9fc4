"""Per-GEMM operand-precision study of the implicit decoder's query path (VERDICT r1 task 4).

CPU emulation: every tensor-core contraction of the query path is evaluated with its operands rounded the way the
kernels round them (fp16 or bf16, hi/lo split or single term) and fp32 accumulation; the result is compared with an
fp64 evaluation of the same network (reference arithmetic: model/shape/implicit.py:25-79,168-184,251-288).

    python tools/precision_study.py [--n 33] [--seed 0] [--out profiles/r2_precision_study.md]

Modes of one GEMM  D = A W^T  (A = activations, W = weights / latent K,V):
    x3  : Ah Wh + Al Wh + Ah Wl      (3 MMA passes)
    x2a : (Ah + Al) Wh               (2 passes; W rounded to one term)
    x2w : Ah (Wh + Wl)               (2 passes; A rounded to one term)
    x1  : Ah Wh                      (1 pass)
TEST/BENCH infrastructure only: imports oracle/.
"""
import argparse
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.implicit import implicit_init, LN_EPS, SKIP_IN  # noqa: E402

GEMMS = ["b0.q", "b0.k", "b0.v", "b0.qk", "b0.pv", "b0.proj", "b0.fc1", "b0.fc2",
         "b1.q", "b1.k", "b1.v", "b1.qk", "b1.pv", "b1.proj", "b1.fc1", "b1.fc2",
         "occ0", "occ1", "occ2", "occ3", "occ4", "occ5", "occ6", "occ7"]


def split(x, dt):
    hi = x.to(dt).to(torch.float32)
    lo = (x - hi).to(dt).to(torch.float32)
    return hi, lo


def mm_emul(A, W, mode):
    """A [.., K] @ W[N, K]^T with the operand rounding of `mode` = (dtype, passes) or None (fp32)."""
    if mode is None:
        return A @ W.transpose(-1, -2)
    dt, p = mode
    Ah, Al = split(A, dt)
    Wh, Wl = split(W, dt)
    Wt_h, Wt_l = Wh.transpose(-1, -2), Wl.transpose(-1, -2)
    if p == "x3":
        return Ah @ Wt_h + (Al @ Wt_h + Ah @ Wt_l)
    if p == "x2a":
        return Ah @ Wt_h + Al @ Wt_h
    if p == "x2w":
        return Ah @ Wt_h + Ah @ Wt_l
    if p == "x1":
        return Ah @ Wt_h
    raise ValueError(p)


def decoder(sd, latent_depth, pts, policy, dtype=torch.float32, heads=8):
    """Query path of Implicit.forward for one image; `policy` maps GEMM name -> mode (missing = exact)."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    latent_depth, pts = latent_depth.to(dtype), pts.to(dtype)

    def mm(name, A, W):
        m = policy.get(name)
        if dtype == torch.float64 or m is None:
            return A @ W.transpose(-1, -2)
        return mm_emul(A, W, m)

    def ln(x, n):
        return F.layer_norm(x, (256,), sd[n + ".weight"], sd[n + ".bias"], LN_EPS)

    C, hd = 256, 32
    scale = hd ** -0.5
    lat = F.linear(latent_depth, sd["latent_proj.weight"], sd["latent_proj.bias"])[0] + sd["pos_embed"][0]   # [L, C]
    x = F.linear(pts, sd["point_proj.proj.weight"], sd["point_proj.proj.bias"])                              # [P, C]
    L = lat.shape[0]
    for b in range(2):
        pre = f"blocks_attn.{b}"
        Wqkv, bqkv = sd[pre + ".attn.qkv.weight"], sd[pre + ".attn.qkv.bias"]
        # latent side (exact here: per-image constant, 0.02 % of the work)
        ql = F.linear(ln(lat, pre + ".norm1"), Wqkv, bqkv)
        kl = ql[:, C:2 * C].reshape(L, heads, hd).permute(1, 0, 2)      # [h, L, d]
        vl = ql[:, 2 * C:].reshape(L, heads, hd).permute(1, 0, 2)
        qlat = ql[:, :C].reshape(L, heads, hd).permute(1, 0, 2)
        # query side
        xn = ln(x, pre + ".norm1")
        q = mm(f"b{b}.q", xn, Wqkv[:C]) + bqkv[:C]
        k = mm(f"b{b}.k", xn, Wqkv[C:2 * C]) + bqkv[C:2 * C]
        v = mm(f"b{b}.v", xn, Wqkv[2 * C:]) + bqkv[2 * C:]
        P = x.shape[0]
        qh = q.reshape(P, heads, hd).permute(1, 0, 2)                   # [h, P, d]
        s_cross = mm(f"b{b}.qk", qh, kl) * scale                        # [h, P, L]
        s_self = (q * k).reshape(P, heads, hd).sum(-1).t().unsqueeze(-1) * scale
        s = torch.cat([s_cross, s_self], -1)
        e = torch.exp(s - s.max(-1, keepdim=True).values)
        den = e.sum(-1, keepdim=True)
        o = mm(f"b{b}.pv", e[..., :L], vl.transpose(-1, -2))           # unnormalised, as the kernel does
        o = (o + e[..., L:] * v.reshape(P, heads, hd).permute(1, 0, 2)) / den
        o = o.permute(1, 0, 2).reshape(P, C)
        x = x + mm(f"b{b}.proj", o, sd[pre + ".attn.proj.weight"]) + sd[pre + ".attn.proj.bias"]
        h = F.gelu(mm(f"b{b}.fc1", ln(x, pre + ".norm2"), sd[pre + ".mlp.fc1.weight"]) + sd[pre + ".mlp.fc1.bias"])
        x = x + mm(f"b{b}.fc2", h, sd[pre + ".mlp.fc2.weight"]) + sd[pre + ".mlp.fc2.bias"]
        if b == 0:
            # latent rows of block 0 (needed as block 1's keys / values)
            al = torch.softmax(qlat @ kl.transpose(-1, -2) * scale, -1)
            ol = (al @ vl).permute(1, 0, 2).reshape(L, C)
            lat = lat + F.linear(ol, sd[pre + ".attn.proj.weight"], sd[pre + ".attn.proj.bias"])
            lat = lat + F.linear(F.gelu(F.linear(ln(lat, pre + ".norm2"), sd[pre + ".mlp.fc1.weight"], sd[pre + ".mlp.fc1.bias"])),
                                 sd[pre + ".mlp.fc2.weight"], sd[pre + ".mlp.fc2.bias"])
    feat = ln(x, "norm")
    inputs = torch.cat([feat, pts], -1)          # kernel K order: [feat | xyz]
    h = None
    r2 = 1.0 / math.sqrt(2.0)
    for l in range(8):
        W = sd[f"impl_mlp.layers.{l}.weight"]
        bl = sd[f"impl_mlp.layers.{l}.bias"]
        if l == 0:
            Wp = torch.cat([W[:, 3:], W[:, :3]], 1)
            z = mm(f"occ{l}", inputs, Wp)
        elif l in SKIP_IN:
            Wx = W[:, :256] * r2
            Wi = torch.cat([W[:, 259:], W[:, 256:259]], 1) * r2
            z = mm(f"occ{l}", torch.cat([h, inputs], -1), torch.cat([Wx, Wi], 1))
        else:
            z = mm(f"occ{l}", h, W)
        h = F.softplus(z + bl, beta=100.0)
    return F.linear(h, sd["impl_mlp.layers.8.weight"], sd["impl_mlp.layers.8.bias"]).squeeze(-1)


def metrics(a, ref):
    a, ref = a.double(), ref.double()
    floor = 0.25 * ref.pow(2).mean().sqrt()
    rel = ((a - ref).abs() / ref.abs().clamp_min(floor)).max().item()
    nw = ((a - ref).pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item()
    flips = int(((a > 0) != (ref > 0)).sum())
    return rel, nw, (a - ref).abs().max().item(), flips


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=33)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    torch.manual_seed(args.seed)
    sd = implicit_init(seed=args.seed)
    g = torch.Generator().manual_seed(args.seed + 1)
    lat = torch.randn(1, 197, 256, generator=g)
    ax = torch.linspace(-1.5, 1.5, args.n)
    pts = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)
    with torch.no_grad():
        ref = decoder(sd, lat, pts, {}, torch.float64)
        lines = []

        def report(tag, policy):
            out = decoder(sd, lat, pts, policy)
            rel, nw, mx, fl = metrics(out, ref)
            passes = sum({"x3": 3, "x2a": 2, "x2w": 2, "x1": 1}[policy[g_][1]] * MACS[g_] for g_ in GEMMS if policy.get(g_)) / sum(MACS.values())
            s = f"| {tag} | {rel:.2e} | {nw:.2e} | {mx:.2e} | {fl} | {passes:.2f} |"
            print(s, flush=True)
            lines.append(s)
            return rel, nw, fl

        hdr = "| policy | parity_rel | normwise | max abs | voxel flips | MMA passes (MAC-weighted) |\n|---|---|---|---|---|---|"
        print(hdr)
        lines.append(hdr)
        report("fp32 everywhere", {})
        for dt, nm in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
            for p in ("x3", "x2a", "x2w", "x1"):
                report(f"all GEMMs {nm} {p}", {g_: (dt, p) for g_ in GEMMS})
        # one GEMM degraded at a time on an fp16 x3 base
        base = {g_: (torch.float16, "x3") for g_ in GEMMS}
        single = {}
        for g_ in GEMMS:
            for p in ("x2a", "x2w", "x1"):
                pol = dict(base)
                pol[g_] = (torch.float16, p)
                single[(g_, p)] = report(f"fp16 x3, {g_} -> {p}", pol)
        # greedy mix: cheapest mode per GEMM whose single-GEMM error stays below a per-GEMM budget
        for budget in (2e-5, 5e-5, 1e-4):
            pol = dict(base)
            for g_ in GEMMS:
                for p in ("x1", "x2a", "x2w"):
                    if single[(g_, p)][0] < budget:
                        pol[g_] = (torch.float16, p)
                        break
            desc = ", ".join(f"{g_}:{pol[g_][1]}" for g_ in GEMMS if pol[g_][1] != "x3")
            report(f"greedy mix (per-GEMM parity_rel < {budget:g}): {desc}", pol)
    if args.out:
        with open(args.out, "w") as f:
            f.write(f"# Per-GEMM precision study (CPU emulation, {args.n}^3 grid, seed {args.seed}; tools/precision_study.py)\n\n")
            f.write("Reference = fp64 evaluation of the same network; fp32 accumulation emulated by fp32 matmuls of the rounded terms.\n\n")
            f.write("\n".join(lines) + "\n")


MACS = {}
for b_ in range(2):
    MACS.update({f"b{b_}.q": 65536, f"b{b_}.k": 65536, f"b{b_}.v": 65536, f"b{b_}.qk": 50432, f"b{b_}.pv": 50432,
                 f"b{b_}.proj": 65536, f"b{b_}.fc1": 262144, f"b{b_}.fc2": 262144})
for l_ in range(8):
    MACS[f"occ{l_}"] = 256 * (259 if l_ == 0 else 515 if l_ in SKIP_IN else 256)

if __name__ == "__main__":
    main()
