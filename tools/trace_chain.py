"""Role-level timelines of the chained tcgen05 kernels (CTA 0): clock64 events of the MMA thread, loader thread 0 and
epilogue warp 4 lane 0, recorded through zs_debug_chain_trace.

Writes gpurun_out/trace_<kernel>_<prec>.txt (events sorted by time, cycles relative to the first event) and a
per-role summary of where the cycles go.  Debug tool -- not part of the product path.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zeroshape_b200 import ops  # noqa: E402
from zeroshape_b200._native import check, lib  # noqa: E402

TAGS_MLP = {
    1: "mma: wait lfull", 2: "mma: lfull ok", 3: "mma: wfull(hi) ok", 4: "mma: hi issued", 5: "mma: wfull(lo) ok",
    6: "mma: chunk issued", 7: "mma: wait tempty", 8: "mma: wait efull", 9: "mma: efull ok",
    10: "ld : wait lempty", 11: "ld : lempty ok", 12: "ld : stored+arrived", 13: "ld : next fetched", 14: "ld : tile start",
    15: "ld : stats+first fetch done",
    20: "epi: wait tfull", 21: "epi: tfull ok", 22: "epi: tmem ld done", 23: "epi: eempty ok", 24: "epi: stored+arrived",
}
TAGS_ATTN = {
    1: "mma: pair start (wait lfull)", 2: "mma: S issued", 3: "mma: PV chunk issued",
    10: "ld : fetched, wait lempty", 11: "ld : lempty ok", 12: "ld : stored+arrived",
    20: "epi: wait sfull", 21: "epi: sfull ok", 22: "epi: pass1 done", 23: "epi: chunk stored", 24: "epi: wait ofull",
    25: "epi: ofull ok", 26: "epi: out stored",
}
TAGS_QKVATTN = {
    1: "mma: qkv chunk: wait lfull", 2: "mma: qkv chunk: lfull ok", 3: "mma: qkv chunk issued", 4: "mma: head start (wait K tile)",
    5: "mma: S issued", 6: "mma: V tile ok (wait oempty, pfull)", 7: "mma: pfull ok", 8: "mma: PV issued",
    10: "ld : wait lempty", 11: "ld : lempty ok", 12: "ld : stored+arrived", 14: "ld : tile start", 15: "ld : stats+first fetch done",
    20: "epi: wait accfull", 21: "epi: accfull ok", 22: "epi: epi-1 done / wait sfull", 23: "epi: sfull ok", 24: "epi: max pass done",
    25: "epi: bar1 passed", 26: "epi: exp pass done, pfull arrived", 27: "epi: bar2 passed", 28: "epi: ofull ok", 29: "epi: out stored",
}
TAGS_PMLP = {
    1: "mma: wait lfull", 2: "mma: lfull ok", 6: "mma: fc1/proj chunk issued", 7: "mma: group start (wait tempty0)", 8: "mma: wait efull",
    9: "mma: efull ok", 4: "mma: fc2 chunk issued",
    10: "ld : wait lempty", 11: "ld : lempty ok", 16: "ld : A chunks done, wait xready", 17: "ld : xready ok",
    20: "epi: wait tfull0", 21: "epi: tfull0 ok", 23: "epi: x' written, stats published", 24: "epi: GELU chunk handed over",
    25: "epi: wait tfull1", 26: "epi: tfull1 ok", 27: "epi: x written",
}
NE = 512


def run_traced(name, tags, fn):
    dev = torch.device("cuda:0")
    tr = torch.zeros(3 * NE, dtype=torch.int64, device=dev)
    for _ in range(2):  # second call = warm
        tr.zero_()
        check(lib.zs_debug_chain_trace(tr.data_ptr()), "trace on")
        try:
            fn()
            torch.cuda.synchronize()
        finally:
            check(lib.zs_debug_chain_trace(None), "trace off")
    ev = []
    for role in range(3):
        for v in tr[role * NE:(role + 1) * NE].tolist():
            if v:
                ev.append(((v >> 8) & ((1 << 56) - 1), role, v & 0xff))
    ev.sort()
    t0 = ev[0][0]
    lines = [f"{t - t0:9d}  r{role}  {tags.get(tag, tag)}" for t, role, tag in ev]
    summ = {}   # per role: time between consecutive events, attributed to the LATER event's tag
    for role in range(3):
        es = [(t, tag) for t, r, tag in ev if r == role]
        for (ta, _), (tb, tagb) in zip(es, es[1:]):
            k = tags.get(tagb, str(tagb))
            c, n = summ.get(k, (0, 0))
            summ[k] = (c + tb - ta, n + 1)
    out = [f"{name}: cycles until each event since the previous event of the same role"]
    for k in sorted(summ):
        c, n = summ[k]
        out.append(f"  {k:32s} total {c:9d}  n {n:4d}  avg {c / max(n, 1):9.1f}")
    out.append(f"  traced span {ev[-1][0] - t0} cycles")
    open(f"gpurun_out/trace_{name}.txt", "w").write("\n".join(out + [""] + lines) + "\n")
    print("\n".join(out))


def main():
    dev = torch.device("cuda:0")
    M = 148 * 128 * 8
    g = torch.Generator().manual_seed(3)
    w1 = (torch.randn(1024, 256, generator=g) / 16).to(dev)
    w2 = (torch.randn(256, 1024, generator=g) / 32).to(dev)
    b1 = torch.zeros(1024, device=dev)
    b2 = torch.zeros(256, device=dev)
    mats = []
    for gi in range(4):
        mats += [w1[256 * gi:256 * (gi + 1), :], w2[:, 256 * gi:256 * (gi + 1)]]
    blob = ops.pack_tiles(mats)
    x0 = torch.randn(M, 256, device=dev)
    os.makedirs("gpurun_out", exist_ok=True)
    which = sys.argv[1:] or ["mlp", "attn"]
    if "qkvattn" in which:
        L, C, H = 197, 256, 8
        lat = torch.randn(L, 2 * C, generator=g).to(dev)
        kb, vb = ops.attn_pack_fused(lat[:, :C], lat[:, C:], H)
        wq = (torch.randn(768, 256, generator=g) / 16).to(dev)
        wb = ops.qkvattn_pack(wq)
        bq = torch.zeros(768, device=dev)
        out = torch.empty(M, C, device=dev)
        for flags in (8, 24):
            run_traced(f"chain_qkvattn_flags{flags}", TAGS_QKVATTN,
                       lambda: ops.chain_qkvattn(x0, wb, bq, kb, vb, L, 32 ** -0.5, flags=flags, out=out))
    if "pmlp" in which:
        a_blk = torch.randn(M // 128, 64, 128, 4, generator=g).to(dev)
        wp = (torch.randn(256, 256, generator=g) / 16).to(dev)
        xp = x0.clone()
        run_traced("chain_pmlp", TAGS_PMLP, lambda: ops.chain_pmlp(xp, a_blk, ops.pack_generic(wp), b2, 1e-6, blob, b1, b2))
    if "mlp" in which and "pmlp" not in which:
        x = x0.clone()
        run_traced("chain_mlp_bf16x3", TAGS_MLP, lambda: ops.chain_mlp(x, None, None, 1e-6, blob, b1, b2, "bf16x3"))
    if "lin" in which:
        wq = (torch.randn(768, 256, generator=g) / 16).to(dev)
        bq = torch.zeros(768, device=dev)
        qblob = ops.pack_generic(wq)
        out = torch.empty(M, 768, device=dev)
        tags = dict(TAGS_MLP)
        tags[24] = "epi: 64-col group stored"
        run_traced("chain_lin_qkv_bf16x3", tags, lambda: ops.chain_lin(x0, qblob, bq, 3, do_ln=True, out=out))
        wp = (torch.randn(256, 256, generator=g) / 16).to(dev)
        pblob = ops.pack_generic(wp)
        xr = x0.clone()
        run_traced("chain_lin_proj_bf16x3", tags, lambda: ops.chain_lin(x0, pblob, bq[:256], 1, res=xr, out=xr))
    if "attn" in which:
        L, C, H = 197, 256, 8
        lat = torch.randn(L, 2 * C, generator=g).to(dev)
        kb, vb = ops.attn_pack_fused(lat[:, :C], lat[:, C:], H)
        qkv = torch.randn(M, 3 * C, generator=g).to(dev)
        out = torch.empty(M, C, device=dev)
        run_traced("chain_attn_bf16x3", TAGS_ATTN, lambda: ops.attn_fused(qkv, kb, vb, L, 32 ** -0.5, "bf16x3", out=out))


if __name__ == "__main__":
    main()
