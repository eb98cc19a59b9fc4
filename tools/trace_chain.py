"""Role-level timeline of chain_mlp_kernel (CTA 0): clock64 events of the MMA thread, loader row 0 and epilogue warp 4.

Writes gpurun_out/trace_chain_mlp_<prec>.txt (events sorted by time, cycles relative to the first event) and a
per-role summary of where the cycles go.  Debug tool -- not part of the product path.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zeroshape_b200 import ops  # noqa: E402
from zeroshape_b200._native import check, lib  # noqa: E402

TAGS = {
    1: "mma: wait lfull", 2: "mma: lfull ok", 3: "mma: wfull(hi) ok", 4: "mma: hi issued", 5: "mma: wfull(lo) ok",
    6: "mma: chunk issued", 7: "mma: wait tempty", 8: "mma: wait efull", 9: "mma: efull ok",
    10: "ld : wait lempty", 11: "ld : lempty ok", 12: "ld : stored+arrived", 13: "ld : next fetched", 14: "ld : tile start",
    15: "ld : stats+first fetch done",
    20: "epi: wait tfull", 21: "epi: tfull ok", 22: "epi: tmem ld done", 23: "epi: eempty ok", 24: "epi: stored+arrived",
}
NE = 512


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
    for prec in ("bf16x3", "bf16"):
        x = x0.clone()
        tr = torch.zeros(3 * NE, dtype=torch.int64, device=dev)
        for _ in range(2):  # second call = warm
            tr.zero_()
            check(lib.zs_chain_mlp_trace(x.data_ptr(), 256, M, 1e-6, blob.data_ptr(), b1.data_ptr(), b2.data_ptr(),
                                         ops.PRECISIONS[prec], tr.data_ptr(), None), "trace")
            torch.cuda.synchronize()
        ev = []
        for role in range(3):
            for v in tr[role * NE:(role + 1) * NE].tolist():
                if v:
                    ev.append(((v >> 8) & ((1 << 56) - 1), role, v & 0xff))
        ev.sort()
        t0 = ev[0][0]
        lines = [f"{t - t0:9d}  r{role}  {TAGS.get(tag, tag)}" for t, role, tag in ev]
        # per-role: time between consecutive events, attributed to the LATER event's tag
        summ = {}
        for role in range(3):
            es = [(t, tag) for t, r, tag in ev if r == role]
            for (ta, _), (tb, tagb) in zip(es, es[1:]):
                k = TAGS.get(tagb, str(tagb))
                c, n = summ.get(k, (0, 0))
                summ[k] = (c + tb - ta, n + 1)
        out = [f"chain_mlp trace, precision {prec}, M={M}; cycles until each event since the previous event of the same role"]
        for k in sorted(summ):
            c, n = summ[k]
            out.append(f"  {k:32s} total {c:9d}  n {n:4d}  avg {c / max(n, 1):9.1f}")
        span = ev[-1][0] - t0
        out.append(f"  traced span {span} cycles")
        open(f"gpurun_out/trace_chain_mlp_{prec}.txt", "w").write("\n".join(out + [""] + lines) + "\n")
        print("\n".join(out))


if __name__ == "__main__":
    main()
