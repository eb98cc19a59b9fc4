#!/usr/bin/env python
"""Device-time breakdown of one round of the brute-force pose search (utils/eval_3D.brute_force_search, method "bvh"):
rotate + normalise (torch), hierarchy build, the two query directions, statistics -- and the dense Chamfer round for scale."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zeroshape_b200 import ops                                    # noqa: E402
from zeroshape_b200.utils import eval_3D                          # noqa: E402
from zeroshape_b200.utils.camera import get_rotation_sphere      # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    n = 10000
    gt = torch.randn(1, n, 3, generator=g)
    gt = (gt / gt.norm(dim=-1, keepdim=True) * torch.tensor([0.5, 0.3, 0.4])).to(dev)
    pred = torch.randn(1, n, 3, generator=g)
    pred = (pred / pred.norm(dim=-1, keepdim=True) * torch.tensor([0.35, 0.5, 0.3]) + 0.02 * torch.randn(1, n, 3, generator=g)).to(dev)
    gt = eval_3D.normalize_pc(gt).contiguous()
    R = get_rotation_sphere(24, 24, 12, [1.0], dev)[1000:1288]
    thr = [0.005, 0.01, 0.02, 0.05, 0.1, 0.2]
    t_rot, rot = timed(lambda: eval_3D.normalize_pc((R @ pred.permute(0, 2, 1)).permute(0, 2, 1)).contiguous())
    gt_bvh = ops.NNBvh(gt)
    gt_order = gt_bvh.morton_order()
    pred_order = ops.NNBvh(pred.contiguous()).morton_order()
    t_build, bvh = timed(lambda: ops.NNBvh(rot))
    t_q1w, _ = timed(lambda: gt_bvh.query(rot, q_order=pred_order, variant=1))
    t_q2w, _ = timed(lambda: bvh.query(gt, batch=R.shape[0], q_order=gt_order, variant=1))
    t_q1, (d1, _) = timed(lambda: gt_bvh.query(rot, q_order=pred_order))
    t_q1u, _ = timed(lambda: gt_bvh.query(rot))
    t_q2, (d2, _) = timed(lambda: bvh.query(gt, batch=R.shape[0], q_order=gt_order))
    t_stats, _ = timed(lambda: ops.chamfer_stats(d1, d2, thr))
    t_dense, _ = timed(lambda: ops.chamfer_nn(rot[:24].contiguous(), gt.expand(24, -1, -1).contiguous()), reps=2)
    print(f"one round of {R.shape[0]} rotations x {n} points (ms): rotate+normalise {t_rot:.3f} | build {t_build:.3f} | "
          f"query pred->gt {t_q1:.3f} (unordered {t_q1u:.3f}) | query gt->pred {t_q2:.3f} | stats {t_stats:.3f} "
          f"|| dense Chamfer of 24 rotations {t_dense:.3f} ms = {t_dense * R.shape[0] / 24:.1f} ms per {R.shape[0]}")
    print(f"warp-cooperative queries: pred->gt {t_q1w:.3f} ms, gt->pred {t_q2w:.3f} ms")
    print(f"mean sqrt(d1) {d1.sqrt().mean().item():.4f}  mean sqrt(d2) {d2.sqrt().mean().item():.4f}")
    for variant in (0, 1):
        ops.BVH_QUERY_VARIANT = variant
        t_v, _ = timed(lambda: eval_3D.brute_force_search(pred[0], gt[0], device=dev), reps=2)
        print(f"brute_force_search, query variant {variant}: {t_v:.1f} ms per shape")
    ops.BVH_QUERY_VARIANT = 0
    t_all, _ = timed(lambda: eval_3D.brute_force_search(pred[0], gt[0], device=dev), reps=2)
    print(f"brute_force_search (6912 rotations): {t_all:.1f} ms per shape")


if __name__ == "__main__":
    main()
