"""Which gradient of the full training iteration goes non-finite first (small batch)?  python tools/diag_train_nan.py [--precision bf16]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--engine", default="tc")
    ap.add_argument("--steps", type=int, default=8)
    args = ap.parse_args()
    from zeroshape_b200 import ops
    from zeroshape_b200.model.depth import dpt_train as DT
    from zeroshape_b200.model.shape.implicit_train import FusedAdamW
    from zeroshape_b200.utils.util import EasyDict
    from test_gpu_graph import synthetic_image_and_mask
    from test_gpu_train import _graph_and_sd
    cuda = torch.device("cuda", 0)
    ops.TRAIN_ENGINE, ops.TRAIN_PRECISION = args.engine, args.precision
    B, N = 2, 512
    rgb, mask = synthetic_image_and_mask(B, 72)
    g = torch.Generator().manual_seed(73)
    depth_gt = (1.5 + 0.3 * torch.rand(B, 1, 224, 224, generator=g)) * mask
    intr = torch.tensor([[1.3875 * 224, 0, 112], [0, 1.3875 * 224, 112], [0, 0, 1.0]]).repeat(B, 1, 1)
    pose = torch.cat([torch.eye(3), torch.tensor([[0.0], [0.0], [1.6]])], dim=1).repeat(B, 1, 1)
    gt_pts = torch.rand(B, N, 3, generator=g) - 0.5
    gt_sdf = gt_pts.norm(dim=-1) - 0.3 - 0.003
    dev = [t.to(cuda) for t in (rgb, mask, depth_gt, intr, pose, gt_pts, gt_sdf)]
    opt, graph, _ = _graph_and_sd(cuda, 71)
    graph.train()
    graph.impl_network.drop_path = 0.0
    params = [p for p in graph.parameters() if p.requires_grad]
    optim = FusedAdamW(params, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05)
    # spy on the geometry backward
    orig = ops.unproject_normalize_bwd

    def spy(depth, mask_, K, seen, scale, dseen):
        dd, dkinv = orig(depth, mask_, K, seen, scale, dseen)
        print("   unproject_normalize_bwd: K finite", bool(torch.isfinite(K).all()), "| dseen finite", bool(torch.isfinite(dseen).all()),
              "absmax", float(dseen.abs().max()), "| dd finite", bool(torch.isfinite(dd).all()), "| dkinv", dkinv.flatten().tolist()[:9],
              "| scale", scale.flatten().tolist(), "| K0", K[0].flatten().tolist())
        return dd, dkinv
    ops.unproject_normalize_bwd = spy
    for it in range(args.steps):
        var = EasyDict(idx=torch.arange(B), rgb_input_map=dev[0], mask_input_map=dev[1], depth_input_map=dev[2], intr=dev[3],
                       pose_gt=dev[4], gt_sample_points=dev[5], gt_sample_sdf=dev[6])
        optim.zero_grad()
        var, loss = graph.forward(opt, var, training=True)
        loss.shape.backward()
        badg = [n for n, p in graph.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
        optim.step()
        badp = [n for n, p in graph.named_parameters() if not torch.isfinite(p).all()]
        print(f"step {it}: loss {loss.shape.item():.5f} | non-finite grads {len(badg)} {badg[:4]} | non-finite params {len(badp)} {badp[:4]}", flush=True)
        if badp:
            break


if __name__ == "__main__":
    main()
