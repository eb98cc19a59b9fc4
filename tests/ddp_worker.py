"""Worker of tests/test_gpu_ddp.py: one rank of a 2-rank `DistributedDataParallel(Graph)` training step
(model/shape_engine.py:71, 248-271), gloo process group so that both ranks can share ONE GPU (NCCL refuses two ranks on a
device; on the multi-GPU box the same code runs with backend "nccl", one rank per GPU -- bench.py --mode train-ddp)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_case(B, seed, dev):
    from zeroshape_b200.utils.util import EasyDict
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(B, 3, 224, 224, generator=g)
    yy, xx = torch.meshgrid(torch.arange(224), torch.arange(224), indexing="ij")
    mask = (((yy - 112) ** 2 + (xx - 108) ** 2) < 78 ** 2).float().view(1, 1, 224, 224).repeat(B, 1, 1, 1)
    rgb = rgb * mask + (1 - mask)
    depth = (1.5 + 0.3 * torch.rand(B, 1, 224, 224, generator=g)) * mask
    intr = torch.tensor([[1.3875 * 224, 0, 112], [0, 1.3875 * 224, 112], [0, 0, 1.0]]).repeat(B, 1, 1)
    pose = torch.cat([torch.eye(3), torch.tensor([[0.0], [0.0], [1.6]])], dim=1).repeat(B, 1, 1)
    pts = torch.rand(B, 512, 3, generator=g) - 0.5
    sdf = pts.norm(dim=-1) - 0.3
    t = dict(rgb_input_map=rgb, mask_input_map=mask, depth_input_map=depth, intr=intr, pose_gt=pose, gt_sample_points=pts,
             gt_sample_sdf=sdf)
    return lambda lo, hi: EasyDict(idx=torch.arange(hi - lo), **{k: v[lo:hi].to(dev) for k, v in t.items()})


def make_opt(dev):
    from zeroshape_b200.utils.util import EasyDict
    return EasyDict(device=dev, H=224, W=224, pretrain=dict(depth=None), optim=dict(fix_dpt=False),
                    arch=dict(num_heads=8, latent_dim=256, win_size=16,
                              depth=dict(encoder="resnet", n_blocks=12, dsp=2, pretrained=None), rgb=dict(encoder=None, n_blocks=12),
                              impl=dict(n_channels=256, att_blocks=2, mlp_ratio=4., posenc_perlayer=False, mlp_layers=8,
                                        posenc_3D=0, skip_in=[2, 4, 6])),
                    loss_weight=dict(depth=None, intr=None, shape=1),
                    training=dict(shape_loss=dict(impt_thres=0.01, impt_weight=1)),
                    eval=dict(vox_res=16, range=[-1.5, 1.5], num_points=1000, brute_force=False, icp=False,
                              f_thresholds=[0.005, 0.01, 0.02, 0.05, 0.1, 0.2]), data=dict(dataset_test="synthetic"))


def build_graph(dev, seed=0):
    from zeroshape_b200.model.compute_graph.graph_shape import Graph
    torch.manual_seed(seed)
    graph = Graph(make_opt(dev)).to(dev).train()
    with torch.no_grad():
        graph.intr_proj.weight.normal_(0, 0.02)
    for m in graph.modules():                      # DropPath off: the two runs must see the same function
        if hasattr(m, "drop_path") and isinstance(getattr(m, "drop_path"), float):
            m.drop_path = 0.0
    return graph


def worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torch.nn.parallel import DistributedDataParallel as DDP
        from zeroshape_b200 import ops
        ops.TRAIN_ENGINE = "f32"
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        graph = build_graph(dev)
        ddp = DDP(graph, device_ids=[0], find_unused_parameters=True)      # shape_engine.py:71
        opt = make_opt(dev)
        var = make_case(world, 7, dev)(rank, rank + 1)                     # DistributedSampler shard: one image per rank
        var, loss = ddp(opt, var, training=True, get_loss=True)            # shape_engine.py:253
        loss.shape.backward()                                              # :258-271; DDP all-reduces (averages) in its hooks
        grads = {n: p.grad.detach().cpu() for n, p in graph.named_parameters() if p.grad is not None}
        if rank == 0:
            torch.save({"grads": grads, "loss": float(loss.shape)}, out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()
