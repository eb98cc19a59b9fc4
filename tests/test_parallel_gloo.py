"""CPU, world_size 2 (gloo): the N>1 path -- slab partitioning of the query grid + all_gather, and the
metric gather -- with the kernels replaced by CPU stand-ins (tests/fake_ops.py)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_cover_grid_exactly():
    from zeroshape_b200.parallel import all_slab_bounds
    for n in (2, 9, 65, 129):
        for world in (1, 2, 3, 4, 8):
            b = all_slab_bounds(n, world)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [x1 - x0 for x0, x1 in b]
            assert max(sizes) - min(sizes) <= 1
    assert all_slab_bounds(129, 8)[0] == (0, 17) and all_slab_bounds(129, 8)[7] == (113, 129)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_ops
        import zeroshape_b200.ops as ops
        for name in dir(fake_ops):
            if not name.startswith("_") and hasattr(ops, name) and callable(getattr(fake_ops, name)) and name != "install":
                setattr(ops, name, getattr(fake_ops, name))
        from zeroshape_b200.model.shape.implicit import Implicit
        from zeroshape_b200.parallel import sharded_grid_occupancy, gather_metrics
        from oracle.implicit import implicit_init
        net = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8,
                       skip_in=[2, 4, 6], pos_perlayer=False)
        net.load_state_dict(implicit_init(seed=7))
        net.eval()
        lat = torch.randn(1, 197, 256, generator=torch.Generator().manual_seed(8))
        full = sharded_grid_occupancy(net, lat, 7, -1.5, 1.5)
        single = net.grid_occupancy(lat, 7, -1.5, 1.5)
        metrics = gather_metrics(torch.full((2, 3), float(rank)))
        # (CPU matmul blocking depends on the batch shape, so slabs agree to rounding here; the real kernels
        #  are bit-identical across slab splits -- asserted in tests/test_gpu_implicit.py)
        ok = torch.allclose(full, single, rtol=0, atol=1e-6) and metrics.shape == (2 * world, 3) and metrics[2 * rank, 0].item() == rank
        torch.save({"ok": ok, "shape": tuple(full.shape)}, os.path.join(out, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_sharded_grid_equals_single_rank(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert res["ok"] and res["shape"] == (1, 7, 7, 7)


def _ddp_worker(rank, world, port, out):
    """DistributedDataParallel over the decoder's custom autograd tape (CPU stand-ins for the kernels): DDP's gradient hooks
    must see every parameter gradient the hand-written backward returns."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_ops
        import zeroshape_b200.ops as ops
        for name in fake_ops.TRAIN_OPS:
            setattr(ops, name, getattr(fake_ops, name))
        from torch.nn.parallel import DistributedDataParallel as DDP
        from zeroshape_b200.model.shape.implicit import Implicit
        from oracle.implicit import implicit_init
        net = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8,
                       skip_in=[2, 4, 6], pos_perlayer=False, drop_path=0.0)
        net.load_state_dict(implicit_init(seed=31))
        net.train()
        ddp = DDP(net)
        g = torch.Generator().manual_seed(9)
        lat, pts = torch.randn(world, 197, 256, generator=g), torch.rand(world, 23, 3, generator=g) * 2 - 1
        wgt = torch.randn(world, 23, generator=g)
        logits, _ = ddp(lat[rank:rank + 1], None, pts[rank:rank + 1])
        (logits * wgt[rank:rank + 1]).sum().backward()
        if rank == 0:
            torch.save({n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}, out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_ddp_over_decoder_tape_gloo(tmp_path):
    """world_size 2: DDP-averaged gradients == mean of the per-sample gradients of torch autograd over the oracle."""
    import socket
    from oracle.implicit import implicit_forward, implicit_init
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "g.pt")
    mp.spawn(_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    sd = implicit_init(seed=31)
    g = torch.Generator().manual_seed(9)
    lat, pts = torch.randn(2, 197, 256, generator=g), torch.rand(2, 23, 3, generator=g) * 2 - 1
    wgt = torch.randn(2, 23, generator=g)
    ref = {k: v.clone().requires_grad_(k != "pos_embed") for k, v in sd.items()}
    logits, _ = implicit_forward(ref, lat, pts)
    ((logits * wgt).sum() / 2).backward()
    assert len(got) == len(ref) - 1
    for k, v in got.items():
        assert ((v - ref[k].grad).norm() / ref[k].grad.norm().clamp_min(1e-30)).item() < 2e-4, k


def _mesh_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_ops
        import zeroshape_b200.ops as ops
        for name in dir(fake_ops):
            if not name.startswith("_") and hasattr(ops, name) and callable(getattr(fake_ops, name)) and not name.startswith("install"):
                setattr(ops, name, getattr(fake_ops, name))
        from zeroshape_b200.model.shape.implicit import Implicit
        from zeroshape_b200.parallel import sharded_meshes, slab_meshes, face_set
        from oracle.implicit import implicit_init
        net = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8,
                       skip_in=[2, 4, 6], pos_perlayer=False)
        net.load_state_dict(implicit_init(seed=7))
        net.eval()
        lat = torch.randn(2, 197, 256, generator=torch.Generator().manual_seed(8))
        n = 9
        meshes = sharded_meshes(net, lat, n, -1.5, 1.5)          # slab decode + per-slab MC + mesh all_gather
        single = slab_meshes(net, lat, n, -1.5, 1.5, 0, 1)        # the unsharded meshes
        ok = all(face_set(v, f, 4) == face_set(sv, sf, 4) and f.shape[0] == sf.shape[0] and f.shape[0] > 0
                 for (v, f), (sv, sf) in zip(meshes, single))
        torch.save({"ok": ok, "faces": [int(f.shape[0]) for _, f in meshes]}, os.path.join(out, f"m{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_sharded_meshes_equal_single_rank_meshes(tmp_path):
    """world_size 2 (gloo): per-slab marching cubes with a one-slice halo + all_gather of the mesh parts reproduces the
    unsharded mesh of every shape as a SET of triangles (seam vertices are duplicated, faces are not)."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_mesh_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = [torch.load(os.path.join(tmp_path, f"m{r}.pt")) for r in range(2)]
    assert all(r["ok"] for r in res) and res[0]["faces"] == res[1]["faces"]
