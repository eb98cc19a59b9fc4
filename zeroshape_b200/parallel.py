"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed; NCCL over NVLink on
the B200 box, gloo in the CPU tests).  SURVEY.md section 8(e).

  A. latency  -- one shape, the (N+1)^3 query grid split into contiguous x-slabs (the reference's own slice
                 axis, utils/eval_3D.py:34-39), every rank decodes its slab, one all_gather of occupancy
                 slabs (8.6 MB fp32 per shape at 129^3) rebuilds the full grid on every rank.
  B. throughput -- shape-per-GPU (what the reference's DistributedSampler already does, data/base.py:12-14):
                 no data-path collective at all; only the tiny metric all_gather of
                 model/shape_engine.py:421-425.  This is what bench.py scales (weak scaling).
Query points are independent given the latents (model/shape/implicit.py:38-46), so no halo is needed for
the occupancy grid itself; per-slab marching cubes would need a 1-slice halo and is not done here (the
full grid is gathered instead -- it is 8.6 MB, ~12 us on NVLink 5).
"""
import torch
import torch.distributed as dist


def slab_bounds(n, world, rank):
    """Contiguous x-slices [x0,x1) of rank `rank`: the first n % world ranks get one extra slice."""
    base, extra = divmod(n, world)
    x0 = rank * base + min(rank, extra)
    return x0, x0 + base + (1 if rank < extra else 0)


def all_slab_bounds(n, world):
    return [slab_bounds(n, world, r) for r in range(world)]


@torch.no_grad()
def sharded_grid_occupancy(net, latent_depth, n, rmin, rmax, sigmoid=True, group=None):
    """Every rank holds the same latents [B,L,C]; returns the full [B,n,n,n] grid on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return net.grid_occupancy(latent_depth, n, rmin, rmax, 0, n, sigmoid=sigmoid)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = all_slab_bounds(n, world)
    x0, x1 = bounds[rank]
    B = latent_depth.shape[0]
    mine = net.grid_occupancy(latent_depth, n, rmin, rmax, x0, x1, sigmoid=sigmoid) if x1 > x0 else \
        latent_depth.new_zeros(B, 0, n, n)
    # all_gather needs equal shapes: pad every slab to the widest one (differs by at most one slice)
    width = max(b[1] - b[0] for b in bounds)
    padded = latent_depth.new_zeros(B, width, n, n)
    padded[:, :x1 - x0] = mine
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:, :b[1] - b[0]] for p, b in zip(parts, bounds)], dim=1)


def gather_metrics(local, group=None):
    """all_gather of per-sample metric tensors [n_local, ...] -> [n_total, ...] (shape_engine.py:413-429)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    parts = [torch.empty_like(local) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, local.contiguous(), group=group)
    return torch.cat(parts, dim=0)
