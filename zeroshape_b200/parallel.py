"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed; NCCL over NVLink on
the B200 box, gloo in the CPU tests).  SURVEY.md section 8(e).

  A. latency  -- one shape, the (N+1)^3 query grid split into contiguous x-slabs (the reference's own slice
                 axis, utils/eval_3D.py:34-39), every rank decodes its slab, one all_gather of occupancy
                 slabs (8.6 MB fp32 per shape at 129^3) rebuilds the full grid on every rank.
  B. throughput -- shape-per-GPU (what the reference's DistributedSampler already does, data/base.py:12-14):
                 no data-path collective at all; only the tiny metric all_gather of
                 model/shape_engine.py:421-425.  This is what bench.py scales (weak scaling).
Query points are independent given the latents (model/shape/implicit.py:38-46), so no halo is needed for
the occupancy grid itself.  `sharded_meshes` is the north-star variant of A (BASELINE config 4): every rank also
evaluates the ONE slice after its slab (the halo), runs marching cubes on its own slab -- every cell of the grid is
then owned by exactly one rank -- and the per-slab MESHES are all-gathered (counts first, then padded vertex / face
buffers; face indices are offset by the vertex prefix sum of the lower ranks; the y/z-edge vertices of a seam slice
exist in both neighbours' parts with bit-identical coordinates, faces are never duplicated).
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


@torch.no_grad()
def slab_meshes(net, latent_depth, n, rmin, rmax, rank, world, iso=0.5):
    """This rank's part of the meshes of all B shapes: occupancy over its x-slab [x0, x1) plus the halo slice x1, marching cubes
    per slab with vertices in GLOBAL index units.  One host sync for the B x 2 counts.  -> [(verts [V,3], faces [F,3] int32)] * B"""
    from . import ops
    x0, x1 = slab_bounds(n, world, rank)
    hi = min(x1 + 1, n)
    B = latent_depth.shape[0]
    if x1 <= x0 or hi - x0 < 2:          # more ranks than cell layers: nothing to own
        dev = latent_depth.device
        return [(torch.zeros(0, 3, device=dev), torch.zeros(0, 3, device=dev, dtype=torch.int32)) for _ in range(B)]
    vols, wss, cnts = [], [], []
    for b in range(B):
        lat = net.prepare_latents(latent_depth[b:b + 1])
        occ = net.grid_occupancy(latent_depth[b:b + 1], n, rmin, rmax, x0, hi, sigmoid=True, lat=lat)[0].contiguous()
        ws, c = ops.marching_cubes_count(occ, iso)
        vols.append(occ); wss.append(ws); cnts.append(c)
    counts = torch.stack(cnts).cpu()           # the one host sync of the mesh path
    return [ops.marching_cubes_emit(vols[b], iso, wss[b], int(counts[b, 0]), int(counts[b, 1]), x_offset=x0) for b in range(B)]


@torch.no_grad()
def gather_meshes(local, group=None):
    """all_gather of per-rank mesh parts [(verts, faces)] * B -> the welded-by-concatenation meshes [(verts, faces)] * B on every
    rank: two small collectives for the sizes, two for the padded vertex / face buffers (NCCL wants equal shapes)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    dev = local[0][0].device
    B = len(local)
    cnt = torch.tensor([[v.shape[0], f.shape[0]] for v, f in local], device=dev, dtype=torch.int64)      # [B,2]
    allc = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt, group=group)
    allc = torch.stack(allc).cpu()                                   # [world,B,2]
    vmax, fmax = int(allc[:, :, 0].sum(dim=1).max()), int(allc[:, :, 1].sum(dim=1).max())
    vbuf = torch.zeros(max(vmax, 1), 3, device=dev, dtype=torch.float32)
    fbuf = torch.zeros(max(fmax, 1), 3, device=dev, dtype=torch.int32)
    vcat = torch.cat([v for v, _ in local]) if vmax else vbuf[:0]
    fcat = torch.cat([f for _, f in local]) if fmax else fbuf[:0]
    vbuf[:vcat.shape[0]] = vcat
    fbuf[:fcat.shape[0]] = fcat
    vall = [torch.empty_like(vbuf) for _ in range(world)]
    fall = [torch.empty_like(fbuf) for _ in range(world)]
    dist.all_gather(vall, vbuf, group=group)
    dist.all_gather(fall, fbuf, group=group)
    voff = torch.cumsum(allc[:, :, 0], dim=1) - allc[:, :, 0]        # start of shape b inside rank r's vertex buffer
    foff = torch.cumsum(allc[:, :, 1], dim=1) - allc[:, :, 1]
    out = []
    for b in range(B):
        vs, fs, base = [], [], 0
        for r in range(world):
            V, F = int(allc[r, b, 0]), int(allc[r, b, 1])
            vs.append(vall[r][int(voff[r, b]):int(voff[r, b]) + V])
            fs.append(fall[r][int(foff[r, b]):int(foff[r, b]) + F] + base)      # indices into the concatenated vertex list
            base += V
        out.append((torch.cat(vs), torch.cat(fs)))
    return out


@torch.no_grad()
def sharded_meshes(net, latent_depth, n, rmin, rmax, iso=0.5, group=None):
    """Meshes of all B shapes (every rank holds the same latents [B,L,C]) from slab-sharded decoding + per-slab marching cubes
    + an all-gather of the mesh parts.  Vertices in global index units (scale like utils/eval_3D.py:252-255 afterwards)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return slab_meshes(net, latent_depth, n, rmin, rmax, 0, 1, iso)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    return gather_meshes(slab_meshes(net, latent_depth, n, rmin, rmax, rank, world, iso), group)


def face_set(verts, faces, decimals=5):
    """Order-free description of a mesh for equality tests: the set of triangles as sorted coordinate triples."""
    import numpy as np
    v = np.round(verts.detach().cpu().double().numpy(), decimals)
    f = faces.detach().cpu().numpy()
    return set(tuple(sorted(map(tuple, v[t]))) for t in f)
