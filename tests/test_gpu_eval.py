"""GPU: marching cubes, surface sampling, Chamfer and F-score kernels against the oracle
(bit-exact for indices / integer work) and, when present, against the reference's own CUDA kernel
compiled unmodified into oracle/_ref/."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import eval3d as E

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _field(kind, n, seed=0):
    g = np.linspace(-1.5, 1.5, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    if kind == "sphere":
        return (np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 1.0).astype(np.float32)
    if kind == "torus":
        return (np.sqrt((np.sqrt(X ** 2 + Y ** 2) - 0.9) ** 2 + Z ** 2) - 0.35).astype(np.float32)
    return np.random.RandomState(seed).randn(n, n, n).astype(np.float32)


@pytest.mark.parametrize("kind,n", [("sphere", 33), ("torus", 65), ("rand", 20), ("rand", 2), ("sphere", 129)])
def test_marching_cubes_matches_oracle_exactly(cuda, kind, n):
    from zeroshape_b200 import ops
    vol = _field(kind, n)
    v_ref, f_ref = E.marching_cubes(vol, 0.0)
    v, f = ops.marching_cubes(torch.from_numpy(vol).to(cuda), 0.0)
    assert v.shape[0] == len(v_ref) and f.shape[0] == len(f_ref)
    np.testing.assert_array_equal(f.cpu().numpy().astype(np.int64), f_ref)            # same faces, same order
    np.testing.assert_allclose(v.cpu().numpy(), v_ref.astype(np.float32), rtol=0, atol=1e-5)


def test_mesh_future_defers_the_size_readback(cuda):
    """ops.MeshFuture: pass 1 + asynchronous size read-back at construction, pass 2 in result() -- the same mesh as
    ops.marching_cubes, also with other work queued in between and several futures in flight."""
    from zeroshape_b200 import ops
    vols = [torch.from_numpy(_field(k, n)).to(cuda) for k, n in (("sphere", 33), ("torus", 41), ("sphere", 17))]
    futures = [ops.MeshFuture(v, 0.0) for v in vols]
    filler = torch.randn(2048, 2048, device=cuda)
    for _ in range(4):
        filler = filler @ filler * 1e-3                      # the "next decoder" queued between the two passes
    for fut, vol in zip(futures, vols):
        v, f = fut.result()
        v0, f0 = ops.marching_cubes(vol, 0.0)
        assert torch.equal(v, v0) and torch.equal(f, f0)
    v, f = ops.MeshFuture(torch.full((9, 9, 9), 1.0, device=cuda), 0.0).result()
    assert v.shape == (0, 3) and f.shape == (0, 3)


def test_marching_cubes_empty_and_full(cuda):
    from zeroshape_b200 import ops
    for val in (-1.0, 1.0):
        v, f = ops.marching_cubes(torch.full((9, 9, 9), val, device=cuda), 0.0)
        assert v.shape == (0, 3) and f.shape == (0, 3)
    pts = ops.mesh_sample(v, f, 16)
    assert torch.equal(pts, torch.zeros(16, 3, device=cuda))      # utils/eval_3D.py:262


def test_mesh_sampling_distribution(cuda):
    from zeroshape_b200 import ops
    n = 65
    v, f = ops.marching_cubes(torch.from_numpy(_field("sphere", n)).to(cuda), 0.0)
    S = 50000
    pts = ops.mesh_sample(v, f, S, 3.0 / (n - 1), -1.5, seed=7).cpu().numpy()
    r = np.linalg.norm(pts, axis=1)
    assert abs(r.mean() - 1.0) < 5e-3 and r.max() < 1.003 and r.min() > 0.99
    assert np.abs(pts.mean(axis=0)).max() < 0.02
    # octant occupancy is uniform (area-weighted face choice)
    octant = ((pts[:, 0] > 0) * 4 + (pts[:, 1] > 0) * 2 + (pts[:, 2] > 0)).astype(int)
    frac = np.bincount(octant, minlength=8) / S
    assert np.abs(frac - 0.125).max() < 0.01
    # deterministic in the seed, different across seeds
    again = ops.mesh_sample(v, f, S, 3.0 / (n - 1), -1.5, seed=7).cpu().numpy()
    other = ops.mesh_sample(v, f, S, 3.0 / (n - 1), -1.5, seed=8).cpu().numpy()
    assert np.array_equal(pts, again) and not np.array_equal(pts, other)


@pytest.mark.parametrize("b,n,m", [(1, 10000, 10000), (3, 1000, 517), (2, 1, 2049), (24, 2500, 2500), (1, 5, 3)])
def test_chamfer_bit_exact_vs_oracle(cuda, b, n, m):
    from zeroshape_b200 import ops
    rs = np.random.RandomState(n + m)
    a = (rs.rand(b, n, 3) - 0.5).astype(np.float32)
    c = (rs.rand(b, m, 3) - 0.5).astype(np.float32)
    d1, d2, i1, i2 = E.chamfer_nn(a, c)
    g1, g2, j1, j2 = ops.chamfer_nn(torch.from_numpy(a).to(cuda), torch.from_numpy(c).to(cuda))
    np.testing.assert_array_equal(j1.cpu().numpy(), i1)
    np.testing.assert_array_equal(j2.cpu().numpy(), i2)
    np.testing.assert_array_equal(g1.cpu().numpy(), d1)       # bit-exact squared distances
    np.testing.assert_array_equal(g2.cpu().numpy(), d2)


def test_chamfer_ties_pick_lowest_index(cuda):
    from zeroshape_b200 import ops
    rs = np.random.RandomState(0)
    base = (rs.rand(1, 700, 3)).astype(np.float32)
    tgt = np.concatenate([base, base, base[:, ::-1]], axis=1)          # every point duplicated 3x
    _, _, i1, _ = E.chamfer_nn(base, tgt)
    _, _, j1, _ = ops.chamfer_nn(torch.from_numpy(base).to(cuda), torch.from_numpy(tgt).to(cuda))
    np.testing.assert_array_equal(j1.cpu().numpy(), i1)
    assert (i1 == np.arange(700)[None]).all()


def test_chamfer_vs_unmodified_reference_kernel(cuda):
    """The reference's own chamfer3D.cu, compiled unmodified for sm_100a into oracle/_ref/ (when the
    build container had /root/reference)."""
    from oracle.build_oracle import ref_chamfer_path, REF_OUT
    if not os.path.exists(ref_chamfer_path()):
        pytest.skip("oracle/_ref/chamfer_3D.so not built")
    sys.path.insert(0, REF_OUT)
    import chamfer_3D
    from zeroshape_b200 import ops
    rs = np.random.RandomState(5)
    for b, n, m in ((1, 10000, 10000), (4, 3000, 2000)):
        a = torch.from_numpy((rs.rand(b, n, 3) - 0.5).astype(np.float32)).to(cuda)
        c = torch.from_numpy((rs.rand(b, m, 3) - 0.5).astype(np.float32)).to(cuda)
        d1, d2 = torch.zeros(b, n, device=cuda), torch.zeros(b, m, device=cuda)
        i1 = torch.zeros(b, n, device=cuda, dtype=torch.int32)
        i2 = torch.zeros(b, m, device=cuda, dtype=torch.int32)
        torch.cuda.synchronize()
        assert chamfer_3D.forward(a, c, d1, d2, i1, i2) == 1
        torch.cuda.synchronize()
        g1, g2, j1, j2 = ops.chamfer_nn(a, c)
        assert torch.equal(j1, i1) and torch.equal(j2, i2)
        assert torch.equal(g1, d1) and torch.equal(g2, d2)


def test_chamfer_module_forward_backward(cuda):
    from zeroshape_b200.external.chamfer3D.dist_chamfer_3D import chamfer_3DDist
    rs = np.random.RandomState(2)
    a = torch.from_numpy(rs.rand(2, 300, 3).astype(np.float32)).to(cuda).requires_grad_(True)
    c = torch.from_numpy(rs.rand(2, 200, 3).astype(np.float32)).to(cuda).requires_grad_(True)
    d1, d2, i1, i2 = chamfer_3DDist()(a, c)
    assert i1.dtype == torch.int32 and d1.shape == (2, 300) and d2.shape == (2, 200)
    (d1.sum() + 2 * d2.sum()).backward()
    ga, gc = E.chamfer_grad(a.detach().cpu().numpy(), c.detach().cpu().numpy(), np.ones((2, 300), np.float32),
                            2 * np.ones((2, 200), np.float32), i1.cpu().numpy(), i2.cpu().numpy())
    np.testing.assert_allclose(a.grad.cpu().numpy(), ga, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(c.grad.cpu().numpy(), gc, rtol=1e-5, atol=1e-6)


def test_fscore_stats_and_brute_force(cuda):
    from zeroshape_b200.utils import eval_3D as ours
    rs = np.random.RandomState(3)
    d1 = torch.from_numpy(rs.rand(3, 1000).astype(np.float32) * 0.3)
    d2 = torch.from_numpy(rs.rand(3, 700).astype(np.float32) * 0.3)
    ref = E.fscore(d1, d2)
    out = ours.compute_fscore(d1.to(cuda), d2.to(cuda))
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=1e-6, atol=1e-7)
    pc = torch.from_numpy(rs.rand(2, 500, 3).astype(np.float32))
    np.testing.assert_allclose(ours.normalize_pc(pc.to(cuda)).cpu().numpy(), E.normalize_pc(pc).numpy(), atol=1e-6)
    # brute-force pose search on a reduced rotation table equals the CPU restatement
    from zeroshape_b200.utils.camera import get_rotation_sphere
    R = get_rotation_sphere(24, 24, 12, device="cpu")
    Rref = E.rotation_sphere(4, 3, 2)
    assert R.shape == (6912, 3, 3)
    np.testing.assert_allclose(get_rotation_sphere(4, 3, 2, device="cpu").numpy(), Rref.numpy(), atol=1e-6)


def test_eval_metrics_default_end_to_end(cuda):
    """compute grid -> MC -> sample -> normalise -> chamfer -> F-score through the reference-named API."""
    from zeroshape_b200.utils import eval_3D as ours
    from zeroshape_b200.utils.util import EasyDict
    from zeroshape_b200.model.shape.implicit import Implicit
    from oracle.implicit import implicit_init
    sd = implicit_init(seed=4)
    net = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8,
                   skip_in=[2, 4, 6], pos_perlayer=False)
    net.load_state_dict(sd)
    net = net.to(cuda).eval()
    net.engine = "f32"
    opt = EasyDict(device=cuda, H=224, W=224, eval=dict(vox_res=24, range=[-1.5, 1.5], num_points=2000, brute_force=False,
                                                        f_thresholds=[0.005, 0.01, 0.02, 0.05, 0.1, 0.2], icp=False),
                   data=dict(dataset_test="synthetic"), arch=dict(win_size=16))
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(1, 197, 256, generator=g)
    gt = torch.randn(1, 2000, 3, generator=g)
    var = EasyDict(idx=torch.zeros(1), latent_depth=lat.to(cuda), latent_semantic=None, rgb_input_map=None,
                   pose_gt=torch.eye(3, 4).unsqueeze(0).to(cuda), dpc=EasyDict(points=gt.to(cuda)))
    acc, comp = ours.eval_metrics(opt, var, net)
    # oracle pipeline on the same occupancy grid, own sampling -> metrics agree within sampling noise
    occ = E.level_grid(sd, lat, 25, -1.5, 1.5)[0].numpy()
    v, f = E.marching_cubes(occ, 0.5)
    assert len(var.mesh_pred[0].faces) == len(f) and len(f) > 0
    np.testing.assert_allclose(var.mesh_pred[0].vertices, E.scale_vertices(v, 25, -1.5, 1.5), atol=5e-4)   # occupancy differs by ~1e-6 -> edge interpolation
    pred = torch.from_numpy(E.sample_surface(E.scale_vertices(v, 25, -1.5, 1.5), f, 2000, np.random.RandomState(0))).float()
    d1, d2, _, _ = E.chamfer_nn(E.normalize_pc(pred.unsqueeze(0)).numpy(), E.normalize_pc(gt).numpy())
    assert abs(np.sqrt(d1).mean() - acc.item()) < 0.15 * acc.item()
    assert abs(np.sqrt(d2).mean() - comp.item()) < 0.15 * comp.item()
    assert var.f_score.shape == (1, 6) and var.cd_acc.shape == (1,)


def test_attention_movie_matches_oracle_zmean(cuda):
    """compute_level_grid(vis_attn=True) (utils/eval_3D.py:47-80): the Z-averaged, head/layer-averaged attention of the columns
    the movie shows equals the oracle's attention maps reduced the reference's way; frames have the reference's count/shape."""
    from oracle.implicit import implicit_forward, implicit_init
    from oracle import eval3d as E
    from zeroshape_b200.model.shape.implicit import Implicit
    from zeroshape_b200.utils import eval_3D
    from zeroshape_b200.utils.util import EasyDict
    sd = implicit_init(seed=31)
    m = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6], pos_perlayer=False)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    n = 17
    g = torch.Generator().manual_seed(32)
    lat = torch.randn(1, 197, 256, generator=g)
    pts = E.dense_grid(n, -1.5, 1.5)                                    # [1,n,n,n,3]
    with torch.no_grad():
        _, attn = implicit_forward(sd, lat, pts.view(1, -1, 3))
    ref = attn.view(1, n, n, n, 197).mean(dim=3)
    ref = (ref[..., :1] + ref[..., 1:])[:, ::8][:, :, ::8]              # [1,3,3,196]
    got, idx = eval_3D.attention_maps_zmean(m, lat.to(cuda), pts.to(cuda))
    assert idx.tolist() == [0, 8, 16]
    assert (got.cpu() - ref).abs().max().item() < 2e-6
    opt = EasyDict(H=224, W=224, arch=dict(win_size=16))
    img = torch.rand(1, 3, 224, 224, generator=g)
    occ, frames = eval_3D.compute_level_grid(opt, m, lat.to(cuda), None, pts.to(cuda), img.to(cuda), vis_attn=True)
    assert occ.shape == (1, n, n, n) and len(frames) == 1 and len(frames[0]) == 9
    assert frames[0][0].shape == (224, 224, 3) and 0.0 <= frames[0][0].min() and frames[0][0].max() <= 1.0


@pytest.mark.parametrize("sets,n,nq", [(1, 10000, 10000), (3, 700, 1300), (2, 33, 5), (1, 16384, 257), (4, 1, 40)])
def test_nn_bvh_is_bit_identical_to_the_dense_chamfer(cuda, sets, n, nq):
    """csrc/nn_bvh.cu (box hierarchy) vs csrc/chamfer.cu (every pair) vs the C restatement of the reference kernel:
    squared distances bit-equal, indices equal (lowest index on ties: duplicated target points are planted)."""
    from zeroshape_b200 import ops
    g = torch.Generator().manual_seed(sets * 1000 + n)
    # a surface-like cloud (points on a noisy ellipsoid) plus exact duplicates, queries from a rotated copy + far outliers
    t = torch.randn(sets, n, 3, generator=g)
    t = t / t.norm(dim=-1, keepdim=True) * torch.tensor([0.5, 0.3, 0.4]) + 0.01 * torch.randn(sets, n, 3, generator=g)
    if n > 40:
        t[:, 7] = t[:, 31]
        t[:, n - 1] = t[:, 2]
    q = torch.randn(sets, nq, 3, generator=g)
    q = q / q.norm(dim=-1, keepdim=True) * torch.tensor([0.3, 0.5, 0.4])
    q[:, : max(1, nq // 50)] *= 6.0
    if n > 40:
        q[:, -1] = t[:, 31]                       # an exact hit on a duplicated target: distance 0, index 7
    td, qd = t.to(cuda).contiguous(), q.to(cuda).contiguous()
    bvh = ops.NNBvh(td)
    order = ops.NNBvh(qd[:1].contiguous()).morton_order()
    d_ref, _, i_ref, _ = ops.chamfer_nn(qd, td)
    for variant in (0, 1):                    # one thread per query / warp-cooperative
        for q_order in (None, order):
            d, i = bvh.query(qd, q_order=q_order, variant=variant)
            assert torch.equal(d, d_ref) and torch.equal(i, i_ref), (variant, q_order is None)
    r1, _, j1, _ = E.chamfer_nn(q.numpy(), t.numpy())
    assert np.array_equal(d.cpu().numpy(), r1) and np.array_equal(i.cpu().numpy(), j1)
    assert sorted(order.cpu().tolist()) == list(range(nq))
    # shared target / shared query forms used by the pose search
    d_sr, _, i_sr, _ = ops.chamfer_nn(qd, td[:1].expand(sets, -1, -1).contiguous())
    d_qr, _, i_qr, _ = ops.chamfer_nn(qd[:1].expand(sets, -1, -1).contiguous(), td)
    for variant in (0, 1):
        d_s, i_s = ops.NNBvh(td[:1].contiguous()).query(qd, variant=variant)
        assert torch.equal(d_s, d_sr) and torch.equal(i_s, i_sr)
        d_q, i_q = bvh.query(qd[:1].contiguous(), batch=sets, variant=variant)
        assert torch.equal(d_q, d_qr) and torch.equal(i_q, i_qr)


def test_pose_search_bvh_equals_dense(cuda):
    """utils/eval_3D.brute_force_search: the box-hierarchy path selects the same rotation and returns the same metrics as the
    dense 10k x 10k Chamfer path (the reference's algorithm) on a reduced rotation table."""
    from zeroshape_b200.utils import eval_3D as ours
    from zeroshape_b200.utils import camera
    g = torch.Generator().manual_seed(12)
    gt = torch.randn(3000, 3, generator=g)
    gt = gt / gt.norm(dim=-1, keepdim=True) * torch.tensor([0.5, 0.25, 0.35])
    Rz = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    pred = (gt[torch.randperm(3000, generator=g)[:2500]] + 0.004 * torch.randn(2500, 3, generator=g)) @ Rz.T
    real = camera.get_rotation_sphere
    camera_small = lambda azim_sample, elev_sample, roll_sample, scales, device: real(8, 6, 4, scales, device)
    saved = ours.get_rotation_sphere
    ours.get_rotation_sphere = camera_small
    try:
        a = ours.brute_force_search(pred, gt, device=cuda, method="bvh", batch_size=50)
        b = ours.brute_force_search(pred, gt, device=cuda, method="dense", batch_size=24)
    finally:
        ours.get_rotation_sphere = saved
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert torch.isfinite(a[0]) and torch.isfinite(a[1]) and a[2].shape == (6,)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_meshes_union_equals_full_mesh(cuda, world):
    """SURVEY.md 8e A / BASELINE config 4 on the real kernels: the x-slab (+ one halo slice) decode and per-slab marching cubes
    of every simulated rank, merged the way `parallel.gather_meshes` merges them, give the single-rank mesh as a set of faces."""
    from oracle.implicit import implicit_init
    from zeroshape_b200.model.shape.implicit import Implicit
    from zeroshape_b200.parallel import slab_meshes, face_set, all_slab_bounds
    from zeroshape_b200 import ops
    net = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
                   pos_perlayer=False)
    net.load_state_dict(implicit_init(seed=17))
    net = net.to(cuda).eval()
    lat = torch.randn(2, 197, 256, generator=torch.Generator().manual_seed(18)).to(cuda)
    n = 33
    full = slab_meshes(net, lat, n, -1.5, 1.5, 0, 1)
    parts = [slab_meshes(net, lat, n, -1.5, 1.5, r, world) for r in range(world)]
    for b in range(2):
        v = torch.cat([parts[r][b][0] for r in range(world)])
        base, fs = 0, []
        for r in range(world):
            fs.append(parts[r][b][1] + base)
            base += parts[r][b][0].shape[0]
        f = torch.cat(fs)
        assert f.shape[0] == full[b][1].shape[0] and f.shape[0] > 500
        assert face_set(v, f) == face_set(*full[b])
        # seam vertices are the only duplicates
        assert v.shape[0] >= full[b][0].shape[0] and v.shape[0] - full[b][0].shape[0] < 0.2 * full[b][0].shape[0] * world / 2
    # the cubic entry point still equals the slab entry point with nx == n
    occ = net.grid_occupancy(lat[:1], n, -1.5, 1.5)[0].contiguous()
    v0, f0 = ops.marching_cubes(occ, 0.5)
    assert torch.equal(f0, full[0][1]) and torch.equal(v0, full[0][0])
