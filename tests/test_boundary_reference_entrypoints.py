"""Boundary: the stock reference entry points run over this package's mirrors.

CPU part (build container, /root/reference present; skipped where it is absent):
  * `zeroshape_b200.install_as_reference_modules()` registers the mirrors under the reference's import names, then the REAL
    `/root/reference/demo.py` and `/root/reference/model/shape_engine.py` are imported (third-party visualisation packages the
    container lacks -- matplotlib, pyrender, imageio, trimesh, mcubes -- are empty stand-in modules: nothing on the tested path
    touches them) and their own code is executed:
        demo.marching_cubes(opt, var, impl_network)                 demo.py:143-153
        Runner.evaluate_batch(opt, var)                             model/shape_engine.py:517-523
    with the kernels replaced by the CPU stand-ins of tests/fake_ops.py (host logic only; kernel numerics are -m gpu tests);
  * `utils.camera` under the shim serves the data loader's calls (data/synthetic.py:139-140: `camera.pose(t=t)`,
    `camera.pose.compose([R, t])` with a numpy R) and equals the real reference module function by function;
  * `utils.eval_3D.ICP / standardize_pc` equal the reference's formulas.
The same run writes tests/golden/entrypoints.npz (tests/golden/make_golden_entrypoints.py); the GPU part replays demo.py's call
sequence on the CUDA kernels and compares with that golden (`/root/reference` does not exist on the GPU box).
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

import fake_ops
from _ref_import import REF_ROOT, reference_available

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "entrypoints.npz")
needs_ref = pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")


def _load_ref_module(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@needs_ref
def test_camera_mirror_equals_reference_module():
    ref = _load_ref_module("_ref_camera", "utils/camera.py")
    from zeroshape_b200.utils import camera as ours
    from zeroshape_b200.utils.util import EasyDict
    g = torch.Generator().manual_seed(0)
    R = torch.linalg.qr(torch.randn(4, 3, 3, generator=g))[0]
    t = torch.randn(4, 3, generator=g)
    for kw in (dict(R=R, t=t), dict(R=R), dict(t=t), dict(t=[0.0, 0.1, 1.5]), dict(R=np.eye(3))):
        assert torch.equal(ours.pose(**kw), ref.pose(**kw))
    P, Q = ref.pose(R=R, t=t), ref.pose(R=R.flip(0), t=t * 2)
    assert torch.equal(ours.pose.invert(P), ref.pose.invert(P))
    assert torch.equal(ours.pose.invert(P, use_inverse=True), ref.pose.invert(P, use_inverse=True))
    assert torch.equal(ours.pose.compose([P, Q, P]), ref.pose.compose([P, Q, P]))
    # the data loader's pattern (data/synthetic.py:135-140): numpy [3,4] rotation composed with a translation pose
    Rt = np.zeros((3, 4)); Rt[:3, :3] = R[0].numpy()
    a = ours.pose.compose([Rt, ours.pose(t=t[0].numpy())])
    b = ref.pose.compose([Rt, ref.pose(t=t[0].numpy())])
    assert torch.equal(a, b) and a.dtype == b.dtype
    X = torch.randn(4, 50, 3, generator=g)
    K = torch.tensor([[300.0, 0, 112], [0, 300.0, 112], [0, 0, 1]]).repeat(4, 1, 1)
    assert torch.equal(ours.to_hom(X), ref.to_hom(X))
    assert torch.equal(ours.world2cam(X, P), ref.world2cam(X, P))
    assert torch.equal(ours.cam2img(X, K), ref.cam2img(X, K))
    opt = EasyDict(device="cpu", H=8, W=8)
    for u, v in zip(ours.proj_points(opt, X, K, P), ref.proj_points(opt, X, K, P)):
        assert torch.equal(u, v)
    assert torch.equal(ours.get_pixel_grid(opt, 5, 7), ref.get_pixel_grid(opt, 5, 7))
    ang = torch.tensor([0.0, 33.0, 120.0, 271.5])
    trig = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1)
    for fn in ("azim_to_rotation_matrix", "elev_to_rotation_matrix", "roll_to_rotation_matrix"):
        for rep, arg in (("angle", ang), ("rad", ang / 50), ("trig", trig)):
            assert torch.equal(getattr(ours, fn)(arg, rep), getattr(ref, fn)(arg, rep)), (fn, rep)
    a = ours.get_rotation_sphere(4, 3, 2, [1.0, 0.5], device="cpu")
    b = ref.get_rotation_sphere(4, 3, 2, [1.0, 0.5], device="cpu")
    assert a.shape == b.shape and (a - b).abs().max() < 1e-6
    sp = torch.randn(2, 64, 3, generator=g)
    mask = (torch.rand(2, 1, 8, 8, generator=g) > 0.3)
    for u, v in zip(ours.valid_norm_fac(sp, mask), ref.valid_norm_fac(sp, mask)):
        assert (u - v).abs().max() < 1e-6


def _stub_visualisation_packages(monkeypatch):
    for name in ("matplotlib", "matplotlib.pyplot", "pyrender", "imageio", "trimesh", "mcubes"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def entrypoint_case(seed=3, vox_res=12):
    """Seeded model + inputs of the entry-point run (shared with the golden generator and the GPU replay)."""
    from oracle.graph_params import graph_shape_param_shapes, seeded_state_dict
    sd = seeded_state_dict(graph_shape_param_shapes(), seed)
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(1, 3, 224, 224, generator=g)
    yy, xx = torch.meshgrid(torch.arange(224), torch.arange(224), indexing="ij")
    mask = (((yy - 112) ** 2 + (xx - 112) ** 2) < 70 ** 2).float().view(1, 1, 224, 224)
    return sd, rgb * mask + (1 - mask), mask, vox_res


_REF_TOP = ("model", "utils", "external", "data", "demo")


def run_reference_entrypoints(monkeypatch):
    """-> dict of outputs produced by the reference's own demo.marching_cubes / Runner.evaluate_batch over our mirrors.
    sys.modules is restored afterwards: other tests import the REAL reference modules under the same names."""
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _REF_TOP}
    try:
        return _run_reference_entrypoints(monkeypatch)
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in _REF_TOP]:
            del sys.modules[k]
        sys.modules.update(saved)


def _run_reference_entrypoints(monkeypatch):
    import zeroshape_b200
    from _ref_import import install_shims, reference_opt
    for name in list(sys.modules):          # a clean import state for the reference's top-level packages
        if name.split(".")[0] in _REF_TOP:
            del sys.modules[name]
    install_shims()
    _stub_visualisation_packages(monkeypatch)
    fake_ops.install(monkeypatch)
    fake_ops.install_eval(monkeypatch)
    zeroshape_b200.install_as_reference_modules()
    import demo                                    # the REAL /root/reference/demo.py
    from model import shape_engine                 # the REAL /root/reference/model/shape_engine.py
    import utils.eval_3D as eval_3D
    import utils.camera as camera
    assert demo.__file__.startswith(REF_ROOT) and shape_engine.__file__.startswith(REF_ROOT)
    assert eval_3D.__name__.startswith("zeroshape_b200") and camera.__name__.startswith("zeroshape_b200")
    assert shape_engine.graph_shape.__name__.startswith("zeroshape_b200"), "the engine must build OUR Graph"
    assert demo.compute_level_grid is eval_3D.compute_level_grid
    sd, rgb, mask, vox_res = entrypoint_case()
    opt = reference_opt("cpu")
    opt.eval.vox_res = vox_res
    graph = shape_engine.graph_shape.Graph(opt)
    graph.load_state_dict(sd, strict=True)
    graph.eval()
    graph.impl_network.engine = "f32"
    with torch.no_grad():       # non-empty iso-surface for the seeded field
        probe = torch.rand(1, 2048, 3, generator=torch.Generator().manual_seed(1)) * 3 - 1.5
    runner = object.__new__(shape_engine.Runner)       # no NCCL / dataset set-up: only the method under test
    runner.graph = graph
    edict = sys.modules["utils.util"].EasyDict
    var = edict(idx=torch.arange(1), rgb_input_map=rgb, mask_input_map=mask, pose_gt=False)
    var = runner.evaluate_batch(opt, var)              # shape_engine.py:517-523 -> Graph.forward
    with torch.no_grad():
        lg, _ = graph.impl_network(var.latent_depth, None, probe, need_attn=False)
    bias_shift = float(lg.median())
    with torch.no_grad():
        graph.impl_network.impl_mlp.layers[-1].bias -= bias_shift
    var = demo.marching_cubes(opt, var, graph.impl_network, visualize_attn=False)     # demo.py:143-153
    mesh = var.mesh_pred[0]
    occ, _ = eval_3D.compute_level_grid(opt, graph.impl_network, var.latent_depth, var.latent_semantic,
                                        eval_3D.get_dense_3D_grid(opt, var), var.rgb_input_map, False)
    # evaluate.py's metric path on the same var (eval_3D.eval_metrics, shape_engine.py:364)
    d = torch.randn(10000, 3, generator=torch.Generator().manual_seed(2))
    var.dpc = edict(points=(d / d.norm(dim=1, keepdim=True) * 0.5).unsqueeze(0))
    var.pose_gt = torch.cat([torch.eye(3), torch.zeros(3, 1)], dim=1).unsqueeze(0)
    opt.eval.num_points = 2000
    eval_3D.eval_metrics(opt, var, graph.impl_network)
    assert var.eval_vox.shape == (1, (vox_res + 1) ** 3, 3)
    # data/synthetic.py:139-140 under the shim
    t = camera.pose(t=np.array([0.0, 0.0, 1.5]))
    Rt = np.zeros((3, 4)); Rt[:3, :3] = np.eye(3)
    assert camera.pose.compose([Rt, t]).shape == (3, 4)
    return dict(depth_pred=var.depth_pred.numpy(), latent_depth=var.latent_depth.numpy(), intr_pred=var.intr_pred.numpy(),
                occ=occ.numpy(), n_vertices=np.int64(len(mesh.vertices)), n_faces=np.int64(len(mesh.faces)),
                vertices=np.asarray(mesh.vertices, dtype=np.float64), faces=np.asarray(mesh.faces, dtype=np.int64),
                bias_shift=np.float64(bias_shift), f_score=var.f_score.numpy(), cd_acc=var.cd_acc.numpy())


@needs_ref
def test_reference_demo_and_engine_entry_points_run_over_the_mirrors(monkeypatch):
    out = run_reference_entrypoints(monkeypatch)
    assert out["n_faces"] > 100 and out["occ"].shape == (1, 13, 13, 13)
    assert 0.05 < (out["occ"] > 0.5).mean() < 0.95
    gold = np.load(GOLD)
    for k in ("depth_pred", "latent_depth", "occ"):
        assert np.abs(out[k] - gold[k]).max() < 1e-5, k
    assert out["n_faces"] == gold["n_faces"] and np.array_equal(out["faces"], gold["faces"])


@needs_ref
def test_icp_and_standardize_equal_reference_formulas(monkeypatch):
    """utils/eval_3D.py:83-91, 271-284: the reference file cannot be imported whole (mcubes / trimesh / CUDA extension), so its
    two functions are compiled from its own source text at test time."""
    import ast
    src = open(os.path.join(REF_ROOT, "utils", "eval_3D.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("ICP", "standardize_pc")]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "ref_eval_3D_excerpt", "exec"), ns)
    fake_ops.install(monkeypatch)
    fake_ops.install_eval(monkeypatch)
    from zeroshape_b200.utils import eval_3D as ours
    ns["chamfer_distance"] = ours.chamfer_distance
    g = torch.Generator().manual_seed(5)
    X2 = torch.randn(2, 300, 3, generator=g)
    ang = 0.2
    Rz = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=torch.float32)
    X1 = X2[:, :250] @ Rz.T + 0.05
    assert (ours.standardize_pc(X2) - ns["standardize_pc"](X2)).abs().max() < 1e-6
    a, b = ours.ICP(None, X1.clone(), X2, num_iter=5), ns["ICP"](None, X1.clone(), X2, num_iter=5)
    assert (a - b).abs().max() < 1e-4


@pytest.mark.gpu
def test_demo_call_sequence_on_gpu_matches_entrypoint_golden(cuda):
    """demo.py:143-153 replayed on the CUDA kernels (get_dense_3D_grid -> compute_level_grid -> .cpu().numpy() ->
    convert_to_explicit) against the golden the REAL reference entry points produced over the CPU stand-ins."""
    from zeroshape_b200.model.compute_graph.graph_shape import Graph
    from zeroshape_b200.utils import eval_3D
    from zeroshape_b200.utils.util import EasyDict
    from _ref_import import reference_opt_dict
    gold = np.load(GOLD)
    sd, rgb, mask, vox_res = entrypoint_case()
    opt = EasyDict(reference_opt_dict(cuda))
    opt.eval.vox_res = vox_res
    graph = Graph(opt)
    graph.load_state_dict(sd, strict=True)
    graph = graph.to(cuda).eval()
    with torch.no_grad():
        graph.impl_network.impl_mlp.layers[-1].bias -= float(gold["bias_shift"])
    var = EasyDict(idx=torch.arange(1), rgb_input_map=rgb.to(cuda), mask_input_map=mask.to(cuda), pose_gt=False)
    var = graph(opt, var, training=False, get_loss=False)
    assert np.abs(var.depth_pred.cpu().numpy() - gold["depth_pred"]).max() < 1e-3 * gold["depth_pred"].max()
    assert np.abs(var.latent_depth.cpu().numpy() - gold["latent_depth"]).max() < 1e-3 * np.abs(gold["latent_depth"]).max()
    points_3D = eval_3D.get_dense_3D_grid(opt, var)
    level_vox, attn_vis = eval_3D.compute_level_grid(opt, graph.impl_network, var.latent_depth, var.latent_semantic, points_3D,
                                                     var.rgb_input_map, False)
    *level_grids, = level_vox.cpu().numpy()
    meshes = eval_3D.convert_to_explicit(opt, level_grids, isoval=0.5, to_pointcloud=False)
    occ = level_vox.cpu().numpy()
    # the latents differ by the encoder's tensor-core rounding (1e-4 relative), the decoder adds < 2e-4: the grid agrees to 1e-3
    assert np.abs(occ - gold["occ"]).max() < 1e-3
    band = np.abs(gold["occ"] - 0.5) > 2e-3
    assert np.array_equal((occ > 0.5)[band], (gold["occ"] > 0.5)[band])
    assert abs(len(meshes[0].faces) - int(gold["n_faces"])) <= 0.02 * int(gold["n_faces"]) + 8
