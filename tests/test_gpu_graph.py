"""GPU: encoder side of the hot path through the reference-named Graph API (DPT-hybrid depth, intrinsics
head, unproject/normalise, CoordEncRes) against the reference-Graph golden vectors and the oracle, then
the whole image -> occupancy grid path.  North-star bar: depth / logits within 1e-3 relative."""
import os

import numpy as np
import pytest
import torch

from oracle import backbone as BB
from oracle import eval3d as E
from oracle.graph_params import graph_shape_param_shapes, seeded_state_dict
from oracle.implicit import implicit_forward

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def synthetic_image_and_mask(B, seed, cx=112, cy=112, radius=80, H=224, W=224):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(B, 3, H, W, generator=g)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    mask = (((yy - cy) ** 2 + (xx - cx) ** 2) < radius ** 2).float().view(1, 1, H, W).repeat(B, 1, 1, 1)
    return rgb * mask + (1 - mask), mask


def make_opt(device):
    from zeroshape_b200.utils.util import EasyDict
    return EasyDict(device=device, H=224, W=224, pretrain=dict(depth=None), optim=dict(fix_dpt=False),
                    arch=dict(num_heads=8, latent_dim=256, win_size=16,
                              depth=dict(encoder="resnet", n_blocks=12, dsp=2, pretrained=None), rgb=dict(encoder=None, n_blocks=12),
                              impl=dict(n_channels=256, att_blocks=2, mlp_ratio=4., posenc_perlayer=False, mlp_layers=8,
                                        posenc_3D=0, skip_in=[2, 4, 6])),
                    eval=dict(vox_res=16, range=[-1.5, 1.5], num_points=1000, brute_force=False, icp=False,
                              f_thresholds=[0.005, 0.01, 0.02, 0.05, 0.1, 0.2]),
                    data=dict(dataset_test="synthetic"))


def _graph(sd, cuda):
    from zeroshape_b200.model.compute_graph.graph_shape import Graph
    g = Graph(make_opt(cuda))
    g.load_state_dict(sd, strict=True)
    return g.to(cuda).eval()


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).abs().max() / b.abs().max()).item()


def test_graph_state_dict_matches_reference_keys(cuda):
    g = np.load(os.path.join(GOLD, "graph_encode.npz"))
    from zeroshape_b200.model.compute_graph.graph_shape import Graph
    sd = Graph(make_opt(cuda)).state_dict()
    assert sorted(sd) == list(g["keys"])
    assert [str(tuple(sd[k].shape)) for k in sorted(sd)] == list(g["shapes"])


def test_graph_forward_matches_reference_golden(cuda):
    from zeroshape_b200.utils.util import EasyDict
    g = np.load(os.path.join(GOLD, "graph_encode.npz"))
    sd = seeded_state_dict(graph_shape_param_shapes(), int(g["weight_seed"]))
    graph = _graph(sd, cuda)
    graph.impl_network.engine = "f32"
    cx, cy, r = [int(v) for v in g["disc"]]
    rgb, mask = synthetic_image_and_mask(1, int(g["image_seed"]), cx, cy, r)
    var = EasyDict(idx=torch.arange(1), rgb_input_map=rgb.to(cuda), mask_input_map=mask.to(cuda), pose_gt=False)
    var = graph.forward(make_opt(cuda), var, training=False, get_loss=False)
    assert var.depth_pred.shape == (1, 1, 224, 224) and var.latent_depth.shape == (1, 197, 256)
    assert np.abs(var.depth_pred.cpu().numpy() - g["depth_pred"]).max() < 1e-3 * g["depth_pred"].max()
    assert _rel(var.intr_pred, torch.from_numpy(g["intr_pred"])) < 1e-5
    assert np.abs(var.seen_points.cpu().numpy()[:, ::7] - g["seen_points"]).max() < 1e-3
    assert _rel(var.latent_depth, torch.from_numpy(g["latent_depth"])) < 1e-3
    logits, _ = graph.impl_network(var.latent_depth, None, torch.from_numpy(g["points"]).to(cuda), need_attn=False)
    assert np.abs(logits.cpu().numpy() - g["logits"]).max() < 1e-3 * np.abs(g["logits"]).max()
    print("depth err", np.abs(var.depth_pred.cpu().numpy() - g["depth_pred"]).max(),
          "latent rel", _rel(var.latent_depth, torch.from_numpy(g["latent_depth"])))


def test_graph_forward_matches_oracle_batch(cuda):
    from zeroshape_b200.utils.util import EasyDict
    sd = seeded_state_dict(graph_shape_param_shapes(), 33)
    graph = _graph(sd, cuda)
    rgb, mask = synthetic_image_and_mask(2, 34, 118, 106, 74)
    with torch.no_grad():
        ref = BB.graph_shape_encode(sd, rgb, mask)
    var = EasyDict(idx=torch.arange(2), rgb_input_map=rgb.to(cuda), mask_input_map=mask.to(cuda), pose_gt=False)
    from zeroshape_b200 import ops
    # bit-faithful FFMA encoder: fp32-grade agreement; tensor-core (bf16x3) encoder: the north-star 1e-3 bar
    for engine, tol in (("f32", 1e-4), ("auto", 1e-3)):
        ops.ENCODER_ENGINE = engine
        try:
            var = EasyDict(idx=torch.arange(2), rgb_input_map=rgb.to(cuda), mask_input_map=mask.to(cuda), pose_gt=False)
            var, loss = graph.forward(make_opt(cuda), var, training=False, get_loss=True)
        finally:
            ops.ENCODER_ENGINE = "auto"
        errs = (_rel(var.depth_pred, ref["depth_pred"]), _rel(var.seen_points, ref["seen_points"]), _rel(var.latent_depth, ref["latent_depth"]))
        print("encoder engine", engine, "depth / seen_points / latent rel err", errs)
        assert max(errs) < tol, (engine, errs)
        assert torch.equal(var.validity_mask.cpu(), ref["validity_mask"])
    # image -> occupancy grid, default (tensor-core) engine: identical thresholded voxels outside the band
    n = 17
    occ = graph.impl_network.grid_occupancy(var.latent_depth, n, -1.5, 1.5).cpu()
    sd_impl = {k[len("impl_network."):]: v for k, v in sd.items() if k.startswith("impl_network.")}
    occ_ref = E.level_grid(sd_impl, ref["latent_depth"], n, -1.5, 1.5)
    assert (occ - occ_ref).abs().max() < 2.5e-4
    band = (occ_ref - 0.5).abs() > 5e-4
    assert torch.equal((occ > 0.5)[band], (occ_ref > 0.5)[band]) and band.float().mean() > 0.9


def test_depth_graph_and_guards(cuda):
    from zeroshape_b200.utils.util import EasyDict
    from zeroshape_b200.model.compute_graph.graph_depth import Graph as DepthGraph
    opt = make_opt(cuda)
    opt.loss_weight = EasyDict(depth=1, intr=1)
    dg = DepthGraph(opt).to(cuda).eval()
    sd = seeded_state_dict(graph_shape_param_shapes(), 35, implicit_prefix=None)
    dg.load_state_dict({k: v for k, v in sd.items() if k.startswith(("dpt_depth.", "intr_"))}, strict=True)
    rgb, mask = synthetic_image_and_mask(1, 36)
    var = EasyDict(idx=torch.arange(1), rgb_input_map=rgb.to(cuda), mask_input_map=mask.to(cuda))
    var = dg.forward(opt, var, training=False, get_loss=False)
    with torch.no_grad():
        d_ref, _ = BB.dpt_depth_forward(sd, rgb, "dpt_depth.")
    assert _rel(var.depth_pred, d_ref) < 1e-3 and var.intr_pred.shape == (1, 3, 3)
    # the training path of this graph is covered by tests/test_gpu_midas.py; eval mode never needs ground truth
    var2, loss = dg.forward(opt, EasyDict(idx=torch.arange(1), rgb_input_map=rgb.to(cuda), mask_input_map=mask.to(cuda)), training=False)
    assert len(loss) == 0 and torch.equal(var2.depth_pred, var.depth_pred)
