"""CPU: the product's HOST logic (module trees, weight packing, layer order, NHWC plumbing, grid slab
loop) executed with CPU stand-ins for the kernels (tests/fake_ops.py) and compared with the oracle.
A bug in the Python orchestration shows up here without a GPU; kernel numerics are tested under -m gpu."""
import os

import numpy as np
import pytest
import torch

import fake_ops
from oracle import backbone as BB
from oracle import eval3d as E
from oracle.graph_params import graph_shape_param_shapes, seeded_state_dict
from oracle.implicit import implicit_forward, implicit_init

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _opt():
    from zeroshape_b200.utils.util import EasyDict
    return EasyDict(device="cpu", H=224, W=224, pretrain=dict(depth=None), optim=dict(fix_dpt=False),
                    arch=dict(num_heads=8, latent_dim=256, win_size=16,
                              depth=dict(encoder="resnet", n_blocks=12, dsp=2, pretrained=None), rgb=dict(encoder=None, n_blocks=12),
                              impl=dict(n_channels=256, att_blocks=2, mlp_ratio=4., posenc_perlayer=False, mlp_layers=8,
                                        posenc_3D=0, skip_in=[2, 4, 6])))


def _img(B, seed):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(B, 3, 224, 224, generator=g)
    yy, xx = torch.meshgrid(torch.arange(224), torch.arange(224), indexing="ij")
    mask = (((yy - 110) ** 2 + (xx - 120) ** 2) < 75 ** 2).float().view(1, 1, 224, 224).repeat(B, 1, 1, 1)
    return rgb * mask + (1 - mask), mask


def test_graph_orchestration_matches_oracle(monkeypatch):
    fake_ops.install(monkeypatch)
    from zeroshape_b200.model.compute_graph.graph_shape import Graph
    from zeroshape_b200.utils.util import EasyDict
    sd = seeded_state_dict(graph_shape_param_shapes(), 41)
    graph = Graph(_opt())
    graph.load_state_dict(sd, strict=True)
    graph.eval()
    rgb, mask = _img(2, 42)
    var = EasyDict(idx=torch.arange(2), rgb_input_map=rgb, mask_input_map=mask, pose_gt=False)
    var = graph.forward(_opt(), var, training=False, get_loss=False)
    with torch.no_grad():
        ref = BB.graph_shape_encode(sd, rgb, mask)
    for k, tol in (("depth_pred", 2e-5), ("seen_points", 1e-4), ("latent_depth", 2e-4)):
        assert (var[k] - ref[k]).abs().max().item() < tol, (k, (var[k] - ref[k]).abs().max().item())
    assert ((var.intr_pred - ref["intr_pred"]).abs() / (ref["intr_pred"].abs() + 1)).max() < 1e-6
    # weight re-pack on in-place update (optimizer step / load_state_dict)
    with torch.no_grad():
        getattr(graph.dpt_depth.scratch.output_conv, "4").bias.add_(0.1)
    var2 = graph.forward(_opt(), EasyDict(idx=torch.arange(2), rgb_input_map=rgb, mask_input_map=mask, pose_gt=False),
                         training=False, get_loss=False)
    assert (var2.depth_pred - (var.depth_pred + 0.1).clamp(0, 1)).abs().max() < 1e-5


def test_implicit_orchestration_matches_oracle(monkeypatch):
    fake_ops.install(monkeypatch)
    from zeroshape_b200.model.shape.implicit import Implicit
    sd = implicit_init(seed=5)
    net = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
                   pos_perlayer=False)
    net.load_state_dict(sd)
    net.eval()
    net.point_chunk = 700          # force several chunks / slabs
    g = torch.Generator().manual_seed(6)
    lat, pts = torch.randn(2, 197, 256, generator=g), torch.rand(2, 1500, 3, generator=g) * 3 - 1.5
    with torch.no_grad():
        ref, ref_attn = implicit_forward(sd, lat, pts)
    out, attn = net(lat, None, pts)
    assert (out - ref).abs().max() < 1e-5 and (attn - ref_attn).abs().max() < 1e-6
    n = 9
    occ = net.grid_occupancy(lat, n, -1.5, 1.5)
    assert (occ - E.level_grid(sd, lat, n, -1.5, 1.5)).abs().max() < 1e-5
    assert torch.equal(torch.cat([net.grid_occupancy(lat, n, -1.5, 1.5, 0, 4), net.grid_occupancy(lat, n, -1.5, 1.5, 4, 9)], 1), occ)
