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


def test_coord_enc_att_matches_reference_golden_and_oracle(cuda):
    """Transformer seen-surface encoder (SURVEY.md section 8a row a7'): the mirror on the CUDA kernels vs the real reference
    module's golden output, and vs the oracle at the deployed geometry (112 x 112 map, 12 blocks)."""
    from test_oracle_coordatt import G as GC, golden_sd
    from zeroshape_b200 import ops
    from zeroshape_b200.model.shape.seen_coord_enc import CoordEncAtt
    mod = CoordEncAtt(embed_dim=256, n_blocks=3, num_heads=8, win_size=8)
    mod.load_state_dict(golden_sd(), strict=True)
    mod = mod.to(cuda).eval()
    ref = torch.from_numpy(GC["out"])
    for engine, tol in (("f32", 2e-5), ("auto", 1e-3)):
        ops.ENCODER_ENGINE = engine
        try:
            out = mod(torch.from_numpy(GC["coord"]).to(cuda), torch.from_numpy(GC["mask"]).to(cuda))
        finally:
            ops.ENCODER_ENGINE = "auto"
        err = (out.cpu() - ref).abs().max().item() / ref.abs().max().item()
        print("CoordEncAtt engine", engine, "rel err vs the reference module", err)
        assert out.shape == ref.shape and err < tol, (engine, err)
    # deployed geometry: H/dsp = W/dsp = 112, win 8 -> 196 window tokens + cls, 12 blocks
    big = CoordEncAtt(embed_dim=256, n_blocks=12, num_heads=8, win_size=8)
    shapes = {k: tuple(v.shape) for k, v in big.state_dict().items()}
    sd = seeded_state_dict(shapes, seed=77, implicit_prefix=None)
    sd["coord_embed.two_d_pos_embed"] = big.state_dict()["coord_embed.two_d_pos_embed"].clone()
    big.load_state_dict(sd, strict=True)
    big = big.to(cuda).eval()
    g = torch.Generator().manual_seed(78)
    coord = torch.randn(2, 112, 112, 3, generator=g) * 0.4
    yy, xx = torch.meshgrid(torch.arange(112), torch.arange(112), indexing="ij")
    mask = (((yy - 56) ** 2 + (xx - 52) ** 2) < 40 ** 2).unsqueeze(0).repeat(2, 1, 1)
    with torch.no_grad():
        want = BB.coord_enc_att_forward({"coord_encoder." + k: v for k, v in sd.items()}, coord * mask.unsqueeze(-1), mask)
    got = big((coord * mask.unsqueeze(-1)).to(cuda), mask.to(cuda))
    assert got.shape == (2, 197, 256) and _rel(got, want) < 1e-3, _rel(got, want)


def test_graph_with_transformer_encoder(cuda):
    """graph_shape.Graph with `arch.depth.encoder != resnet` (dsp 2, windows of 16 // 2 = 8 pixels on the 112 x 112 resampled XYZ map):
    latent_depth equals the oracle's interpolate_coordmap + CoordEncAtt on the graph's own seen surface (graph_shape.py:141-150)."""
    import torch.nn.functional as F
    from zeroshape_b200.model.compute_graph.graph_shape import Graph
    from zeroshape_b200.utils.util import EasyDict
    opt = make_opt(cuda)
    opt.arch.depth.encoder = "transformer"
    opt.arch.depth.n_blocks = 2
    opt.arch.depth.dsp = 2
    torch.manual_seed(5)
    graph = Graph(opt).to(cuda).eval()
    with torch.no_grad():
        getattr(graph.dpt_depth.scratch.output_conv, "4").bias.fill_(0.5)      # depth inside (0, 1) for the random-init estimator
    rgb, mask = synthetic_image_and_mask(2, 91, 116, 108, 70)
    var = EasyDict(idx=torch.arange(2), rgb_input_map=rgb.to(cuda), mask_input_map=mask.to(cuda), pose_gt=False)
    var = graph.forward(opt, var, training=False, get_loss=False)
    assert var.latent_depth.shape == (2, 197, 256) and torch.isfinite(var.latent_depth).all()
    seen = var.seen_points.cpu().view(2, 224, 224, 3).permute(0, 3, 1, 2)
    m = (mask > 0.5).float()
    cv = F.interpolate(seen * m, (112, 112), mode="bilinear", align_corners=False)
    mk = F.interpolate(m, (112, 112), mode="bilinear", align_corners=False)
    mb = (mk > 0.5).float()
    coord = (cv / (mk + 1.e-6)) * mb
    sd = {k: v.detach().cpu() for k, v in graph.state_dict().items() if k.startswith("coord_encoder.")}
    with torch.no_grad():
        want = BB.coord_enc_att_forward(sd, coord.permute(0, 2, 3, 1).contiguous(), mb.squeeze(1) > 0.5)
    assert _rel(var.latent_depth, want) < 1e-3, _rel(var.latent_depth, want)
    logits, _ = graph.impl_network(var.latent_depth, None, torch.rand(2, 50, 3, device=cuda) - 0.5, need_attn=False)
    assert logits.shape == (2, 50) and torch.isfinite(logits).all()


def test_graphed_inference_encoder_equals_the_eager_one(cuda):
    """Graph.forward in eval mode replays the image -> latents encoder from a CUDA graph (graph_shape.py `_encode_graphed`): the
    same bits as launching it op by op, for new inputs through the same capture, for another batch size, and after a weight
    update (the capture is keyed on the parameters' version counters and re-made)."""
    from zeroshape_b200 import ops
    from zeroshape_b200._native import lib
    from zeroshape_b200.utils.util import EasyDict
    sd = seeded_state_dict(graph_shape_param_shapes(), 41)
    graph = _graph(sd, cuda)
    opt = make_opt(cuda)

    def run(rgb, mask, graphed, model=None):
        ops.ENCODER_CUDA_GRAPH = graphed
        try:
            var = EasyDict(idx=torch.arange(rgb.shape[0]), rgb_input_map=rgb.to(cuda), mask_input_map=mask.to(cuda), pose_gt=False)
            with torch.no_grad():
                var = (model or graph).forward(opt, var, training=False, get_loss=False)
        finally:
            ops.ENCODER_CUDA_GRAPH = True
        return {k: var[k].clone() for k in ("depth_pred", "intr_pred", "seen_points", "latent_depth", "validity_mask")}

    def same(a, b):
        return all(torch.equal(a[k], b[k]) for k in a)
    x1, x2, x3 = (synthetic_image_and_mask(1, 42), synthetic_image_and_mask(1, 43, 100, 120, 60), synthetic_image_and_mask(2, 44, 118, 106, 74))
    e1, e2, e3 = run(*x1, False), run(*x2, False), run(*x3, False)
    n0 = lib.zs_launch_count()
    g1 = run(*x1, True)                                     # warm-up + capture + first replay
    assert len(graph._encoder_graphs) == 1
    n1 = lib.zs_launch_count()
    g2 = run(*x2, True)                                     # replay only, new input
    per_replay = lib.zs_launch_count() - n1
    assert per_replay > 100 and len(graph._encoder_graphs) == 1, per_replay      # the replayed launches are accounted for
    g3 = run(*x3, True)                                     # another batch size: its own capture
    assert len(graph._encoder_graphs) == 2
    assert same(e1, g1) and same(e2, g2) and same(e3, g3)
    assert not torch.equal(g1["latent_depth"], g2["latent_depth"])
    # weight update -> new capture, new (correct) numbers
    with torch.no_grad():
        graph.intr_proj.bias.add_(0.05)
        dict(graph.dpt_depth.named_parameters())["scratch.output_conv.4.bias"].add_(0.01)
    e4, g4 = run(*x1, False), run(*x1, True)
    assert same(e4, g4) and not torch.equal(g4["depth_pred"], g1["depth_pred"])
    # the captures live outside the module: deepcopy / pickling of the Graph still work, the copy starts without captures
    import copy
    import io
    twin = copy.deepcopy(graph)
    assert len(twin._encoder_graphs) == 0 and len(graph._encoder_graphs) >= 1
    buf = io.BytesIO()
    torch.save(graph, buf)
    buf.seek(0)
    loaded = torch.load(buf, weights_only=False)
    with torch.no_grad():                                   # the copies own their weights (and their packed images of them)
        dict(twin.dpt_depth.named_parameters())["scratch.output_conv.4.bias"].add_(0.02)
    t4, l4 = run(*x1, True, twin), run(*x1, True, loaded)
    assert same(l4, g4) and not torch.equal(t4["depth_pred"], g4["depth_pred"]) and same(run(*x1, True), g4)
    # the outputs handed out are copies: a later replay does not overwrite them
    keep = g4["latent_depth"].clone()
    run(*x2, True)
    assert torch.equal(keep, g4["latent_depth"])
