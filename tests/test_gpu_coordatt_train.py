"""GPU: training of the transformer seen-surface encoder (model/shape/seen_coord_att_train.py) -- every parameter gradient and the
gradient w.r.t. the XYZ map against torch autograd over the oracle restatement (pinned to the reference module), on both training
engines; and one full Graph training step with `arch.depth.encoder != resnet`."""
import numpy as np
import pytest
import torch

from oracle import backbone as BB
from oracle.graph_params import seeded_state_dict

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("engine", ["f32", "tc"])
def test_coord_enc_att_gradients_match_oracle_autograd(cuda, engine):
    from zeroshape_b200 import ops
    from zeroshape_b200.model.depth import dpt_train as T
    from zeroshape_b200.model.shape import seen_coord_att_train as CAT
    from zeroshape_b200.model.shape.seen_coord_enc import CoordEncAtt
    if engine != "f32" and ops.device_cc() != 100:
        pytest.skip("tcgen05 kernels need sm_100")
    mod = CoordEncAtt(embed_dim=256, n_blocks=2, num_heads=8, win_size=8, drop_path=0.0)
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    sd = seeded_state_dict(shapes, seed=41, implicit_prefix=None)
    sd["coord_embed.two_d_pos_embed"] = mod.state_dict()["coord_embed.two_d_pos_embed"].clone()
    mod.load_state_dict(sd, strict=True)
    mod = mod.to(cuda).train()
    g = torch.Generator().manual_seed(42)
    B, H, W = 2, 32, 24
    coord = torch.randn(B, H, W, 3, generator=g) * 0.4
    mask = torch.rand(B, H, W, generator=g) < 0.7
    mask[0, :8, :8] = False
    wgt = torch.randn(B, 1 + (H // 8) * (W // 8), 256, generator=g)
    # oracle autograd
    sd_ref = {"coord_encoder." + k: (v.clone().requires_grad_(k != "coord_embed.two_d_pos_embed")) for k, v in sd.items()}
    coord_r = (coord * mask.unsqueeze(-1)).clone().requires_grad_(True)
    out_ref = BB.coord_enc_att_forward(sd_ref, coord_r, mask)
    (out_ref * wgt).sum().backward()
    # ours, on the tape (so that the gradient w.r.t. the XYZ map is visible)
    saved = (ops.TRAIN_ENGINE, ops.TRAIN_PRECISION)
    ops.TRAIN_ENGINE, ops.TRAIN_PRECISION = ("f32" if engine == "f32" else "tc"), "bf16x3"
    try:
        with torch.no_grad():
            tp = T.Tape()
            cd = (coord * mask.unsqueeze(-1)).to(cuda).contiguous()
            out = CAT.train_forward(tp, mod, cd, mask.float().to(cuda).contiguous())
            tp.add(out, wgt.to(cuda))
            tp.backward()
            dcoord = tp.pop(cd)
        # and through the autograd bridge (parameters only)
        mod.zero_grad()
        out2 = mod(cd, mask.to(cuda))
        (out2 * wgt.to(cuda)).sum().backward()
    finally:
        ops.TRAIN_ENGINE, ops.TRAIN_PRECISION = saved
    tol = 2e-4 if engine == "f32" else 3e-3
    assert _rel(out, out_ref) < (1e-5 if engine == "f32" else 1e-3)
    worst = ("", 0.0)
    for name, p in mod.named_parameters():
        if name == "coord_embed.two_d_pos_embed":
            continue
        gref = sd_ref["coord_encoder." + name].grad
        r = _rel(tp.pgrads[id(p)], gref)
        worst = max(worst, (name, r), key=lambda t: t[1])
        assert r < tol, (name, r)
        assert _rel(p.grad, gref) < tol, name
    m3 = mask.unsqueeze(-1)
    assert _rel(dcoord.cpu() * m3, coord_r.grad * m3) < tol
    print(f"[{engine}] CoordEncAtt worst parameter-gradient error {worst}, d xyz {_rel(dcoord.cpu() * m3, coord_r.grad * m3):.2e}")


def test_full_graph_training_step_with_transformer_encoder(cuda):
    from zeroshape_b200.model.compute_graph.graph_shape import Graph
    from zeroshape_b200.model.shape.implicit_train import FusedAdamW
    from zeroshape_b200.utils.util import EasyDict
    from test_gpu_graph import make_opt, synthetic_image_and_mask
    opt = make_opt(cuda)
    opt.arch.depth.encoder, opt.arch.depth.n_blocks, opt.arch.depth.dsp = "transformer", 2, 2
    opt.loss_weight = EasyDict(depth=None, intr=None, shape=1)
    opt.training = EasyDict(shape_loss=EasyDict(impt_thres=0.01, impt_weight=1))
    torch.manual_seed(3)
    graph = Graph(opt).to(cuda).train()
    with torch.no_grad():
        getattr(graph.dpt_depth.scratch.output_conv, "4").bias.fill_(0.5)
        graph.intr_proj.weight.normal_(0, 0.02)
    B, N = 2, 256
    rgb, mask = synthetic_image_and_mask(B, 71)
    g = torch.Generator().manual_seed(72)
    depth_gt = (1.5 + 0.3 * torch.rand(B, 1, 224, 224, generator=g)) * mask
    intr = torch.tensor([[1.3875 * 224, 0, 112], [0, 1.3875 * 224, 112], [0, 0, 1.0]]).repeat(B, 1, 1)
    pose = torch.cat([torch.eye(3), torch.tensor([[0.0], [0.0], [1.6]])], dim=1).repeat(B, 1, 1)
    gt_pts = torch.rand(B, N, 3, generator=g) - 0.5
    gt_sdf = gt_pts.norm(dim=-1) - 0.3

    def batch():
        return EasyDict(idx=torch.arange(B), rgb_input_map=rgb.to(cuda), mask_input_map=mask.to(cuda), depth_input_map=depth_gt.to(cuda),
                        intr=intr.to(cuda), pose_gt=pose.to(cuda), gt_sample_points=gt_pts.to(cuda), gt_sample_sdf=gt_sdf.to(cuda))
    trainable = [p for p in graph.parameters() if p.requires_grad]
    optim = FusedAdamW(trainable, lr=1e-5, betas=(0.9, 0.95), weight_decay=0.05)
    losses = []
    for it in range(3):
        var, loss = graph.forward(opt, batch(), training=True)
        optim.zero_grad()
        loss.shape.backward()
        if it == 0:
            names = {n for n, p in graph.named_parameters() if p.grad is not None and torch.isfinite(p.grad).all()}
            for prefix in ("coord_encoder.coord_embed.pos_embed", "coord_encoder.coord_embed.invalid_coord_token", "coord_encoder.blocks.1.mlp.fc2",
                           "coord_encoder.cls_token", "dpt_depth.scratch.output_conv.4", "intr_proj", "impl_network.impl_mlp.layers.8"):
                assert any(n.startswith(prefix) for n in names), prefix
        optim.step()
        losses.append(loss.shape.item())
    print("transformer-encoder graph, shape loss per step:", [round(v, 4) for v in losses])
    assert all(np.isfinite(losses))
