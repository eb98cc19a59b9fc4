"""GPU: the MiDaS depth loss kernels (csrc/midas.cu through ops.midas_loss / utils.loss.Loss.depth_loss) against the golden
vectors of the REAL reference module (tests/golden/midas.npz: loss value and autograd gradient) and against the oracle
restatement at full map size; and the depth-only compute graph in train mode (`train.py options/depth.yaml`)."""
import os

import numpy as np
import pytest
import torch

from oracle.midas import midas_loss, midas_loss_grad

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "midas.npz"))


@pytest.mark.parametrize("name", ["small", "odd", "empty_image"])
@pytest.mark.parametrize("alpha", [0.1, 0.0])
def test_midas_kernel_matches_reference_golden(cuda, name, alpha):
    from zeroshape_b200 import ops
    pred, gt, mask = (torch.from_numpy(G[f"{name}_{k}"]).to(cuda) for k in ("pred", "gt", "mask"))
    tag = f"{name}_a{int(alpha * 10)}"
    loss, grad = ops.midas_loss(pred, gt, mask, alpha=alpha, inverse_depth=True)
    ref_loss, ref_grad = float(G[f"{tag}_loss"]), torch.from_numpy(G[f"{tag}_grad"]).double()
    assert abs(loss.item() - ref_loss) < 5e-6 * max(1.0, abs(ref_loss)), (loss.item(), ref_loss)
    err = (grad.cpu().double() - ref_grad).abs().max().item()
    assert err < 2e-5 * ref_grad.abs().max().item() + 1e-9, err


def test_midas_mask_shrink_matches_reference_golden(cuda):
    from zeroshape_b200 import ops
    from zeroshape_b200.utils.loss import Loss
    pred, gt, mask = (torch.from_numpy(G[f"shrink_{k}"]).to(cuda) for k in ("pred", "gt", "mask"))
    assert np.array_equal(ops.erode_mask(mask).cpu().numpy(), G["shrink_eroded"])
    lossfn = Loss({"training": {"depth_loss": {"grad_reg": 0.1, "depth_inv": True, "mask_shrink": True}}})
    p = pred.clone().requires_grad_(True)
    loss = lossfn.depth_loss(p, gt, mask)
    loss.backward()
    assert abs(loss.item() - float(G["shrink_loss"])) < 5e-6 * abs(float(G["shrink_loss"]))
    ref = torch.from_numpy(G["shrink_grad"]).double()
    assert (p.grad.cpu().double() - ref).abs().max().item() < 2e-5 * ref.abs().max().item()


def test_midas_kernel_matches_oracle_at_full_size(cuda):
    from zeroshape_b200 import ops
    from zeroshape_b200.utils.loss import Loss
    g = torch.Generator().manual_seed(77)
    B, H, W = 3, 224, 224
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    mask = (((yy - 112) ** 2 + (xx - 108) ** 2) < 80 ** 2).float().view(1, 1, H, W).repeat(B, 1, 1, 1)
    mask[2] = (torch.rand(1, H, W, generator=g) < 0.3).float()
    pred = 0.15 + 0.7 * torch.rand(B, 1, H, W, generator=g)
    gt = (1.5 + 0.3 * torch.rand(B, 1, H, W, generator=g)) * mask
    ref = midas_loss(pred, gt, mask, alpha=0.1)
    ref_grad = midas_loss_grad(pred, gt, mask, alpha=0.1)
    loss, grad = ops.midas_loss(pred.to(cuda), gt.to(cuda), mask.to(cuda), alpha=0.1)
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item()), (loss.item(), ref.item())
    err = (grad.cpu().double() - ref_grad).abs().max().item()
    assert err < 1e-4 * ref_grad.abs().max().item(), err
    # the reference-named entry point, through autograd
    lossfn = Loss({"training": {"depth_loss": {"grad_reg": 0.1, "depth_inv": True, "mask_shrink": False}}})
    p = pred.to(cuda).requires_grad_(True)
    out = lossfn.depth_loss(p, gt.to(cuda), mask.to(cuda))
    (out * 3.0).backward()
    assert torch.allclose(p.grad, grad * 3.0, rtol=1e-6, atol=1e-12) and abs(out.item() - loss.item()) < 1e-7


@pytest.mark.parametrize("tag,cap", [("dm", None), ("dmcap", 1.6)])
def test_depth_metric_matches_reference_golden(cuda, tag, cap):
    """utils.eval_depth.DepthMetric (one launch) vs the real reference class (golden) and, at full size, vs the oracle restatement."""
    from oracle.midas import depth_metrics
    from zeroshape_b200.utils.eval_depth import DepthMetric
    pred, mask, gt = (torch.from_numpy(G[k]).to(cuda) for k in ("odd_pred", "odd_mask", "dm_gt"))
    dm = DepthMetric(thresholds=[1.02, 1.05, 1.1, 1.2], depth_cap=cap)
    assert dm.metric_keys == ["d>1.02", "d>1.05", "d>1.1", "d>1.2", "rmse", "l1_err", "abs_rel"]
    metrics, depth = dm.compute_metrics(pred, gt, mask)
    got = np.stack([metrics[k].cpu().numpy() for k in dm.metric_keys], axis=1)
    np.testing.assert_allclose(got, G[f"{tag}_metrics"], rtol=3e-5, atol=1e-7)
    np.testing.assert_allclose(depth.cpu().numpy(), G[f"{tag}_depth"], rtol=3e-5)
    g = torch.Generator().manual_seed(5)
    B, H, W = 3, 224, 224
    p2 = 0.2 + 0.6 * torch.rand(B, 1, H, W, generator=g)
    g2 = 0.9 + 1.2 * torch.rand(B, 1, H, W, generator=g)
    m2 = (torch.rand(B, 1, H, W, generator=g) < 0.45).float()
    ref, dref = depth_metrics(p2, g2, m2, thresholds=[1.02, 1.05, 1.1, 1.2], depth_cap=cap)
    out, dout = dm.compute_metrics(p2.to(cuda), g2.to(cuda), m2.to(cuda))
    for k in dm.metric_keys:
        np.testing.assert_allclose(out[k].cpu().numpy(), ref[k].numpy(), rtol=5e-5, atol=1e-7)
    np.testing.assert_allclose(dout.cpu().numpy(), dref.numpy(), rtol=5e-5)


def test_depth_graph_training_step(cuda):
    """graph_depth.Graph.forward(training=True) with options/depth.yaml's loss weights (depth 1, intr 10): the losses equal the
    oracle's on the graph's own outputs, every parameter that takes part gets a finite gradient, AdamW steps lower the loss."""
    from zeroshape_b200 import ops
    from zeroshape_b200.model.compute_graph.graph_depth import Graph
    from zeroshape_b200.model.shape.implicit_train import FusedAdamW
    from zeroshape_b200.utils.util import EasyDict
    from test_gpu_graph import make_opt, synthetic_image_and_mask
    opt = make_opt(cuda)
    opt.loss_weight = EasyDict(depth=1, intr=10)
    opt.training = EasyDict(depth_loss=EasyDict(grad_reg=0.1, depth_inv=True, mask_shrink=False))
    torch.manual_seed(0)
    graph = Graph(opt).to(cuda).train()
    with torch.no_grad():
        graph.intr_proj.weight.normal_(0, 0.02)
    B = 2
    rgb, mask = synthetic_image_and_mask(B, 11)
    g = torch.Generator().manual_seed(12)
    depth_gt = (1.5 + 0.3 * torch.rand(B, 1, 224, 224, generator=g)) * mask
    intr = torch.tensor([[1.3875 * 224, 0, 112], [0, 1.3875 * 224, 112], [0, 0, 1.0]]).repeat(B, 1, 1)

    def batch():
        return EasyDict(idx=torch.arange(B), rgb_input_map=rgb.to(cuda), mask_input_map=mask.to(cuda), depth_input_map=depth_gt.to(cuda),
                        intr=intr.to(cuda))
    trainable = [p for p in graph.parameters() if p.requires_grad]
    optim = FusedAdamW(trainable, lr=3e-6, betas=(0.9, 0.95), weight_decay=0.05)      # random init: a larger step kills the output ReLU
    totals = []
    for it in range(4):
        var, loss = graph.forward(opt, batch(), training=True)
        total = opt.loss_weight.depth * loss.depth + opt.loss_weight.intr * loss.intr
        if it == 0:
            ref_d = midas_loss(var.depth_pred.detach().cpu(), depth_gt, mask, alpha=0.1)
            assert abs(loss.depth.item() - ref_d.item()) < 1e-4 * abs(ref_d.item()), (loss.depth.item(), ref_d.item())
            dist = ((var.seen_points_pred.detach() - var.seen_points_gt) ** 2).sum(-1)
            ref_i = (dist * var.validity_mask).sum() / (var.validity_mask.sum() + 1e-8)
            assert abs(loss.intr.item() - ref_i.item()) < 1e-5 * max(1e-6, abs(ref_i.item()))
        optim.zero_grad()
        total.backward()
        if it == 0:
            used = [p for p in trainable if p.grad is not None]
            assert len(used) > 370 and all(torch.isfinite(p.grad).all() for p in used), len(used)
            names = {n for n, p in graph.named_parameters() if p.grad is not None}
            for prefix in ("dpt_depth.scratch.output_conv.4", "dpt_depth.pretrained.model.patch_embed.backbone.stem.conv", "intr_head.0.linear1", "intr_proj"):
                assert any(n.startswith(prefix) for n in names), prefix
        optim.step()
        totals.append(total.item())
    print("depth-graph total loss per step:", [round(v, 4) for v in totals])
    assert min(totals[1:]) < totals[0] and all(np.isfinite(totals))
