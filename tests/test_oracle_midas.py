"""CPU: the oracle restatement of the reference's MiDaS depth loss (oracle/midas.py) and its closed-form gradient against
golden vectors produced by the REAL reference module + torch autograd (tests/golden/midas.npz, make_golden_midas.py)."""
import os

import numpy as np
import pytest
import torch

from oracle.midas import depth_metrics, erode_mask, midas_loss, midas_loss_grad

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "midas.npz"))


@pytest.mark.parametrize("name", ["small", "odd", "empty_image"])
@pytest.mark.parametrize("alpha", [0.1, 0.0])
def test_midas_oracle_matches_reference(name, alpha):
    pred, gt, mask = (torch.from_numpy(G[f"{name}_{k}"]) for k in ("pred", "gt", "mask"))
    tag = f"{name}_a{int(alpha * 10)}"
    loss = midas_loss(pred, gt, mask, alpha=alpha)
    assert abs(float(loss) - float(G[f"{tag}_loss"])) < 2e-6 * max(1.0, abs(float(G[f"{tag}_loss"])))
    grad = midas_loss_grad(pred, gt, mask, alpha=alpha)
    ref = torch.from_numpy(G[f"{tag}_grad"]).double()
    assert (grad - ref).abs().max().item() < 1e-5 * ref.abs().max().item() + 1e-9, (grad - ref).abs().max().item()


def test_midas_oracle_gradient_matches_autograd_of_itself():
    g = torch.Generator().manual_seed(5)
    pred = (0.3 + 0.5 * torch.rand(2, 1, 18, 22, generator=g)).double().requires_grad_(True)
    gt = (1.0 + torch.rand(2, 1, 18, 22, generator=g)).double()
    mask = (torch.rand(2, 1, 18, 22, generator=g) < 0.7).float()
    midas_loss(pred, gt * mask, mask).backward()
    an = midas_loss_grad(pred.detach().float(), (gt * mask).float(), mask)
    assert (an - pred.grad).abs().max().item() < 1e-5 * pred.grad.abs().max().item()


def test_midas_oracle_mask_shrink():
    pred, gt, mask = (torch.from_numpy(G[f"shrink_{k}"]) for k in ("pred", "gt", "mask"))
    er = erode_mask(mask)
    assert np.array_equal(er.numpy(), G["shrink_eroded"]) and 0 < er.sum() < mask.sum()
    assert abs(float(midas_loss(pred, gt, er)) - float(G["shrink_loss"])) < 2e-6 * abs(float(G["shrink_loss"]))
    ref = torch.from_numpy(G["shrink_grad"]).double()
    assert (midas_loss_grad(pred, gt, er) - ref).abs().max().item() < 1e-5 * ref.abs().max().item()


@pytest.mark.parametrize("tag,cap", [("dm", None), ("dmcap", 1.6)])
def test_depth_metric_oracle_matches_reference(tag, cap):
    pred, mask, gt = torch.from_numpy(G["odd_pred"]), torch.from_numpy(G["odd_mask"]), torch.from_numpy(G["dm_gt"])
    metrics, depth = depth_metrics(pred, gt, mask, thresholds=[1.02, 1.05, 1.1, 1.2], depth_cap=cap)
    got = np.stack([v.numpy() for v in metrics.values()], axis=1)
    np.testing.assert_allclose(got, G[f"{tag}_metrics"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(depth.numpy(), G[f"{tag}_depth"], rtol=2e-5)
