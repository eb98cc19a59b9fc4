"""Golden vectors for the MiDaS depth loss: the REAL reference module (model/depth/midas_loss.py, imported from
/root/reference in the build container) evaluated forward + autograd on seeded inputs -> tests/golden/midas.npz.
Test infrastructure; the .npz is committed (the GPU box has no /root/reference)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ZEROSHAPE_REFERENCE", "/root/reference")


def cases():
    g = torch.Generator().manual_seed(2024)
    out = []
    for name, B, H, W, fill in (("small", 2, 20, 24, 0.6), ("odd", 3, 37, 29, 0.4), ("empty_image", 2, 16, 16, 0.5)):
        pred = (0.2 + 0.6 * torch.rand(B, 1, H, W, generator=g))
        gt = (0.8 + 1.4 * torch.rand(B, 1, H, W, generator=g))
        mask = (torch.rand(B, 1, H, W, generator=g) < fill).float()
        if name == "empty_image":
            mask[1] = 0
        out.append((name, pred, gt * mask, mask))
    return out


def main():
    sys.path.insert(0, REF)
    from model.depth.midas_loss import MidasLoss
    store = {}
    for name, pred, gt, mask in cases():
        for alpha in (0.1, 0.0):
            fn = MidasLoss(alpha=alpha, inverse_depth=True, shrink_mask=False)
            p = pred.clone().requires_grad_(True)
            loss = fn(p, gt, mask)
            loss.backward()
            tag = f"{name}_a{int(alpha * 10)}"
            store[f"{tag}_loss"] = loss.detach().numpy()
            store[f"{tag}_grad"] = p.grad.numpy()
        store[f"{name}_pred"], store[f"{name}_gt"], store[f"{name}_mask"] = pred.numpy(), gt.numpy(), mask.numpy()
    # mask shrinking (training.depth_loss.mask_shrink): a blocky mask so that some 4x4 blocks survive the erosion
    g = torch.Generator().manual_seed(99)
    pred = 0.2 + 0.6 * torch.rand(2, 1, 24, 28, generator=g)
    gt = 0.8 + 1.4 * torch.rand(2, 1, 24, 28, generator=g)
    mask = torch.nn.functional.interpolate((torch.rand(2, 1, 6, 7, generator=g) < 0.7).float(), (24, 28), mode="nearest")
    mask[:, :, 5, 9] = 0
    fn = MidasLoss(alpha=0.1, inverse_depth=True, shrink_mask=True)
    p = pred.clone().requires_grad_(True)
    loss = fn(p, gt * mask, mask)
    loss.backward()
    store.update(shrink_pred=pred.numpy(), shrink_gt=(gt * mask).numpy(), shrink_mask=mask.numpy(), shrink_loss=loss.detach().numpy(),
                 shrink_grad=p.grad.numpy(), shrink_eroded=fn.erode_mask(mask).float().numpy())
    # DepthMetric (utils/eval_depth.py) on the "odd" case (depth in [0.2, 0.8], gt > 0 inside the mask), with and without a depth cap
    from utils.eval_depth import DepthMetric
    name, pred, gt, mask = cases()[1]
    gt = torch.where(mask > 0.5, gt, torch.ones_like(gt))          # the reference divides by target inside the mask only
    for tag, cap in (("dm", None), ("dmcap", 1.6)):
        dm = DepthMetric(thresholds=[1.02, 1.05, 1.1, 1.2], depth_cap=cap)
        metrics, depth = dm.compute_metrics(pred, gt, mask)
        store[f"{tag}_metrics"] = np.stack([metrics[k].numpy() for k in dm.metric_keys], axis=1)
        store[f"{tag}_depth"] = depth.numpy()
    store["dm_gt"] = gt.numpy()
    np.savez_compressed(os.path.join(HERE, "midas.npz"), **store)
    print("wrote", os.path.join(HERE, "midas.npz"), {k: v.shape for k, v in store.items() if k.endswith("loss")})


if __name__ == "__main__":
    main()
