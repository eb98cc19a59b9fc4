"""CPU: the HOST logic of the transformer seen-surface encoder's training tape (model/shape/seen_coord_att_train.py: op order, closures,
gradient routing through the window selection / cls concatenation / front end) with the kernels replaced by per-op torch stand-ins, against
torch autograd over the oracle restatement (pinned to the reference module).  The kernels themselves are checked on the GPU
(tests/test_gpu_coordatt_train.py)."""
import torch

import fake_ops
from oracle import backbone as BB
from oracle.graph_params import seeded_state_dict


def test_coord_att_tape_routes_every_gradient(monkeypatch):
    fake_ops.install_train(monkeypatch)
    from zeroshape_b200.model.depth import dpt_train as T
    from zeroshape_b200.model.shape import seen_coord_att_train as CAT
    from zeroshape_b200.model.shape.seen_coord_enc import CoordEncAtt
    mod = CoordEncAtt(embed_dim=64, n_blocks=2, num_heads=4, win_size=4, drop_path=0.0)
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    sd = seeded_state_dict(shapes, seed=41, implicit_prefix=None)
    sd["coord_embed.two_d_pos_embed"] = mod.state_dict()["coord_embed.two_d_pos_embed"].clone()
    mod.load_state_dict(sd, strict=True)
    mod.train()
    g = torch.Generator().manual_seed(42)
    B, H, W = 2, 12, 8
    mask = torch.rand(B, H, W, generator=g) < 0.7
    coord = torch.randn(B, H, W, 3, generator=g) * 0.4 * mask.unsqueeze(-1)
    wgt = torch.randn(B, 1 + (H // 4) * (W // 4), 64, generator=g)
    sd_ref = {"coord_encoder." + k: v.clone().requires_grad_(k != "coord_embed.two_d_pos_embed") for k, v in sd.items()}
    coord_r = coord.clone().requires_grad_(True)
    out_ref = BB.coord_enc_att_forward(sd_ref, coord_r, mask, heads=4, ws=4)
    (out_ref * wgt).sum().backward()
    with torch.no_grad():
        tp = T.Tape()
        cd = coord.contiguous()
        out = CAT.train_forward(tp, mod, cd, mask.float().contiguous())
        assert (out - out_ref).abs().max().item() < 1e-5
        tp.add(out, wgt)
        tp.backward()
        dcoord = tp.pop(cd)
    for name, p in mod.named_parameters():
        if name == "coord_embed.two_d_pos_embed":
            continue
        assert id(p) in tp.pgrads, name
        gref = sd_ref["coord_encoder." + name].grad
        assert ((tp.pgrads[id(p)] - gref).norm() / gref.norm()).item() < 1e-4, name
    m3 = mask.unsqueeze(-1)
    assert ((dcoord * m3 - coord_r.grad * m3).norm() / (coord_r.grad * m3).norm()).item() < 1e-4
    # DropPath in train mode: a zero keep-probability draw drops a whole sample's residual branch but the output stays finite
    mod.drop_path = 0.5
    with torch.no_grad():
        out_dp = CAT.train_forward(T.Tape(), mod, cd, mask.float().contiguous())
    assert torch.isfinite(out_dp).all() and not torch.allclose(out_dp, out)
