"""CPU: the HOST logic of the implicit decoder's training step (model/shape/implicit_train.py: saved activations, gradient routing
through both attention blocks, the latent branch, the skip connections of the occupancy MLP) with the kernels replaced by per-op torch
stand-ins (tests/fake_ops.py), against torch autograd over the oracle restatement of the reference's Implicit.forward.  The kernels
themselves are checked on the GPU (tests/test_gpu_train.py)."""
import torch

import fake_ops
from oracle.implicit import implicit_forward, implicit_init


def test_decoder_tape_matches_oracle_autograd(monkeypatch):
    fake_ops.install_train(monkeypatch)
    from zeroshape_b200.model.shape import implicit_train as IT
    from zeroshape_b200.model.shape.implicit import Implicit
    sd = implicit_init(seed=21)
    net = Implicit(196, latent_dim=256, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8, skip_in=[2, 4, 6],
                   pos_perlayer=False, drop_path=0.0)
    net.load_state_dict(sd)
    net.train()
    g = torch.Generator().manual_seed(3)
    B, P = 2, 37
    lat = torch.randn(B, 197, 256, generator=g)
    pts = torch.rand(B, P, 3, generator=g) * 2 - 1
    wgt = torch.randn(B, P, generator=g)
    sd_ref = {k: v.clone().requires_grad_(k != "pos_embed") for k, v in sd.items()}
    lat_ref = lat.clone().requires_grad_(True)
    logits_ref, _ = implicit_forward(sd_ref, lat_ref, pts)
    (logits_ref * wgt).sum().backward()
    with torch.no_grad():
        logits, tape = IT.train_forward(net, lat, pts)
        assert (logits - logits_ref).abs().max().item() < 1e-4
        G, dz = IT.train_backward(net, tape, wgt)
    for name, p in net.named_parameters():
        if name == "pos_embed":
            continue
        gref = sd_ref[name].grad
        got = G.get(p)
        assert got is not None, name
        assert ((got - gref).norm() / gref.norm().clamp_min(1e-30)).item() < 2e-4, name
    assert ((dz - lat_ref.grad).norm() / lat_ref.grad.norm()).item() < 2e-4
