"""CPU: the oracle restatement of the transformer seen-surface encoder (oracle/backbone.py::coord_enc_att_forward) against the
golden output of the REAL reference module (tests/golden/coordatt.npz); key / shape parity of the product mirror's state_dict
(parameter containers only -- no compute without a GPU) and of its fixed sin-cos table."""
import os

import numpy as np
import torch

from oracle import backbone as BB
from oracle.graph_params import seeded_state_dict

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "coordatt.npz"))


def golden_sd(prefix=""):
    shapes = {k: eval(s) for k, s in zip(G["keys"], G["shapes"])}
    sd = seeded_state_dict(shapes, seed=int(G["weight_seed"]), implicit_prefix=None)
    sd["coord_embed.two_d_pos_embed"] = torch.from_numpy(G["sincos"])
    return {prefix + k: v for k, v in sd.items()}


def test_coord_enc_att_oracle_matches_reference():
    sd = golden_sd("coord_encoder.")
    out = BB.coord_enc_att_forward(sd, torch.from_numpy(G["coord"]), torch.from_numpy(G["mask"]), "coord_encoder.", heads=8, ws=8)
    ref = torch.from_numpy(G["out"])
    assert out.shape == ref.shape and (out - ref).abs().max().item() < 2e-5 * ref.abs().max().item()


def test_mirror_state_dict_and_sincos_table():
    from zeroshape_b200.model.shape.seen_coord_enc import CoordEncAtt
    mod = CoordEncAtt(embed_dim=256, n_blocks=3, num_heads=8, win_size=8)
    sd = mod.state_dict()
    assert sorted(sd) == list(G["keys"])
    assert [str(tuple(sd[k].shape)) for k in sorted(sd)] == list(G["shapes"])
    np.testing.assert_allclose(sd["coord_embed.two_d_pos_embed"].numpy(), G["sincos"], atol=1e-6)
    assert not mod.coord_embed.two_d_pos_embed.requires_grad
