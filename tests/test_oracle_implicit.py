"""CPU: the implicit-decoder oracle against golden vectors produced by the real reference module
(tests/golden/make_golden.py, model/shape/implicit.py), and against the reference itself when
/root/reference is present."""
import os

import numpy as np
import pytest
import torch

from _ref_import import fill_deterministic, reference_available, install_shims
from oracle.implicit import implicit_forward, implicit_init, implicit_param_shapes, sincos_pos_embed_2d

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _weights(seed):
    base = implicit_init(seed=0, recentre=False)
    return fill_deterministic(base, seed)


def test_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "implicit.npz"))
    sd = _weights(int(g["seed"]))
    with torch.no_grad():
        logits, attn = implicit_forward(sd, torch.from_numpy(g["latent"]), torch.from_numpy(g["points"]))
    # same op sequence as the reference on the same machine type -> bit-exact; allow 1e-6 across CPUs
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(attn.numpy(), g["attn"], rtol=0, atol=1e-7)


def test_param_names_shapes_and_pos_embed_match_reference():
    g = np.load(os.path.join(GOLD, "implicit_init.npz"))
    shapes = implicit_param_shapes()
    assert sorted(shapes) == list(g["keys"])
    assert [str(tuple(shapes[k])) for k in sorted(shapes)] == list(g["shapes"])
    pe = torch.from_numpy(sincos_pos_embed_2d(256, 14)).float().unsqueeze(0).numpy()
    np.testing.assert_array_equal(pe, g["pos_embed"])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present")
def test_oracle_matches_live_reference_module():
    install_shims()
    from model.shape.implicit import Implicit
    m = Implicit(196, latent_dim=256, semantic=False, n_channels=256, n_blocks_attn=2, n_layers_mlp=8,
                 num_heads=8, posenc_3D=0, mlp_ratio=4., skip_in=[2, 4, 6], pos_perlayer=False).eval()
    sd = fill_deterministic(m, 5)
    g = torch.Generator().manual_seed(6)
    lat, pts = torch.randn(1, 197, 256, generator=g), torch.rand(1, 257, 3, generator=g) * 3 - 1.5
    with torch.no_grad():
        a, b = m(lat, None, pts)
        c, d = implicit_forward(sd, lat, pts)
    assert torch.equal(a, c) and torch.equal(b, d)


def test_recentred_init_has_both_signs():
    sd = implicit_init(seed=0)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        lg, _ = implicit_forward(sd, torch.randn(1, 197, 256, generator=g), torch.rand(1, 2000, 3, generator=g) * 3 - 1.5)
    frac = (lg > 0).float().mean().item()
    assert 0.2 < frac < 0.8
