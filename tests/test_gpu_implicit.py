"""GPU: implicit decoder (product, through the C ABI) against the oracle and the committed
reference golden vectors.  Tolerance: 1e-3 relative (north_star) -- the f32 engine is held to 1e-5."""
import os

import numpy as np
import pytest
import torch

from _ref_import import fill_deterministic
from oracle.implicit import implicit_forward, implicit_init
from oracle import eval3d as E

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _module(sd, cuda, engine):
    from zeroshape_b200.model.shape.implicit import Implicit
    m = Implicit(196, latent_dim=256, semantic=False, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8,
                 posenc_3D=0, mlp_ratio=4., skip_in=[2, 4, 6], pos_perlayer=False)
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda).eval()
    m.engine = engine
    return m


def rel_err(a, b):
    from parity import parity_rel
    return parity_rel(a, b)


def test_state_dict_is_reference_compatible(cuda):
    g = np.load(os.path.join(GOLD, "implicit_init.npz"))
    from zeroshape_b200.model.shape.implicit import Implicit
    m = Implicit(196, latent_dim=256, semantic=False, n_channels=256, n_blocks_attn=2, n_layers_mlp=8, num_heads=8,
                 posenc_3D=0, mlp_ratio=4., skip_in=[2, 4, 6], pos_perlayer=False)
    sd = m.state_dict()
    assert sorted(sd) == list(g["keys"])
    assert [str(tuple(sd[k].shape)) for k in sorted(sd)] == list(g["shapes"])
    np.testing.assert_array_equal(sd["pos_embed"].numpy(), g["pos_embed"])
    assert not m.pos_embed.requires_grad


def test_f32_engine_matches_reference_golden(cuda):
    g = np.load(os.path.join(GOLD, "implicit.npz"))
    sd = fill_deterministic(implicit_init(0, recentre=False), int(g["seed"]))
    m = _module(sd, cuda, "f32")
    logits, attn = m(torch.from_numpy(g["latent"]).to(cuda), None, torch.from_numpy(g["points"]).to(cuda))
    assert rel_err(logits, torch.from_numpy(g["logits"])) < 1e-4      # plain-fp32 engine: 10x inside the 1e-3 bar
    np.testing.assert_allclose(attn.cpu().numpy(), g["attn"], rtol=0, atol=2e-7)


@pytest.mark.parametrize("P", [1, 127, 4096])
def test_f32_engine_matches_oracle_ragged(cuda, P):
    sd = implicit_init(seed=1)
    m = _module(sd, cuda, "f32")
    g = torch.Generator().manual_seed(P)
    lat, pts = torch.randn(2, 197, 256, generator=g), torch.rand(2, P, 3, generator=g) * 3 - 1.5
    with torch.no_grad():
        ref, ref_attn = implicit_forward(sd, lat, pts)
    logits, attn = m(lat.to(cuda), None, pts.to(cuda))
    assert rel_err(logits, ref) < 2e-4      # plain-fp32 engine: 5x inside the 1e-3 bar
    assert (attn.cpu() - ref_attn).abs().max() < 2e-7
    lg2, none = m(lat.to(cuda), None, pts.to(cuda), need_attn=False)
    assert none is None and torch.equal(lg2, logits)


def test_grid_occupancy_matches_reference_slice_loop(cuda):
    sd = implicit_init(seed=2)
    m = _module(sd, cuda, "f32")
    g = torch.Generator().manual_seed(9)
    lat = torch.randn(2, 197, 256, generator=g)
    n = 17
    ref = E.level_grid(sd, lat, n, -1.5, 1.5)
    occ = m.grid_occupancy(lat.to(cuda), n, -1.5, 1.5)
    assert occ.shape == (2, n, n, n)
    assert (occ.cpu() - ref).abs().max() < 2e-6
    # identical thresholded voxel grid outside a |logit| < 1e-4 band
    band = (ref - 0.5).abs() > 2.5e-5
    assert torch.equal((occ.cpu() > 0.5)[band], (ref > 0.5)[band])
    assert 0.05 < (ref > 0.5).float().mean() < 0.95
    # slab sharding invariance (multi-GPU partitioning A): any split of x gives the same grid
    parts = [m.grid_occupancy(lat.to(cuda), n, -1.5, 1.5, x0, x1) for x0, x1 in ((0, 5), (5, 6), (6, 17))]
    assert torch.equal(torch.cat(parts, dim=1), occ)
    # dense grid op == torch.linspace/meshgrid of the reference
    from zeroshape_b200 import ops
    assert torch.equal(ops.dense_grid(129, -1.5, 1.5, 0, 2, cuda).cpu(), E.dense_grid(129, -1.5, 1.5)[0, :2])


def test_query_point_gradients_are_refused(cuda):
    """The training path differentiates w.r.t. parameters and latents; gradients w.r.t. the query points do not exist."""
    m = _module(implicit_init(0), cuda, "f32")
    lat = torch.randn(1, 197, 256, device=cuda)
    with pytest.raises(NotImplementedError):
        m(lat, None, torch.rand(1, 8, 3, device=cuda, requires_grad=True))
