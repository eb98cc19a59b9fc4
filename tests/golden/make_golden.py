"""Generate the committed golden vectors by running the REAL reference modules (imported from
/root/reference through the shims of _ref_import.py) on seeded inputs.  Runs only in the build
container (the GPU box has no /root/reference); the .npz files it writes are committed.

    python tests/golden/make_golden.py [name ...]

Weights are never stored: every fixture records the seed, and tests regenerate identical weights
with `_ref_import.fill_deterministic` (pure function of sorted key names, shapes and seed).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _ref_import import install_shims, fill_deterministic  # noqa: E402


def g_implicit():
    """model/shape/implicit.py Implicit, shipped config (options/shape.yaml:19-44)."""
    from model.shape.implicit import Implicit
    torch.manual_seed(0)
    m = Implicit(196, latent_dim=256, semantic=False, n_channels=256, n_blocks_attn=2, n_layers_mlp=8,
                 num_heads=8, posenc_3D=0, mlp_ratio=4., skip_in=[2, 4, 6], pos_perlayer=False).eval()
    fill_deterministic(m, seed=11)
    g = torch.Generator().manual_seed(12)
    latent = torch.randn(2, 197, 256, generator=g)
    pts = torch.rand(2, 96, 3, generator=g) * 3 - 1.5
    with torch.no_grad():
        logits, attn = m(latent, None, pts)
    return dict(seed=11, latent=latent.numpy(), points=pts.numpy(), logits=logits.numpy(), attn=attn.numpy())


def g_implicit_init():
    """Reference init scheme (implicit.py:235-249) at torch.manual_seed(0): pins pos_embed and statistics."""
    from model.shape.implicit import Implicit
    torch.manual_seed(0)
    m = Implicit(196, latent_dim=256, semantic=False, n_channels=256, n_blocks_attn=2, n_layers_mlp=8,
                 num_heads=8, posenc_3D=0, mlp_ratio=4., skip_in=[2, 4, 6], pos_perlayer=False).eval()
    sd = m.state_dict()
    return dict(pos_embed=sd["pos_embed"].numpy(),
                keys=np.array(sorted(sd.keys())), shapes=np.array([str(tuple(sd[k].shape)) for k in sorted(sd.keys())]))


GENERATORS = {"implicit": g_implicit, "implicit_init": g_implicit_init}


def main(names):
    install_shims()
    import make_golden_extra
    GENERATORS.update(make_golden_extra.GENERATORS)
    for name in names or sorted(GENERATORS):
        out = GENERATORS[name]()
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main(sys.argv[1:])
