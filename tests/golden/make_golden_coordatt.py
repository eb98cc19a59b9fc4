"""Golden vectors for the transformer seen-surface encoder: the REAL reference module
(model/shape/seen_coord_enc.py CoordEncAtt, imported from /root/reference through the timm shim of _ref_import.py) with seeded
weights on a seeded XYZ map -> tests/golden/coordatt.npz (inputs, output, key/shape list, the fixed sin-cos table)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def inputs(B=2, H=32, W=32, seed=7):
    g = torch.Generator().manual_seed(seed)
    coord = torch.randn(B, H, W, 3, generator=g) * 0.4
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    mask = (((yy - H / 2) ** 2 + (xx - W / 2.2) ** 2) < (0.36 * H) ** 2).unsqueeze(0).repeat(B, 1, 1)
    mask[1, :9, :] = False                        # a fully invalid window row
    return coord * mask.unsqueeze(-1), mask


def main():
    from _ref_import import install_shims
    install_shims()
    from model.shape.seen_coord_enc import CoordEncAtt
    from oracle.graph_params import seeded_state_dict
    torch.manual_seed(0)
    mod = CoordEncAtt(embed_dim=256, n_blocks=3, num_heads=8, win_size=8).eval()
    sincos = mod.coord_embed.two_d_pos_embed.detach().clone()
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    sd = seeded_state_dict(shapes, seed=31, implicit_prefix=None)
    sd["coord_embed.two_d_pos_embed"] = sincos                     # the fixed table stays what the module computed
    mod.load_state_dict(sd, strict=True)
    coord, mask = inputs()
    with torch.no_grad():
        out = mod(coord.clone(), mask)
    keys = sorted(shapes)
    np.savez_compressed(os.path.join(HERE, "coordatt.npz"), coord=coord.numpy(), mask=mask.numpy(), out=out.numpy(), sincos=sincos.numpy(),
                        keys=np.array(keys), shapes=np.array([str(shapes[k]) for k in keys]), weight_seed=31)
    print("wrote coordatt.npz", out.shape, float(out.abs().mean()))


if __name__ == "__main__":
    main()
