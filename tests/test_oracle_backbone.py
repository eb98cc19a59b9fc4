"""CPU: encoder-side oracle (DPT-hybrid depth + intrinsics + geometry glue + CoordEncRes) against
 (1) golden vectors produced by the reference's own Graph.forward (tests/golden/graph_encode.npz),
 (2) the live reference when /root/reference is present,
 (3) the independent `transformers` BiT implementation for the third-party (timm) ResNetV2 arithmetic."""
import os

import numpy as np
import pytest
import torch

from _ref_import import reference_available
from oracle import backbone as BB
from oracle.graph_params import graph_shape_param_shapes, seeded_state_dict
from oracle.implicit import implicit_forward

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def synthetic_image_and_mask(B, seed, cx=112, cy=112, radius=80, H=224, W=224):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(B, 3, H, W, generator=g)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    mask = (((yy - cy) ** 2 + (xx - cx) ** 2) < radius ** 2).float().view(1, 1, H, W).repeat(B, 1, 1, 1)
    return rgb * mask + (1 - mask), mask


def test_state_dict_keys_match_reference_graph():
    g = np.load(os.path.join(GOLD, "graph_encode.npz"))
    shapes = graph_shape_param_shapes()
    assert sorted(shapes) == list(g["keys"])
    assert [str(tuple(shapes[k])) for k in sorted(shapes)] == list(g["shapes"])


def test_encoder_oracle_matches_reference_graph_golden():
    g = np.load(os.path.join(GOLD, "graph_encode.npz"))
    sd = seeded_state_dict(graph_shape_param_shapes(), int(g["weight_seed"]))
    cx, cy, r = [int(v) for v in g["disc"]]
    rgb, mask = synthetic_image_and_mask(1, int(g["image_seed"]), cx, cy, r)
    with torch.no_grad():
        out = BB.graph_shape_encode(sd, rgb, mask)
        logits, _ = implicit_forward(sd, out["latent_depth"], torch.from_numpy(g["points"]), prefix="impl_network.")
    np.testing.assert_allclose(out["depth_pred"].numpy(), g["depth_pred"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(out["intr_pred"].numpy(), g["intr_pred"], rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(out["seen_points"].numpy()[:, ::7], g["seen_points"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(out["latent_depth"].numpy(), g["latent_depth"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=1e-4)
    assert 0.2 < g["depth_pred"].mean() < 0.8 and np.abs(g["latent_depth"]).max() < 50     # non-degenerate fixture


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present")
def test_encoder_oracle_matches_live_reference_graph():
    from _ref_import import install_fake_timm_factory, reference_opt
    install_fake_timm_factory()
    from model.compute_graph.graph_shape import Graph
    from utils.util import EasyDict as edict
    opt = reference_opt()
    graph = Graph(opt).eval()
    sd = seeded_state_dict(graph_shape_param_shapes(), 5)
    graph.load_state_dict(sd, strict=True)
    rgb, mask = synthetic_image_and_mask(2, 6, 120, 100, 70)
    var = edict(idx=torch.arange(2), rgb_input_map=rgb, mask_input_map=mask, pose_gt=False)
    with torch.no_grad():
        var = graph.forward(opt, var, training=False, get_loss=False)
        out = BB.graph_shape_encode(sd, rgb, mask)
    for k, tol in (("depth_pred", 2e-5), ("seen_points", 5e-5), ("latent_depth", 1e-4)):
        assert (var[k] - out[k]).abs().max().item() < tol, k
    assert ((var.intr_pred - out["intr_pred"]).abs() / (out["intr_pred"].abs() + 1)).max().item() < 1e-6


def test_resnetv2_restatement_vs_transformers_bit():
    """timm's weight-standardised SAME convs + GroupNorm stem/stages are third-party code absent from
    /root/reference; `transformers.models.bit` is an independent port of the same network."""
    transformers = pytest.importorskip("transformers")
    from transformers import BitConfig, BitBackbone
    cfg = BitConfig(layer_type="bottleneck", depths=[3, 4, 9], hidden_sizes=[256, 512, 1024], embedding_size=64,
                    num_groups=32, global_padding="SAME", embedding_dynamic_padding=True, hidden_act="relu",
                    out_features=["stage1", "stage2", "stage3"], num_channels=3)
    torch.manual_seed(0)
    hf = BitBackbone(cfg).eval()
    sd = seeded_state_dict(graph_shape_param_shapes(), 7, implicit_prefix=None)
    pre = "dpt_depth.pretrained.model.patch_embed.backbone."
    hf_sd = hf.state_dict()
    mapped = {}
    for k in hf_sd:
        kk = k.replace("bit.embedder.convolution.", "stem.conv.").replace("bit.embedder.norm.", "stem.norm.")
        kk = kk.replace("bit.encoder.stages.", "stages.").replace(".layers.", ".blocks.")
        if pre + kk not in sd:
            continue
        mapped[k] = sd[pre + kk]
    missing = [k for k in hf_sd if k not in mapped and "num_batches" not in k]
    assert not missing, missing[:5]
    hf.load_state_dict(mapped, strict=False)
    x = torch.rand(1, 3, 224, 224, generator=torch.Generator().manual_seed(8)) * 2 - 1
    with torch.no_grad():
        want = hf(x).feature_maps
        got = BB.resnetv2_stem_stages(sd, x, pre)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert (a - b).abs().max().item() < 2e-4 * max(1.0, b.abs().max().item())
