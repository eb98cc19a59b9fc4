"""More golden generators (imported by make_golden.py): the reference's full encoder-side Graph.forward,
executed for real on CPU.  The only stand-in is the timm backbone class (FakeTimmHybridViT, restated
from timm 0.6.12 because the package is absent); every line of /root/reference glue runs unmodified:
model/compute_graph/graph_shape.py:115-148, model/depth/{dpt_depth,blocks,vit}.py, utils/camera.py,
utils/util.py, utils/layers.py, model/shape/seen_coord_enc.py (torchvision resnet50)."""
import numpy as np
import torch

GENERATORS = {}


def register(name):
    def deco(fn):
        GENERATORS[name] = fn
        return fn
    return deco

from _ref_import import install_fake_timm_factory, reference_opt


def synthetic_image_and_mask(B, seed, cx=112, cy=112, radius=80, H=224, W=224):
    """uniform-noise RGB inside a filled disc, white background."""
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(B, 3, H, W, generator=g)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    mask = (((yy - cy) ** 2 + (xx - cx) ** 2) < radius ** 2).float().view(1, 1, H, W).repeat(B, 1, 1, 1)
    return rgb * mask + (1 - mask), mask


@register("graph_encode")
def g_graph_encode():
    install_fake_timm_factory()
    from model.compute_graph.graph_shape import Graph
    from utils.util import EasyDict as edict
    from oracle.graph_params import graph_shape_param_shapes, seeded_state_dict
    opt = reference_opt()
    torch.manual_seed(0)
    graph = Graph(opt).eval()
    shapes = graph_shape_param_shapes()
    ref_sd = graph.state_dict()
    assert sorted(shapes) == sorted(ref_sd)
    assert all(tuple(ref_sd[k].shape) == tuple(shapes[k]) for k in shapes)
    sd = seeded_state_dict(shapes, seed=21)
    graph.load_state_dict(sd, strict=True)
    rgb, mask = synthetic_image_and_mask(1, seed=22, cx=104, cy=118, radius=78)
    var = edict(idx=torch.arange(1), rgb_input_map=rgb, mask_input_map=mask, pose_gt=False)
    with torch.no_grad():
        var = graph.forward(opt, var, training=False, get_loss=False)
        pts = torch.rand(1, 64, 3, generator=torch.Generator().manual_seed(23)) * 2 - 1
        logits, _ = graph.impl_network(var.latent_depth, None, pts)
    return dict(weight_seed=21, image_seed=22, disc=np.array([104, 118, 78]),
                keys=np.array(sorted(ref_sd.keys())), shapes=np.array([str(tuple(ref_sd[k].shape)) for k in sorted(ref_sd)]),
                depth_pred=var.depth_pred.numpy(), intr_pred=var.intr_pred.numpy(),
                seen_points=var.seen_points.numpy()[:, ::7].copy(), latent_depth=var.latent_depth.numpy(),
                points=pts.numpy(), logits=logits.numpy())
