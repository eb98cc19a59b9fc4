"""Host-side mirror of the reference's `model/compute_graph/graph_shape.py` Graph: same constructor,
sub-module attribute names (dpt_depth, intr_head, intr_pool, intr_proj, coord_encoder, rgb_encoder,
impl_network, loss_fns), state_dict keys and `forward(opt, var, training, get_loss)` contract; all layer
math runs in libzeroshape_b200.so.

    forward fills var.{latent_semantic, depth_pred, intr_pred, validity_mask, seen_points, latent_depth, pose}
    (reference: model/compute_graph/graph_shape.py:115-192)

Training, decoder slice: with a GT batch (`gt_sample_points`, `gt_sample_sdf`,
`depth_input_map`, `intr`, `pose_gt`) the forward also fills seen_points_gt / gt_points_cam / gt_surf_points /
pred_sample_occ (graph_shape.py:155-185) and `compute_loss` returns the shape (BCE) and intrinsics losses
(utils/loss.py:18-41).  Gradients exist for every module: with `optim.fix_dpt` the depth estimator runs on the inference kernels and only
coord_encoder + impl_network train; otherwise the whole encoder side runs on one tape (model/depth/dpt_train.py) and the
shape loss reaches dpt_depth / intr_head / intr_proj through the unprojected, normalised seen surface.  The MiDaS depth loss (model/depth/midas_loss.py)
belongs to the depth-engine row and raises.
"""
import weakref

import torch
import torch.nn as nn

from ... import ops
from ...utils.layers import Bottleneck_Conv
from ...utils.util import EasyDict as edict
from ..depth.dpt_depth import DPTDepthModel
from ..shape.implicit import Implicit
from ..shape.seen_coord_enc import CoordEncAtt, CoordEncRes
from ...utils.loss import Loss


_ENCODER_GRAPHS = weakref.WeakKeyDictionary()     # Graph instance -> its captured encoder graphs (see Graph._encode_graphed)


class Graph(nn.Module):

    def __init__(self, opt):
        super().__init__()
        self.intr_feat_channels = 768
        self.intr_head = nn.Sequential(Bottleneck_Conv(768, kernel_size=3), Bottleneck_Conv(768, kernel_size=3))
        self.intr_pool = nn.AdaptiveAvgPool2d((1, 1))          # attribute kept; pooling runs in the library
        self.intr_proj = nn.Linear(768, 3)
        nn.init.zeros_(self.intr_proj.weight)                  # graph_shape.py:27-28
        nn.init.zeros_(self.intr_proj.bias)
        self.dpt_depth = DPTDepthModel(backbone="vitb_rn50_384")
        self.load_pretrained_depth(opt)
        if opt.optim.fix_dpt:
            for m in (self.dpt_depth, self.intr_head, self.intr_proj):
                for p in m.parameters():
                    p.requires_grad_(False)
        if opt.arch.depth.encoder == "resnet":
            opt.arch.depth.dsp = 1                              # graph_shape.py:41-43
            self.coord_encoder = CoordEncRes(opt)
        else:                                                   # graph_shape.py:44-46 (inference only here)
            self.coord_encoder = CoordEncAtt(embed_dim=opt.arch.latent_dim, n_blocks=opt.arch.depth.n_blocks,
                                             num_heads=opt.arch.num_heads, win_size=opt.arch.win_size // opt.arch.depth.dsp)
        if opt.arch.rgb.encoder:
            raise NotImplementedError("RGB branch is 'not used in final model' (graph_shape.py:48)")
        self.rgb_encoder = None
        feat_res = opt.H // opt.arch.win_size
        self.impl_network = Implicit(feat_res ** 2, latent_dim=opt.arch.latent_dim, semantic=False,
                                     n_channels=opt.arch.impl.n_channels, n_blocks_attn=opt.arch.impl.att_blocks,
                                     n_layers_mlp=opt.arch.impl.mlp_layers, num_heads=opt.arch.num_heads,
                                     posenc_3D=opt.arch.impl.posenc_3D, mlp_ratio=opt.arch.impl.mlp_ratio,
                                     skip_in=opt.arch.impl.skip_in, pos_perlayer=opt.arch.impl.posenc_perlayer)
        self.loss_fns = Loss(opt)

    def load_pretrained_depth(self, opt):
        """graph_shape.py:69-87 (checkpoint bootstrap of the depth sub-network)."""
        def child(sd, prefix):
            return {k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + ".")}
        if opt.pretrain.depth:
            ckpt = torch.load(opt.pretrain.depth, map_location="cpu")
            self.dpt_depth.load_state_dict(child(ckpt["graph"], "dpt_depth"))
            self.intr_head.load_state_dict(child(ckpt["graph"], "intr_head"))
            self.intr_proj.load_state_dict(child(ckpt["graph"], "intr_proj"))
        elif opt.arch.depth.pretrained:
            ckpt = torch.load(opt.arch.depth.pretrained, map_location="cpu")
            self.dpt_depth.load_state_dict(ckpt["model_state_dict"])

    def intr_param2mtx(self, opt, intr_params):
        """[B,3] -> [B,3,3] (graph_shape.py:89-113)."""
        return ops.intr_param2mtx(intr_params.float().contiguous(), opt.H, opt.W)

    # ---- the inference encoder (image -> depth, intrinsics, seen surface, latents) replayed from a CUDA graph ----------------
    # ~300 launches of 5-40 us of device time each: issued one by one from Python the HOST is the limiter (a floor of ~25 us
    # per layer whatever its size, tools/diag_encoder_layers.py).  Everything on the path is stream-ordered, so the eval-mode
    # forward of a given input shape is captured once and replayed; the capture is dropped whenever a parameter or buffer
    # changes (version counters / storage), the precision policy changes, or an op timer is active.
    @property
    def _encoder_graphs(self):
        """{(input shapes, device): capture}; kept outside the module's __dict__ so that pickling / deepcopy of the Graph
        (torch.save(graph), EMA copies) never meets a CUDAGraph object."""
        return _ENCODER_GRAPHS.setdefault(self, {})

    def _encoder_signature(self):
        sig = [ops.ENCODER_ENGINE, ops.ENCODER_PRECISION]
        for m in (self.dpt_depth, self.intr_head, self.intr_proj, self.coord_encoder):
            for t in list(m.parameters()) + list(m.buffers()):
                sig.append(t._version)
                sig.append(t.data_ptr())
        return tuple(sig)

    def _encoder_graph_ok(self, var, full_train):
        if full_train or not ops.ENCODER_CUDA_GRAPH or getattr(self, "_in_encoder_capture", False):
            return False
        rgb = var.rgb_input_map
        return (isinstance(self.coord_encoder, CoordEncRes) and not self.coord_encoder.training and not self.dpt_depth.training
                and isinstance(rgb, torch.Tensor) and rgb.is_cuda and not torch.cuda.is_current_stream_capturing())

    def _encode_graphed(self, opt, var):
        from ..._native import lib
        rgb = var.rgb_input_map.float().contiguous()
        mask = var.mask_input_map.float().contiguous()
        key = (tuple(rgb.shape), tuple(mask.shape), rgb.device.index)
        sig = self._encoder_signature()
        cache = self._encoder_graphs
        ent = cache.get(key)
        if ent is None or ent["sig"] != sig:
            cache.pop(key, None)
            while len(cache) >= 4:                      # every capture owns its activations: keep the four most recent shapes
                cache.pop(next(iter(cache)))
            s_rgb, s_mask = rgb.clone(), mask.clone()

            def run():
                v = edict(idx=torch.arange(rgb.shape[0]), rgb_input_map=s_rgb, mask_input_map=s_mask, pose_gt=False)
                self._in_encoder_capture = True
                try:
                    v = self.forward(opt, v, training=False, get_loss=False)
                finally:
                    self._in_encoder_capture = False
                return {"depth_pred": v.depth_pred, "intr_pred": v.intr_pred, "seen_points": v.seen_points,
                        "latent_depth": v.latent_depth, "validity_mask": v.validity_mask, "mean": self.last_mean,
                        "scale": self.last_scale}
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run()                                   # packs the weights, sizes the workspace pools
            torch.cuda.current_stream().wait_stream(side)
            n0 = lib.zs_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                out = run()
            ent = {"sig": self._encoder_signature(), "graph": g, "rgb": s_rgb, "mask": s_mask, "out": out,
                   "launches": int(lib.zs_launch_count() - n0)}
            cache[key] = ent
        ent["rgb"].copy_(rgb)
        ent["mask"].copy_(mask)
        ent["graph"].replay()
        lib.zs_launch_count_add(ent["launches"])
        return {k: v.clone() for k, v in ent["out"].items()}     # the static outputs are overwritten by the next replay

    def forward(self, opt, var, training=False, get_loss=True):
        batch_size = len(var.idx)
        depth_params = [p for m in (self.dpt_depth, self.intr_head, self.intr_proj) for p in m.parameters()]
        full_train = training and torch.is_grad_enabled() and any(p.requires_grad for p in depth_params)
        graphed = self._encoder_graph_ok(var, full_train)
        if graphed:
            with torch.no_grad():
                enc = self._encode_graphed(opt, var)
            var.latent_semantic = None
            var.depth_pred, var.intr_pred, var.seen_points = enc["depth_pred"], enc["intr_pred"], enc["seen_points"]
            var.latent_depth, var.validity_mask = enc["latent_depth"], enc["validity_mask"]
            self.last_mean, self.last_scale = enc["mean"], enc["scale"]
            mask = var.mask_input_map.float().contiguous()
        if full_train:
            # default options/shape.yaml (fix_dpt: false): the shape loss reaches the depth estimator through the seen surface.
            # One tape from the image to latent_depth (model/depth/dpt_train.py), hand-written backward for every layer.
            if not self.coord_encoder.training:
                raise NotImplementedError("zeroshape_b200 Graph: training the depth estimator with the seen-surface encoder in "
                                          "eval mode is not supported (call graph.train())")
            from ..depth.dpt_train import EncoderTrainFn
            enc_params = depth_params + list(self.coord_encoder.parameters())
            var.latent_semantic = None
            mask = var.mask_input_map.float().contiguous()
            var.validity_mask = (mask > 0.5).float().view(batch_size, -1)
            var.depth_pred, var.intr_pred, var.seen_points, var.latent_depth = EncoderTrainFn.apply(
                self, opt, var.rgb_input_map, mask, *enc_params)
        with torch.no_grad():
            if not full_train and not graphed:
                var.latent_semantic = None
                var.depth_pred = self.dpt_depth(var.rgb_input_map, get_feat=False)
                feat = self.dpt_depth.last_feat_nhwc                                   # layer_4 [B,7,7,768] NHWC
                if not hasattr(self, "_intr_cache"):
                    from ...packing import PackCache
                    self._intr_cache = PackCache(self.intr_head)
                self._intr_cache.refresh()
                f = self.intr_head[0].run_nhwc(feat, self._intr_cache, "h0")
                f = self.intr_head[1].run_nhwc(f, self._intr_cache, "h1")
                intr_params = ops.linear(ops.avgpool_nhwc(f), self.intr_proj.weight, self.intr_proj.bias)
                var.intr_pred = self.intr_param2mtx(opt, intr_params)
                mask = var.mask_input_map.float().contiguous()
                var.validity_mask = (mask > 0.5).float().view(batch_size, -1)
                # unproject + masked mean / max-norm + normalise + zero background: one launch, no host sync
                var.seen_points, self.last_mean, self.last_scale = ops.unproject_normalize(var.depth_pred, mask, var.intr_pred)
                if isinstance(self.coord_encoder, CoordEncAtt):
                    # graph_shape.py:141-150: mask-aware resampling of the XYZ map to H/dsp x W/dsp, NHWC + boolean mask
                    from ...utils.util import interpolate_coordmap
                    seen_map = var.seen_points.view(batch_size, opt.H, opt.W, 3).permute(0, 3, 1, 2).contiguous()
                    seen_dsp, mask_dsp = interpolate_coordmap(seen_map, var.mask_input_map.float(),
                                                              (opt.H // opt.arch.depth.dsp, opt.W // opt.arch.depth.dsp))
                    coord, att_mask = seen_dsp.permute(0, 2, 3, 1).contiguous(), mask_dsp.squeeze(1) > 0.5
                    if not self.coord_encoder.training:
                        var.latent_depth = self.coord_encoder(coord, att_mask)
                else:
                    # interpolate_coordmap at dsp=1 (utils/util.py:336-345) is the identity resample followed by / (1 + 1e-6)
                    coord = ops.axpby(var.seen_points.view(batch_size, opt.H, opt.W, 3), 1.0 / (1.0 + 1.e-6))
                    if not self.coord_encoder.training:
                        var.latent_depth = self.coord_encoder.forward_nhwc(coord)
            var.pose = var.pose_gt if "pose_gt" in var else False
            if "gt_sample_points" in var and "gt_sample_sdf" in var:
                # graph_shape.py:157-182: normalising factors from the GT seen surface, GT points -> camera frame -> normalised
                gt_pts, self.gt_mean, self.gt_scale = ops.unproject_normalize(var.depth_input_map.float(), mask, var.intr.float())
                var.seen_points_gt = gt_pts
                R_gt, T_gt = var.pose_gt[:, :, :3].float(), var.pose_gt[:, :, 3:].float()
                cam = (R_gt @ var.gt_sample_points.float().permute(0, 2, 1) + T_gt).permute(0, 2, 1)   # [B,N,3] (tiny; host glue)
                var.gt_points_cam = ((cam - self.gt_mean.unsqueeze(1)) / self.gt_scale.view(-1, 1, 1)).contiguous()
                idx = torch.topk(var.gt_sample_sdf.abs(), k=min(100, var.gt_sample_sdf.shape[1]), dim=1, largest=False)[1]
                var.gt_surf_points = torch.gather(var.gt_points_cam, 1, idx.unsqueeze(-1).repeat(1, 1, 3))
        if self.coord_encoder.training and not full_train and isinstance(self.coord_encoder, CoordEncAtt):
            # train mode, depth estimator frozen (optim.fix_dpt): DropPath + hand-written backward (seen_coord_att_train.py)
            var.latent_depth = self.coord_encoder(coord, att_mask)
        elif self.coord_encoder.training and not full_train:
            # train mode: batch-statistics BatchNorm, differentiable w.r.t. the encoder parameters (seen_coord_enc_train.py)
            var.latent_depth = self.coord_encoder.forward_nhwc(coord)
        if "gt_sample_points" in var and "gt_sample_sdf" in var:
            # graph_shape.py:185 -- differentiable when the decoder is in train mode (implicit_train.ImplicitTrainFn)
            var.pred_sample_occ, _ = self.impl_network(var.latent_depth, None, var.gt_points_cam, need_attn=False)
        if get_loss:
            return var, self.compute_loss(opt, var, training)
        return var

    def compute_loss(self, opt, var, training=False):
        """graph_shape.py:194-202."""
        loss = edict()
        lw = opt.get("loss_weight", None) if isinstance(opt, dict) else getattr(opt, "loss_weight", None)
        if lw is None:
            return loss
        if lw.get("depth") is not None:
            loss.depth = self.loss_fns.depth_loss(var.depth_pred, var.depth_input_map, var.mask_input_map)
        if lw.get("intr") is not None and training:
            loss.intr = self.loss_fns.intr_loss(var.seen_points, var.seen_points_gt, var.validity_mask)
        if lw.get("shape") is not None and training:
            loss.shape = self.loss_fns.shape_loss(var.pred_sample_occ, var.gt_sample_sdf)
        return loss
