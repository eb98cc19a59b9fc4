"""Device time of the whole inference encoder (Graph.forward, eval mode: image -> depth / intrinsics / seen surface / latents) at
batch 1 and 8, launched op by op and replayed from its CUDA graph.  ZS_GEMM_SPLITK=n sets the split-K granularity (chunks per split)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from zeroshape_b200 import ops
from zeroshape_b200.model.compute_graph.graph_shape import Graph
from zeroshape_b200.utils.util import EasyDict

dev = torch.device("cuda", 0)
opt = bench.make_opt(dev, 128)
torch.manual_seed(0)
graph = Graph(opt).to(dev).eval()
for B in (1, 8):
    rgb, mask = bench.synthetic_images(B, 100)
    rgb, mask = rgb.to(dev), mask.to(dev)
    for graphed in (False, True):
        ops.ENCODER_CUDA_GRAPH = graphed

        def run():
            var = EasyDict(idx=torch.arange(B), rgb_input_map=rgb, mask_input_map=mask, pose_gt=False)
            with torch.no_grad():
                return graph.forward(opt, var, training=False, get_loss=False)
        for _ in range(4):
            run()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(20):
            run()
        t1.record()
        torch.cuda.synchronize()
        print(f"splitk={os.environ.get('ZS_GEMM_SPLITK', 'default')} batch {B} {'graph ' if graphed else 'eager '}: {t0.elapsed_time(t1) / 20:7.3f} ms per forward", flush=True)
