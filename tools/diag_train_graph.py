"""Where does a CUDA-graph capture of the training iteration get invalidated?  Wraps every op of zeroshape_b200.ops (any thread)
and every torch function of the capturing thread with a capture-status probe (cuStreamIsCapturing) and reports the first call after
which the capture is no longer active.   python tools/diag_train_graph.py [--batch 2]"""
import argparse
import ctypes
import os
import sys
import traceback
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from torch.overrides import TorchFunctionMode

cuda_drv = ctypes.CDLL("libcuda.so.1")
STREAM = [None]
FOUND = [False]


def status():
    st = ctypes.c_int(-1)
    rc = cuda_drv.cuStreamIsCapturing(ctypes.c_void_p(STREAM[0]), ctypes.byref(st))
    return rc, st.value


def probe(what):
    if FOUND[0] or STREAM[0] is None:
        return
    rc, st = status()
    if rc != 0 or st != 1:
        FOUND[0] = True
        print(f"\n==== capture no longer active after `{what}` (rc {rc}, status {st}: 0 none, 1 active, 2 invalidated) ====", flush=True)
        traceback.print_stack(limit=14)


class Probe(TorchFunctionMode):
    def __torch_function__(self, func, types_, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        probe(getattr(func, "__name__", str(func)))
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--points", type=int, default=512)
    ap.add_argument("--precision", default="bf16")
    args = ap.parse_args()
    import bench
    from zeroshape_b200 import ops
    from zeroshape_b200.model.compute_graph.graph_shape import Graph
    from zeroshape_b200.model.shape.implicit_train import FusedAdamW
    from zeroshape_b200.utils.util import EasyDict
    dev = torch.device("cuda", 0)
    ops.TRAIN_ENGINE, ops.TRAIN_PRECISION = "tc", args.precision
    B, N = args.batch, args.points
    opt = bench.make_opt(dev, 128)
    opt.loss_weight = EasyDict(depth=None, intr=None, shape=1)
    opt.training = EasyDict(shape_loss=EasyDict(impt_thres=0.01, impt_weight=1))
    torch.manual_seed(0)
    graph = Graph(opt).to(dev).train()
    with torch.no_grad():
        graph.intr_proj.weight.normal_(0, 0.02)
    optim = FusedAdamW([p for p in graph.parameters() if p.requires_grad], lr=3e-5, betas=(0.9, 0.95), weight_decay=0.05, capturable=True)
    rgb, mask = bench.synthetic_images(B, 2000)
    g = torch.Generator().manual_seed(3)
    depth = (1.5 + 0.3 * torch.rand(B, 1, 224, 224, generator=g)) * mask
    intr = torch.tensor([[1.3875 * 224, 0, 112], [0, 1.3875 * 224, 112], [0, 0, 1.0]]).repeat(B, 1, 1)
    pose = torch.cat([torch.eye(3), torch.tensor([[0.0], [0.0], [1.6]])], dim=1).repeat(B, 1, 1)
    pts = torch.rand(B, N, 3, generator=g) - 0.5
    sdf = pts.norm(dim=-1) - 0.3 - 0.003
    inputs = [t.to(dev) for t in (rgb, mask, depth, intr, pose, pts, sdf)]

    def iteration(phase):
        var = EasyDict(idx=torch.arange(B), rgb_input_map=inputs[0], mask_input_map=inputs[1], depth_input_map=inputs[2], intr=inputs[3],
                       pose_gt=inputs[4], gt_sample_points=inputs[5], gt_sample_sdf=inputs[6])
        optim.zero_grad()
        var, loss = graph.forward(opt, var, training=True)
        probe("Graph.forward (end)")
        if phase >= 1:
            try:
                loss.shape.backward()
            except Exception:
                print("---- backward raised:", flush=True)
                traceback.print_exc()
                raise
            probe("backward (end)")
        if phase >= 2:
            optim.step()
            probe("optim.step (end)")
        return loss.shape

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            iteration(2)
    torch.cuda.synchronize()
    # wrap the ops
    for name, fn in list(vars(ops).items()):
        if isinstance(fn, types.FunctionType) and fn.__module__ == ops.__name__ and not name.startswith("_"):
            def wrapped(*a, __fn=fn, __name=name, **k):
                out = __fn(*a, **k)
                probe("ops." + __name)
                return out
            setattr(ops, name, wrapped)
    for phase, label in ((0, "forward"), (1, "forward + backward"), (2, "forward + backward + optimizer")):
        FOUND[0] = False
        gr = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(gr):
                STREAM[0] = torch.cuda.current_stream().cuda_stream
                with Probe():
                    iteration(phase)
            STREAM[0] = None
            torch.cuda.synchronize()
            gr.replay()
            torch.cuda.synchronize()
            print(f"capture of {label}: OK", flush=True)
        except Exception as e:                                   # noqa: BLE001
            STREAM[0] = None
            print(f"capture of {label}: FAILED", flush=True)
            traceback.print_exc()
            break


if __name__ == "__main__":
    main()
