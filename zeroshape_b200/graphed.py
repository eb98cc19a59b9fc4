"""A whole training iteration (model/shape_engine.py:180-199 `train_iteration`: zero_grad, Graph.forward, loss.backward,
optim.step) captured ONCE in a CUDA graph and replayed per step.

The step of options/shape.yaml is ~4,600 kernel launches at batch 32; issued one by one from Python the host is the
limiter for part of the step (the SM clock sits at its maximum with no power cap while the inference path, which keeps the
GPU full, runs into the cap).  Everything on the path is stream-ordered -- no host synchronisation in the forward tapes, the
hand-written backward passes or the optimizer (FusedAdamW(capturable=True): step-dependent scalars in device memory, tensor
table in the kernel parameters) -- so the iteration is one graph launch.

    optim = FusedAdamW(params, ..., capturable=True)
    def iteration(rgb, mask, ...):               # device tensors in, loss tensor out; exactly the reference's train_iteration
        optim.zero_grad()
        var, loss = graph.forward(opt, make_var(rgb, mask, ...), training=True)
        loss.all.backward()
        optim.step()
        return loss.all
    step = GraphedTrainStep(iteration, optim, example_inputs=(rgb, mask, ...))
    for batch in loader:
        loss = step(*batch)                      # batch: host (pinned) or device tensors of the captured shapes

Static shapes only (the data loader of the reference drops the last incomplete batch: `drop_last` in data/base.py).  The learning
rate of every parameter group is re-read before each replay, so torch.optim.lr_scheduler objects keep working.
"""
import torch


class GraphedTrainStep:
    def __init__(self, iteration, optim, example_inputs, warmup=3):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedTrainStep: no CUDA device (zeroshape_b200 has no CPU path)")
        if not getattr(optim, "capturable", False):
            raise ValueError("GraphedTrainStep: the optimizer must be FusedAdamW(..., capturable=True)")
        self.optim = optim
        self.inputs = [t.detach().clone().cuda() if not t.is_cuda else t.detach().clone() for t in example_inputs]
        # a few eager iterations on a side stream first (lazy initialisations, workspace pools, packed-weight caches,
        # optimizer state), as for any whole-network capture
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                iteration(*self.inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from ._native import lib
        self._lib = lib
        n0 = lib.zs_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):      # the stream of the warm-up: AccumulateGrad nodes keep theirs
            self.loss = iteration(*self.inputs)
        self.launches_per_replay = int(lib.zs_launch_count() - n0)     # launches of this library recorded in the graph
        self.replays = 0

    def __call__(self, *inputs):
        if len(inputs) != len(self.inputs):
            raise ValueError(f"GraphedTrainStep: expected {len(self.inputs)} inputs, got {len(inputs)}")
        for dst, src in zip(self.inputs, inputs):
            if src is not dst:
                if src.shape != dst.shape:
                    raise ValueError(f"GraphedTrainStep: captured shape {tuple(dst.shape)}, got {tuple(src.shape)}")
                dst.copy_(src, non_blocking=True)
        self.optim.prepare_replay()
        self.graph.replay()
        self._lib.zs_launch_count_add(self.launches_per_replay)
        self.optim.mark_updated()
        self.replays += 1
        return self.loss
