"""GPU: `DistributedDataParallel(Graph)` (model/shape_engine.py:71) over the custom autograd tapes.

Two ranks (gloo, both on cuda:0 -- NCCL needs one device per rank; the multi-GPU run of the same step is
`bench.py --mode train-ddp` under torchrun) each take one image of a 2-image batch, run Graph.forward(training=True),
backward; DDP's bucketed all-reduce averages the gradients.  They must equal the mean of the two single-image gradients
computed without DDP in this process."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_ddp_training_step_matches_single_rank_gradient_mean(cuda, tmp_path):
    import ddp_worker as W
    from zeroshape_b200 import ops
    out = str(tmp_path / "ddp_rank0.pt")
    mp.spawn(W.worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    saved = ops.TRAIN_ENGINE
    ops.TRAIN_ENGINE = "f32"
    try:
        graph = W.build_graph(cuda)
        opt = W.make_opt(cuda)
        case = W.make_case(2, 7, cuda)
        ref, losses = {}, []
        for r in range(2):
            graph.zero_grad(set_to_none=True)
            var, loss = graph(opt, case(r, r + 1), training=True, get_loss=True)
            loss.shape.backward()
            losses.append(float(loss.shape))
            for n, p in graph.named_parameters():
                if p.grad is not None:
                    ref[n] = ref.get(n, 0) + p.grad.detach().cpu() / 2
    finally:
        ops.TRAIN_ENGINE = saved
    assert abs(got["loss"] - losses[0]) < 1e-5 * max(1.0, abs(losses[0]))
    assert set(got["grads"]) == set(ref) and len(ref) > 300
    worst = 0.0
    for n, g in ref.items():
        d = (got["grads"][n].double() - g.double()).norm() / g.double().norm().clamp_min(1e-20)
        worst = max(worst, float(d))
        assert d < 2e-4, (n, float(d))        # atomics in the weight-gradient reductions reorder fp32 sums run to run
    print(f"DDP(Graph) 2-rank step: {len(ref)} parameter gradients equal the single-rank mean (worst relative {worst:.2e})")
