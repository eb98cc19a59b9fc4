"""CPU: the switches around the CUDA-graph paths (no capture happens here): the eval-mode encoder is only replayed for CUDA inputs
with the flag on, an active OpTimer turns the flag off and restores it, the capture cache lives outside the module."""
import copy

import torch


def test_encoder_graph_guards_and_optimer_flag():
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_graph import make_opt
    from zeroshape_b200 import ops
    from zeroshape_b200.model.compute_graph.graph_shape import Graph
    from zeroshape_b200.utils.util import EasyDict
    graph = Graph(make_opt(torch.device("cpu"))).eval()
    var = EasyDict(idx=torch.arange(1), rgb_input_map=torch.rand(1, 3, 224, 224), mask_input_map=torch.ones(1, 1, 224, 224))
    assert ops.ENCODER_CUDA_GRAPH is True
    assert graph._encoder_graph_ok(var, full_train=False) is False          # host tensors: never captured
    assert graph._encoder_graph_ok(var, full_train=True) is False
    with ops.OpTimer():
        assert ops.ENCODER_CUDA_GRAPH is False                               # the per-op timer needs the individual launches
    assert ops.ENCODER_CUDA_GRAPH is True
    assert graph._encoder_graphs == {} and "_encoder_graphs" not in graph.__dict__
    twin = copy.deepcopy(graph)
    assert twin._encoder_graphs == {} and twin._encoder_graphs is not graph._encoder_graphs
    sig = graph._encoder_signature()
    with torch.no_grad():
        graph.intr_proj.bias.add_(1.0)                                       # an in-place update bumps the version counter
    assert graph._encoder_signature() != sig
