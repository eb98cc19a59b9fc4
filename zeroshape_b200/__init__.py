"""zeroshape_b200 -- B200 (sm_100a) implementation of ZeroShape's per-image hot path.

Host side: Python/PyTorch modules that mirror the reference's module surface
(`model.compute_graph.graph_shape.Graph`, `model.shape.implicit.Implicit`, `utils.eval_3D`,
`external.chamfer3D.dist_chamfer_3D.chamfer_3DDist`); math: hand-written CUDA in
`libzeroshape_b200.so` behind the C ABI of `include/zeroshape_b200.h`.

`import zeroshape_b200` itself is light (so `python -m zeroshape_b200.build` can run before the
library exists); every functional sub-module imports `_native`, which raises if the CUDA library
is missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"


def install_as_reference_modules():
    """Register this package's modules under the reference's import names (`model.*`, `utils.*`,
    `external.*`) so the stock train.py / demo.py / evaluate.py import them unchanged.
    See INTEGRATION.md."""
    import importlib
    import sys
    mapping = {
        "model.shape.implicit": "zeroshape_b200.model.shape.implicit",
        "model.compute_graph.graph_shape": "zeroshape_b200.model.compute_graph.graph_shape",
        "model.compute_graph.graph_depth": "zeroshape_b200.model.compute_graph.graph_depth",
        "utils.eval_3D": "zeroshape_b200.utils.eval_3D",
        "utils.camera": "zeroshape_b200.utils.camera",
        "utils.loss": "zeroshape_b200.utils.loss",
        "utils.eval_depth": "zeroshape_b200.utils.eval_depth",
        "external.chamfer3D.dist_chamfer_3D": "zeroshape_b200.external.chamfer3D.dist_chamfer_3D",
    }
    for ref_name, ours in mapping.items():
        sys.modules[ref_name] = importlib.import_module(ours)
