"""Generates tests/golden/entrypoints.npz: outputs of the REAL reference entry points (demo.marching_cubes, demo.py:143-153;
Runner.evaluate_batch, model/shape_engine.py:517-523; utils.eval_3D.eval_metrics) executed in the build container over this
package's mirrors with CPU stand-ins for the kernels (tests/fake_ops.py).  Run from the repo root:

    python tests/golden/make_golden_entrypoints.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


class _Patch:
    """The two pytest.MonkeyPatch methods the test helper uses."""
    def setattr(self, obj, name, value):
        setattr(obj, name, value)

    def setitem(self, d, k, v):
        d[k] = v

    def delitem(self, d, k):
        del d[k]


if __name__ == "__main__":
    from test_boundary_reference_entrypoints import run_reference_entrypoints
    out = run_reference_entrypoints(_Patch())
    np.savez_compressed(os.path.join(HERE, "entrypoints.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
