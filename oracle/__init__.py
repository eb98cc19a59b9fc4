"""CPU oracle for the ZeroShape hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Everything under ``oracle/`` is a plain fp32 (PyTorch-CPU / numpy / C) restatement of the
reference algorithm for the path named in BASELINE.json, written as *functions over a
state_dict* so that the same weights can be fed to the oracle and to the CUDA product.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package.  The product (``zeroshape_b200``) never does;
it fails loudly if its CUDA library is missing.

Parity status (see DESIGN.md "Oracle pinning"):
  * implicit decoder, CoordEncRes, Bottleneck_Conv, DPT scratch/fusion/head, camera glue,
    interpolate_coordmap, losses: pinned against the real reference modules imported from
    /root/reference (through import shims) by ``tests/golden/make_golden.py``; the resulting
    input/output vectors are committed under ``tests/golden/``.
  * timm hybrid ViT backbone (third-party, timm==0.6.12, not in /root/reference): restated from
    the published architecture, cross-checked against the independent ``transformers`` DPT-hybrid
    implementation; the reference repo itself holds no golden vector for it -> "parity unpinned"
    by the reference, pinned only by that cross-check.
  * marching cubes / surface sampling (PyMCubes 0.1.4 / trimesh 4.0.8, absent): restated from the
    published algorithm; known-answer tests on analytic fields; "parity unpinned" by the reference.
  * chamfer: C restatement of external/chamfer3D/chamfer3D.cu, checked against scipy cKDTree and,
    on the GPU box, against the reference .cu itself compiled into oracle/_ref/.
"""
