"""CPU oracle for the vit-tf feature-volume hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``vittf_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and there only as the checker or as the timed CPU
baseline, never as the product path.

Every function restates (in plain PyTorch fp32 / NumPy fp64 on the CPU) the
algorithm of the reference at the cited ``file:line`` under ``/root/reference``.

Pinning status:
  * ``feature_volume``  – pinned against the reference's own ``infer.compute_qkv``
    + merge loop run in the build container (``oracle/make_golden.py`` ->
    ``tests/golden/feat_*.npz``).  The DINO ViT itself lives in an un-vendored
    third-party repo (facebookresearch/dino, unpinned ``main``); it is restated
    from its published architecture in ``dino_vit.py`` -> ViT arithmetic is
    "parity unpinned" beyond our own restatement.
  * ``similarity``      – pinned against ``predict_ntf.compute_similarities`` and
    ``infer.sample_features3d`` outputs (``tests/golden/sim_*.npz``).
  * ``bls``             – pinned against ``bilateral_solver3d.apply_bilateral_solver3d``
    outputs (``tests/golden/bls_*.npz``).
"""
