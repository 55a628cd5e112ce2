"""Similarity passes: CUDA-event timings (L2 flushed) at the benchmark shapes.  VITTF_SIM_CELLS=1 selects the
previous per-cell up-sampling kernel for A/B."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os
from vittf_b200 import ops  # noqa: E402
VM = os.environ.get('VITTF_SIM_UP_TC') is not None      # voxel-major dots + the tcgen05 up-sampling kernel
from tools.microbench import timeit  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for F, n, N, A, C, zr in ((384, 64, 128, 32, 4, None), (768, 64, 512, 32, 16, None), (384, 64, 256, 32, 8, None), (384, 64, 256, 8, 8, None), (384, 64, 256, 32, 8, (96, 128)),
                          (768, 64, 512, 64, 16, None), (384, 128, 512, 64, 8, None)):
    feats = torch.nn.functional.normalize(torch.randn(F, n, n, n, device="cuda"), dim=0).half()
    protos = torch.nn.functional.normalize(torch.randn(A, F, device="cuda"), dim=-1)
    offs = torch.arange(0, A + 1, A // C, dtype=torch.int32, device="cuda")
    z0, z1 = zr if zr else (0, N)
    t1, _ = timeit(lambda: ops.sim_lowres(feats, protos, voxel_major=VM), flush=flush)
    low = ops.sim_lowres(feats, protos, voxel_major=VM)
    out = torch.empty(C, N, N, z1 - z0, device="cuda")
    t2, _ = timeit(lambda: ops.sim_upsample(low[0], low[1], (n, n, n), offs, (N, N, N), 0, 0.25, 2.0, z0, z1, out=out, layout=low[2], n_protos=A), flush=flush)
    alg = feats.numel() * 2 + out.numel() * 4 + A * F * 4
    print(f"F={F} {n}^3->{N}^3 A={A} C={C} z=[{z0},{z1}): lowres {t1 * 1e3:.0f} us, upsample {t2 * 1e3:.0f} us, "
          f"{out.numel() / C / ((t1 + t2) * 1e-3) / 1e9:.1f} Gvox/s, {alg / (t1 + t2) / 1e6:.0f} GB/s algorithmic "
          f"({alg / (t1 + t2) / 1e6 / 6542.1:.3f} of HBM peak)", flush=True)
    del feats, out, low
