"""Patch embedding at the benchmark shapes (64 slices of a 256^3 / 512^3 uint8 volume -> 512^2 images): CUDA-event timings."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vittf_b200 import ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n, D in ((256, 384), (512, 768)):
    vol = torch.randint(0, 255, (n, n, n), dtype=torch.uint8, device="cuda")
    mm = ops.minmax(vol)
    pw = torch.randn(64, D, device="cuda") * 0.1
    pb = torch.randn(D, device="cuda") * 0.1
    pos = torch.randn(4097, D, device="cuda") * 0.02
    for ax in ("z", "y", "x"):
        t, _ = timeit(lambda: ops.patch_embed(vol, ax, 0, 64, 512, 512, 8, mm, pw, pb, pos), flush=flush)
        print(f"{n}^3 D={D} axis {ax}: {t * 1e3:.0f} us per 64 slices ({64 * 4097 * D * 4 / t / 1e6:.0f} GB/s written)", flush=True)
