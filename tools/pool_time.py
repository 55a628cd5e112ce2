"""pool_axis at the cfg2 / cfg3 shapes: CUDA-event timings, L2 flushed."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vittf_b200 import ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for S, D in ((256, 384), (512, 768)):
    k = torch.randn(S, 4096, D, device="cuda").half()
    for ax in ("z", "y", "x"):
        t, _ = timeit(lambda: ops.pool_axis(k, 64, 64, ax, 64), flush=flush)
        print(f"S={S} D={D} axis {ax}: {t * 1e3:.0f} us, {k.numel() * 2 / t / 1e6:.0f} GB/s read", flush=True)
    del k
