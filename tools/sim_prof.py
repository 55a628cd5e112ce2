"""One similarity pass at the cfg2 shape (384 x 64^3 fp16 -> 256^3, A=32, C=8) for `ncu`."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os
from vittf_b200 import ops  # noqa: E402
VM = os.environ.get('VITTF_SIM_UP_TC') is not None      # voxel-major dots + the tcgen05 up-sampling kernel

F, n, N, A, C = 384, 64, 256, 32, 8
feats = torch.nn.functional.normalize(torch.randn(F, n, n, n, device="cuda"), dim=0).half()
protos = torch.nn.functional.normalize(torch.randn(A, F, device="cuda"), dim=-1)
offs = torch.arange(0, A + 1, A // C, dtype=torch.int32, device="cuda")
out = torch.empty(C, N, N, N, device="cuda")
for _ in range(2):
    low = ops.sim_lowres(feats, protos, voxel_major=VM)
    ops.sim_upsample(low[0], low[1], (n, n, n), offs, (N, N, N), 0, 0.25, 2.0, 0, N, out=out, layout=low[2], n_protos=A)
torch.cuda.synchronize()
print("ok")
