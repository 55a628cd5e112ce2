"""One bilateral-solver refinement (configs[4]: 8 noisy class maps over one grey reference, sigma 7/5/5) for `ncu`.
usage: python tools/bls_prof.py [size=512]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vittf_b200 import synth  # noqa: E402
from vittf_b200.bilateral_solver3d import solve_many  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
r8, lab = synth.ct_volume(size, n_shells=8, seed=0)
gen = torch.Generator().manual_seed(2)
t = torch.stack([((lab == c).float() * 0.8 + 0.2 * torch.rand(lab.shape, generator=gen)).clamp(0, 1) for c in range(8)]).cuda()
r8 = r8.cuda()
gp = dict(sigma_spatial=7, sigma_luma=5, sigma_chroma=5)
for _ in range(2):
    out, iters = solve_many(t, r8, None, gp)
torch.cuda.synchronize()
print("ok", iters.tolist())
