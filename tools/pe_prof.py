"""One patch-embed launch at the cfg2 shape for `ncu`."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vittf_b200 import ops  # noqa: E402

vol = torch.randint(0, 255, (256, 256, 256), dtype=torch.uint8, device="cuda")
mm = ops.minmax(vol)
pw = torch.randn(64, 384, device="cuda") * 0.1
pb = torch.randn(384, device="cuda") * 0.1
pos = torch.randn(4097, 384, device="cuda") * 0.02
for _ in range(2):
    out = ops.patch_embed(vol, "y", 0, 64, 512, 512, 8, mm, pw, pb, pos)
torch.cuda.synchronize()
print("ok")
