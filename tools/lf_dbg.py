import sys, torch
sys.path.insert(0, '.')
from vittf_b200 import ops
feats = torch.randn(96, 16, 16, 16, device="cuda").half()
protos = torch.randn(8, 96, device="cuda")
d, g = ops.sim_lowres(feats, protos)
torch.cuda.synchronize()
ref = torch.einsum('fxyz,af->axyz', feats.float(), protos)
print("dots err", (d.view_as(ref) - ref).abs().max().item())
