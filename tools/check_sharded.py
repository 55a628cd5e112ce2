"""torchrun --nproc-per-node N tools/check_sharded.py : the sharded pipeline (slices over ranks, one all-reduce
per axis, z-slab similarity) must reproduce the single-GPU result bit-exactly on every rank."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vittf_b200 import pipeline, synth  # noqa: E402
from vittf_b200.dino import build_dino  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
vol, _ = synth.ct_volume((64, 64, 64), n_shells=4, seed=3)
ann = synth.annotations((64, 64, 64), 4, 3, seed=1)
model = build_dino("vits8", seed=0)
v = vol.to(dev)
f1, s1, l1, zr1 = pipeline.volume_to_similarity(v, model, ann, 8, 16, 4)                       # un-sharded
fN, sN, lN, zrN = pipeline.volume_to_similarity(v, model, ann, 8, 16, 4, rank=rank, world=world)
torch.cuda.synchronize()
ok = torch.equal(f1, fN) and torch.equal(s1[..., zrN[0]:zrN[1]], sN) and torch.equal(l1[..., zrN[0]:zrN[1]], lN)
# solver refinement: slab-local pixel passes + all-reduced grid vectors vs the one-pass solve (fp64 atomics: 1e-6)
ref8 = (vol.float() / vol.float().max() * 255).to(torch.uint8).to(dev) if vol.dtype != torch.uint8 else v
r1 = pipeline.refine_similarity(s1, ref8, (0, 64))
rN = pipeline.refine_similarity(sN, ref8, zrN)
torch.cuda.synchronize()
bls_err = (r1[..., zrN[0]:zrN[1]] - rN).abs().max().item()
ok = ok and bls_err < 1e-6
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED_OK" if flag.item() == 1 else "SHARDED_MISMATCH", "world", world, "z-range rank0", zrN, "bls err", bls_err)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
