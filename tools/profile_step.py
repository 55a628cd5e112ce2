"""One ViT forward over a batch of slices (the unit bench.py repeats 96x3 times per volume) + one similarity
pass: short enough for `ncu`.  Usage: python tools/profile_step.py [batch] [arch]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vittf_b200 import ops, pipeline, synth  # noqa: E402
from vittf_b200.dino import build_dino  # noqa: E402
from vittf_b200.vit import engine_for  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
arch = sys.argv[2] if len(sys.argv) > 2 else "vits8"
dev = torch.device("cuda", 0)
vol, _ = synth.ct_volume((256, 256, max(16, batch)), n_shells=8, seed=0)
v = vol.to(dev)
model = build_dino(arch, seed=0)
eng = engine_for(model, dev, max_batch=batch)
mm = ops.minmax(v)
for _ in range(2):
    k = eng.k_features(v, "z", 0, batch, 512, 512, mm)
torch.cuda.synchronize()
feats, _ = synth.class_features(384, (64, 64, 64), 8, seed=0)
feats = feats.to(dev)
ann = synth.annotations(256, 8, 4, seed=0)
protos = pipeline.prototypes(feats, ann, (256, 256, 256))
from vittf_b200.similarity import class_offsets, similarity_maps  # noqa: E402
s = similarity_maps(feats, protos, class_offsets(ann, dev), (256, 256, 256), mode="ns")
torch.cuda.synchronize()
print("ok", k.shape, s.shape)
