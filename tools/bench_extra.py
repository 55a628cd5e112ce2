"""Secondary benchmarks of BASELINE.json configs[3] (similarity sweep) and configs[4] (bilateral solver), one GPU.
Writes gpurun_out/bench_extra.json.  CUDA-event timing, 3 warm-ups, median of 7, inputs larger than L2 or L2
flushed between iterations."""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vittf_b200 import ops, pipeline, synth  # noqa: E402
from vittf_b200.bilateral_solver3d import solve_many  # noqa: E402
from vittf_b200.similarity import similarity_maps  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()) if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}


def timed(fn, iters=7, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


res = {"similarity_sweep": [], "bilateral_solver": []}
# ---- configs[3]: 384-d 128^3 fp16 feature volume -> 512^3, A = 1..64 prototypes, C = min(A, 8) classes
feats, protos_c = synth.class_features(384, (128, 128, 128), 8, seed=0)
feats = feats.to(dev)
g = torch.Generator().manual_seed(1)
for A in (1, 2, 4, 8, 16, 32, 64):
    C = min(A, 8)
    p = torch.nn.functional.normalize(protos_c.repeat((A + 7) // 8, 1)[:A] + 0.05 * torch.randn(A, 384, generator=g), dim=-1).to(dev)
    offs = torch.tensor([round(i * A / C) for i in range(C + 1)], dtype=torch.int32, device=dev)
    out = None

    def run():
        global out
        out = similarity_maps(feats, p, offs, (512, 512, 512), mode="ns")
    ms = timed(run)
    alg = feats.numel() * 2 + C * 512 ** 3 * 4 + A * 384 * 4
    res["similarity_sweep"].append({"A": A, "C": C, "ms": ms, "gvoxel_per_s": 512 ** 3 / ms / 1e6, "algorithmic_gb": alg / 1e9,
                                    "hbm_gbs": alg / ms / 1e6, "frac_of_measured_hbm": alg / ms / 1e6 / peaks["hbm_gbs"]})
    print(res["similarity_sweep"][-1], flush=True)
    del out
# ---- configs[4]: solver refinement of 8 class maps at 256^3 (sigma 7/5/5, Sobel confidence, defaults otherwise)
for size in (128, 256, 512):
    r8, lab = synth.ct_volume(size, n_shells=8, seed=0)
    # noisy class maps (like real similarity maps): a piecewise-constant target aligned with the reference is a fixed
    # point of the solver and would time zero PCG iterations
    gen = torch.Generator().manual_seed(2)
    t = torch.stack([((lab == c).float() * 0.8 + 0.2 * torch.rand(lab.shape, generator=gen)).clamp(0, 1) for c in range(8)]).to(dev)
    r8 = r8.to(dev)
    gp = dict(sigma_spatial=7, sigma_luma=5, sigma_chroma=5)
    ms = timed(lambda: solve_many(t, r8, None, gp), iters=5)
    out, iters = solve_many(t, r8, None, gp)
    res["bilateral_solver"].append({"size": size, "classes": 8, "ms_all_classes": ms, "ms_per_class": ms / 8,
                                    "pcg_iters": iters.tolist(), "pixel_bytes_per_class": size ** 3 * 9})
    print(res["bilateral_solver"][-1], flush=True)
# CPU port of the solver on this host (one class, bounded: 128^3)
from oracle import bls  # noqa: E402
r8c, labc = synth.ct_volume(128, n_shells=8, seed=0)
tc = ((labc == 1).float() * 0.8 + 0.2 * torch.rand(labc.shape, generator=torch.Generator().manual_seed(2))).clamp(0, 1)[None]
t0 = time.perf_counter()
bls.solve_sparse(tc, r8c.expand(3, -1, -1, -1), grid_params=dict(sigma_spatial=7, sigma_luma=5, sigma_chroma=5))
res["bilateral_solver_cpu_port"] = {"size": 128, "classes": 1, "ms": (time.perf_counter() - t0) * 1e3, "kind": "port (np.unique + CSR + scipy cg)"}
print(res["bilateral_solver_cpu_port"], flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/bench_extra.json").write_text(json.dumps(res, indent=1))
