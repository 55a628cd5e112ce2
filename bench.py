#!/usr/bin/env python
"""Benchmark of the vit-tf feature-volume hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg3|cfg2|cfg1|tiny|cfg4|cfg5]

Default workload = BASELINE.json's metric config, configs[2]: 512^3 volume, ViT-B/8, 16 classes (fits one GPU).
One "step" = one whole volume: raw voxels -> 3-axis ViT K-feature volume (merged, fp16) -> prototype
similarity (north-star order) for C classes -> argmax label volume.  Prints ONE JSON line (rank 0):
  value  ms per volume, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e    the same through the public API from a pinned HOST volume: H2D of the volume, D2H of the fp16
         feature volume (what infer.py saves), of the uint8 half-resolution per-class maps (what
         compute_similarities returns, predict_ntf.py:95-100) and of the uint8 label volume inside the timed region
`--workload cfg4` = configs[3], the similarity-only sweep (384-d 128^3 features -> 512^3, 1..64 prototypes; value =
Gvoxel/s at 64 prototypes); `--workload cfg5` = configs[4], bilateral-solver refinement of 8 class maps at 512^3 (256^3
reported beside it), z-slab sharded.
  roofline      the dominant kernel (flash attention, tensor-bound), timed live with CUDA events on the
                launching stream inside the engine (vittf_vit_timing_*)
  cpu_baseline  the oracle port of the reference's CPU path timed on this host on a bounded sample
`--impl reference` times that CPU port alone (rank 0 only).  N > 1: one process per GPU under torchrun;
the slices of each axis are sharded over the ranks (strong scaling of ONE volume), one all-reduce per axis.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (volume size, arch, fos, classes, annotations per class, batch)
    "tiny": (32, "vits8", 8, 4, 2, 8),
    "cfg1": (128, "vits8", 64, 4, 8, 64),
    "cfg2": (256, "vits8", 64, 8, 4, 64),
    "cfg3": (512, "vitb8", 64, 16, 2, 64),
}
METRIC = "ms per 512^3 volume end-to-end (ViT feats + similarity)"      # BASELINE.json:metric; other workloads name their volume in config
WORKLOAD_TEXT = {
    "tiny": "smoke: 32^3 uint8 phantom, ViT-S/8 random init, 64^2 images, 4 classes",
    "cfg1": "configs[0]: 128^3 phantom, ViT-S/8 random init, 3-axis 512^2 slices (384 images), similarity for 4 classes",
    "cfg2": "configs[1]: 256^3 CT-shaped uint8 volume, ViT-S/8 bf16 (384-d), 3-axis 512^2 slices (768 images), "
            "NS similarity for 8 classes at 256^3",
    "cfg3": "configs[2]: 512^3 uint8 volume, ViT-B/8 bf16 (768-d), 3-axis 512^2 slices (1536 images), NS similarity "
            "for 16 classes at 512^3",
}
ARCH = {"vits8": (384, 12, 6, 8), "vitb8": (768, 12, 12, 8)}


def needed_flops_per_image(arch, tokens):
    """(L-1) full blocks + the K projection of the last block."""
    d, l, _, p = ARCH[arch]
    patch = 2 * (tokens - 1) * (3 * p * p) * d
    block = 24 * tokens * d * d + 4 * tokens * tokens * d
    return patch + (l - 1) * block + 2 * tokens * d * d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower() == "active"})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


def build_inputs(workload):
    from vittf_b200 import synth
    from vittf_b200.dino import build_dino
    size, arch, fos, n_cls, per_cls, batch = WORKLOADS[workload]
    vol, _ = synth.ct_volume(size, n_shells=n_cls, seed=0)            # uint8 CT-shaped phantom
    ann = synth.annotations(size, n_cls, per_cls, seed=0)
    model = build_dino(arch, seed=0)
    return vol, ann, model, dict(size=size, arch=arch, fos=fos, classes=n_cls, per_class=per_cls, batch=batch)


# --------------------------------------------------------------------------------------------- CPU port
_CPU_CACHE = {}


def cpu_reference(workload, budget_s=20.0, step=0):
    """The reference's CPU path (oracle port, fp32, all host threads) on a bounded sample of the
    workload, extrapolated linearly to ms per volume (slices are independent units).  Inputs and the
    model are built once per process; `step` rotates the slicing axis of the sampled image."""
    from oracle import dino_vit, feature_volume as ofv, similarity as osim
    from vittf_b200 import synth
    size, arch, fos, n_cls, per_cls, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if workload not in _CPU_CACHE:
        vol, _ = synth.ct_volume(size, n_shells=n_cls, seed=0)
        ann = synth.annotations(size, n_cls, per_cls, seed=0)
        model = dino_vit.build(arch, seed=0)
        im_sz, f_sz = ofv.image_sizes(tuple(vol.shape), 8, fos)
        feats, _ = synth.class_features(ARCH[arch][0], f_sz, n_cls, seed=0, dtype=torch.float16)
        _CPU_CACHE[workload] = (vol, ann, model, im_sz, f_sz, feats)
    vol, ann, model, im_sz, f_sz, feats = _CPU_CACHE[workload]
    # (i) ViT: time whole slice images, batch 1 like the reference default (infer.py:302), until 60 % of the budget is spent
    t_img, n_img = 0.0, 0
    for k in range(3):
        ax = ("z", "y", "x")[(step + k) % 3]
        imgs = ofv.slice_images(vol, ax)[size // 2:size // 2 + 1]
        r, c = ofv.AXIS_IMAGE_DIMS[ax]
        x = torch.nn.functional.interpolate(imgs, size=(im_sz[r], im_sz[c]), mode="nearest")
        t0 = time.perf_counter()
        ofv.hooked_qkv(model, x)
        t_img += time.perf_counter() - t0
        n_img += 1
        if t_img > budget_s * 0.6:
            break
    vit_ms = t_img / n_img * 3 * size * 1e3
    # (ii) similarity, NS composition, a few output z-planes
    pts = torch.cat(list(ann.values()))
    protos = osim.sample_prototypes(feats.float(), osim.rel_coords(pts, vol.shape), "bilinear")
    offs = [0]
    for v in ann.values():
        offs.append(offs[-1] + v.size(0))
    slab = 4 if budget_s >= 10 and size <= 256 else 1
    sim_key = (workload, "sim")
    if step < 2 or sim_key not in _CPU_CACHE:          # < 2 % of the total: sampled on the first two steps, then reused
        t0 = time.perf_counter()
        osim.ns_composite(feats, protos, offs, (size, size, size), slab=slab, z_range=(size // 2, size // 2 + slab))
        _CPU_CACHE[sim_key] = (time.perf_counter() - t0) * (size / slab) * 1e3
    sim_ms = _CPU_CACHE[sim_key]
    return {"value": vit_ms + sim_ms, "unit": "ms", "cores": cores, "kind": "port",
            "sample": f"ViT: {n_img} slice image(s) of {im_sz[0]}x{im_sz[1]} (batch 1) extrapolated x{3 * size}/{n_img}; "
                      f"similarity: {slab} of {size} output z-planes extrapolated; fp32, {cores} threads",
            "vit_ms": vit_ms, "similarity_ms": sim_ms}


# ------------------------------------------------------------------------- configs[3] / configs[4] workloads
SWEEP_A = (1, 2, 4, 8, 16, 32, 64)
CFG4_TEXT = ("configs[3]: similarity-only sweep, 384-d 128^3 fp16 feature volume up-sampled to 512^3 on the fly (NS order), "
             "1..64 prototypes in min(A, 8) classes, fp32 maps; headline = 64 prototypes / 8 classes")
CFG5_TEXT = ("configs[4]: bilateral_solver3d refinement (sigma 7/5/5, Sobel confidence, lam 256, 25 PCG iterations) of 8 noisy "
             "class maps over one grey uint8 reference at 512^3 (256^3 reported beside it), z-slab sharded")


def _peaks():
    pk = ROOT / "MEASURED_PEAKS.json"
    return json.loads(pk.read_text()) if pk.exists() else {}


def _sweep_inputs(A, protos_c, f_dim=384):
    g = torch.Generator().manual_seed(1 + A)
    C = min(A, 8)
    p = torch.nn.functional.normalize(protos_c.repeat((A + 7) // 8, 1)[:A] + 0.05 * torch.randn(A, f_dim, generator=g), dim=-1)
    offs = [round(i * A / C) for i in range(C + 1)]
    return p, offs, C


def cpu_reference_sweep(lr=128, out=512, A=64, slab=2):
    """configs[3] on the host: the NS composition of the reference's torch ops (oracle port) on `slab` output z-planes."""
    from oracle import similarity as osim
    from vittf_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    feats, protos_c = synth.class_features(384, (lr,) * 3, 8, seed=0)
    p, offs, C = _sweep_inputs(A, protos_c)
    t0 = time.perf_counter()
    osim.ns_composite(feats, p, offs, (out,) * 3, slab=slab, z_range=(out // 2, out // 2 + slab))
    ms = (time.perf_counter() - t0) * (out / slab) * 1e3
    return {"value": out ** 3 / ms / 1e6, "unit": "Gvoxel/s", "cores": cores, "kind": "port", "ms": ms,
            "sample": f"{slab} of {out} output z-planes of the {lr}^3 -> {out}^3 map set (A={A}, C={C}) extrapolated; fp32, {cores} threads"}


def _solver_inputs(size, n_cls=8):
    from vittf_b200 import synth
    r8, lab = synth.ct_volume(size, n_shells=n_cls, seed=0)
    gen = torch.Generator().manual_seed(2)
    # noisy class maps (like real similarity maps): a piecewise-constant target aligned with the reference is a fixed
    # point of the solver and would time zero PCG iterations
    t = torch.stack([((lab == c).float() * 0.8 + 0.2 * torch.rand(lab.shape, generator=gen)).clamp(0, 1) for c in range(n_cls)])
    return r8, t


def cpu_reference_solver(size=512, n_cls=8, sample=128):
    """configs[4] on the host: the sparse-matrix port of bilateral_solver3d (np.unique + CSR + scipy cg) for ONE class at
    `sample`^3, extrapolated linearly in voxels and classes."""
    from oracle import bls
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    r8, t = _solver_inputs(sample, n_cls)
    t0 = time.perf_counter()
    bls.solve_sparse(t[1:2], r8.expand(3, -1, -1, -1), grid_params=dict(sigma_spatial=7, sigma_luma=5, sigma_chroma=5))
    one = (time.perf_counter() - t0) * 1e3
    ms = one * n_cls * (size / sample) ** 3
    return {"value": ms, "unit": "ms", "cores": 1, "kind": "port", "one_class_ms_at_sample": one,
            "sample": f"one class at {sample}^3 (np.unique + CSR + scipy cg, single-threaded numpy/scipy) extrapolated x{n_cls} classes "
                      f"x{(size // sample) ** 3} voxels"}


def _timed(fn, steps, dev, world, dist):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        fn()
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def bench_similarity_sweep(args, rank, world, dev, dist):
    from vittf_b200 import _lib, ops, pipeline, synth
    from vittf_b200 import dist as vdist
    from vittf_b200.similarity import similarity_maps
    lr, out = 128, 512
    feats_h, protos_c = synth.class_features(384, (lr,) * 3, 8, seed=0)
    feats_h = feats_h.pin_memory()
    feats = feats_h.to(dev)
    xr = vdist.x_range(out, world, rank)          # x-slabs: pass 1 (dots + Gram) shards with the maps, no exchange
    peaks = _peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    warm = max(3, args.warmup)
    lib = _lib.load()
    sweep, head = [], None
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    for A in SWEEP_A:
        p_h, offs_l, C = _sweep_inputs(A, protos_c)
        p = p_h.to(dev)
        offs = torch.tensor(offs_l, dtype=torch.int32, device=dev)

        def run():
            return similarity_maps(feats, p, offs, (out,) * 3, mode="ns", x_range=xr)
        for _ in range(warm):
            run()
        is_head = A == SWEEP_A[-1]
        if is_head and rank == 0:
            sampler.start()
        lib.vittf_launch_count_reset()
        ms = _timed(run, args.steps if is_head else 3, dev, world, dist)
        launches = lib.vittf_launch_count()
        # features read once, maps written once, prototypes read once (whole job over all ranks)
        alg = feats.numel() * 2 + C * out ** 3 * 4 + A * 384 * 4
        row = {"A": A, "C": C, "ms": ms, "gvoxel_per_s": out ** 3 / ms / 1e6, "algorithmic_gb": alg / 1e9,
               "hbm_gbs_per_gpu": alg / world / ms / 1e6, "frac_of_measured_hbm": alg / world / ms / 1e6 / hbm}
        sweep.append(row)
        if is_head:
            maps_host = None

            def e2e():
                nonlocal maps_host
                f = feats_h.to(dev, non_blocking=True)                       # H2D of the cached feature volume
                sims = similarity_maps(f, p, offs, (out,) * 3, mode="ns", x_range=xr)
                q, _ = pipeline.quantized_maps(sims, (0, out), out)          # (slab bounds are even: half-resolution rows stay local)
                if maps_host is None:
                    maps_host = torch.empty(q.shape, dtype=torch.uint8).pin_memory()
                maps_host.copy_(q, non_blocking=True)                        # what compute_similarities returns
            e2e()
            ms_e2e = _timed(e2e, args.steps, dev, world, dist)
            head = dict(row, launches=int(launches), ms_e2e=ms_e2e, d2h=maps_host.numel())
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return
    outj = {"metric": "similarity Gvoxel/s (512^3 output voxels per second, 64 prototypes / 8 classes)", "value": head["gvoxel_per_s"],
            "unit": "Gvoxel/s", "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": head["ms"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16 features, f32 accumulate / maps",
            "data": "synthetic",
            "config": {"workload": CFG4_TEXT, "parallelism": f"output x-slabs over {world} GPU(s): each rank evaluates the low-res planes under its "
                       "slab +- 1 (no exchange), feature volume replicated",
                       "l2": "1.61 GB of features and 4.3 GB of maps per pass exceed the 126 MB L2"},
            "e2e": {"value": out ** 3 / head["ms_e2e"] / 1e6, "unit": "Gvoxel/s", "h2d_bytes_per_step": feats_h.numel() * 2,
                    "d2h_bytes_per_step": head["d2h"], "ms": head["ms_e2e"]},
            "gpu_launches": head["launches"], "clocks": clocks,
            "roofline": {"kernel": "similarity stage = sim_lowres (dots + Gram) + sim_upsample (tcgen05 cell tiles)", "bound": "hbm",
                         "achieved": head["hbm_gbs_per_gpu"], "peak": hbm, "unit": "GB/s", "frac": head["frac_of_measured_hbm"],
                         "traffic": None, "algorithmic_bytes_per_launch": head["algorithmic_gb"] * 1e9 / world,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"},
            "sweep": sweep}
    if not args.no_cpu_baseline:
        try:
            outj["cpu_baseline"] = cpu_reference_sweep()
        except Exception as e:
            outj["cpu_baseline"] = {"error": repr(e)}
    _emit(outj)


def bench_solver(args, rank, world, dev, dist):
    from vittf_b200 import _lib, pipeline
    from vittf_b200 import dist as vdist
    peaks = _peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    warm = max(3, args.warmup)
    lib = _lib.load()
    res = {}
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    for size in (256, 512):
        r8_h, t_h = _solver_inputs(size)
        zr = vdist.z_range(size, world, rank)
        t_slab_h = t_h[..., zr[0]:zr[1]].contiguous().pin_memory()
        r8_h = r8_h.contiguous().pin_memory()
        del t_h
        t_slab, r8 = t_slab_h.to(dev), r8_h.to(dev)

        def run():
            return pipeline.refine_similarity(t_slab, r8, zr)
        for _ in range(warm):
            run()
        is_head = size == 512
        if is_head and rank == 0:
            sampler.start()
        lib.vittf_launch_count_reset()
        ms = _timed(run, args.steps, dev, world, dist)
        launches = lib.vittf_launch_count()
        out_host = torch.empty(t_slab.shape, dtype=torch.float32).pin_memory()

        def e2e():
            r = r8_h.to(dev, non_blocking=True)
            t = t_slab_h.to(dev, non_blocking=True)
            out_host.copy_(pipeline.refine_similarity(t, r, zr), non_blocking=True)
        e2e()
        ms_e2e = _timed(e2e, args.steps, dev, world, dist)
        npix = size ** 3
        alg = 8 * npix * (4 + 4) + 2 * npix      # targets read, results written, reference read by Sobel/splat and by slice
        res[size] = {"ms": ms, "ms_per_class": ms / 8, "ms_e2e": ms_e2e, "launches": int(launches) // max(1, args.steps),
                     "pixel_pass_bytes": alg, "pixel_pass_gbs_per_gpu": alg / world / ms / 1e6,
                     "h2d": t_slab_h.numel() * 4 + r8_h.numel(), "d2h": out_host.numel() * 4}
        del t_slab, r8, t_slab_h, out_host
        torch.cuda.empty_cache()
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return
    h = res[512]
    outj = {"metric": "ms per bilateral_solver3d refinement of 8 class maps at 512^3", "value": h["ms"], "unit": "ms",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": h["ms"], "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64 grid, f32 maps", "data": "synthetic",
            "config": {"workload": CFG5_TEXT, "parallelism": f"pixel passes on z-slabs over {world} GPU(s), grid vectors all-reduced, grid "
                       "problem replicated", "l2": "4.3 GB of maps per pass exceed the 126 MB L2; 21 M-cell fp64 grid vectors (168 MB each) do too"},
            "e2e": {"value": h["ms_e2e"], "unit": "ms", "h2d_bytes_per_step": h["h2d"], "d2h_bytes_per_step": h["d2h"]},
            "gpu_launches": h["launches"] * args.steps, "clocks": clocks,
            "roofline": {"kernel": "solver pixel passes (Sobel + splat + slice); the PCG on the grid is the rest of the time", "bound": "hbm",
                         "achieved": h["pixel_pass_gbs_per_gpu"], "peak": hbm, "unit": "GB/s", "frac": h["pixel_pass_gbs_per_gpu"] / hbm,
                         "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"},
            "sizes": {str(k): v for k, v in res.items()}}
    if not args.no_cpu_baseline:
        try:
            outj["cpu_baseline"] = cpu_reference_solver()
        except Exception as e:
            outj["cpu_baseline"] = {"error": repr(e)}
    _emit(outj)


# --------------------------------------------------------------------------------------------- main
_JSON_FD = None


def _emit(obj):
    """The ONE JSON line, written to the process's original stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    # NCCL (and anything else native) writes banners straight to file descriptor 1: keep the original stdout for the
    # JSON line only and send every other write to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=list(WORKLOADS) + ["cfg4", "cfg5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=0, help="slices per ViT forward (default: per workload)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    extra = args.workload in ("cfg4", "cfg5")
    if not extra:
        size, arch, fos, n_cls, per_cls, batch = WORKLOADS[args.workload]
        if args.batch > 0:
            batch = args.batch
        config = {"workload": WORKLOAD_TEXT[args.workload], "volume": f"{size}^3 uint8", "backbone": arch,
                  "feature_output_size": fos, "classes": n_cls, "prototypes": n_cls * per_cls, "slice_batch": batch,
                  "parallelism": (f"slices sharded over {world} GPU(s), per-axis slab all-gather, z-slab similarity" if world > 1
                                  else "single GPU"),
                  "maps": "fp32 per-class maps stay in HBM; e2e brings back the fp16 feature volume, the uint8 half-resolution maps "
                          "(predict_ntf.py:95-100) and the uint8 label volume",
                  "l2": f"per-step working set (K-feature staging {2 * 64 * 4096 * ARCH[arch][0] * size // fos / 1e9:.1f} GB/axis, "
                        f"{n_cls * size ** 3 * 4 / 1e9:.2f} GB of maps) exceeds the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, args.steps)
        n_runs = max(0, min(args.warmup, 1)) + steps
        if extra:
            fn = cpu_reference_sweep if args.workload == "cfg4" else cpu_reference_solver
            vals = [fn() for _ in range(min(n_runs, 3))][-min(steps, 2):]
            v = sum(x["value"] for x in vals) / len(vals)
            cb = dict(vals[-1])
            cb["value"] = v
            hib = args.workload == "cfg4"
            _emit({"impl": "reference", "metric": ("similarity Gvoxel/s (512^3 output voxels per second, 64 prototypes / 8 classes)" if hib
                                                   else "ms per bilateral_solver3d refinement of 8 class maps at 512^3"),
                   "value": v, "unit": cb["unit"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                   "ms_per_step": cb.get("ms", v), "higher_is_better": hib, "scaling": "strong", "vs_baseline": None,
                   "dtype": "f32" if hib else "f64", "data": "synthetic",
                   "config": {"workload": CFG4_TEXT if hib else CFG5_TEXT}, "cpu_baseline": cb,
                   "e2e": {"value": v, "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
            return 0
        vals = [cpu_reference(args.workload, budget_s=120.0 / n_runs, step=i) for i in range(n_runs)][-steps:]
        v = sum(x["value"] for x in vals) / len(vals)
        cb = dict(vals[-1])
        cb["value"] = v
        _emit(({"impl": "reference", "metric": METRIC, "value": v,
                          "unit": "ms", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": v,
                          "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    if not torch.cuda.is_available():
        _emit({"error": "no CUDA device: vittf_b200 has no CPU fallback"})
        return 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if extra:
        (bench_similarity_sweep if args.workload == "cfg4" else bench_solver)(args, rank, world, dev, dist)
        if world > 1:
            dist.destroy_process_group()
        return 0
    from vittf_b200 import _lib, ops, pipeline
    from vittf_b200.vit import engine_for
    from vittf_b200.infer import image_sizes, _max_tokens

    vol, ann, model, _ = build_inputs(args.workload)
    vol_host = vol.contiguous().pin_memory()
    vol_dev = vol_host.to(dev)
    im_sz, f_sz = image_sizes(tuple(vol.shape), 8, fos)
    engine = engine_for(model, dev, max_batch=batch, max_tokens=_max_tokens(im_sz, 8))
    tokens = 1 + (im_sz[0] // 8) * (im_sz[1] // 8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return pipeline.volume_to_similarity(vol_dev, model, ann, 8, fos, batch, rank=rank, world=world)

    n_feat = ARCH[arch][0] * f_sz[0] * f_sz[1] * f_sz[2]
    feat_share = n_feat // world if n_feat % world == 0 else n_feat               # the feature volume is replicated: every rank
    feat_host = torch.empty(feat_share, dtype=torch.float16).pin_memory()          # brings back 1/world of it over its own link
    lab_host = maps_host = None

    def step_e2e():
        nonlocal lab_host, maps_host
        v = pipeline.distribute_volume(vol_host, dev, rank, world)                # H2D of the raw volume (1/world per rank + all-gather)
        feats, sims, labels, zr = pipeline.volume_to_similarity(v, model, ann, 8, fos, batch, rank=rank, world=world)
        q, _ = pipeline.quantized_maps(sims, zr, size)                            # uint8 half-resolution maps of this rank's slab
        if lab_host is None:
            lab_host = torch.empty(labels.shape, dtype=torch.uint8).pin_memory()
            maps_host = torch.empty(q.shape, dtype=torch.uint8).pin_memory()
        if feat_share != n_feat:
            feat_host.copy_(feats.view(-1)[rank * feat_share:(rank + 1) * feat_share], non_blocking=True)   # what infer.py saves
        elif rank == 0:
            feat_host.copy_(feats.view(-1), non_blocking=True)
        maps_host.copy_(q, non_blocking=True)                                    # what compute_similarities returns
        lab_host.copy_(labels, non_blocking=True)                                # the step's result
        return labels

    def timed(fn, steps):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            fn()
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(max(3, args.warmup)):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    engine.timing(True)
    engine.read_timing()
    load = _lib.load()
    load.vittf_launch_count_reset()
    ms_dev = timed(step_device, args.steps)
    launches = load.vittf_launch_count()
    timing = engine.read_timing()
    engine.timing(False)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # similarity stage alone (second half of BASELINE.json's metric): Gvoxel/s of output voxels
    feats, sims, labels, zr = step_device()
    protos = pipeline.prototypes(feats, ann, tuple(vol.shape))
    from vittf_b200.similarity import class_offsets, similarity_maps
    offs = class_offsets(ann, dev)

    def sim_only():
        return similarity_maps(feats, protos, offs, tuple(vol.shape), mode="ns", z_range=zr)
    for _ in range(2):                       # the first call allocates the (up to 8.6 GB) output: keep cudaMalloc out of the timing
        sim_only()
    torch.cuda.synchronize()
    ms_sim = timed(sim_only, max(3, args.steps))
    sim_bytes = feats.numel() * 2 + n_cls * size * size * (zr[1] - zr[0]) * 4 + protos.numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained"
    att_ms, att_n = timing["attention"]
    d, l, h, _ = ARCH[arch]
    # QK^T + PV of every (image, layer) this rank ran in the timed steps, over the summed duration of its attention launches
    # (a launch holds `batch` images only when the rank's slices per axis are a multiple of it)
    from vittf_b200 import dist as vdist
    local_imgs = 0
    for ax_len, n_out in zip(vol.shape, f_sz):
        a, b = vdist.slices_for_slabs(int(ax_len), int(n_out), *vdist.slab_range(int(n_out), world, rank))
        local_imgs += b - a
    att_flops_timed = 4.0 * h * tokens * tokens * 64 * local_imgs * (l - 1) * args.steps
    att_avg_ms = att_ms / max(1, att_n)
    achieved = att_flops_timed / (att_ms * 1e-3) / 1e12 if att_n else None
    gemm_ms, gemm_n = timing["gemm"]
    n_img = 3 * size
    total_flops = needed_flops_per_image(arch, tokens) * n_img
    # DRAM bytes of one attention launch from the committed `ncu --set full` captures (dram__bytes_read.sum +
    # dram__bytes_write.sum), keyed by the launch shape; the algorithmic bytes of a launch (q, k, V^T read once, output
    # written once) are 4 * B * tokens * D * 2.  Shapes without a capture (other N: fewer slices per launch) report null.
    traffic, traffic_src = None, None
    tt = ROOT / "profiles" / "attention_traffic.json"
    if tt.exists() and world == 1:
        ent = json.loads(tt.read_text()).get(f"heads{h}_batch{batch}_tokens{tokens}")
        if ent:
            traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
    sim_obj = {"ms": ms_sim, "gvoxel_per_s": size * size * (zr[1] - zr[0]) * world / (ms_sim * 1e-3) / 1e9,
               "hbm_gbs": sim_bytes / (ms_sim * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs", 6650.0),
               "frac": sim_bytes / (ms_sim * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0),
               "algorithmic_bytes": sim_bytes, "bound": "hbm",
               "what": "second half of BASELINE.json's metric: sim_lowres + sim_upsample of this rank's z-slab, timed alone"}
    out = {
        "metric": METRIC, "value": ms_dev, "unit": "ms", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_dev, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
        "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": vol_host.numel(),
                "d2h_bytes_per_step": n_feat * 2 + (lab_host.numel() + maps_host.numel()) * world,
                "bytes_note": "whole job: every rank copies 1/world of the volume in and 1/world of the fp16 feature volume plus its own "
                              "slab of uint8 maps and labels out"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"kernel": "attention_kernel (tcgen05 flash attention, hd 64)", "bound": "tensor",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None,
                     "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read + write)",
                     "traffic_source": traffic_src, "algorithmic_bytes_per_launch": 4 * batch * tokens * d * 2, "peak_source": peak_src, "avg_launch_ms": att_avg_ms, "launches_timed": int(att_n),
                     "share_of_step": att_ms / (ms_dev * args.steps) if ms_dev else None,
                     "gemm_share_of_step": gemm_ms / (ms_dev * args.steps) if ms_dev else None,
                     "similarity": sim_obj},
        "vit_tflops_needed": total_flops / world / (ms_dev * 1e-3) / 1e12 * world,
        "similarity": sim_obj,
    }
    if not args.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_reference(args.workload)
        except Exception as e:  # the baseline is a reported number; never let it hide the GPU result
            out["cpu_baseline"] = {"error": repr(e)}
    _emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
