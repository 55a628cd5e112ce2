#!/usr/bin/env python
"""Benchmark of the vit-tf feature-volume hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2|cfg1|cfg3|tiny]

One "step" = one whole volume: raw voxels -> 3-axis ViT K-feature volume (merged, fp16) -> prototype
similarity (north-star order) for C classes -> argmax label volume.  Prints ONE JSON line (rank 0):
  value  ms per volume, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e    the same through the public API from a pinned HOST volume: H2D of the volume, D2H of the fp16
         feature volume (what infer.py saves) and of the uint8 label volume inside the timed region
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
def cpu_reference(workload, budget_s=20.0):
    """The reference's CPU path (oracle port, fp32, all host threads) on a bounded sample of the
    workload, extrapolated linearly to ms per volume (slices are independent units)."""
    from oracle import dino_vit, feature_volume as ofv, similarity as osim
    from vittf_b200 import synth
    size, arch, fos, n_cls, per_cls, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    vol, _ = synth.ct_volume(size, n_shells=n_cls, seed=0)
    ann = synth.annotations(size, n_cls, per_cls, seed=0)
    model = dino_vit.build(arch, seed=0)
    im_sz, f_sz = ofv.image_sizes(tuple(vol.shape), 8, fos)
    # (i) ViT: time n images per axis, batch 1 like the reference default (infer.py:302)
    t_img, n_img = 0.0, 0
    for ax in ("z", "y", "x"):
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
    # (ii) similarity, NS composition, a few output z-slabs
    feats, _ = synth.class_features(ARCH[arch][0], f_sz, n_cls, seed=0, dtype=torch.float16)
    pts = torch.cat(list(ann.values()))
    protos = osim.sample_prototypes(feats.float(), osim.rel_coords(pts, vol.shape), "bilinear")
    offs = [0]
    for v in ann.values():
        offs.append(offs[-1] + v.size(0))
    slab = 4
    t0 = time.perf_counter()
    osim.ns_composite(feats, protos, offs, (size, size, size), slab=slab, z_range=(size // 2, size // 2 + slab))
    sim_ms = (time.perf_counter() - t0) * (size / slab) * 1e3
    return {"value": vit_ms + sim_ms, "unit": "ms", "cores": cores, "kind": "port",
            "sample": f"ViT: {n_img} slice images of {im_sz[0]}x{im_sz[1]} (batch 1) extrapolated x{3 * size}/{n_img}; "
                      f"similarity: {slab} of {size} output z-planes extrapolated; fp32, {cores} threads",
            "vit_ms": vit_ms, "similarity_ms": sim_ms}


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
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=0, help="slices per ViT forward (default: per workload)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    size, arch, fos, n_cls, per_cls, batch = WORKLOADS[args.workload]
    if args.batch > 0:
        batch = args.batch
    config = {"workload": WORKLOAD_TEXT[args.workload], "volume": f"{size}^3 uint8", "backbone": arch,
              "feature_output_size": fos, "classes": n_cls, "prototypes": n_cls * per_cls, "slice_batch": batch,
              "parallelism": f"slices sharded over {world} GPU(s), z-slab similarity" if world > 1 else "single GPU",
              "l2": "per-step working set (K-feature staging 0.8 GB/axis, 537 MB maps) exceeds the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, args.steps)
        vals = [cpu_reference(args.workload, budget_s=60.0 / (steps + max(0, args.warmup))) for _ in range(max(0, min(args.warmup, 1)) + steps)][-steps:]
        v = sum(x["value"] for x in vals) / len(vals)
        cb = dict(vals[-1])
        cb["value"] = v
        _emit(({"impl": "reference", "metric": "ms per volume end-to-end (ViT feats + similarity)", "value": v,
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

    feat_host = torch.empty((ARCH[arch][0],) + tuple(f_sz), dtype=torch.float16).pin_memory()
    lab_host = None

    def step_e2e():
        nonlocal lab_host
        v = vol_host.to(dev, non_blocking=True)                                  # H2D of the raw volume
        feats, sims, labels, zr = pipeline.volume_to_similarity(v, model, ann, 8, fos, batch, rank=rank, world=world)
        if lab_host is None:
            lab_host = torch.empty(labels.shape, dtype=torch.uint8).pin_memory()
        if rank == 0:
            feat_host.copy_(feats, non_blocking=True)                           # what infer.py saves
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
    # DRAM bytes of one attention launch from the committed `ncu --set full` capture (profiles/r1_ncu_attention_v7.json:
    # dram__bytes_read.sum + dram__bytes_write.sum = 610.4 + 182.7 MB at ViT-S/8, 64 slices of 4097 tokens); the algorithmic
    # bytes of that launch (q, k, V^T read once, output written once) are 4 * B * tokens * D * 2 = 805 MB
    traffic = 793.1e6 if (arch == "vits8" and batch == 64 and tokens == 4097 and world == 1) else None
    out = {
        "metric": "ms per volume end-to-end (ViT feats + similarity)", "value": ms_dev, "unit": "ms", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_dev, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
        "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": vol_host.numel(),
                "d2h_bytes_per_step": feat_host.numel() * 2 + (lab_host.numel() if lab_host is not None else 0)},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"kernel": "attention_kernel (tcgen05 flash attention, hd 64)", "bound": "tensor",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved / peak_tf) if achieved else None,
                     "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read + write)",
                     "traffic_source": "profiles/r1_ncu_attention_v7.json" if traffic else None, "peak_source": peak_src, "avg_launch_ms": att_avg_ms, "launches_timed": int(att_n),
                     "share_of_step": att_ms / (ms_dev * args.steps) if ms_dev else None,
                     "gemm_share_of_step": gemm_ms / (ms_dev * args.steps) if ms_dev else None},
        "vit_tflops_needed": total_flops / world / (ms_dev * 1e-3) / 1e12 * world,
        "similarity": {"ms": ms_sim, "gvoxel_per_s": size * size * (zr[1] - zr[0]) * world / (ms_sim * 1e-3) / 1e9,
                       "hbm_gbs": sim_bytes / (ms_sim * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs", 6650.0),
                       "frac": sim_bytes / (ms_sim * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0)},
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
